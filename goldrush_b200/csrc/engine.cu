// libgoldrush_b200 — context, device memory management and the C ABI (include/goldrush_b200.h)
// over the kernels in kernels_*.cuh.  One context = one CUDA device, one stream, one host thread.
#include "common.cuh"
#include "kernels_decode.cuh"
#include "kernels_filter.cuh"
#include "kernels_ntcard.cuh"
#include "kernels_select.cuh"
#include "batch_common.cuh"
#include "kernels_query.cuh"
#include "kernels_commit.cuh"
#include "comm.cuh"
#include "kernels_probe.cuh"
#include "kernels_polish.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace {
thread_local std::string g_create_error;
const uint64_t kSeedBase[4] = { 0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL,
                                0x20323ed082572324ULL, 0x295549f54be24456ULL };
}


// ---- process-wide device allocation cache (declared in common.cuh) ---------------------------
#include <map>
#include <mutex>
#include <unordered_map>
namespace {
struct GrbPool
{
  std::mutex mu;
  struct Block
  {
    size_t bytes;
    int device;
  };
  std::unordered_map<void*, Block> live;                     // handed out
  std::multimap<std::pair<int, size_t>, void*> idle;         // (device, bytes) -> block
  size_t idle_bytes = 0;
  bool enabled = true;
  GrbPool()
  {
    if (const char* e = getenv("GRB_POOL")) {
      enabled = strcmp(e, "0") != 0;
    }
  }
  static size_t round_up(size_t b)
  {
    const size_t g = b >= (1u << 20) ? (2u << 20) : 512; // 2 MiB pages for anything big
    return (std::max<size_t>(b, 1) + g - 1) / g * g;
  }
  void trim_locked(int device)
  {
    for (auto it = idle.begin(); it != idle.end();) {
      if (device < 0 || it->first.first == device) {
        cudaFree(it->second);
        idle_bytes -= it->first.second;
        it = idle.erase(it);
      } else {
        ++it;
      }
    }
  }
};
GrbPool&
pool()
{
  static GrbPool* p = new GrbPool; // leaked on purpose: no CUDA calls during static destruction
  return *p;
}
} // namespace

cudaError_t
grb_pool_alloc(void** out, size_t bytes)
{
  GrbPool& P = pool();
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t need = GrbPool::round_up(bytes);
  std::lock_guard<std::mutex> lk(P.mu);
  if (P.enabled) {
    // smallest idle block that fits without wasting more than a quarter of it
    auto it = P.idle.lower_bound({ dev, need });
    if (it != P.idle.end() && it->first.first == dev && it->first.second <= need + need / 4 + (2u << 20)) {
      *out = it->second;
      P.live[*out] = GrbPool::Block{ it->first.second, dev };
      P.idle_bytes -= it->first.second;
      P.idle.erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, need);
  if (e != cudaSuccess && P.idle_bytes) {
    cudaGetLastError();
    P.trim_locked(dev); // give the cache back to the driver and try once more
    e = cudaMalloc(out, need);
  }
  if (e == cudaSuccess) {
    P.live[*out] = GrbPool::Block{ need, dev };
  }
  return e;
}

void
grb_pool_free(void* p)
{
  if (!p) {
    return;
  }
  GrbPool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  auto it = P.live.find(p);
  if (it == P.live.end()) {
    cudaFree(p);
    return;
  }
  const GrbPool::Block b = it->second;
  P.live.erase(it);
  if (P.enabled) {
    P.idle.insert({ { b.device, b.bytes }, p });
    P.idle_bytes += b.bytes;
  } else {
    cudaFree(p);
  }
}

extern "C" void
grb_release_cached_memory(void)
{
  GrbPool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  P.trim_locked(-1);
}

extern "C" uint64_t
grb_cached_memory_bytes(void)
{
  GrbPool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  return P.idle_bytes;
}

// pinned + mapped upload arenas of destroyed contexts, kept for the next one (all of one size)
static std::mutex g_arena_mu;
static std::vector<void*> g_arena_idle;
static void*
grb_arena_take(size_t cap)
{
  {
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (!g_arena_idle.empty()) {
      void* p = g_arena_idle.back();
      g_arena_idle.pop_back();
      return p;
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
static void
grb_arena_give(void* p)
{
  std::lock_guard<std::mutex> lk(g_arena_mu);
  g_arena_idle.push_back(p);
}

// Host ranges page-locked by grb_host_pin, in pieces of kPinPiece bytes each registered on its own:
// one cudaMemcpyAsync must not span two registrations (or a registered and an unregistered part),
// so the ingest copies are cut at the piece boundaries.
static const size_t kPinPiece = (size_t)1 << 30;
struct GrbPin
{
  const char* base;
  size_t bytes;
};
static std::mutex g_pin_mu;
static std::vector<GrbPin> g_pins;

// host <-> device copy of n bytes whose HOST side is [host, host + n), cut at the piece boundaries
// of a range page-locked by grb_host_pin
static cudaError_t
grb_copy_host(void* dev, const char* host, size_t n, bool to_device, cudaStream_t s)
{
  GrbPin pin{ nullptr, 0 };
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (const GrbPin& q : g_pins) {
      if (host + n > q.base && host < q.base + q.bytes) {
        pin = q;
        break;
      }
    }
  }
  size_t done = 0;
  while (done < n) {
    const char* at = host + done;
    size_t len = n - done;
    if (pin.base) {
      if (at < pin.base) {
        len = std::min<size_t>(len, (size_t)(pin.base - at));
      } else if (at < pin.base + pin.bytes) {
        const size_t in_piece = kPinPiece - (size_t)(at - pin.base) % kPinPiece;
        len = std::min(len, std::min<size_t>(in_piece, (size_t)(pin.base + pin.bytes - at)));
      }
    }
    const cudaError_t e = to_device
                            ? cudaMemcpyAsync((char*)dev + done, at, len, cudaMemcpyHostToDevice, s)
                            : cudaMemcpyAsync(const_cast<char*>(at), (const char*)dev + done, len, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) {
      return e;
    }
    done += len;
  }
  return cudaSuccess;
}

static cudaError_t
grb_copy_h2d(void* dst, const char* src, size_t n, cudaStream_t s)
{
  return grb_copy_host(dst, src, n, true, s);
}

struct grb_ctx
{
  grb_params p{};
  std::vector<std::string> seeds;
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_ms = 0;
  uint64_t launches = 0;
  int sm_count = 148;
  uint64_t tile_frames = 0; // frames of a full tile: tile_length + kmer_size - span of pattern 0

  GrbSeedTables h_seed{};
  GrbSeedTables* d_seed = nullptr;
  ulonglong2* d_gtab = nullptr; // grouped half-hash tables: L[ng * 256] then R[ng * 256]
  uint32_t gt_groups = 0;

  // ---- read store ----
  uint64_t n_reads = 0;
  uint64_t n_words = 0;       // packed words in use
  uint64_t ingested_bytes = 0; // running offset of the concatenated input
  uint64_t origin_bytes = 0;   // offset of this rank's first byte within the whole input
  uint64_t own_first = 0, own_count = 0; // after grb_reads_allgather: the reads this rank ingested
  bool gathered = false;
  std::vector<uint32_t> h_len;
  std::vector<uint64_t> h_word_off;
  std::vector<uint8_t> h_flags;
  std::vector<grb_read_meta> h_meta;
  DevBuf<uint64_t> d_bases;
  DevBuf<uint32_t> d_nmask;
  DevBuf<uint64_t> d_word_off;
  DevBuf<uint32_t> d_len;
  DevBuf<uint8_t> d_flags;
  bool flags_dirty = false;
  // ingest read-ahead (grb_reads_readahead): the next chunk's bytes travel over PCIe on a second
  // stream while the current chunk's decode kernels run
  const char* ra_base = nullptr; // host range the caller will ingest chunk by chunk
  size_t ra_total = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ra_ready = nullptr, ra_free = nullptr;
  DevBuf<uint8_t> d_raw2;
  const char* pf_host = nullptr; // host range now (being) copied into d_raw2
  size_t pf_len = 0, pf_dev_off = 0;
  int ra_mapped = -1;            // 1: the host range is device-readable (pinned + mapped)
  const uint8_t* ra_dev_base = nullptr;
  // ingest scratch
  DevBuf<uint8_t> d_raw;
  DevBuf<uint32_t> d_blk_cnt;
  DevBuf<uint64_t> d_blk_off;
  DevBuf<uint32_t> d_nl;
  DevBuf<grb_read_meta> d_meta;
  DevBuf<uint32_t> d_wpr;
  DevBuf<uint64_t> d_wpr_off;
  DevBuf<uint32_t> d_err;

  // ---- filter ----
  GrbFilterDev filt{};
  bool filter_alloc = false, finalized = false;
  uint64_t blocks_cap = 0, slots_cap = 0; // bytes allocated behind filt.blocks / filt.slots
  DevBuf<uint32_t> d_chunk_read;
  DevBuf<uint64_t> d_chunk_first;
  DevBuf<uint32_t> fill_lists, fill_cursor; // partitioned fill (kernels_filter.cuh)
  bool fill_attr = false;

  // ---- selection loop ----
  GrbSelParams prm{};
  GrbSelScratch sc{};
  GrbSelState* d_state = nullptr;
  bool sel_init = false;
  bool sel_finished = false;
  uint64_t sc_tiles = 0; // capacity of the per-read scratch, in tiles
  uint64_t sc_tab = 0;   // capacity of the insert table, entries
  DevBuf<uint64_t> b_stash;
  DevBuf<uint32_t> b_best_id, b_best_count, b_n_cand, b_cand_id, b_cand_cnt, b_tile_id, b_snap;
  DevBuf<uint8_t> b_tile_as;
  DevBuf<GrbReadPlan> b_plan;
  DevBuf<uint64_t> b_tab_key, b_tab_mask;
  DevBuf<grb_decision> d_dec;
  size_t query_smem = 0, query2_smem = 0;
  // batch engine (kernels_query.cuh + kernels_commit.cuh); GRB_ENGINE=serial keeps the
  // one-read-at-a-time loop of kernels_select.cuh as the in-tree cross-check
  bool batch_mode = true;
  bool batch_reads_fixed = false; // GRB_BATCH_READS given: no adaptation to the genome size
  uint32_t batch_reads = 320;  // reads per speculative batch (A/B on B200: 64..800, profiles/README.md)
  uint32_t batch_tiles = 8192; // tile budget per batch (a single longer read still forms a batch)
  DevBuf<uint64_t> bb_stash;
  DevBuf<uint32_t> bb_best_id, bb_best_count, bb_hits, bb_miss;
  DevBuf<uint32_t> bb_uq, bb_nu, bb_cm, bb_sp_nas;
  DevBuf<GrbReadPlan> bb_sp_plan;
  DevBuf<uint32_t> bb_sp_adv, bb_rd_hits, bb_rd_miss, bb_rd_q;
  // chunk descriptors, two sets: chunk k + 1 is described and queued while chunk k still runs
  DevBuf<uint64_t> bb_read_idx[2], bb_cm_off[2];
  DevBuf<uint32_t> bb_tile_first[2], bb_tile_read[2];
  DevBuf<uint64_t> bb_dec_idx[2];
  int bb_set = 0;                 // the set the batches being launched read
  GrbSelState* h_state = nullptr; // pinned: loop state as of the end of each of the two chunks in flight
  cudaEvent_t ev_chunk[2] = { nullptr, nullptr };
  // per-batch query outputs and the probe index of the commit
  uint64_t b2_cap_tiles = 0, b2_ix_entries = 0;
  uint32_t b2_cap_reads = 0;
  bool b2_attr = false;
  size_t b2_smem_max = 0;
  DevBuf<uint32_t> b2_vk, b2_vc, b2_ix_sidx, b2_counters, b2_c_slot, b2_c_probe, b2_c_next, b2_c_sidx;
  DevBuf<unsigned long long> b2_ix;
  DevBuf<uint32_t> b2_fbits, b2_fl_n, b2_fl, b2_fr;
  DevBuf<GrbReadPlan> b2_plan_out;
  // ordered commit (kernels_commit.cuh)
  bool b3_attr = false, b3_stage_attr = false;
  uint32_t b3_ctas = 0, b3_dcap = 0;
  uint32_t b3_bs = 512; // threads per CTA of k3_fix: 512, or 256 (two CTAs per SM; measured slower)
  uint64_t b3_cap_tiles = 0, b3_cmat_cap = 0;
  DevBuf<GrbShared3> b3_shared;
  DevBuf<uint32_t> b3_bm, b3_cand, b3_ix_cnt;
  uint32_t b3_bm_bits = 0;
  DevBuf<uint32_t> b3_c_pos, b3_m_fill, b3_m_key, b3_m_ci, b3_m_seen, b3_np_adv, b3_np_nas, b3_cmat_g;
  DevBuf<int32_t> b3_np_dh, b3_d_vals;
  DevBuf<GrbReadPlan> b3_np;
  DevBuf<GrbFixCtl> b3_ctl;
  DevBuf<unsigned long long> b3_d_keys, b3_barrier;
  DevBuf<uint32_t> b3_dbg; // GRB_FIX_DEBUG=<file>: per-read records of the commit's re-validation phase

  // ---- multi-GPU (comm.cuh): the process-wide NCCL communicator, when this context's device is
  // the one it was created on and it spans more than one rank ----
  GrbComm* comm = nullptr;
  bool shard_query = false;  // pass-2 query sharding: off unless GRB_SHARD_QUERY=1 (the
                             // exchange costs what it saves)
  DevBuf<uint64_t> comm_tmp; // all-gather landing zone of the pass-1 OR-reduce
  int fail_nccl(ncclResult_t r, const char* what)
  {
    err = std::string(what) + ": " + comm->api.GetErrorString(r);
    return GRB_ERR_CUDA;
  }

  // ---- per-kernel-class device timing (grb_profile_enable / grb_kernel_time) ----
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_free;
  struct ProfPending
  {
    int k;
    cudaEvent_t a, b;
  };
  std::vector<ProfPending> prof_pending;
  cudaEvent_t prof_open = nullptr;
  double prof_ms[GRB_K_COUNT] = {};
  uint64_t prof_n[GRB_K_COUNT] = {};
  cudaEvent_t prof_event()
  {
    cudaEvent_t e = nullptr;
    if (!prof_free.empty()) {
      e = prof_free.back();
      prof_free.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
  void kbegin()
  {
    if (prof_on) {
      prof_open = prof_event();
      cudaEventRecord(prof_open, stream);
    }
  }
  void kend(int k, uint64_t n_launches = 1)
  {
    launches += n_launches;
    if (prof_on) {
      cudaEvent_t b = prof_event();
      cudaEventRecord(b, stream);
      prof_pending.push_back(ProfPending{ k, prof_open, b });
      prof_n[k] += n_launches;
      prof_open = nullptr;
    }
  }
  // call after the stream has been synchronised
  void kflush()
  {
    for (const ProfPending& q : prof_pending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, q.a, q.b) == cudaSuccess) {
        prof_ms[q.k] += ms;
      }
      prof_free.push_back(q.a);
      prof_free.push_back(q.b);
    }
    prof_pending.clear();
  }

  int fail(int code, const std::string& msg)
  {
    err = msg;
    return code;
  }

  // ---- small host -> device uploads that must not queue behind the ingest read-ahead ----
  // While a 1 GB read-ahead copy occupies the host-to-device copy engine, every cudaMemcpyAsync of
  // a descriptor array on the compute stream waits for it (measured: zero overlap).  During that
  // window descriptors are staged in a pinned, mapped arena and pulled by a small kernel instead.
  uint8_t* up_host = nullptr;      // pinned + mapped arena
  uint8_t* up_dev = nullptr;       // its device view
  size_t up_cap = 0, up_used = 0;  // up_used is reset whenever the stream has been synchronised
  cudaError_t upload(void* dst, const void* src, size_t bytes, cudaStream_t s)
  {
    if (bytes == 0) {
      return cudaSuccess;
    }
    static const bool arena_on = !(getenv("GRB_UPLOAD_ARENA") && strcmp(getenv("GRB_UPLOAD_ARENA"), "0") == 0);
    const bool window = arena_on && pf_host != nullptr; // a read-ahead copy may be in flight
    if (window && !up_host) {
      const size_t cap = (size_t)32 << 20;
      up_host = (uint8_t*)grb_arena_take(cap); // page-locking 32 MB costs tens of ms: kept per process
      if (up_host && cudaHostGetDevicePointer((void**)&up_dev, up_host, 0) == cudaSuccess) {
        up_cap = cap;
      } else {
        cudaGetLastError();
        up_host = nullptr;
      }
    }
    const size_t at = ((up_used + 15) & ~(size_t)15) + ((uintptr_t)dst & 15);
    if (!window || !up_host || at + bytes > up_cap) {
      return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    }
    memcpy(up_host + at, src, bytes);
    up_used = at + bytes;
    k_copy_host<<<8, 256, 0, s>>>((uint8_t*)dst, up_dev + at, bytes);
    launches += 1;
    return cudaGetLastError();
  }
  GrbReadsDev reads_dev() const
  {
    return GrbReadsDev{ d_bases.p, d_nmask.p, d_word_off.p, d_len.p, d_flags.p };
  }
  void tic() { cudaEventRecord(ev0, stream); }
  void toc()
  {
    cudaEventRecord(ev1, stream);
    cudaEventSynchronize(ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    last_ms = ms;
  }
};

namespace {

int
build_seed_tables(grb_ctx* c)
{
  const auto& s = c->seeds;
  const std::string& p0 = s[0];
  if (p0.size() % 2 != 0 || p0.empty()) {
    return c->fail(GRB_ERR_ARG, "seed pattern 0 must have even length (left + right halves)");
  }
  const size_t half = p0.size() / 2;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] != p0.substr(0, half) + std::string(i, '0') + p0.substr(half)) {
      return c->fail(GRB_ERR_ARG, "seed patterns must be left + i zeros + right "
                                  "(spaced_seeds.cpp:63-66)");
    }
    for (char ch : s[i]) {
      if (ch != '0' && ch != '1') {
        return c->fail(GRB_ERR_ARG, "seed pattern holds a character other than 0/1");
      }
    }
  }
  if (p0.size() + s.size() - 1 > GRB_MAX_SPAN || half > 32) {
    return c->fail(GRB_ERR_ARG, "seed span + patterns - 1 exceeds 64 bases");
  }
  if (s.size() > GRB_MAX_PATTERNS) {
    return c->fail(GRB_ERR_ARG, "more than 8 seed patterns");
  }
  GrbSeedTables& t = c->h_seed;
  memset(&t, 0, sizeof t);
  t.k = (uint32_t)p0.size();
  t.h = (uint32_t)s.size();
  t.half = (uint32_t)half;
  for (size_t q = 0; q < p0.size(); ++q) {
    if (p0[q] == '1') {
      if (t.n_care >= GRB_MAX_WEIGHT) {
        return c->fail(GRB_ERR_ARG, "seed weight exceeds 64");
      }
      const unsigned j = t.n_care++;
      t.care[j] = (uint8_t)q;
      if (q < half) {
        t.n_left = t.n_care;
      }
      for (unsigned b = 0; b < 4; ++b) {
        t.fwd[j][b] = grb_srol_any(kSeedBase[b], t.k - 1 - (unsigned)q);
        t.rev[j][b] = grb_srol_any(kSeedBase[3 - b], (unsigned)q);
      }
    }
  }
  if (t.n_care == 0) {
    return c->fail(GRB_ERR_ARG, "seed pattern has no care position");
  }
  return GRB_OK;
}

// grouped half-hash tables (nthash.cuh): entry [g][v] XORs the contributions of the care positions
// whose window offset lies in [4g, 4g + 4), for the 4 bases encoded in byte v
std::vector<ulonglong2>
build_group_tables(const GrbSeedTables& t, uint32_t* n_groups)
{
  const uint32_t ng = (t.half + 3) / 4;
  std::vector<ulonglong2> tab((size_t)2 * ng * 256, make_ulonglong2(0, 0));
  for (uint32_t j = 0; j < t.n_care; ++j) {
    const bool left = j < t.n_left;
    const uint32_t o = left ? t.care[j] : t.care[j] - t.half;
    const uint32_t g = o / 4, sh = 2 * (o % 4);
    ulonglong2* dst = tab.data() + (size_t)(left ? 0 : ng) * 256 + (size_t)g * 256;
    for (uint32_t v = 0; v < 256; ++v) {
      const uint32_t b = (v >> sh) & 3u;
      dst[v].x ^= t.fwd[j][b];
      dst[v].y ^= t.rev[j][b];
    }
  }
  *n_groups = ng;
  return tab;
}

inline unsigned
grid_for(uint64_t n, unsigned bs, unsigned cap)
{
  const uint64_t g = (n + bs - 1) / bs;
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(g, cap));
}

uint64_t
next_pow2(uint64_t x)
{
  uint64_t p = 1;
  while (p < x) {
    p <<= 1;
  }
  return p;
}

} // namespace

extern "C" {

void
grb_params_default(grb_params* p)
{
  memset(p, 0, sizeof *p);
  p->assigned_max = 1;
  p->unassigned_min = 5;
  p->tile_length = 1000;
  p->block_size = 10;
  p->min_length = 20000;
  p->hash_num = 3;
  p->occupancy = 0.1;
  p->ratio = 0.9;
  p->max_paths = 1;
  p->threshold = 10;
  p->phred_min = 0;
  p->phred_delta = 5;
}

const char*
grb_last_error(const grb_ctx* ctx)
{
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

uint64_t
grb_launch_count(const grb_ctx* ctx)
{
  return ctx->launches;
}

double
grb_last_device_ms(const grb_ctx* ctx)
{
  return ctx->last_ms;
}

int
grb_create(const grb_params* p, grb_ctx** out)
{
  *out = nullptr;
  grb_ctx* c = new grb_ctx;
  auto bail = [&](int code, const std::string& m) {
    g_create_error = m;
    delete c;
    return code;
  };
  c->p = *p;
  if (!p->seeds || p->hash_num == 0) {
    return bail(GRB_ERR_ARG, "grb_create: seeds / hash_num missing");
  }
  for (uint64_t i = 0; i < p->hash_num; ++i) {
    c->seeds.emplace_back(p->seeds[i]);
  }
  c->p.seeds = nullptr;
  if (p->tile_length == 0 || p->block_size == 0 || p->kmer_size == 0) {
    return bail(GRB_ERR_ARG, "grb_create: tile_length, block_size and kmer_size must be non-zero");
  }
  int rc = build_seed_tables(c);
  if (rc != GRB_OK) {
    return bail(rc, c->err);
  }
  if (p->tile_length < c->h_seed.k + c->h_seed.h - 1) {
    return bail(GRB_ERR_ARG, "grb_create: tile_length shorter than the longest seed span");
  }
  // make_seed_pattern builds pattern 0 from two halves of k / 2 positions (spaced_seeds.cpp:28,
  // 58-60): its span is k - 1 for an odd -k, and the reference then dies on
  // assert(m_sseeds[0].size() == kmerSize) (MIBloomFilter.hpp:180) once pass 1 is done.  Refused
  // here, before any work.
  if (p->kmer_size != c->h_seed.k) {
    return bail(GRB_ERR_ARG, "grb_create: seed pattern 0 must span kmer_size (an odd -k gives seeds "
                             "of span k - 1, which the reference rejects: MIBloomFilter.hpp:180)");
  }
  c->tile_frames = p->tile_length + p->kmer_size - c->h_seed.k;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    return bail(GRB_ERR_CUDA, std::string("no usable CUDA device: ") +
                                (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  if (p->device < 0 || p->device >= ndev) {
    return bail(GRB_ERR_ARG, "grb_create: device ordinal out of range");
  }
  c->device = p->device;
  if ((e = cudaSetDevice(c->device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess) {
    return bail(GRB_ERR_CUDA, std::string("CUDA init: ") + cudaGetErrorString(e));
  }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
  // reads per speculative batch: two per SM, so that the per-read phase of the commit kernel (one
  // CTA per read, one CTA per SM) splits evenly; 296 on B200 (A/B 64 .. 800: profiles/README.md)
  c->batch_reads = (uint32_t)std::min(1024, std::max(64, 2 * c->sm_count));
  // L2 -> DRAM fetch granularity: measured on B200 (profiles/sector_roofline_r01.json, ncu
  // dram__sectors_read of k2_query) every L2 miss moves a whole 128-byte line whatever this limit
  // says, and asking for 32 bytes only slowed the L2-sized structures down.  The device default is
  // kept; GRB_L2_FETCH=32|64|128 sets the limit for A/B measurements.
  if (const char* g = getenv("GRB_L2_FETCH")) {
    const long v = strtol(g, nullptr, 10);
    if (v == 32 || v == 64 || v == 128) {
      if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v) != cudaSuccess) {
        cudaGetLastError(); // not fatal: the limit is only a performance hint
      }
    }
  }
  // GRB_ENGINE=serial keeps the one-read-at-a-time loop (kernels_select.cuh) for A/B checks;
  // GRB_BATCH_READS overrides the speculative batch size
  if (const char* e = getenv("GRB_ENGINE")) {
    c->batch_mode = strcmp(e, "serial") != 0;
  }
  if (const char* e = getenv("GRB_BATCH_TILES")) { // tile budget of a batch (26-bit probe index caps it)
    const long v = strtol(e, nullptr, 10);
    if (v > 0 && v <= (1 << 20)) {
      c->batch_tiles = (uint32_t)v;
    }
  }
  if (const char* e = getenv("GRB_BATCH_READS")) {
    const long v = strtol(e, nullptr, 10);
    if (v > 0 && v <= 65536) {
      c->batch_reads = (uint32_t)v;
      c->batch_reads_fixed = true;
    }
  }
  {
    GrbComm& g = grb_comm();
    std::lock_guard<std::mutex> lk(g.mu);
    const char* off = getenv("GRB_COMM");
    if (g.comm && g.world > 1 && g.device == c->device && !(off && strcmp(off, "0") == 0)) {
      c->comm = &g;
      // Each batch's speculative query can be sharded over the ranks (GRB_SHARD_QUERY=1), but it is
      // not the default: measured on B200 (profiles/README.md, round 2) the all-gather of the
      // per-tile results costs what the sharded query saves at every N (cfg2, 8 GPUs: query 134 ->
      // 23 ms, exchange 113 ms), and a collective inside every batch couples the ranks' launch threads:
      // on a host with four cores per rank the human-scale run took 30.8 s end to end with the query
      // sharded and 18.2 s with every rank querying the whole batch.
      const char* sq = getenv("GRB_SHARD_QUERY");
      c->shard_query = sq ? strcmp(sq, "1") == 0 : false;
    }
  }
  if ((e = grb_pool_alloc((void**)&c->d_seed, sizeof(GrbSeedTables))) != cudaSuccess ||
      (e = cudaMemcpy(c->d_seed, &c->h_seed, sizeof(GrbSeedTables), cudaMemcpyHostToDevice)) !=
        cudaSuccess) {
    return bail(GRB_ERR_CUDA, std::string("seed table upload: ") + cudaGetErrorString(e));
  }
  {
    const std::vector<ulonglong2> gt = build_group_tables(c->h_seed, &c->gt_groups);
    if ((e = grb_pool_alloc((void**)&c->d_gtab, gt.size() * sizeof(ulonglong2))) != cudaSuccess ||
        (e = cudaMemcpy(c->d_gtab, gt.data(), gt.size() * sizeof(ulonglong2),
                        cudaMemcpyHostToDevice)) != cudaSuccess) {
      return bail(GRB_ERR_CUDA, std::string("group table upload: ") + cudaGetErrorString(e));
    }
  }
  // 10^(-q/10) with glibc pow, indexed by the raw quality byte (calc_phred_average.cpp:17-20;
  // `char` is signed on this platform, as in the reference build)
  double tab[256];
  for (int b = 0; b < 256; ++b) {
    const int phred_score = (int)((char)b - 33);
    tab[b] = pow(10.0, -phred_score / 10.0);
  }
  if ((e = cudaMemcpyToSymbol(c_delog, tab, sizeof tab)) != cudaSuccess) {
    return bail(GRB_ERR_CUDA, std::string("phred table upload: ") + cudaGetErrorString(e));
  }
  *out = c;
  return GRB_OK;
}

void
grb_destroy(grb_ctx* c)
{
  if (!c) {
    return;
  }
  cudaSetDevice(c->device);
  if (c->stream) {
    cudaStreamSynchronize(c->stream);
  }
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->ra_ready);
    cudaEventDestroy(c->ra_free);
  }
  if (c->up_host) {
    grb_arena_give(c->up_host);
  }
  if (c->h_state) {
    cudaFreeHost(c->h_state);
    cudaEventDestroy(c->ev_chunk[0]);
    cudaEventDestroy(c->ev_chunk[1]);
  }
  if (c->b3_dbg.p && getenv("GRB_FIX_DEBUG")) {
    std::vector<uint32_t> h(c->b3_dbg.cap);
    if (cudaMemcpy(h.data(), c->b3_dbg.p, h.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess) {
      if (FILE* f = fopen(getenv("GRB_FIX_DEBUG"), "wb")) {
        const size_t n = std::min<size_t>(h[0], (h.size() - 4) / 4);
        fwrite(h.data(), 4, 4 + 4 * n, f);
        fclose(f);
      }
    }
  }
  grb_pool_free(c->filt.blocks);
  grb_pool_free(c->filt.slots);
  grb_pool_free(c->d_state);
  grb_pool_free(c->d_seed);
  grb_pool_free(c->d_gtab);
  if (c->ev0) {
    cudaEventDestroy(c->ev0);
  }
  if (c->ev1) {
    cudaEventDestroy(c->ev1);
  }
  c->kflush();
  for (cudaEvent_t e : c->prof_free) {
    cudaEventDestroy(e);
  }
  cudaStream_t s = c->stream;
  delete c;
  if (s) {
    cudaStreamDestroy(s);
  }
}

int
grb_sync(grb_ctx* c)
{
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->kflush();
  return GRB_OK;
}

int
grb_stream(grb_ctx* c, void** cuda_stream)
{
  *cuda_stream = (void*)c->stream;
  return GRB_OK;
}

// ---- multi-GPU: process-wide communicator (one process per GPU) ----
int
grb_comm_unique_id(uint8_t* out128)
{
  GrbComm& g = grb_comm();
  std::lock_guard<std::mutex> lk(g.mu);
  if (!grb_nccl_load(g)) {
    g_create_error = g.err;
    return GRB_ERR_STATE;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in the ABI this binds");
  ncclUniqueId id;
  const ncclResult_t r = g.api.GetUniqueId(&id);
  if (r != ncclSuccess) {
    g_create_error = g.err = std::string("ncclGetUniqueId: ") + g.api.GetErrorString(r);
    return GRB_ERR_CUDA;
  }
  memcpy(out128, &id, 128);
  return GRB_OK;
}

int
grb_comm_init(const uint8_t* id128, int rank, int world, int device)
{
  GrbComm& g = grb_comm();
  std::lock_guard<std::mutex> lk(g.mu);
  if (world < 1 || rank < 0 || rank >= world || !id128) {
    g_create_error = g.err = "grb_comm_init: rank / world out of range";
    return GRB_ERR_ARG;
  }
  if (g.comm) {
    g_create_error = g.err = "grb_comm_init: a communicator already exists (grb_comm_destroy first)";
    return GRB_ERR_STATE;
  }
  if (!grb_nccl_load(g)) {
    g_create_error = g.err;
    return GRB_ERR_STATE;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_error = g.err = std::string("grb_comm_init: ") + cudaGetErrorString(e);
    return GRB_ERR_CUDA;
  }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  const ncclResult_t r = g.api.CommInitRank(&g.comm, world, id, rank);
  if (r != ncclSuccess) {
    g.comm = nullptr;
    g_create_error = g.err = std::string("ncclCommInitRank: ") + g.api.GetErrorString(r);
    return GRB_ERR_CUDA;
  }
  g.rank = rank;
  g.world = world;
  g.device = device;
  return GRB_OK;
}

void
grb_comm_destroy(void)
{
  GrbComm& g = grb_comm();
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.comm) {
    g.api.CommDestroy(g.comm);
  }
  g.comm = nullptr;
  g.rank = 0;
  g.world = 1;
  g.device = -1;
}

int
grb_comm_info(const grb_ctx* c, int* rank, int* world)
{
  // c == NULL: the process-wide communicator; else what this context actually uses
  if (c) {
    *rank = c->comm ? c->comm->rank : 0;
    *world = c->comm ? c->comm->world : 1;
  } else {
    GrbComm& g = grb_comm();
    std::lock_guard<std::mutex> lk(g.mu);
    *rank = g.rank;
    *world = g.world;
  }
  return GRB_OK;
}

int
grb_query_sharded(const grb_ctx* c)
{
  return c->comm && c->shard_query ? 1 : 0;
}

int
grb_profile_enable(grb_ctx* c, int on)
{
  cudaSetDevice(c->device);
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->kflush();
  c->prof_on = on != 0;
  for (int k = 0; k < GRB_K_COUNT; ++k) {
    c->prof_ms[k] = 0;
    c->prof_n[k] = 0;
  }
  return GRB_OK;
}

int
grb_kernel_time(grb_ctx* c, int kclass, double* ms, uint64_t* n_launches)
{
  if (kclass < 0 || kclass >= GRB_K_COUNT) {
    return c->fail(GRB_ERR_ARG, "grb_kernel_time: unknown kernel class");
  }
  cudaSetDevice(c->device);
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->kflush();
  *ms = c->prof_ms[kclass];
  *n_launches = c->prof_n[kclass];
  return GRB_OK;
}

// ------------------------------------------------------------------------------------------
// K1: ingest
// ------------------------------------------------------------------------------------------
uint64_t
grb_reads_count(const grb_ctx* c)
{
  return c->n_reads;
}

void
grb_reads_clear(grb_ctx* c)
{
  c->n_reads = 0;
  c->n_words = 0;
  c->ingested_bytes = 0;
  c->origin_bytes = 0;
  c->own_first = c->own_count = 0;
  c->gathered = false;
  c->h_len.clear();
  c->h_word_off.clear();
  c->h_flags.clear();
  c->h_meta.clear();
  // a read-ahead copy still in flight belongs to the input that is being dropped
  if (c->copy_stream) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->copy_stream);
  }
  c->ra_base = nullptr;
  c->ra_total = 0;
  c->pf_host = nullptr;
  c->ra_mapped = -1;
}

int
grb_host_pin(void* p, size_t n, size_t* pinned_bytes)
{
  size_t done = 0;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    while (done < n) {
      const size_t len = std::min(kPinPiece, n - done);
      if (cudaHostRegister((char*)p + done, len, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError(); // not an error of the caller's later CUDA calls
        break;
      }
      done += len;
    }
    if (done) {
      g_pins.push_back(GrbPin{ (const char*)p, done });
    }
  }
  if (pinned_bytes) {
    *pinned_bytes = done;
  }
  return GRB_OK;
}

void
grb_host_unpin(void* p, size_t pinned_bytes)
{
  std::lock_guard<std::mutex> lk(g_pin_mu);
  for (size_t done = 0; done < pinned_bytes; done += kPinPiece) {
    cudaHostUnregister((char*)p + done);
  }
  cudaGetLastError();
  for (size_t i = 0; i < g_pins.size(); ++i) {
    if (g_pins[i].base == (const char*)p) {
      g_pins.erase(g_pins.begin() + i);
      break;
    }
  }
}

int
grb_reads_reserve(grb_ctx* c, uint64_t fastq_bytes)
{
  cudaSetDevice(c->device);
  if (c->n_reads != 0) {
    return GRB_OK; // only a hint, and only before the first record
  }
  // a record is at least 2 * len + 6 bytes, so bases <= fastq_bytes / 2; +1 word per read for the
  // word alignment of every read (reads of 64 bases or more: <= bases / 64 extra words)
  const uint64_t words = fastq_bytes / 64 + fastq_bytes / 128 + 64;
  GRB_CUDA(c, c->d_bases.reserve_exact(c->n_words + words + 4, c->stream));
  GRB_CUDA(c, c->d_nmask.reserve_exact(c->n_words + words + 4, c->stream));
  return GRB_OK;
}

int
grb_reads_readahead(grb_ctx* c, const char* base, size_t total)
{
  cudaSetDevice(c->device);
  if (c->copy_stream) {
    GRB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  }
  c->ra_base = base;
  c->ra_total = base ? total : 0;
  c->pf_host = nullptr;
  c->ra_mapped = -1;
  return GRB_OK;
}

int
grb_reads_ingest_fastq(grb_ctx* c, const char* bytes, size_t n, int final, size_t* consumed)
{
  cudaSetDevice(c->device);
  *consumed = 0;
  if (n == 0) {
    return GRB_OK;
  }
  if (n > (1ull << 31)) {
    return c->fail(GRB_ERR_ARG, "grb_reads_ingest_fastq: chunk larger than 2 GiB");
  }
  if (c->gathered) {
    return c->fail(GRB_ERR_STATE, "grb_reads_ingest_fastq after grb_reads_allgather");
  }
  if (c->ingested_bytes == c->origin_bytes && c->n_reads == 0 && bytes[0] != '@') {
    return c->fail(GRB_ERR_FORMAT, "Gold Path requires fastq format");
  }
  cudaStream_t s = c->stream;
  c->tic();
  const uint64_t padded = (n + 4095) / 4096 * 4096;
  const uint64_t n_blk = padded / 4096;
  GRB_CUDA(c, c->d_raw.reserve(padded, 0, s));
  GRB_CUDA(c, cudaMemsetAsync(c->d_raw.p + (padded - 4096), 0, 4096, s));
  if (c->pf_host && bytes >= c->pf_host && bytes + n <= c->pf_host + c->pf_len) {
    // the chunk was read ahead: a device-to-device copy re-aligns it to the start of d_raw
    GRB_CUDA(c, cudaStreamWaitEvent(s, c->ra_ready, 0));
    GRB_CUDA(c, cudaMemcpyAsync(c->d_raw.p, c->d_raw2.p + c->pf_dev_off + (bytes - c->pf_host), n,
                                cudaMemcpyDeviceToDevice, s));
  } else {
    GRB_CUDA(c, grb_copy_h2d(c->d_raw.p, bytes, n, s));
  }
  c->pf_host = nullptr;
  if (c->ra_base && bytes >= c->ra_base && bytes + n < c->ra_base + c->ra_total) {
    // start the next chunk's copy: it begins a little before this chunk's end because the caller
    // re-sends the last partial record (records longer than the overlap fall back to a plain copy)
    const size_t kOverlap = (size_t)16 << 20;
    const char* lo = bytes + n - std::min(kOverlap, n);
    const size_t len = std::min<size_t>((size_t)(c->ra_base + c->ra_total - lo), n + kOverlap);
    if (!c->copy_stream) {
      GRB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      GRB_CUDA(c, cudaEventCreateWithFlags(&c->ra_ready, cudaEventDisableTiming));
      GRB_CUDA(c, cudaEventCreateWithFlags(&c->ra_free, cudaEventDisableTiming));
    }
    if (len > c->d_raw2.cap) {
      GRB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
      c->d_raw2.release();
      GRB_CUDA(c, c->d_raw2.reserve(n + 2 * kOverlap + 16, 0, s));
    }
    GRB_CUDA(c, cudaEventRecord(c->ra_free, s)); // d_raw2 is free once the copy above has run
    GRB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ra_free, 0));
    // default: the copy engine (full PCIe rate; the descriptor uploads of the launches that run
    // meanwhile go through grb_ctx::upload).  GRB_READAHEAD_KERNEL=1 pulls the bytes with a copy
    // kernel from pinned + mapped host memory instead (measured slower: ~40 GB/s against 51)
    c->pf_dev_off = 0;
    const uint8_t* dev_view = nullptr;
    if (c->ra_mapped < 0) {
      cudaPointerAttributes pa{};
      const char* e = getenv("GRB_READAHEAD_KERNEL");
      c->ra_mapped = 0;
      if ((e && strcmp(e, "1") == 0) && cudaPointerGetAttributes(&pa, c->ra_base) == cudaSuccess &&
          pa.type == cudaMemoryTypeHost && pa.devicePointer) {
        c->ra_mapped = 1;
        c->ra_dev_base = (const uint8_t*)pa.devicePointer;
      }
      cudaGetLastError();
    }
    if (c->ra_mapped == 1) {
      dev_view = c->ra_dev_base + (lo - c->ra_base);
      c->pf_dev_off = (size_t)((uintptr_t)dev_view & 15);
      k_copy_host<<<64, 256, 0, c->copy_stream>>>(c->d_raw2.p + c->pf_dev_off, dev_view, len);
      c->launches += 1;
    } else {
      GRB_CUDA(c, grb_copy_h2d(c->d_raw2.p, lo, len, c->copy_stream));
    }
    GRB_CUDA(c, cudaEventRecord(c->ra_ready, c->copy_stream));
    c->pf_host = lo;
    c->pf_len = len;
  }
  GRB_CUDA(c, c->d_blk_cnt.reserve(n_blk, 0, s));
  GRB_CUDA(c, c->d_blk_off.reserve(n_blk + 1, 0, s));
  k_nl_count<<<(unsigned)n_blk, 256, 0, s>>>(c->d_raw.p, c->d_blk_cnt.p);
  k_scan_u32<<<1, 1024, 0, s>>>(c->d_blk_cnt.p, c->d_blk_off.p, n_blk);
  c->launches += 2;
  uint64_t n_nl = 0;
  GRB_CUDA(c, cudaMemcpyAsync(&n_nl, c->d_blk_off.p + n_blk, 8, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  uint64_t n_lines = n_nl;
  if (final && bytes[n - 1] != '\n') {
    ++n_lines; // last line without a trailing newline
  }
  const uint64_t n_rec = n_lines / 4;
  if (n_rec == 0) {
    c->toc();
    return GRB_OK;
  }
  GRB_CUDA(c, c->d_nl.reserve(n_nl + 1, 0, s));
  k_nl_write<<<(unsigned)n_blk, 256, 0, s>>>(c->d_raw.p, c->d_blk_off.p, c->d_nl.p);
  GRB_CUDA(c, c->d_meta.reserve(n_rec, 0, s));
  GRB_CUDA(c, c->d_wpr.reserve(n_rec, 0, s));
  GRB_CUDA(c, c->d_wpr_off.reserve(n_rec + 1, 0, s));
  GRB_CUDA(c, c->d_err.reserve(1, 0, s));
  GRB_CUDA(c, cudaMemsetAsync(c->d_err.p, 0, 4, s));
  k_records<<<grid_for(n_rec, 128, 1u << 30), 128, 0, s>>>(c->d_raw.p, n, c->d_nl.p, n_nl, n_rec,
                                                          c->ingested_bytes, c->d_meta.p,
                                                          c->d_wpr.p, c->d_err.p);
  k_scan_u32<<<1, 1024, 0, s>>>(c->d_wpr.p, c->d_wpr_off.p, n_rec);
  c->launches += 3;
  uint64_t new_words = 0;
  uint32_t fmt_err = 0;
  uint32_t last_nl = 0;
  GRB_CUDA(c, cudaMemcpyAsync(&new_words, c->d_wpr_off.p + n_rec, 8, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaMemcpyAsync(&fmt_err, c->d_err.p, 4, cudaMemcpyDeviceToHost, s));
  if (4 * n_rec - 1 < n_nl) {
    GRB_CUDA(c, cudaMemcpyAsync(&last_nl, c->d_nl.p + (4 * n_rec - 1), 4, cudaMemcpyDeviceToHost, s));
  }
  GRB_CUDA(c, cudaStreamSynchronize(s));
  if (fmt_err) {
    return c->fail(GRB_ERR_FORMAT, "malformed FASTQ record (expected '@' header and '+' separator)");
  }
  // grow the store (+3 words so a 64-base window may read past the last read)
  GRB_CUDA(c, c->d_bases.reserve(c->n_words + new_words + 4, c->n_words, s));
  GRB_CUDA(c, c->d_nmask.reserve(c->n_words + new_words + 4, c->n_words, s));
  GRB_CUDA(c, c->d_word_off.reserve(c->n_reads + n_rec, c->n_reads, s));
  GRB_CUDA(c, c->d_len.reserve(c->n_reads + n_rec, c->n_reads, s));
  GRB_CUDA(c, c->d_flags.reserve(c->n_reads + n_rec, c->n_reads, s));
  k_pack<<<(unsigned)n_rec, 256, 0, s>>>(c->d_raw.p, c->ingested_bytes, c->d_meta.p,
                                         c->d_wpr_off.p, c->n_words, c->d_bases.p, c->d_nmask.p);
  GRB_CUDA(c, cudaMemsetAsync(c->d_bases.p + c->n_words + new_words, 0, 4 * 8, s));
  k_phred<<<grid_for(n_rec, 64, 1u << 30), 64, 0, s>>>(c->d_raw.p, c->ingested_bytes, c->d_meta.p,
                                                      n_rec);
  c->launches += 2;
  const size_t old = c->h_meta.size();
  c->h_meta.resize(old + n_rec);
  std::vector<uint64_t> woff(n_rec + 1);
  GRB_CUDA(c, cudaMemcpyAsync(c->h_meta.data() + old, c->d_meta.p, n_rec * sizeof(grb_read_meta),
                              cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaMemcpyAsync(woff.data(), c->d_wpr_off.p, (n_rec + 1) * 8, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  c->h_len.resize(old + n_rec);
  c->h_word_off.resize(old + n_rec);
  c->h_flags.resize(old + n_rec, 0);
  for (uint64_t i = 0; i < n_rec; ++i) {
    c->h_len[old + i] = c->h_meta[old + i].len;
    c->h_word_off[old + i] = c->n_words + woff[i];
    c->h_flags[old + i] = c->h_meta[old + i].non_acgt ? 4 : 0;
  }
  c->up_used = 0; // the stream was synchronised above: earlier arena contents have been consumed
  GRB_CUDA(c, c->upload(c->d_word_off.p + old, c->h_word_off.data() + old, n_rec * 8, s));
  GRB_CUDA(c, c->upload(c->d_len.p + old, c->h_len.data() + old, n_rec * 4, s));
  GRB_CUDA(c, c->upload(c->d_flags.p + old, c->h_flags.data() + old, n_rec, s));
  c->toc();
  GRB_CUDA(c, cudaGetLastError());
  c->n_reads += n_rec;
  c->n_words += new_words;
  *consumed = (4 * n_rec - 1 < n_nl) ? (size_t)last_nl + 1 : n;
  c->ingested_bytes += *consumed;
  return GRB_OK;
}

int
grb_reads_set_origin(grb_ctx* c, uint64_t byte_offset)
{
  if (c->n_reads != 0 || c->ingested_bytes != c->origin_bytes) {
    return c->fail(GRB_ERR_STATE, "grb_reads_set_origin after the first grb_reads_ingest_fastq");
  }
  c->origin_bytes = byte_offset;
  c->ingested_bytes = byte_offset;
  return GRB_OK;
}

int
grb_reads_own_range(const grb_ctx* c, uint64_t* first, uint64_t* count)
{
  *first = c->gathered ? c->own_first : 0;
  *count = c->gathered ? c->own_count : c->n_reads;
  return GRB_OK;
}

// Every rank holds the reads of its own slice of the input; afterwards every rank holds all of
// them, in rank (= file) order.  The packed stores travel device to device (one NCCL broadcast per
// rank and array, grouped); the per-read metadata is small and goes host -> device -> all -> host.
int
grb_reads_allgather(grb_ctx* c)
{
  cudaSetDevice(c->device);
  if (c->gathered) {
    return c->fail(GRB_ERR_STATE, "grb_reads_allgather called twice");
  }
  c->gathered = true;
  c->own_first = 0;
  c->own_count = c->n_reads;
  if (!c->comm) {
    return GRB_OK;
  }
  GrbComm& g = *c->comm;
  const int W = g.world, me = g.rank;
  cudaStream_t s = c->stream;
  if (c->copy_stream) {
    GRB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  }
  c->tic();
  // ---- how much every rank brings ----
  DevBuf<uint64_t> d_cnt;
  GRB_CUDA(c, d_cnt.reserve(2 * (size_t)W, 0, s));
  const uint64_t mine[2] = { c->n_reads, c->n_words };
  GRB_CUDA(c, cudaMemcpyAsync(d_cnt.p + 2 * me, mine, 16, cudaMemcpyHostToDevice, s));
  ncclResult_t r = g.api.AllGather(d_cnt.p + 2 * me, d_cnt.p, 2, ncclUint64, g.comm, s);
  if (r != ncclSuccess) {
    return c->fail_nccl(r, "ncclAllGather (read counts)");
  }
  std::vector<uint64_t> cnt(2 * (size_t)W);
  GRB_CUDA(c, cudaMemcpyAsync(cnt.data(), d_cnt.p, cnt.size() * 8, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  std::vector<uint64_t> r0(W + 1, 0), w0(W + 1, 0);
  for (int q = 0; q < W; ++q) {
    r0[q + 1] = r0[q] + cnt[2 * q];
    w0[q + 1] = w0[q] + cnt[2 * q + 1];
  }
  const uint64_t n_reads = r0[W], n_words = w0[W];
  // ---- packed bases + masks, device to device ----
  DevBuf<uint64_t> nb;
  DevBuf<uint32_t> nm;
  DevBuf<grb_read_meta> dm;
  GRB_CUDA(c, nb.reserve_exact(n_words + 4, s));
  GRB_CUDA(c, nm.reserve_exact(n_words + 4, s));
  GRB_CUDA(c, dm.reserve_exact(std::max<uint64_t>(1, n_reads), s));
  if (c->n_words) {
    GRB_CUDA(c, cudaMemcpyAsync(nb.p + w0[me], c->d_bases.p, c->n_words * 8, cudaMemcpyDeviceToDevice, s));
    GRB_CUDA(c, cudaMemcpyAsync(nm.p + w0[me], c->d_nmask.p, c->n_words * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (c->n_reads) {
    GRB_CUDA(c, cudaMemcpyAsync(dm.p + r0[me], c->h_meta.data(), c->n_reads * sizeof(grb_read_meta),
                                cudaMemcpyHostToDevice, s));
  }
  GRB_CUDA(c, cudaMemsetAsync(nb.p + n_words, 0, 4 * 8, s));
  GRB_CUDA(c, cudaMemsetAsync(nm.p + n_words, 0, 4 * 4, s));
  r = g.api.GroupStart();
  for (int q = 0; q < W && r == ncclSuccess; ++q) {
    if (cnt[2 * q + 1]) {
      r = g.api.Broadcast(nb.p + w0[q], nb.p + w0[q], cnt[2 * q + 1] * 8, ncclUint8, q, g.comm, s);
      if (r == ncclSuccess) {
        r = g.api.Broadcast(nm.p + w0[q], nm.p + w0[q], cnt[2 * q + 1] * 4, ncclUint8, q, g.comm, s);
      }
    }
    if (r == ncclSuccess && cnt[2 * q]) {
      r = g.api.Broadcast(dm.p + r0[q], dm.p + r0[q], cnt[2 * q] * sizeof(grb_read_meta), ncclUint8, q,
                          g.comm, s);
    }
  }
  const ncclResult_t r2 = g.api.GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess) {
    return c->fail_nccl(r != ncclSuccess ? r : r2, "ncclBroadcast (read store)");
  }
  c->h_meta.resize(n_reads);
  if (n_reads) {
    GRB_CUDA(c, cudaMemcpyAsync(c->h_meta.data(), dm.p, n_reads * sizeof(grb_read_meta),
                                cudaMemcpyDeviceToHost, s));
  }
  GRB_CUDA(c, cudaStreamSynchronize(s));
  std::swap(c->d_bases.p, nb.p);
  std::swap(c->d_bases.cap, nb.cap);
  std::swap(c->d_nmask.p, nm.p);
  std::swap(c->d_nmask.cap, nm.cap);
  // ---- per-read arrays: every read starts on a word boundary, so the offsets follow from the lengths ----
  c->h_len.resize(n_reads);
  c->h_word_off.resize(n_reads);
  c->h_flags.assign(n_reads, 0);
  uint64_t w = 0;
  for (uint64_t i = 0; i < n_reads; ++i) {
    c->h_len[i] = c->h_meta[i].len;
    c->h_word_off[i] = w;
    c->h_flags[i] = c->h_meta[i].non_acgt ? 4 : 0;
    w += (c->h_meta[i].len + 31) / 32;
  }
  if (w != n_words) {
    return c->fail(GRB_ERR_STATE, "grb_reads_allgather: word count does not match the read lengths");
  }
  GRB_CUDA(c, c->d_word_off.reserve(std::max<uint64_t>(1, n_reads), 0, s));
  GRB_CUDA(c, c->d_len.reserve(std::max<uint64_t>(1, n_reads), 0, s));
  GRB_CUDA(c, c->d_flags.reserve(std::max<uint64_t>(1, n_reads), 0, s));
  if (n_reads) {
    GRB_CUDA(c, cudaMemcpyAsync(c->d_word_off.p, c->h_word_off.data(), n_reads * 8, cudaMemcpyHostToDevice, s));
    GRB_CUDA(c, cudaMemcpyAsync(c->d_len.p, c->h_len.data(), n_reads * 4, cudaMemcpyHostToDevice, s));
    GRB_CUDA(c, cudaMemcpyAsync(c->d_flags.p, c->h_flags.data(), n_reads, cudaMemcpyHostToDevice, s));
  }
  c->toc();
  c->own_first = r0[me];
  c->own_count = cnt[2 * me];
  c->n_reads = n_reads;
  c->n_words = n_words;
  c->launches += 1;
  return GRB_OK;
}

int
grb_comm_allgather_host(grb_ctx* c, const void* send, uint64_t n, void* out, uint64_t out_cap,
                        uint64_t* sizes)
{
  cudaSetDevice(c->device);
  if (!c->comm) {
    if (n > out_cap) {
      return c->fail(GRB_ERR_ARG, "grb_comm_allgather_host: output buffer too small");
    }
    memcpy(out, send, n);
    if (sizes) {
      sizes[0] = n;
    }
    return GRB_OK;
  }
  GrbComm& g = *c->comm;
  const int W = g.world, me = g.rank;
  cudaStream_t s = c->stream;
  DevBuf<uint64_t> d_cnt;
  GRB_CUDA(c, d_cnt.reserve((size_t)W, 0, s));
  GRB_CUDA(c, cudaMemcpyAsync(d_cnt.p + me, &n, 8, cudaMemcpyHostToDevice, s));
  ncclResult_t r = g.api.AllGather(d_cnt.p + me, d_cnt.p, 1, ncclUint64, g.comm, s);
  if (r != ncclSuccess) {
    return c->fail_nccl(r, "ncclAllGather (sizes)");
  }
  std::vector<uint64_t> cnt((size_t)W);
  GRB_CUDA(c, cudaMemcpyAsync(cnt.data(), d_cnt.p, (size_t)W * 8, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  uint64_t total = 0, most = 0;
  for (int q = 0; q < W; ++q) {
    total += cnt[q];
    most = std::max(most, cnt[q]);
    if (sizes) {
      sizes[q] = cnt[q];
    }
  }
  if (total > out_cap) {
    return c->fail(GRB_ERR_ARG, "grb_comm_allgather_host: output buffer too small");
  }
  if (most == 0) {
    return GRB_OK;
  }
  const uint64_t pad = (most + 15) / 16 * 16;
  DevBuf<uint8_t> stage;
  GRB_CUDA(c, stage.reserve_exact(pad * (uint64_t)W, s));
  if (n) {
    GRB_CUDA(c, cudaMemcpyAsync(stage.p + pad * me, send, n, cudaMemcpyHostToDevice, s));
  }
  r = g.api.AllGather(stage.p + pad * me, stage.p, pad, ncclUint8, g.comm, s);
  if (r != ncclSuccess) {
    return c->fail_nccl(r, "ncclAllGather (host strings)");
  }
  uint64_t at = 0;
  for (int q = 0; q < W; ++q) {
    if (cnt[q]) {
      GRB_CUDA(c, cudaMemcpyAsync((char*)out + at, stage.p + pad * q, cnt[q], cudaMemcpyDeviceToHost, s));
    }
    at += cnt[q];
  }
  GRB_CUDA(c, cudaStreamSynchronize(s));
  return GRB_OK;
}

// (f4) GoldPolish targeted Bloom filters: jobs = (batch, k) pairs, run in waves sized to the device
// memory their counting filters need (10 MiB each in the reference)
int
grb_polish_fill_batches(grb_ctx* c, const grb_polish_params* p, uint32_t n_batches,
                        const uint64_t* batch_first, const char* seqs, const uint64_t* seq_off,
                        const uint32_t* thresholds, uint8_t* out_bfs)
{
  cudaSetDevice(c->device);
  if (!p || p->n_k == 0 || p->hash_num == 0 || p->hash_num > 8 || p->cbf_bytes < 2 || p->bf_bytes < 1) {
    return c->fail(GRB_ERR_ARG, "grb_polish_fill_batches: n_k, hash_num in 1..8, cbf_bytes >= 2, bf_bytes >= 1");
  }
  if (n_batches == 0) {
    return GRB_OK;
  }
  cudaStream_t s = c->stream;
  const uint64_t n_reads = batch_first[n_batches];
  const uint64_t n_bytes = seq_off[n_reads];
  for (uint64_t r = 0; r < n_reads; ++r) {
    if (thresholds[r] < 4) { // utils.cpp:105-107
      return c->fail(GRB_ERR_ARG, "grb_polish_fill_batches: kmer_threshold must be greater than or equal to 4");
    }
  }
  DevBuf<char> d_seqs;
  DevBuf<uint64_t> d_off, d_first;
  DevBuf<uint32_t> d_thr, d_k;
  DevBuf<int> d_status;
  GRB_CUDA(c, d_seqs.reserve_exact(std::max<uint64_t>(n_bytes, 1), s));
  GRB_CUDA(c, d_off.reserve_exact(n_reads + 1, s));
  GRB_CUDA(c, d_first.reserve_exact((uint64_t)n_batches + 1, s));
  GRB_CUDA(c, d_thr.reserve_exact(std::max<uint64_t>(n_reads, 1), s));
  GRB_CUDA(c, d_k.reserve_exact(p->n_k, s));
  GRB_CUDA(c, d_status.reserve_exact(1, s));
  GRB_CUDA(c, cudaMemcpyAsync(d_seqs.p, seqs, n_bytes, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemcpyAsync(d_off.p, seq_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemcpyAsync(d_first.p, batch_first, ((uint64_t)n_batches + 1) * 8, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemcpyAsync(d_thr.p, thresholds, n_reads * 4, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemcpyAsync(d_k.p, p->k_values, (uint64_t)p->n_k * 4, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemsetAsync(d_status.p, 0, 4, s));
  // per-k hash tables of the warp kernel (polish_core.h); k above 64 falls back to the thread kernel
  bool k_too_long = false;
  for (uint32_t i = 0; i < p->n_k; ++i) {
    k_too_long = k_too_long || p->k_values[i] > GRB_P_MAX_K || p->k_values[i] == 0;
  }
  DevBuf<GrbPolishPair> d_tables;
  if (!k_too_long) {
    std::vector<GrbPolishPair> tabs((size_t)p->n_k * GRB_P_GROUPS * 256);
    for (uint32_t i = 0; i < p->n_k; ++i) {
      grb_p_build_table(p->k_values[i], tabs.data() + (size_t)i * GRB_P_GROUPS * 256);
    }
    GRB_CUDA(c, d_tables.reserve_exact(tabs.size(), s));
    GRB_CUDA(c, cudaMemcpyAsync(d_tables.p, tabs.data(), tabs.size() * sizeof(GrbPolishPair), cudaMemcpyHostToDevice, s));
    GRB_CUDA(c, cudaStreamSynchronize(s));
  }
  size_t free_b = 0, total_b = 0;
  GRB_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  const uint64_t per_job = p->cbf_bytes + 2 * p->bf_bytes;
  const uint64_t n_jobs = (uint64_t)n_batches * p->n_k;
  // as many jobs in flight as three quarters of the free device memory hold: a job is one warp,
  // so the number of resident counting filters IS the parallelism.  A call that would fit one wave
  // is still cut in up to four (of at least 1024 jobs, which fill the part) so that a wave's Bloom
  // filters travel back to the host under the next wave's kernel.
  const uint64_t budget = free_b / 4 * 3;
  uint64_t wave = std::max<uint64_t>(1, std::min<uint64_t>(n_jobs, budget / per_job));
  const uint64_t kMinWave = 1024;
  if (n_jobs >= 2 * kMinWave) {
    wave = std::min(wave, std::max(kMinWave, (n_jobs + 3) / 4));
  }
  DevBuf<uint8_t> d_cbf, d_bf[2];
  GRB_CUDA(c, d_cbf.reserve_exact(wave * p->cbf_bytes, s));
  GRB_CUDA(c, d_bf[0].reserve_exact(wave * p->bf_bytes, s));
  GRB_CUDA(c, d_bf[1].reserve_exact(wave * p->bf_bytes, s));
  if (!c->copy_stream) {
    GRB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    GRB_CUDA(c, cudaEventCreateWithFlags(&c->ra_ready, cudaEventDisableTiming));
    GRB_CUDA(c, cudaEventCreateWithFlags(&c->ra_free, cudaEventDisableTiming));
  }
  const cudaStream_t cs = c->copy_stream;
  struct Events
  {
    cudaEvent_t filled[2] = { nullptr, nullptr }, copied[2] = { nullptr, nullptr };
    ~Events()
    {
      for (int i = 0; i < 2; ++i) {
        if (filled[i]) {
          cudaEventDestroy(filled[i]);
        }
        if (copied[i]) {
          cudaEventDestroy(copied[i]);
        }
      }
    }
  } ev;
  for (int i = 0; i < 2; ++i) {
    GRB_CUDA(c, cudaEventCreateWithFlags(&ev.filled[i], cudaEventDisableTiming));
    GRB_CUDA(c, cudaEventCreateWithFlags(&ev.copied[i], cudaEventDisableTiming));
  }
  c->tic();
  uint64_t wi = 0;
  for (uint64_t j0 = 0; j0 < n_jobs; j0 += wave, ++wi) {
    const uint64_t n = std::min(wave, n_jobs - j0);
    const int b = (int)(wi & 1);
    if (wi >= 2) {
      GRB_CUDA(c, cudaStreamWaitEvent(s, ev.copied[b], 0)); // the copy of wave wi - 2 has left this buffer
    }
    GRB_CUDA(c, cudaMemsetAsync(d_cbf.p, 0, n * p->cbf_bytes, s));
    GRB_CUDA(c, cudaMemsetAsync(d_bf[b].p, 0, n * p->bf_bytes, s));
    GrbPolishWave w;
    w.seqs = d_seqs.p;
    w.off = d_off.p;
    w.thr = d_thr.p;
    w.batch_first = d_first.p;
    w.k_values = d_k.p;
    w.n_k = p->n_k;
    w.hash_num = p->hash_num;
    w.job0 = (uint32_t)j0;
    w.n_jobs = (uint32_t)n;
    w.cbf = d_cbf.p;
    w.bf = d_bf[b].p;
    w.cbf_bytes = p->cbf_bytes;
    w.cbf_inv = (uint64_t)(((unsigned __int128)1 << 64) / p->cbf_bytes);
    w.bf_bytes = p->bf_bytes;
    w.bf_inv = (uint64_t)(((unsigned __int128)1 << 64) / (p->bf_bytes * 8));
    w.status = d_status.p;
    // GRB_POLISH=thread: one sequential thread per job (the first version, kept as the cross-check);
    // default: one warp per job, 32 k-mers at a time when their counters are pairwise distinct
    static const bool per_thread = getenv("GRB_POLISH") && strcmp(getenv("GRB_POLISH"), "thread") == 0;
    if (per_thread || k_too_long || p->cbf_bytes >= (1ull << 56) || (p->bf_bytes & 3) != 0) {
      k_polish_fill<<<(unsigned)((n + 31) / 32), 32, 0, s>>>(w);
    } else {
      k_polish_fill_warp<<<(unsigned)((n + GRB_PW_WARPS - 1) / GRB_PW_WARPS), 32 * GRB_PW_WARPS, 0, s>>>(
        w, d_tables.p);
    }
    c->launches += 1;
    GRB_CUDA(c, cudaEventRecord(ev.filled[b], s));
    GRB_CUDA(c, cudaStreamWaitEvent(cs, ev.filled[b], 0));
    GRB_CUDA(c, grb_copy_host(d_bf[b].p, (const char*)out_bfs + j0 * p->bf_bytes, n * p->bf_bytes, false, cs));
    GRB_CUDA(c, cudaEventRecord(ev.copied[b], cs));
  }
  for (int b = 0; b < 2 && (uint64_t)b < wi; ++b) {
    GRB_CUDA(c, cudaStreamWaitEvent(s, ev.copied[b], 0));
  }
  c->toc();
  int status = 0;
  GRB_CUDA(c, cudaMemcpy(&status, d_status.p, 4, cudaMemcpyDeviceToHost));
  GRB_CUDA(c, cudaGetLastError());
  if (status != 0) {
    return c->fail(GRB_ERR_ARG, "grb_polish_fill_batches: a job refused its input");
  }
  return GRB_OK;
}

int
grb_comm_exchange_host(grb_ctx* c, const grb_host_msg* sends, uint32_t n_sends, const grb_host_msg* recvs,
                       uint32_t n_recvs)
{
  cudaSetDevice(c->device);
  const int me = c->comm ? c->comm->rank : 0;
  // messages to oneself: matched in list order, copied on the host
  {
    uint32_t ri = 0;
    for (uint32_t i = 0; i < n_sends; ++i) {
      if (sends[i].peer != me) {
        continue;
      }
      while (ri < n_recvs && recvs[ri].peer != me) {
        ++ri;
      }
      if (ri == n_recvs || recvs[ri].bytes != sends[i].bytes) {
        return c->fail(GRB_ERR_ARG, "grb_comm_exchange_host: unmatched message to self");
      }
      memcpy(recvs[ri].ptr, sends[i].ptr, sends[i].bytes);
      ++ri;
    }
  }
  if (!c->comm) {
    return GRB_OK;
  }
  GrbComm& g = *c->comm;
  cudaStream_t s = c->stream;
  uint64_t out_bytes = 0, in_bytes = 0;
  for (uint32_t i = 0; i < n_sends; ++i) {
    out_bytes += sends[i].peer != me ? (sends[i].bytes + 15) / 16 * 16 : 0;
  }
  for (uint32_t i = 0; i < n_recvs; ++i) {
    in_bytes += recvs[i].peer != me ? (recvs[i].bytes + 15) / 16 * 16 : 0;
  }
  DevBuf<uint8_t> d_out, d_in;
  GRB_CUDA(c, d_out.reserve_exact(std::max<uint64_t>(out_bytes, 16), s));
  GRB_CUDA(c, d_in.reserve_exact(std::max<uint64_t>(in_bytes, 16), s));
  uint64_t at = 0;
  for (uint32_t i = 0; i < n_sends; ++i) {
    if (sends[i].peer != me && sends[i].bytes) {
      GRB_CUDA(c, cudaMemcpyAsync(d_out.p + at, sends[i].ptr, sends[i].bytes, cudaMemcpyHostToDevice, s));
    }
    at += sends[i].peer != me ? (sends[i].bytes + 15) / 16 * 16 : 0;
  }
  ncclResult_t r = g.api.GroupStart();
  at = 0;
  for (uint32_t i = 0; i < n_sends && r == ncclSuccess; ++i) {
    if (sends[i].peer != me) {
      if (sends[i].bytes) {
        r = g.api.Send(d_out.p + at, sends[i].bytes, ncclUint8, sends[i].peer, g.comm, s);
      }
      at += (sends[i].bytes + 15) / 16 * 16;
    }
  }
  at = 0;
  for (uint32_t i = 0; i < n_recvs && r == ncclSuccess; ++i) {
    if (recvs[i].peer != me) {
      if (recvs[i].bytes) {
        r = g.api.Recv(d_in.p + at, recvs[i].bytes, ncclUint8, recvs[i].peer, g.comm, s);
      }
      at += (recvs[i].bytes + 15) / 16 * 16;
    }
  }
  const ncclResult_t r2 = g.api.GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess) {
    return c->fail_nccl(r != ncclSuccess ? r : r2, "ncclSend / ncclRecv (host messages)");
  }
  at = 0;
  for (uint32_t i = 0; i < n_recvs; ++i) {
    if (recvs[i].peer != me) {
      if (recvs[i].bytes) {
        GRB_CUDA(c, cudaMemcpyAsync(recvs[i].ptr, d_in.p + at, recvs[i].bytes, cudaMemcpyDeviceToHost, s));
      }
      at += (recvs[i].bytes + 15) / 16 * 16;
    }
  }
  GRB_CUDA(c, cudaStreamSynchronize(s));
  c->launches += 1;
  return GRB_OK;
}

int
grb_reads_get_meta(grb_ctx* c, uint64_t first, uint64_t count, grb_read_meta* out)
{
  if (first + count > c->n_reads) {
    return c->fail(GRB_ERR_ARG, "grb_reads_get_meta: range past the end of the read store");
  }
  memcpy(out, c->h_meta.data() + first, count * sizeof(grb_read_meta));
  return GRB_OK;
}

int
grb_reads_set_flags(grb_ctx* c, uint64_t first, uint64_t count, const uint8_t* flags)
{
  cudaSetDevice(c->device);
  if (first + count > c->n_reads) {
    return c->fail(GRB_ERR_ARG, "grb_reads_set_flags: range past the end of the read store");
  }
  for (uint64_t i = 0; i < count; ++i) {
    c->h_flags[first + i] = (uint8_t)((c->h_flags[first + i] & 4u) | (flags[i] & 3u));
  }
  GRB_CUDA(c, c->upload(c->d_flags.p + first, c->h_flags.data() + first, count, c->stream));
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->up_used = 0;
  return GRB_OK;
}

int
grb_phred_sums(grb_ctx* c, const char* qual, size_t n, double* first_half_sum, double* total_sum)
{
  cudaSetDevice(c->device);
  cudaStream_t s = c->stream;
  GRB_CUDA(c, c->d_raw.reserve(n + 16, 0, s));
  GRB_CUDA(c, c->d_meta.reserve(1, 0, s));
  grb_read_meta m{};
  m.qual_off = 0;
  m.qual_len = (uint32_t)n;
  GRB_CUDA(c, cudaMemcpyAsync(c->d_raw.p, qual, n, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemcpyAsync(c->d_meta.p, &m, sizeof m, cudaMemcpyHostToDevice, s));
  k_phred<<<1, 32, 0, s>>>(c->d_raw.p, 0, c->d_meta.p, 1);
  c->launches += 1;
  GRB_CUDA(c, cudaMemcpyAsync(&m, c->d_meta.p, sizeof m, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  *first_half_sum = m.phred_first_half_sum;
  *total_sum = m.phred_total_sum;
  return GRB_OK;
}

// ------------------------------------------------------------------------------------------
// K5: ntcard
// ------------------------------------------------------------------------------------------
int
grb_estimate_cardinality(grb_ctx* c, uint64_t input_bytes, uint64_t* per_pattern, uint64_t* total)
{
  cudaSetDevice(c->device);
  cudaStream_t s = c->stream;
  const unsigned rBits = 27;
  const unsigned sBits = input_bytes < 50000000000ULL ? 7 : 11; // ntcard.hpp:182-183
  const uint64_t rBuck = 1ull << rBits;
  const unsigned h = c->h_seed.h;
  for (uint64_t i = 0; i < c->n_reads; ++i) {
    if (c->h_len[i] < c->h_seed.k + h - 1) {
      return c->fail(GRB_ERR_ARG, "SeedNtHash: sequence length is smaller than k");
    }
  }
  DevBuf<uint32_t> counters; // [h][2][rBuck], 32-bit so the uint16 wrap of ntcard.hpp:82,92 is exact
  DevBuf<uint32_t> valid;    // [n_reads][h] valid windows of reads that hold non-ACGT bytes
  DevBuf<unsigned long long> zeros;
  GRB_CUDA(c, counters.reserve((size_t)h * 2 * rBuck, 0, s));
  GRB_CUDA(c, valid.reserve(std::max<uint64_t>(1, c->n_reads * h), 0, s));
  GRB_CUDA(c, zeros.reserve(h * 2, 0, s));
  GRB_CUDA(c, cudaMemsetAsync(counters.p, 0, (size_t)h * 2 * rBuck * 4, s));
  GRB_CUDA(c, cudaMemsetAsync(valid.p, 0, std::max<uint64_t>(1, c->n_reads * h) * 4, s));
  GRB_CUDA(c, cudaMemsetAsync(zeros.p, 0, h * 2 * 8, s));
  c->tic();
  // chunk table over ALL reads (ntcard.hpp:203-205 hashes every record, unfiltered)
  std::vector<uint32_t> chunk_read;
  std::vector<uint64_t> chunk_first(c->n_reads);
  for (uint64_t r = 0; r < c->n_reads; ++r) {
    chunk_first[r] = chunk_read.size();
    const uint64_t nc = ((uint64_t)c->h_len[r] + GRB_FILL_CHUNK - 1) / GRB_FILL_CHUNK;
    chunk_read.insert(chunk_read.end(), nc, (uint32_t)r);
  }
  if (!chunk_read.empty()) {
    GRB_CUDA(c, c->d_chunk_read.reserve(chunk_read.size(), 0, s));
    GRB_CUDA(c, c->d_chunk_first.reserve(c->n_reads, 0, s));
    GRB_CUDA(c, cudaMemcpyAsync(c->d_chunk_read.p, chunk_read.data(), chunk_read.size() * 4,
                                cudaMemcpyHostToDevice, s));
    GRB_CUDA(c, cudaMemcpyAsync(c->d_chunk_first.p, chunk_first.data(), c->n_reads * 8,
                                cudaMemcpyHostToDevice, s));
    k_ntcard_count<<<grid_for(chunk_read.size(), 1, c->sm_count * 8), 256, 0, s>>>(
      c->reads_dev(), c->d_seed, c->d_chunk_read.p, c->d_chunk_first.p, chunk_read.size(),
      counters.p, valid.p, sBits, rBits);
    k_ntcard_tail<<<grid_for(c->n_reads * h, 128, 1u << 30), 128, 0, s>>>(
      c->reads_dev(), c->d_seed, c->n_reads, counters.p, valid.p, sBits, rBits);
    c->launches += 2;
  }
  k_ntcard_zeros<<<c->sm_count * 4, 256, 0, s>>>(counters.p, (uint64_t)h * 2, rBuck, zeros.p);
  c->launches += 1;
  std::vector<unsigned long long> hz(h * 2);
  GRB_CUDA(c, cudaMemcpyAsync(hz.data(), zeros.p, h * 2 * 8, cudaMemcpyDeviceToHost, s));
  c->toc();
  GRB_CUDA(c, cudaGetLastError());
  uint64_t sum = 0;
  for (unsigned i = 0; i < h; ++i) { // ntcard.hpp:127-139 compEst, only F0 is consumed (:265-270)
    const double pMean0 = ((double)hz[2 * i] + (double)hz[2 * i + 1]) / (1.0 * 2);
    const double F0Mean =
      (double)(ssize_t)((rBits * log(2) - log(pMean0)) * 1.0 * ((size_t)1 << (sBits + rBits)));
    const uint64_t f0 = (uint64_t)(size_t)F0Mean;
    if (per_pattern) {
      per_pattern[i] = f0;
    }
    sum += f0;
  }
  *total = sum;
  return GRB_OK;
}

// ------------------------------------------------------------------------------------------
// filter
// ------------------------------------------------------------------------------------------
int
grb_filter_alloc(grb_ctx* c, uint64_t filter_bits)
{
  cudaSetDevice(c->device);
  if (filter_bits < 64) {
    return c->fail(GRB_ERR_ARG, "grb_filter_alloc: filter smaller than 64 bits");
  }
  c->filt.bits = filter_bits;
  c->filt.inv = (uint64_t)((((__uint128_t)1) << 64) / filter_bits);
  c->filt.n_blocks = (filter_bits + GRB_BLK_BITS - 1) / GRB_BLK_BITS;
  c->filt.pop = 0;
  // device allocations are kept across runs of one context (cudaMalloc / cudaFree of tens of GB
  // cost more than the rank build)
  if (c->filt.n_blocks * 32 > c->blocks_cap) {
    if (c->filt.blocks) {
      cudaStreamSynchronize(c->stream);
      grb_pool_free(c->filt.blocks);
      c->filt.blocks = nullptr;
      c->blocks_cap = 0;
    }
    GRB_CUDA(c, grb_pool_alloc((void**)&c->filt.blocks, c->filt.n_blocks * 32));
    c->blocks_cap = c->filt.n_blocks * 32;
  }
  GRB_CUDA(c, cudaMemsetAsync(c->filt.blocks, 0, c->filt.n_blocks * 32, c->stream));
  c->filter_alloc = true;
  c->finalized = false;
  c->sel_init = false;
  c->sel_finished = false;
  return GRB_OK;
}

int
grb_build_bitvector_range(grb_ctx* c, uint64_t first, uint64_t count)
{
  cudaSetDevice(c->device);
  if (!c->filter_alloc || c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_build_bitvector: needs grb_filter_alloc and no finalize yet");
  }
  if (first + count > c->n_reads) {
    return c->fail(GRB_ERR_ARG, "grb_build_bitvector_range: range past the end of the read store");
  }
  cudaStream_t s = c->stream;
  std::vector<uint32_t> chunk_read;
  std::vector<uint64_t> chunk_first(c->n_reads, 0);
  const uint32_t need = c->h_seed.k + c->h_seed.h - 1;
  for (uint64_t r = first; r < first + count; ++r) {
    if (!(c->h_flags[r] & GRB_READ_PASS1)) {
      continue;
    }
    if (c->h_len[r] < need) { // btllib::SeedNtHash refuses sequences shorter than the seed
      return c->fail(GRB_ERR_ARG, "SeedNtHash: sequence length is smaller than k");
    }
    if (c->h_flags[r] & 4u) {
      return c->fail(GRB_ERR_ARG, "a read with non-ACGT bases is flagged GRB_READ_PASS1");
    }
    chunk_first[r] = chunk_read.size();
    const uint64_t nc = ((uint64_t)c->h_len[r] + GRB_FILL_CHUNK - 1) / GRB_FILL_CHUNK;
    chunk_read.insert(chunk_read.end(), nc, (uint32_t)r);
  }
  c->tic();
  if (!chunk_read.empty()) {
    GRB_CUDA(c, c->d_chunk_read.reserve(chunk_read.size(), 0, s));
    GRB_CUDA(c, c->d_chunk_first.reserve(c->n_reads, 0, s));
    GRB_CUDA(c, c->upload(c->d_chunk_read.p, chunk_read.data(), chunk_read.size() * 4, s));
    GRB_CUDA(c, c->upload(c->d_chunk_first.p, chunk_first.data(), c->n_reads * 8, s));
    c->tic();
    // Partitioned fill (kernels_filter.cuh): worth it once the vector outgrows L2; the direct
    // kernel stays for small filters, for h > 4 and for A/B runs (GRB_FILL=direct|part).
    const uint64_t h = c->h_seed.h;
    uint32_t pshift = 27;
    if (const char* e = getenv("GRB_FILL_PSHIFT")) { // tests: many partitions on a small filter
      const long v = strtol(e, nullptr, 10);
      if (v >= 10 && v <= 31) {
        pshift = (uint32_t)v;
      }
    }
    const char* bs_env = getenv("GRB_FILL_BS");
    const bool bs512 = !(bs_env && strcmp(bs_env, "1024") == 0); // measured faster on B200 (60 vs 65 ms)
    uint64_t part_limit = 1024;
    if (const char* e = getenv("GRB_FILL_MAXPART")) {
      const long v = strtol(e, nullptr, 10);
      if (v >= 1 && (uint64_t)v <= part_limit) {
        part_limit = (uint64_t)v;
      }
    }
    while (((c->filt.bits + (1ull << pshift) - 1) >> pshift) > part_limit) {
      ++pshift;
    }
    const uint32_t n_part = (uint32_t)((c->filt.bits + (1ull << pshift) - 1) >> pshift);
    const char* mode = getenv("GRB_FILL");
    bool part = h <= GRB_PART_H && pshift < 32 && n_part >= 2 && c->filt.n_blocks * 32 > (48ull << 20);
    if (mode && strcmp(mode, "direct") == 0) {
      part = false;
    }
    if (mode && strcmp(mode, "part") == 0) {
      part = h <= GRB_PART_H && pshift < 32;
    }
    if (!part) {
      c->kbegin();
      k_fill_bits<<<grid_for(chunk_read.size(), 1, c->sm_count * 16), 256, 0, s>>>(
        c->reads_dev(), c->d_seed, c->filt, c->d_chunk_read.p, c->d_chunk_first.p, chunk_read.size());
      c->kend(GRB_K_FILL);
    } else {
      // list space: GRB_FILL_SCRATCH_MB (default 2048) split evenly over the partitions; a round
      // takes as many chunks as fill a full-size partition's list to 1/1.05 of its capacity
      uint64_t scratch_mb = 2048;
      if (const char* e = getenv("GRB_FILL_SCRATCH_MB")) {
        const long v = strtol(e, nullptr, 10);
        if (v >= 16 && v <= 65536) {
          scratch_mb = (uint64_t)v;
        }
      }
      const uint64_t max_probes = (uint64_t)chunk_read.size() * GRB_FILL_CHUNK * h;
      const double share = std::min(1.0, (double)(1ull << pshift) / (double)c->filt.bits);
      uint64_t cap = std::min<uint64_t>((scratch_mb << 20) / 4 / n_part,
                                        (uint64_t)((double)max_probes * share * 1.05) + 65536 + 4096);
      cap = std::min<uint64_t>(std::max<uint64_t>(cap, 131072), 0xFFFFF000ull / n_part); // 32-bit list index
      const uint64_t round_probes =
        std::max<uint64_t>((uint64_t)((double)(cap - 65536) / 1.05 / share), GRB_FILL_CHUNK * h);
      const uint64_t round_chunks = std::max<uint64_t>(1, round_probes / (GRB_FILL_CHUNK * h));
      if (const char* e = getenv("GRB_FILL_CAP")) { // tests: force the list-overflow fallback
        const long v = strtol(e, nullptr, 10);
        if (v >= 1 && (uint64_t)v < cap) {
          cap = (uint64_t)v;
        }
      }
      GRB_CUDA(c, c->fill_lists.reserve(cap * n_part, 0, s));
      GRB_CUDA(c, c->fill_cursor.reserve(n_part + 1, 0, s));
      GrbFillPart fp{ c->fill_lists.p, c->fill_cursor.p, n_part, pshift, (uint32_t)cap, 0 };
      const size_t dyn = (size_t)c->gt_groups * 256 * 32 +
                         ((size_t)(bs512 ? 1024 : GRB_FILL_CHUNK) * h * 2 + 3 * (size_t)n_part + 1) * 4;
      const int smem_optin = (int)((size_t)GRB_MAX_GROUPS * 256 * 32 +
                                   ((size_t)GRB_FILL_CHUNK * GRB_PART_H * 2 + 3 * 1024 + 1) * 4);
      if (!c->fill_attr) {
        GRB_CUDA(c, cudaFuncSetAttribute(k_fill_part<1024, 2048, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_optin));
        GRB_CUDA(c, cudaFuncSetAttribute(k_fill_part<512, 1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_optin));
        c->fill_attr = true;
      }
      // two 512-thread CTAs per SM over 1024-position sub-chunks (default), or GRB_FILL_BS=1024:
      // one 1024-thread CTA over the whole chunk
      c->kbegin();
      uint64_t n_launch = 0;
      for (uint64_t c0 = 0; c0 < chunk_read.size(); c0 += round_chunks) {
        const uint64_t nc = std::min<uint64_t>(round_chunks, chunk_read.size() - c0);
        GRB_CUDA(c, cudaMemsetAsync(fp.cursor, 0, ((size_t)n_part + 1) * 4, s));
        if (bs512) {
          k_fill_part<512, 1024, 2><<<grid_for(nc * 2, 1, c->sm_count * 8), 512, dyn, s>>>(
            c->reads_dev(), c->d_seed, c->d_gtab, c->gt_groups, c->filt, c->d_chunk_read.p,
            c->d_chunk_first.p, c0, nc, fp);
        } else {
          k_fill_part<1024, 2048, 1><<<grid_for(nc, 1, c->sm_count * 4), 1024, dyn, s>>>(
            c->reads_dev(), c->d_seed, c->d_gtab, c->gt_groups, c->filt, c->d_chunk_read.p,
            c->d_chunk_first.p, c0, nc, fp);
        }
        k_fill_apply<<<c->sm_count * 8, 256, ((size_t)n_part + 1) * 4, s>>>(c->filt, fp);
        n_launch += 2;
      }
      c->kend(GRB_K_FILL, n_launch);
    }
  }
  c->toc();
  GRB_CUDA(c, cudaGetLastError());
  return GRB_OK;
}

// Pass 1 on W GPUs: this rank's share of the reads, then the OR-reduce of the W bit vectors
// (all-gather slice by slice + k_or_gathered; the rank word of every block is still zero here).
static int
or_reduce_bitvector(grb_ctx* c)
{
  GrbComm& g = *c->comm;
  cudaStream_t s = c->stream;
  const uint64_t n_words = c->filt.n_blocks * 4;
  const uint64_t slice = std::min<uint64_t>(n_words, (64ull << 20) / 8);
  GRB_CUDA(c, c->comm_tmp.reserve(slice * g.world, 0, s));
  c->kbegin();
  uint64_t n_launch = 0;
  for (uint64_t off = 0; off < n_words; off += slice) {
    const uint64_t n = std::min(slice, n_words - off);
    uint64_t* mine = (uint64_t*)c->filt.blocks + off;
    const ncclResult_t r = g.api.AllGather(mine, c->comm_tmp.p, n * 8, ncclUint8, g.comm, s);
    if (r != ncclSuccess) {
      return c->fail_nccl(r, "ncclAllGather (bit vector)");
    }
    k_or_gathered<<<grid_for(n, 256, c->sm_count * 8), 256, 0, s>>>(mine, c->comm_tmp.p, n, n,
                                                                    g.rank, g.world);
    n_launch += 2;
  }
  c->kend(GRB_K_GATHER, n_launch);
  GRB_CUDA(c, cudaGetLastError());
  return GRB_OK;
}

// OR-reduce of the replicas' bit vectors for callers that sharded pass 1 themselves with
// grb_build_bitvector_range (grb_run_path does, chunk by chunk under the ingest); no-op on one GPU
int
grb_bitvector_or_reduce(grb_ctx* c)
{
  cudaSetDevice(c->device);
  if (!c->filter_alloc || c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_bitvector_or_reduce: needs grb_filter_alloc and no finalize yet");
  }
  if (!c->comm) {
    return GRB_OK;
  }
  c->tic();
  const int rc = or_reduce_bitvector(c);
  if (rc == GRB_OK) {
    c->toc();
  }
  return rc;
}

int
grb_build_bitvector(grb_ctx* c)
{
  if (!c->comm) {
    return grb_build_bitvector_range(c, 0, c->n_reads);
  }
  // equal shares of the flagged bases, cut at read boundaries (every rank computes the same cut)
  uint64_t total = 0;
  for (uint64_t r = 0; r < c->n_reads; ++r) {
    total += (c->h_flags[r] & GRB_READ_PASS1) ? c->h_len[r] : 0;
  }
  const int W = c->comm->world, me = c->comm->rank;
  uint64_t lo = c->n_reads, hi = c->n_reads, acc = 0;
  const uint64_t b_lo = total / W * me + std::min<uint64_t>(me, total % W);
  const uint64_t b_hi = total / W * (me + 1) + std::min<uint64_t>(me + 1, total % W);
  for (uint64_t r = 0; r < c->n_reads; ++r) {
    if (lo == c->n_reads && acc >= b_lo) {
      lo = r;
    }
    if (acc >= b_hi) {
      hi = r;
      break;
    }
    acc += (c->h_flags[r] & GRB_READ_PASS1) ? c->h_len[r] : 0;
  }
  if (me == W - 1) {
    hi = c->n_reads;
  }
  if (lo > hi) {
    lo = hi;
  }
  cudaEvent_t t0 = c->prof_event();
  cudaEventRecord(t0, c->stream);
  int rc = grb_build_bitvector_range(c, lo, hi - lo);
  if (rc == GRB_OK) {
    rc = or_reduce_bitvector(c);
  }
  if (rc == GRB_OK) { // device time of the whole phase, exchange included
    cudaEventRecord(c->ev1, c->stream);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, c->ev1);
    c->last_ms = ms;
  }
  c->prof_free.push_back(t0);
  return rc;
}

int
grb_finalize_bitvector(grb_ctx* c, uint64_t* pop)
{
  cudaSetDevice(c->device);
  if (!c->filter_alloc || c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_finalize_bitvector: needs grb_filter_alloc, once");
  }
  cudaStream_t s = c->stream;
  const uint64_t per_cta = 256ull * GRB_RANK_ITEMS;
  const uint64_t n_cta = (c->filt.n_blocks + per_cta - 1) / per_cta;
  DevBuf<uint32_t> partial;
  DevBuf<uint64_t> partial_off;
  GRB_CUDA(c, partial.reserve(n_cta, 0, s));
  GRB_CUDA(c, partial_off.reserve(n_cta + 1, 0, s));
  c->tic();
  c->kbegin();
  k_rank_partial<<<(unsigned)n_cta, 256, 0, s>>>(c->filt.blocks, c->filt.n_blocks, partial.p);
  k_scan_u32<<<1, 1024, 0, s>>>(partial.p, partial_off.p, n_cta);
  k_rank_write<<<(unsigned)n_cta, 256, 0, s>>>(c->filt.blocks, c->filt.n_blocks, partial_off.p);
  c->kend(GRB_K_RANK, 3);
  uint64_t total = 0;
  GRB_CUDA(c, cudaMemcpyAsync(&total, partial_off.p + n_cta, 8, cudaMemcpyDeviceToHost, s));
  c->toc();
  GRB_CUDA(c, cudaGetLastError());
  c->filt.pop = total;
  // m_data + m_counts (MIBloomFilter.hpp:165-184, MIBFConstructSupport.hpp:175-181), zeroed
  if ((total + 1) * sizeof(GrbSlot) > c->slots_cap) {
    if (c->filt.slots) {
      cudaStreamSynchronize(s);
      grb_pool_free(c->filt.slots);
      c->filt.slots = nullptr;
      c->slots_cap = 0;
    }
    GRB_CUDA(c, grb_pool_alloc((void**)&c->filt.slots, (total + 1) * sizeof(GrbSlot)));
    c->slots_cap = (total + 1) * sizeof(GrbSlot);
  }
  GRB_CUDA(c, cudaMemsetAsync(c->filt.slots, 0, (total + 1) * sizeof(GrbSlot), s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  c->finalized = true;
  if (pop) {
    *pop = total;
  }
  return GRB_OK;
}

int
grb_reset_ids(grb_ctx* c)
{
  cudaSetDevice(c->device);
  if (!c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_reset_ids before grb_finalize_bitvector");
  }
  GRB_CUDA(c, cudaMemsetAsync(c->filt.slots, 0, (c->filt.pop + 1) * sizeof(GrbSlot),
                              c->stream));
  return GRB_OK;
}

int
grb_bitvector_device(grb_ctx* c, void** dev_ptr, uint64_t* bytes)
{
  if (!c->filter_alloc) {
    return c->fail(GRB_ERR_STATE, "grb_bitvector_device before grb_filter_alloc");
  }
  *dev_ptr = c->filt.blocks;
  *bytes = c->filt.n_blocks * 32;
  return GRB_OK;
}

int
grb_or_words(grb_ctx* c, void* dst, const void* src, uint64_t n_words)
{
  cudaSetDevice(c->device);
  k_or_words<<<grid_for(n_words, 256, c->sm_count * 8), 256, 0, c->stream>>>(
    (uint64_t*)dst, (const uint64_t*)src, n_words);
  c->launches += 1;
  GRB_CUDA(c, cudaGetLastError());
  return GRB_OK;
}

// ------------------------------------------------------------------------------------------
// TEST HOOK: the grouped half-hash evaluation of nthash.cuh (the very functions k2_query and
// k_fill_part inline: table construction, grb_lo64, grb_group_half, grb_combine) run on the HOST,
// so that the product's hash formulation is checked against the oracle on a machine without a GPU.
// out[frame * h + pattern] with the stale-tail rule of multiLensfrHashIterator.hpp:49-68.
// ACGT (upper case) only.  Nothing in the product path calls this.
// ------------------------------------------------------------------------------------------
struct GrbWordArray
{
  const uint64_t* w;
  __host__ __device__ uint64_t operator()(uint64_t i) const { return w[i]; }
};

int
grb_test_group_hash_host(const char* const* seeds, uint32_t h, const char* seq, size_t n, uint64_t* out)
{
  grb_ctx tmp;
  for (uint32_t i = 0; i < h; ++i) {
    tmp.seeds.emplace_back(seeds[i]);
  }
  if (h == 0 || build_seed_tables(&tmp) != GRB_OK) {
    return GRB_ERR_ARG;
  }
  const GrbSeedTables& t = tmp.h_seed;
  if (n < t.k + t.h - 1) {
    return GRB_ERR_ARG;
  }
  uint32_t ng = 0;
  const std::vector<ulonglong2> tab = build_group_tables(t, &ng);
  const ulonglong2* tl_ = tab.data();
  const ulonglong2* tr_ = tab.data() + (size_t)ng * 256;
  std::vector<uint64_t> words(n / 32 + 4, 0);
  for (size_t i = 0; i < n; ++i) {
    uint64_t code;
    switch (seq[i]) {
      case 'A': code = 0; break;
      case 'C': code = 1; break;
      case 'G': code = 2; break;
      case 'T': code = 3; break;
      default: return GRB_ERR_ARG;
    }
    words[i >> 5] |= code << (2 * (i & 31));
  }
  const GrbWordArray word{ words.data() };
  const uint64_t frames = n - t.k + 1;
  for (uint64_t f = 0; f < frames; ++f) {
    for (uint32_t i = 0; i < t.h; ++i) {
      const uint64_t n_i = n - (t.k + i) + 1;
      const uint64_t p = f < n_i ? f : n_i - 1;
      const ulonglong2 l = grb_group_half(tl_, ng, grb_lo64(word, p));
      const ulonglong2 r = grb_group_half(tr_, ng, grb_lo64(word, p + t.half + i));
      out[f * t.h + i] = grb_combine(i, l.x, l.y, r.x, r.y);
    }
  }
  return GRB_OK;
}

// ------------------------------------------------------------------------------------------
// probe microbenchmark (kernels_probe.cuh)
// ------------------------------------------------------------------------------------------
int
grb_probe_bench(grb_ctx* c, uint64_t filter_bits, double fill, uint32_t h, uint64_t n_probes,
                uint64_t seed, int reps, grb_probe_bench_result* out)
{
  cudaSetDevice(c->device);
  if (h == 0 || h > GRB_MAX_PATTERNS || !(fill > 0.0 && fill < 0.95) || reps < 1 || n_probes < h) {
    return c->fail(GRB_ERR_ARG, "grb_probe_bench: h in 1..8, fill in (0, 0.95), reps >= 1");
  }
  memset(out, 0, sizeof *out);
  int rc = grb_filter_alloc(c, filter_bits);
  if (rc != GRB_OK) {
    return rc;
  }
  cudaStream_t s = c->stream;
  // keys whose h hashes set a fraction `fill` of the bits: 1 - exp(-n h / m) = fill
  const uint64_t n_fill = std::max<uint64_t>(1, (uint64_t)(-log(1.0 - fill) * (double)filter_bits / h));
  k_probe_fill<<<c->sm_count * 16, 256, 0, s>>>(c->filt, n_fill, h, seed);
  c->launches += 1;
  uint64_t pop = 0;
  if ((rc = grb_finalize_bitvector(c, &pop)) != GRB_OK) {
    return rc;
  }
  k_probe_ids<<<c->sm_count * 16, 256, 0, s>>>(c->filt.slots, pop, seed ^ 0x5bd1e995u);
  c->launches += 1;
  DevBuf<unsigned long long> sum;
  GRB_CUDA(c, sum.reserve(2, 0, s));
  GRB_CUDA(c, cudaMemsetAsync(sum.p, 0, 16, s));
  const uint64_t n_keys = n_probes / h;
  const unsigned grid = (unsigned)c->sm_count * 16;
  cudaEvent_t e0 = c->prof_event(), e1 = c->prof_event();
  float best_q = 1e30f, best_i = 1e30f;
  for (int r = 0; r <= reps; ++r) { // first round is the warm-up
    cudaEventRecord(e0, s);
    k_probe_query<<<grid, 256, 0, s>>>(c->filt, n_keys, n_fill, h, seed, ~seed + 977ull * r * n_keys, sum.p);
    cudaEventRecord(e1, s);
    GRB_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r) {
      best_q = std::min(best_q, ms);
    }
  }
  // one-line-per-probe variant over the same memory (the slot array read as 128-byte lines)
  float best_l = 1e30f;
  const uint64_t n_lines = (pop + 1) * sizeof(GrbSlot) / 128;
  for (int r = 0; r <= reps && n_lines > 0; ++r) {
    cudaEventRecord(e0, s);
    k_probe_line<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(c->filt.slots), n_lines, n_keys, h,
                                      seed, seed * 17 + 311ull * r * n_keys, sum.p);
    cudaEventRecord(e1, s);
    GRB_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r) {
      best_l = std::min(best_l, ms);
    }
  }
  for (int r = 0; r <= reps; ++r) {
    cudaEventRecord(e0, s);
    k_probe_insert<<<grid, 256, 0, s>>>(c->filt, n_keys, n_fill, h, seed, seed * 31 + 131ull * r * n_keys, 7u + r);
    cudaEventRecord(e1, s);
    GRB_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r) {
      best_i = std::min(best_i, ms);
    }
  }
  c->launches += 3 * (uint64_t)(reps + 1);
  c->prof_free.push_back(e0);
  c->prof_free.push_back(e1);
  unsigned long long hsum[2] = { 0, 0 };
  GRB_CUDA(c, cudaMemcpyAsync(hsum, sum.p, 16, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  GRB_CUDA(c, cudaGetLastError());
  out->query_ms = best_q;
  out->insert_ms = best_i;
  out->line_query_ms = n_lines ? best_l : 0.0;
  out->line_bytes = n_lines * 128;
  out->probes = n_keys * h;
  out->pop = pop;
  out->filter_bits = filter_bits;
  out->footprint_bytes = c->filt.n_blocks * 32 + (pop + 1) * sizeof(GrbSlot);
  out->checksum = hsum[0];
  out->probes_missed = hsum[1];
  out->keys_filled = n_fill;
  return GRB_OK;
}

// ------------------------------------------------------------------------------------------
// parity exports
// ------------------------------------------------------------------------------------------
int
grb_copy_bitvector(grb_ctx* c, uint64_t* words)
{
  cudaSetDevice(c->device);
  if (!c->filter_alloc) {
    return c->fail(GRB_ERR_STATE, "grb_copy_bitvector before grb_filter_alloc");
  }
  const uint64_t n_words = (c->filt.bits + 63) / 64;
  DevBuf<uint64_t> tmp;
  GRB_CUDA(c, tmp.reserve(n_words, 0, c->stream));
  k_export_plain<<<grid_for(n_words, 256, c->sm_count * 8), 256, 0, c->stream>>>(c->filt.blocks,
                                                                                n_words, tmp.p);
  c->launches += 1;
  GRB_CUDA(c, cudaMemcpyAsync(words, tmp.p, n_words * 8, cudaMemcpyDeviceToHost, c->stream));
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GRB_OK;
}

int
grb_load_bitvector(grb_ctx* c, const uint64_t* words)
{
  cudaSetDevice(c->device);
  if (!c->filter_alloc || c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_load_bitvector: needs grb_filter_alloc and no finalize yet");
  }
  const uint64_t n_words = (c->filt.bits + 63) / 64;
  DevBuf<uint64_t> tmp;
  GRB_CUDA(c, tmp.reserve(n_words, 0, c->stream));
  GRB_CUDA(c, cudaMemcpyAsync(tmp.p, words, n_words * 8, cudaMemcpyHostToDevice, c->stream));
  GRB_CUDA(c, cudaMemsetAsync(c->filt.blocks, 0, c->filt.n_blocks * 32, c->stream));
  k_import_plain<<<grid_for(n_words, 256, c->sm_count * 8), 256, 0, c->stream>>>(c->filt.blocks,
                                                                                n_words, tmp.p);
  c->launches += 1;
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  return GRB_OK;
}

int
grb_rank(grb_ctx* c, const uint64_t* pos, size_t n, uint64_t* rank, uint8_t* bit)
{
  cudaSetDevice(c->device);
  if (!c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_rank before grb_finalize_bitvector");
  }
  for (size_t i = 0; i < n; ++i) {
    if (pos[i] >= c->filt.bits) {
      return c->fail(GRB_ERR_ARG, "grb_rank: position past the end of the filter");
    }
  }
  DevBuf<uint64_t> dpos, drank;
  DevBuf<uint8_t> dbit;
  cudaStream_t s = c->stream;
  GRB_CUDA(c, dpos.reserve(n, 0, s));
  GRB_CUDA(c, drank.reserve(n, 0, s));
  GRB_CUDA(c, dbit.reserve(n, 0, s));
  GRB_CUDA(c, cudaMemcpyAsync(dpos.p, pos, n * 8, cudaMemcpyHostToDevice, s));
  k_rank_query<<<grid_for(n, 256, 1u << 30), 256, 0, s>>>(c->filt, dpos.p, n, drank.p, dbit.p);
  c->launches += 1;
  GRB_CUDA(c, cudaMemcpyAsync(rank, drank.p, n * 8, cudaMemcpyDeviceToHost, s));
  if (bit) {
    GRB_CUDA(c, cudaMemcpyAsync(bit, dbit.p, n, cudaMemcpyDeviceToHost, s));
  }
  GRB_CUDA(c, cudaStreamSynchronize(s));
  return GRB_OK;
}

int
grb_get_ids(grb_ctx* c, const uint64_t* rank, size_t n, uint32_t* ids, uint32_t* counts)
{
  cudaSetDevice(c->device);
  if (!c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_get_ids before grb_finalize_bitvector");
  }
  for (size_t i = 0; i < n; ++i) {
    if (rank[i] >= c->filt.pop) {
      return c->fail(GRB_ERR_ARG, "grb_get_ids: rank past the end of the ID array");
    }
  }
  DevBuf<uint64_t> dr;
  DevBuf<uint32_t> di, dc;
  cudaStream_t s = c->stream;
  GRB_CUDA(c, dr.reserve(n, 0, s));
  GRB_CUDA(c, di.reserve(n, 0, s));
  GRB_CUDA(c, dc.reserve(n, 0, s));
  GRB_CUDA(c, cudaMemcpyAsync(dr.p, rank, n * 8, cudaMemcpyHostToDevice, s));
  k_get_slots<<<grid_for(n, 256, 1u << 30), 256, 0, s>>>(c->filt.slots, dr.p, n, di.p, dc.p);
  c->launches += 1;
  GRB_CUDA(c, cudaMemcpyAsync(ids, di.p, n * 4, cudaMemcpyDeviceToHost, s));
  if (counts) {
    GRB_CUDA(c, cudaMemcpyAsync(counts, dc.p, n * 4, cudaMemcpyDeviceToHost, s));
  }
  GRB_CUDA(c, cudaStreamSynchronize(s));
  return GRB_OK;
}

int
grb_set_ids(grb_ctx* c, const uint64_t* rank, size_t n, const uint32_t* ids, const uint32_t* counts)
{
  cudaSetDevice(c->device);
  if (!c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_set_ids before grb_finalize_bitvector");
  }
  for (size_t i = 0; i < n; ++i) {
    if (rank[i] >= c->filt.pop) {
      return c->fail(GRB_ERR_ARG, "grb_set_ids: rank past the end of the ID array");
    }
  }
  DevBuf<uint64_t> dr;
  DevBuf<uint32_t> di, dc;
  cudaStream_t s = c->stream;
  GRB_CUDA(c, dr.reserve(n, 0, s));
  GRB_CUDA(c, di.reserve(n, 0, s));
  GRB_CUDA(c, dc.reserve(n, 0, s));
  GRB_CUDA(c, cudaMemcpyAsync(dr.p, rank, n * 8, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemcpyAsync(di.p, ids, n * 4, cudaMemcpyHostToDevice, s));
  GRB_CUDA(c, cudaMemcpyAsync(dc.p, counts, n * 4, cudaMemcpyHostToDevice, s));
  k_set_slots<<<grid_for(n, 256, 1u << 30), 256, 0, s>>>(c->filt.slots, dr.p, n, di.p, dc.p);
  c->launches += 1;
  GRB_CUDA(c, cudaStreamSynchronize(s));
  return GRB_OK;
}

int
grb_hash_sequence(grb_ctx* c, const char* seq, size_t n, uint64_t* out)
{
  cudaSetDevice(c->device);
  const unsigned h = c->h_seed.h, k = c->h_seed.k;
  if (n < k + h - 1) {
    return c->fail(GRB_ERR_ARG, "SeedNtHash: sequence length is smaller than k");
  }
  std::vector<uint64_t> words((n + 31) / 32 + 4, 0);
  for (size_t i = 0; i < n; ++i) {
    unsigned code;
    switch (seq[i]) {
      case 'A': case 'a': code = 0; break;
      case 'C': case 'c': code = 1; break;
      case 'G': case 'g': code = 2; break;
      case 'T': case 't': code = 3; break;
      default: return c->fail(GRB_ERR_ARG, "grb_hash_sequence: non-ACGT base");
    }
    words[i >> 5] |= (uint64_t)code << (2 * (i & 31));
  }
  const uint64_t frames = n - k + 1;
  DevBuf<uint64_t> dw, dout;
  cudaStream_t s = c->stream;
  GRB_CUDA(c, dw.reserve(words.size(), 0, s));
  GRB_CUDA(c, dout.reserve(frames * h, 0, s));
  GRB_CUDA(c, cudaMemcpyAsync(dw.p, words.data(), words.size() * 8, cudaMemcpyHostToDevice, s));
  k_hash_sequence<<<grid_for(frames, 128, c->sm_count * 16), 128, 0, s>>>(dw.p, (uint32_t)n,
                                                                         c->d_seed, dout.p);
  c->launches += 1;
  GRB_CUDA(c, cudaMemcpyAsync(out, dout.p, frames * h * 8, cudaMemcpyDeviceToHost, s));
  GRB_CUDA(c, cudaStreamSynchronize(s));
  return GRB_OK;
}

// ------------------------------------------------------------------------------------------
// selection loop
// ------------------------------------------------------------------------------------------
static int
sel_prepare(grb_ctx* c, uint64_t max_len)
{
  cudaStream_t s = c->stream;
  const uint64_t T = c->p.tile_length, h = c->h_seed.h, k = c->h_seed.k;
  const uint64_t F = c->tile_frames;
  if (!c->sel_init) {
    GrbSelParams& q = c->prm;
    q.tile_len = (uint32_t)T;
    q.kmer = (uint32_t)c->p.kmer_size;
    q.tile_frames = (uint32_t)F;
    q.k = (uint32_t)k;
    q.h = (uint32_t)h;
    q.cand_cap = (uint32_t)(F * h / 3 + 1);
    // A tile votes for at most F * h distinct ids.  The batch engine never adds to the per-tile tables
    // after the query (its commit keeps deltas apart), so it needs no slack beyond a load factor
    // below 3/4; the serial engine keeps the 2x sizing.
    q.table_size = (uint32_t)next_pow2(2 * F * h);
    if (c->batch_mode) {
      const char* e = getenv("GRB_VOTE_TABLE");
      if (!(e && strcmp(e, "wide") == 0)) {
        q.table_size = (uint32_t)next_pow2(F * h + F * h / 3 + 1);
      }
    }
    q.sw_words = (uint32_t)((T + k + 63) / 32 + 4);
    q.silver = c->p.silver_path;
    q.pad = 0;
    q.threshold = c->p.threshold;
    q.unassigned_min = c->p.unassigned_min;
    q.assigned_max = c->p.assigned_max;
    q.block_size = c->p.block_size;
    q.max_paths = c->p.max_paths;
    q.target_bases = (uint64_t)(c->p.ratio * c->p.genome_size); // goldrush_path.cpp:1223
    c->query_smem = sizeof(GrbSeedTables) + (size_t)q.sw_words * 8 + (size_t)q.table_size * 8;
    if (c->query_smem > 200 * 1024) {
      return c->fail(GRB_ERR_ARG, "tile_length * hash_num too large for the shared-memory vote "
                                  "table (limit 12288)");
    }
    GRB_CUDA(c, cudaFuncSetAttribute(k_query<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)c->query_smem));
    if (!c->d_state) {
      GRB_CUDA(c, grb_pool_alloc((void**)&c->d_state, sizeof(GrbSelState)));
    }
    GrbSelState st{};
    st.curr_path = 1;
    GRB_CUDA(c, cudaMemcpyAsync(c->d_state, &st, sizeof st, cudaMemcpyHostToDevice, s));
    GRB_CUDA(c, cudaStreamSynchronize(s));
    GRB_CUDA(c, c->b_plan.reserve(1, 0, s));
    c->sel_init = true;
    c->sel_finished = false;
  }
  const uint64_t tiles = std::max<uint64_t>(1, max_len / T);
  if (tiles > c->sc_tiles) {
    const uint64_t cap = c->prm.cand_cap;
    GRB_CUDA(c, c->b_stash.reserve(tiles * F * h, 0, s));
    GRB_CUDA(c, c->b_best_id.reserve(tiles, 0, s));
    GRB_CUDA(c, c->b_best_count.reserve(tiles, 0, s));
    GRB_CUDA(c, c->b_n_cand.reserve(tiles, 0, s));
    GRB_CUDA(c, c->b_cand_id.reserve(tiles * cap, 0, s));
    GRB_CUDA(c, c->b_cand_cnt.reserve(tiles * cap, 0, s));
    GRB_CUDA(c, c->b_tile_id.reserve(tiles, 0, s));
    GRB_CUDA(c, c->b_tile_as.reserve(tiles, 0, s));
    GRB_CUDA(c, c->b_snap.reserve(tiles + 2, 0, s));
    c->sc_tiles = tiles;
  }
  const uint64_t round_tiles = std::min<uint64_t>(tiles, 64 * c->p.block_size);
  const uint64_t tab = next_pow2(2 * round_tiles * F * h);
  if (tab > c->sc_tab) {
    c->b_tab_key.release();
    c->b_tab_mask.release();
    GRB_CUDA(c, c->b_tab_key.reserve(tab, 0, s));
    GRB_CUDA(c, c->b_tab_mask.reserve(tab, 0, s));
    GRB_CUDA(c, cudaMemsetAsync(c->b_tab_key.p, 0xFF, c->b_tab_key.cap * 8, s));
    GRB_CUDA(c, cudaMemsetAsync(c->b_tab_mask.p, 0, c->b_tab_mask.cap * 8, s));
    c->sc_tab = tab;
  }
  c->sc = GrbSelScratch{ c->b_stash.p,   c->b_best_id.p, c->b_best_count.p, c->b_n_cand.p,
                         c->b_cand_id.p, c->b_cand_cnt.p, c->b_tile_id.p,   c->b_tile_as.p,
                         c->b_snap.p,    c->b_plan.p,     c->b_tab_key.p,   c->b_tab_mask.p };
  return GRB_OK;
}

static void
launch_read(grb_ctx* c, uint64_t r, uint64_t dec_idx, grb_decision* d_dec)
{
  cudaStream_t s = c->stream;
  const uint64_t T = c->p.tile_length, h = c->h_seed.h;
  const uint64_t tiles = c->h_len[r] / T;
  const unsigned qgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, 1024));
  c->kbegin();
  k_query<512><<<qgrid, 512, c->query_smem, s>>>(c->reads_dev(), c->d_seed, c->filt, c->prm, c->sc,
                                                 c->d_state, r);
  c->kend(GRB_K_QUERY);
  c->kbegin();
  k_decide<<<1, 32, 0, s>>>(c->reads_dev(), c->prm, c->sc, c->d_state, d_dec, r, dec_idx);
  c->kend(GRB_K_DECIDE);
  const uint64_t blocks = (tiles + c->p.block_size - 1) / c->p.block_size;
  const uint64_t rounds = std::max<uint64_t>(1, (blocks + 63) / 64);
  const uint64_t round_tiles = std::min<uint64_t>(std::max<uint64_t>(tiles, 1), 64 * c->p.block_size);
  const uint32_t tab = (uint32_t)next_pow2(2 * round_tiles * c->tile_frames * h);
  const unsigned cgrid = grid_for(round_tiles * c->tile_frames * h, 256, c->sm_count * 4);
  const unsigned agrid = grid_for(tab, 256, c->sm_count * 4);
  for (uint64_t round = 0; round < rounds; ++round) {
    c->kbegin();
    k_insert_collect<<<cgrid, 256, 0, s>>>(c->reads_dev(), c->prm, c->sc, c->sc.stash, c->d_state,
                                           r, (uint32_t)round, tab);
    k_insert_apply<<<agrid, 256, 0, s>>>(c->filt, c->sc, c->d_state, r, (uint32_t)round, tab);
    c->kend(GRB_K_INSERT, 2);
  }
}

// ---- batch engine -------------------------------------------------------------------------
struct BatchPlan
{
  std::vector<uint64_t> read_idx;   // visited reads of the chunk, in order
  std::vector<uint64_t> dec_idx;    // their index in the decisions array
  std::vector<uint32_t> tile_first; // per batch: local prefix (nb + 1 entries each), concatenated
  std::vector<uint32_t> tile_read;  // per batch: b of each tile, concatenated
  std::vector<uint64_t> cm_off;     // per read: offset of its count matrix within the batch
  struct Batch
  {
    uint32_t read0, nb, tf0, tr0, n_bt;
    uint32_t max_tiles; // longest read of the batch, in tiles
    uint64_t cm_words;  // sum of tiles^2
  };
  std::vector<Batch> batches;
};

// ---- batch engine: query-side buffers ------------------------------------------------------
static int
query_prepare(grb_ctx* c, uint64_t max_batch_tiles, uint64_t max_read_tiles, uint64_t max_cm_words,
               uint32_t max_batch_reads)
{
  cudaStream_t s = c->stream;
  const uint64_t T = c->tile_frames, h = c->h_seed.h; // T: frames per tile (stride of the per-probe buffers)
  if (max_read_tiles > 2048) {
    return c->fail(GRB_ERR_ARG, "a read spans more than 2048 tiles: raise the tile length");
  }
  if (max_batch_tiles > c->b2_cap_tiles) {
    const uint64_t n = max_batch_tiles;
    const uint64_t n_probe = n * T * h;
    if (n_probe >= (1ull << 26)) {
      return c->fail(GRB_ERR_ARG, "batch too large for the 26-bit probe index");
    }
    c->bb_stash.release();
    c->b2_vk.release();
    c->b2_vc.release();
    c->b2_ix.release();
    c->b2_ix_sidx.release();
    GRB_CUDA(c, c->bb_stash.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b2_vk.reserve(n * c->prm.table_size, 0, s));
    GRB_CUDA(c, c->b2_vc.reserve(n * c->prm.table_size, 0, s));
    c->b2_ix_entries = next_pow2(n_probe + n_probe / 2);
    GRB_CUDA(c, c->b2_ix.reserve(c->b2_ix_entries, 0, s));
    GRB_CUDA(c, c->b2_ix_sidx.reserve(c->b2_ix_entries, 0, s));
    GRB_CUDA(c, c->b2_c_slot.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b2_c_probe.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b2_c_next.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b2_c_sidx.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b2_fbits.reserve(n * T / 32 + 2, 0, s));
    GRB_CUDA(c, c->b2_fl.reserve(n * T, 0, s));
    GRB_CUDA(c, c->b2_fr.reserve(n * T * (2 + h), 0, s));
    GRB_CUDA(c, c->bb_best_id.reserve(n, 0, s));
    GRB_CUDA(c, c->bb_best_count.reserve(n, 0, s));
    GRB_CUDA(c, c->bb_hits.reserve(n, 0, s));
    GRB_CUDA(c, c->bb_miss.reserve(n, 0, s));
    GRB_CUDA(c, c->bb_uq.reserve(n, 0, s));
    GRB_CUDA(c, c->b2_counters.reserve(8, 0, s));
    c->b2_cap_tiles = n;
  }
  if (max_batch_reads > c->b2_cap_reads) {
    const uint32_t nb = max_batch_reads;
    GRB_CUDA(c, c->bb_sp_nas.reserve(nb, 0, s));
    GRB_CUDA(c, c->bb_sp_plan.reserve(nb, 0, s));
    GRB_CUDA(c, c->bb_sp_adv.reserve(nb, 0, s));
    GRB_CUDA(c, c->bb_rd_hits.reserve(nb, 0, s));
    GRB_CUDA(c, c->bb_rd_miss.reserve(nb, 0, s));
    GRB_CUDA(c, c->bb_rd_q.reserve(nb, 0, s));
    GRB_CUDA(c, c->bb_nu.reserve(nb, 0, s));
    GRB_CUDA(c, c->b2_fl_n.reserve(nb, 0, s));
    GRB_CUDA(c, c->b2_plan_out.reserve(nb, 0, s));
    c->b2_cap_reads = nb;
  }
  GRB_CUDA(c, c->bb_cm.reserve(std::max<uint64_t>(1, max_cm_words), 0, s));
  if (!c->b2_attr) {
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
    c->b2_smem_max = (size_t)max_optin - 2048; // static shared memory of the kernels
    GRB_CUDA(c, cudaFuncSetAttribute(k2_cmat<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)c->b2_smem_max));
    c->query2_smem = (size_t)c->gt_groups * 256 * 32 + (size_t)c->prm.sw_words * 8 +
                     (size_t)c->prm.table_size * 8;
    if (c->query2_smem > c->b2_smem_max) {
      return c->fail(GRB_ERR_ARG, "tile_length x hash_num too large for the shared-memory vote table");
    }
    GRB_CUDA(c, cudaFuncSetAttribute(k2_query<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)c->query2_smem));
    c->b2_attr = true;
  }
  return GRB_OK;
}

// ---- batch engine: commit-side buffers and the per-batch launch sequence -------------------
static const uint32_t kFixDeltaSmem = 8192; // shared-memory delta table entries of k3_fix

static int
commit_prepare(grb_ctx* c, uint64_t max_batch_tiles, uint64_t max_read_tiles, uint64_t max_cm_words,
               uint32_t max_batch_reads)
{
  cudaStream_t s = c->stream;
  const uint64_t T = c->tile_frames, h = c->h_seed.h; // T: frames per tile
  int rc = query_prepare(c, max_batch_tiles, max_read_tiles, max_cm_words, max_batch_reads);
  if (rc != GRB_OK) {
    return rc;
  }
  if (max_batch_reads > 1024) {
    return c->fail(GRB_ERR_ARG, "GRB_BATCH_READS above 1024");
  }
  if (!c->b3_attr) {
    GRB_CUDA(c, cudaFuncSetAttribute(k3_fix<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)c->b2_smem_max));
    GRB_CUDA(c, cudaFuncSetAttribute(k3_fix<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)c->b2_smem_max));
    int per_sm = 0;
    GRB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_fix<512>, 512,
                                                              c->b2_smem_max));
    if (per_sm < 1) {
      return c->fail(GRB_ERR_CUDA, "k3_fix cannot be resident on this device");
    }
    // The re-validation phase of the commit is one CTA per read.  GRB_FIX_BS=256 runs two
    // 256-thread CTAs per SM (one read each) instead of one 512-thread CTA per SM (two reads in
    // turn): measured slower on cfg2 (commit 82 -> 108 ms), because a batch waits for its slowest
    // read and that read takes twice as long with half the threads.  Per-CTA scratch is sized for
    // two CTAs per SM either way.
    if (const char* e = getenv("GRB_FIX_BS")) {
      c->b3_bs = strcmp(e, "256") == 0 ? 256u : 512u;
    }
    c->b3_ctas = (uint32_t)c->sm_count * 2;
    c->b3_dcap = (uint32_t)next_pow2(4 * T * h + 64);
    GRB_CUDA(c, c->b3_d_keys.reserve((size_t)c->b3_ctas * c->b3_dcap, 0, s));
    GRB_CUDA(c, c->b3_d_vals.reserve((size_t)c->b3_ctas * c->b3_dcap, 0, s));
    GRB_CUDA(c, c->b3_ctl.reserve(1, 0, s));
    GRB_CUDA(c, c->b3_barrier.reserve(1, 0, s));
    if (getenv("GRB_FIX_DEBUG")) {
      GRB_CUDA(c, c->b3_dbg.reserve_exact((size_t)4 + 4 * (1u << 21), s));
      GRB_CUDA(c, cudaMemsetAsync(c->b3_dbg.p, 0, 16, s));
    }
    c->b3_attr = true;
  }
  if (max_batch_tiles > c->b3_cap_tiles) {
    const uint64_t n_probe = max_batch_tiles * T * h;
    GRB_CUDA(c, c->b3_shared.reserve(n_probe / 2 + 1, 0, s));
    GRB_CUDA(c, c->b3_m_fill.reserve(n_probe / 2 + 1, 0, s));
    GRB_CUDA(c, c->b3_c_pos.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b3_cand.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b3_ix_cnt.reserve(c->b2_ix_entries, 0, s));
    // rank bit map of a batch: 32 bits per probe keeps the false-positive share near 3 %, and
    // 2^28 bits (32 MB) is what stays resident in L2 next to the streams
    c->b3_bm_bits = (uint32_t)std::min<uint64_t>(1ull << 28, next_pow2(32 * n_probe));
    GRB_CUDA(c, c->b3_bm.reserve(c->b3_bm_bits / 32 + 1, 0, s));
    GRB_CUDA(c, c->b3_m_key.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b3_m_ci.reserve(n_probe, 0, s));
    GRB_CUDA(c, c->b3_m_seen.reserve(n_probe, 0, s));
    c->b2_fr.release();
    GRB_CUDA(c, c->b2_fr.reserve(max_batch_tiles * T * (2 + 2 * h), 0, s));
    c->b3_cap_tiles = max_batch_tiles;
  }
  GRB_CUDA(c, c->b3_np.reserve(max_batch_reads, 0, s));
  GRB_CUDA(c, c->b3_np_adv.reserve(max_batch_reads, 0, s));
  GRB_CUDA(c, c->b3_np_nas.reserve(max_batch_reads, 0, s));
  GRB_CUDA(c, c->b3_np_dh.reserve(max_batch_reads, 0, s));
  if (max_read_tiles > 160 && max_read_tiles * max_read_tiles > c->b3_cmat_cap) {
    c->b3_cmat_cap = max_read_tiles * max_read_tiles;
    GRB_CUDA(c, c->b3_cmat_g.reserve(c->b3_cmat_cap * c->b3_ctas, 0, s));
  }
  return GRB_OK;
}

static int
launch_batch(grb_ctx* c, const BatchPlan::Batch& b, grb_decision* d_dec)
{
  cudaStream_t s = c->stream;
  const uint64_t T = c->tile_frames, h = c->h_seed.h; // T: frames per tile
  GrbBatchDev bd{};
  bd.read_idx = c->bb_read_idx[c->bb_set].p + b.read0;
  bd.tile_first = c->bb_tile_first[c->bb_set].p + b.tf0;
  bd.tile_read = c->bb_tile_read[c->bb_set].p + b.tr0;
  bd.nb = b.nb;
  bd.n_bt = b.n_bt;
  bd.stash = c->bb_stash.p;
  bd.best_id = c->bb_best_id.p;
  bd.best_count = c->bb_best_count.p;
  bd.tile_hits = c->bb_hits.p;
  bd.tile_miss = c->bb_miss.p;
  bd.uq = c->bb_uq.p;
  bd.nu = c->bb_nu.p;
  bd.cm = c->bb_cm.p;
  bd.cm_off = c->bb_cm_off[c->bb_set].p + b.read0;
  bd.sp_n_as = c->bb_sp_nas.p;
  bd.sp_plan = c->bb_sp_plan.p;
  bd.sp_adv = c->bb_sp_adv.p;
  bd.rd_hits = c->bb_rd_hits.p;
  bd.rd_miss = c->bb_rd_miss.p;
  bd.rd_queries = c->bb_rd_q.p;
  GrbB2 b2{};
  b2.vk = c->b2_vk.p;
  b2.vc = c->b2_vc.p;
  b2.table_size = c->prm.table_size;
  GrbB3 b3{};
  b3.vk = c->b2_vk.p;
  b3.vc = c->b2_vc.p;
  b3.table_size = c->prm.table_size;
  b3.d_cap = c->b3_dcap;
  b3.ix_tab = c->b2_ix.p;
  b3.bm = c->b3_bm.p;
  b3.cand = c->b3_cand.p;
  b3.ix_cnt = c->b3_ix_cnt.p;
  b3.t_probe = c->b2_c_next.p;
  b3.t_slot = c->b2_c_slot.p;
  const uint64_t n_probe = (uint64_t)b.n_bt * T * h;
  const uint32_t bm_bits = (uint32_t)std::min<uint64_t>(c->b3_bm_bits, std::max<uint64_t>(1024, next_pow2(32 * n_probe)));
  b3.bm_mask = bm_bits - 1;
  const uint64_t ix_entries = std::min<uint64_t>(c->b2_ix_entries, next_pow2(n_probe + n_probe / 2 + 64));
  b3.ix_mask = ix_entries - 1;
  b3.ix_sidx = c->b2_ix_sidx.p;
  b3.counters = c->b2_counters.p;
  b3.c_probe = c->b2_c_probe.p;
  b3.c_sidx = c->b2_c_sidx.p;
  b3.c_pos = c->b3_c_pos.p;
  b3.shared = c->b3_shared.p;
  b3.m_fill = c->b3_m_fill.p;
  b3.m_key = c->b3_m_key.p;
  b3.m_ci = c->b3_m_ci.p;
  b3.m_seen = c->b3_m_seen.p;
  b3.fbits = c->b2_fbits.p;
  b3.fl_n = c->b2_fl_n.p;
  b3.fl = c->b2_fl.p;
  b3.fr = c->b2_fr.p;
  b3.np = c->b3_np.p;
  b3.np_adv = c->b3_np_adv.p;
  b3.np_nas = c->b3_np_nas.p;
  b3.np_dh = c->b3_np_dh.p;
  b3.plan_out = c->b2_plan_out.p;
  b3.ctl = c->b3_ctl.p;
  b3.d_keys = c->b3_d_keys.p;
  b3.d_vals = c->b3_d_vals.p;
  b3.cmat_g = c->b3_cmat_g.p;
  b3.barrier = c->b3_barrier.p;
  b3.dbg = c->b3_dbg.p;
  b3.dbg_cap = c->b3_dbg.p ? (uint32_t)((c->b3_dbg.cap - 4) / 4) : 0u;

  const uint32_t n_cap = std::max<uint32_t>(b.max_tiles, 1);
  const uint32_t us = (uint32_t)next_pow2(2 * (uint64_t)n_cap);
  const uint32_t cm_smem = n_cap <= 160 ? 1u : 0u;
  const size_t as_pad = ((size_t)(n_cap + 15) / 16) * 16;
  const size_t cmat_smem = (size_t)6 * n_cap * 4 + 8 + (size_t)2 * us * 4 + as_pad +
                           (cm_smem ? (size_t)n_cap * n_cap * 4 : 0);
  // delta table of the re-validation phase: 8192 entries (GRB_FIX_DELTA=16384 doubles it when the
  // batch's longest read leaves room; measured slower on cfg2, commit 76 -> 84 ms: clearing and
  // scanning the table costs every read more than the shorter probe sequences save the few full ones)
  uint32_t dc = kFixDeltaSmem;
  if (const char* e = getenv("GRB_FIX_DELTA")) {
    dc = strcmp(e, "16384") == 0 ? 2 * kFixDeltaSmem : kFixDeltaSmem;
  }
  auto fix_bytes = [&](uint32_t d) {
    return (size_t)d * 12 + (size_t)n_cap * 8 + ((size_t)2 * us + (size_t)8 * n_cap + 2) * 4 + as_pad +
           (cm_smem ? (size_t)n_cap * n_cap * 4 : 0);
  };
  if (fix_bytes(dc) > c->b2_smem_max) {
    dc = kFixDeltaSmem;
  }
  const size_t fix_smem = fix_bytes(dc);
  if (cmat_smem > c->b2_smem_max || fix_smem > c->b2_smem_max) {
    return c->fail(GRB_ERR_ARG, "a read spans too many tiles for the commit kernel's shared "
                                "memory: raise the tile length");
  }
  static_assert(sizeof(GrbReadPlan) % 16 == 0 && sizeof(GrbFixCtl) == 16, "k3_reset stores uint4");
  k3_reset<<<(unsigned)c->sm_count * 4, 256, 0, s>>>(b3, c->d_state, bm_bits / 32,
                                                     (uint32_t)((uint64_t)b.n_bt * T / 32 + 2), b.nb);
  c->launches += 1;
  if (!c->comm || !c->shard_query) {
    c->kbegin();
    k2_query<512><<<grid_for(b.n_bt, 1, 1u << 20), 512, c->query2_smem, s>>>(
      c->reads_dev(), c->d_seed, c->d_gtab, c->gt_groups, c->filt, c->prm, bd, b2, c->d_state, 0u,
      b.n_bt);
    c->kend(GRB_K_QUERY);
  } else {
    // W GPUs: this rank's share of the batch's tiles, then one NCCL group gathers every per-tile
    // output in place (the buffers are sized for W * chunk tiles, grb_select_reads)
    GrbComm& g = *c->comm;
    const GrbShare sh = grb_share(b.n_bt, g.rank, g.world);
    c->kbegin();
    if (sh.hi > sh.lo) {
      k2_query<512><<<grid_for(sh.hi - sh.lo, 1, 1u << 20), 512, c->query2_smem, s>>>(
        c->reads_dev(), c->d_seed, c->d_gtab, c->gt_groups, c->filt, c->prm, bd, b2, c->d_state,
        (uint32_t)sh.lo, (uint32_t)sh.hi);
    }
    c->kend(GRB_K_QUERY, sh.hi > sh.lo ? 1 : 0);
    c->kbegin();
    const uint64_t ts = c->prm.table_size;
    struct Piece
    {
      void* base;
      uint64_t bytes_per_tile;
    };
    const Piece pieces[] = { { bd.stash, T * h * 8 },  { b2.vk, ts * 4 },        { b2.vc, ts * 4 },
                             { bd.best_id, 4 },        { bd.best_count, 4 },     { bd.tile_hits, 4 },
                             { bd.tile_miss, 4 } };
    ncclResult_t r = g.api.GroupStart();
    for (const Piece& pc : pieces) {
      if (r != ncclSuccess) {
        break;
      }
      const uint64_t n = sh.chunk * pc.bytes_per_tile;
      r = g.api.AllGather((const char*)pc.base + (uint64_t)g.rank * n, pc.base, n, ncclUint8, g.comm, s);
    }
    const ncclResult_t r2 = g.api.GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) {
      return c->fail_nccl(r != ncclSuccess ? r : r2, "ncclAllGather (batch query results)");
    }
    c->kend(GRB_K_GATHER, 1);
  }
  c->kbegin();
  k2_cmat<256><<<b.nb, 256, cmat_smem, s>>>(c->reads_dev(), c->prm, bd, b2, c->d_state, n_cap, us,
                                           cm_smem);
  c->kend(GRB_K_SMOOTH);
  c->kbegin();
  const unsigned g_small = (unsigned)c->sm_count * 8;
  const unsigned g_tiles = grid_for(b.n_bt, 1, (unsigned)c->sm_count * 8);
  const size_t stage_smem = (size_t)T * h * 4;
  if (2 * stage_smem > 48 * 1024 && !c->b3_stage_attr) {
    GRB_CUDA(c, cudaFuncSetAttribute(k3_mark, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)stage_smem));
    GRB_CUDA(c, cudaFuncSetAttribute(k3_members, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(2 * stage_smem)));
    c->b3_stage_attr = true;
  }
  k3_mark<<<g_tiles, 256, stage_smem, s>>>(c->reads_dev(), c->prm, bd, b3, c->d_state);
  k3_set_clear<<<g_small, 256, 0, s>>>(b3, c->d_state);
  k3_dupset<<<g_small, 256, 0, s>>>(bd, b3, c->d_state);
  k3_members<<<g_tiles, 256, 2 * stage_smem, s>>>(c->reads_dev(), c->prm, bd, b3, c->d_state);
  k3_open<<<g_small, 256, 0, s>>>(c->filt, b3, c->d_state);
  k3_conf<<<g_small, 256, 0, s>>>(c->reads_dev(), c->prm, bd, b3, c->d_state);
  k3_seg<<<g_small, 256, 0, s>>>(b3, c->d_state);
  k3_scatter<<<g_small, 256, 0, s>>>(c->prm, bd, b3, c->d_state);
  k3_sort<<<g_small, 256, 0, s>>>(b3, c->d_state);
  k3_frames<<<grid_for((uint64_t)b.nb * GRB_FRAME_SPLIT, 1, g_small * 2), 256, 0, s>>>(
    c->filt, c->prm, bd, b3, c->d_state);
  c->kend(GRB_K_DEDUPE, 10);
  {
    GrbReadsDev reads = c->reads_dev();
    GrbSelState* st = c->d_state;
    grb_decision* dec = d_dec;
    const uint64_t* dec_idx = c->bb_dec_idx[c->bb_set].p + b.read0;
    uint32_t us_a = us, n_cap_a = n_cap, dc_a = dc, cm_a = cm_smem;
    void* args[] = { &reads, &c->prm, &bd, &b3, &st, &dec, &dec_idx, &us_a, &n_cap_a, &dc_a, &cm_a };
    c->kbegin();
    int per_sm = 1;
    if (c->b3_bs == 256) {
      GRB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_fix<256>, 256, fix_smem));
    }
    if (c->b3_bs == 256 && per_sm >= 2) {
      GRB_CUDA(c, cudaLaunchCooperativeKernel((void*)k3_fix<256>, dim3(2 * c->sm_count), dim3(256), args,
                                              fix_smem, s));
    } else {
      GRB_CUDA(c, cudaLaunchCooperativeKernel((void*)k3_fix<512>, dim3(c->sm_count), dim3(512), args,
                                              fix_smem, s));
    }
    c->kend(GRB_K_COMMIT);
  }
  c->kbegin();
  k3_bulk<<<grid_for(b.n_bt, 1, c->sm_count * 16), 256, 0, s>>>(c->reads_dev(), c->filt, c->prm, bd,
                                                                b3, c->d_state);
  c->kend(GRB_K_INSERT);
  GRB_CUDA(c, cudaGetLastError());
  return GRB_OK;
}

int
grb_select_reads(grb_ctx* c, uint64_t first, uint64_t count, grb_decision* decisions,
                 grb_path_stats* stats, uint32_t stats_cap, uint32_t* n_stats, int* finished)
{
  cudaSetDevice(c->device);
  if (n_stats) {
    *n_stats = 0;
  }
  if (!c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_select_reads before grb_finalize_bitvector");
  }
  if (first + count > c->n_reads) {
    return c->fail(GRB_ERR_ARG, "grb_select_reads: range past the end of the read store");
  }
  cudaStream_t s = c->stream;
  uint64_t max_len = 0;
  for (uint64_t r = first; r < first + count; ++r) {
    if (c->h_flags[r] & GRB_READ_PASS2) {
      if (c->h_flags[r] & 4u) {
        return c->fail(GRB_ERR_ARG, "a read with non-ACGT bases is flagged GRB_READ_PASS2");
      }
      max_len = std::max<uint64_t>(max_len, c->h_len[r]);
    }
  }
  int rc = sel_prepare(c, max_len);
  if (rc != GRB_OK) {
    return rc;
  }
  GRB_CUDA(c, c->d_dec.reserve(std::max<uint64_t>(1, count), 0, s));
  GRB_CUDA(c, cudaMemsetAsync(c->d_dec.p, 0, std::max<uint64_t>(1, count) * sizeof(grb_decision), s));
  c->tic();
  const uint64_t end = first + count;
  uint64_t i = first;
  uint64_t stop_at = c->sel_finished ? first : end; // reads at or past it were never reached
  // A batch that covers a large part of the genome shares most of its ranks with itself and the
  // ordered commit degenerates: keep a batch below about half the genome (cfg1, 5 Mbp: 296 reads
  // per batch 57 ms, 128 reads 41 ms; cfg2 is capped by two reads per SM either way).
  uint32_t batch_reads = c->batch_reads;
  if (!c->batch_reads_fixed && max_len > 0 && c->p.genome_size > 0) {
    const uint64_t fit = c->p.genome_size / (2 * max_len);
    batch_reads = (uint32_t)std::min<uint64_t>(batch_reads, std::max<uint64_t>(32, fit));
  }
  // batches per chunk = between two looks at the loop state (path rollover / exit).  Chunks are
  // pipelined two deep (below), so the device does not wait for the host between them; after a
  // rollover the launches already queued return at once (state->halt) and the loop resumes at the
  // read after it: a shorter chunk wastes fewer such launches (five rollovers per silver run)
  static const uint64_t chunk_batches = [] {
    const char* e = getenv("GRB_CHUNK_BATCHES");
    const long v = e ? strtol(e, nullptr, 10) : 0;
    return (uint64_t)(v > 0 && v <= 1024 ? v : 4);
  }();
  const uint64_t kChunk = c->batch_mode ? chunk_batches * batch_reads : 256;
  if (c->batch_mode) {
    // Chunks of batches are pipelined two deep: chunk k + 1 is planned, described and queued while
    // chunk k runs, and only then is chunk k's loop state looked at (path rollover / exit).  A rollover
    // inside chunk k makes every launch queued after it return at once (state->halt), and the loop
    // resumes at the read after it -- the device never waits for the host between chunks.
    if (!c->h_state) {
      GRB_CUDA(c, cudaHostAlloc((void**)&c->h_state, 2 * sizeof(GrbSelState), cudaHostAllocDefault));
      GRB_CUDA(c, cudaEventCreateWithFlags(&c->ev_chunk[0], cudaEventDisableTiming));
      GRB_CUDA(c, cudaEventCreateWithFlags(&c->ev_chunk[1], cudaEventDisableTiming));
    }
    auto launch_chunk = [&](uint64_t from, int set, uint64_t* next) -> int {
      uint64_t launched = 0, j = from;
      c->bb_set = set;
      // cut the next kChunk visited reads into batches and describe them to the device
      BatchPlan bp;
      const uint64_t T = c->p.tile_length;
      uint64_t max_bt = 0, max_cm = 0;
      uint32_t max_nb = 0;
      // the commit indexes a batch's probes with 26 bits
      const uint64_t tile_budget = std::max<uint64_t>(
        1, std::min<uint64_t>(c->batch_tiles, ((1ull << 26) - 1) / (c->tile_frames * c->h_seed.h)));
      for (; j < end && launched < kChunk; ++j) {
        if (!(c->h_flags[j] & GRB_READ_PASS2)) {
          continue;
        }
        const uint32_t tiles = (uint32_t)(c->h_len[j] / T);
        if (bp.batches.empty() || bp.batches.back().nb >= batch_reads ||
            (bp.batches.back().n_bt && bp.batches.back().n_bt + tiles > tile_budget)) {
          if (!bp.batches.empty()) {
            bp.tile_first.push_back(bp.batches.back().n_bt);
          }
          bp.batches.push_back(BatchPlan::Batch{ (uint32_t)bp.read_idx.size(), 0,
                                                 (uint32_t)bp.tile_first.size(),
                                                 (uint32_t)bp.tile_read.size(), 0, 0, 0 });
        }
        BatchPlan::Batch& b = bp.batches.back();
        bp.read_idx.push_back(j);
        bp.dec_idx.push_back(j - first);
        bp.cm_off.push_back(b.cm_words);
        b.cm_words += (uint64_t)tiles * tiles;
        b.max_tiles = std::max(b.max_tiles, tiles);
        max_cm = std::max(max_cm, b.cm_words);
        bp.tile_first.push_back(b.n_bt);
        bp.tile_read.insert(bp.tile_read.end(), tiles, b.nb);
        b.nb += 1;
        b.n_bt += tiles;
        max_bt = std::max<uint64_t>(max_bt, b.n_bt);
        max_nb = std::max(max_nb, b.nb);
        ++launched;
      }
      if (!bp.batches.empty()) {
        bp.tile_first.push_back(bp.batches.back().n_bt);
        // with W ranks every per-tile buffer holds W equal shares (the last one padded)
        const uint64_t pad_bt = (c->comm && c->shard_query) ? (uint64_t)c->comm->world : 0;
        rc = commit_prepare(c, std::max<uint64_t>(max_bt, 1) + pad_bt, max_len / T, max_cm, max_nb);
        if (rc != GRB_OK) {
          return rc;
        }
        GRB_CUDA(c, c->bb_read_idx[set].reserve(bp.read_idx.size(), 0, s));
        GRB_CUDA(c, c->bb_tile_first[set].reserve(bp.tile_first.size(), 0, s));
        GRB_CUDA(c, c->bb_tile_read[set].reserve(std::max<size_t>(1, bp.tile_read.size()), 0, s));
        GRB_CUDA(c, cudaMemcpyAsync(c->bb_read_idx[set].p, bp.read_idx.data(), bp.read_idx.size() * 8,
                                    cudaMemcpyHostToDevice, s));
        GRB_CUDA(c, c->bb_dec_idx[set].reserve(bp.dec_idx.size(), 0, s));
        GRB_CUDA(c, cudaMemcpyAsync(c->bb_dec_idx[set].p, bp.dec_idx.data(), bp.dec_idx.size() * 8,
                                    cudaMemcpyHostToDevice, s));
        GRB_CUDA(c, c->bb_cm_off[set].reserve(bp.cm_off.size(), 0, s));
        GRB_CUDA(c, cudaMemcpyAsync(c->bb_cm_off[set].p, bp.cm_off.data(), bp.cm_off.size() * 8,
                                    cudaMemcpyHostToDevice, s));
        GRB_CUDA(c, cudaMemcpyAsync(c->bb_tile_first[set].p, bp.tile_first.data(),
                                    bp.tile_first.size() * 4, cudaMemcpyHostToDevice, s));
        if (!bp.tile_read.empty()) {
          GRB_CUDA(c, cudaMemcpyAsync(c->bb_tile_read[set].p, bp.tile_read.data(),
                                      bp.tile_read.size() * 4, cudaMemcpyHostToDevice, s));
        }
        for (const BatchPlan::Batch& b : bp.batches) {
          rc = launch_batch(c, b, c->d_dec.p);
          if (rc != GRB_OK) {
            return rc;
          }
        }
      }
      GRB_CUDA(c, cudaMemcpyAsync(&c->h_state[set], c->d_state, sizeof(GrbSelState), cudaMemcpyDeviceToHost, s));
      GRB_CUDA(c, cudaEventRecord(c->ev_chunk[set], s));
      *next = j;
      return GRB_OK;
    };
    struct Queued
    {
      bool valid;
      uint64_t next_i;
      int set;
    };
    Queued prev{ false, 0, 0 };
    int set = 0;
    while (true) {
      Queued cur{ false, 0, set };
      if (i < end && !c->sel_finished) {
        uint64_t j = i;
        rc = launch_chunk(i, set, &j);
        if (rc != GRB_OK) {
          return rc;
        }
        cur = Queued{ true, j, set };
      }
      if (prev.valid) {
        GRB_CUDA(c, cudaEventSynchronize(c->ev_chunk[prev.set]));
        GrbSelState st = c->h_state[prev.set];
        if (st.halt) {
          if (cur.valid) { // queued behind the halt: nothing of it ran
            GRB_CUDA(c, cudaEventSynchronize(c->ev_chunk[cur.set]));
          }
          if (st.n_snap && stats && n_stats && *n_stats < stats_cap) {
            stats[(*n_stats)++] = st.snap;
          }
          if (st.finished) {
            c->sel_finished = true;
            stop_at = st.halt_read + 1;
            break;
          }
          // rollover: reset_counts + reset_ID_vector, then resume after the read that triggered it
          GRB_CUDA(c, cudaMemsetAsync(c->filt.slots, 0, (c->filt.pop + 1) * sizeof(GrbSlot), s));
          st.halt = 0;
          st.n_snap = 0;
          GRB_CUDA(c, cudaMemcpyAsync(c->d_state, &st, sizeof st, cudaMemcpyHostToDevice, s));
          GRB_CUDA(c, cudaStreamSynchronize(s));
          i = st.halt_read + 1;
          prev.valid = false;
          continue;
        }
      }
      if (!cur.valid) {
        break;
      }
      prev = cur;
      i = cur.next_i;
      set ^= 1;
    }
  } else {
    while (i < end && !c->sel_finished) {
      uint64_t launched = 0, j = i;
      for (; j < end && launched < kChunk; ++j) {
        if (c->h_flags[j] & GRB_READ_PASS2) {
          launch_read(c, j, j - first, c->d_dec.p);
          ++launched;
        }
      }
      GrbSelState st;
      GRB_CUDA(c, cudaMemcpyAsync(&st, c->d_state, sizeof st, cudaMemcpyDeviceToHost, s));
      GRB_CUDA(c, cudaStreamSynchronize(s));
      c->kflush();
      if (st.halt) {
        if (st.n_snap && stats && n_stats && *n_stats < stats_cap) {
          stats[(*n_stats)++] = st.snap;
        }
        if (st.finished) {
          c->sel_finished = true;
          stop_at = st.halt_read + 1;
          break;
        }
        // rollover: reset_counts + reset_ID_vector, then resume after the read that triggered it
        GRB_CUDA(c, cudaMemsetAsync(c->filt.slots, 0, (c->filt.pop + 1) * sizeof(GrbSlot), s));
        st.halt = 0;
        st.n_snap = 0;
        GRB_CUDA(c, cudaMemcpyAsync(c->d_state, &st, sizeof st, cudaMemcpyHostToDevice, s));
        GRB_CUDA(c, cudaStreamSynchronize(s));
        i = st.halt_read + 1;
        continue;
      }
      i = j;
    }
  }
  GRB_CUDA(c, cudaMemcpyAsync(decisions, c->d_dec.p, count * sizeof(grb_decision),
                              cudaMemcpyDeviceToHost, s));
  c->toc();
  c->kflush(); // toc() waited for the stream: every launch group's event pair is complete
  GRB_CUDA(c, cudaGetLastError());
  for (uint64_t r = first; r < end; ++r) {
    grb_decision& d = decisions[r - first];
    if (r >= stop_at) {
      memset(&d, 0, sizeof d);
    } else if (!(c->h_flags[r] & GRB_READ_PASS2)) {
      memset(&d, 0, sizeof d);
      d.verdict = GRB_SKIPPED;
    }
  }
  if (finished) {
    *finished = c->sel_finished ? 1 : 0;
  }
  return GRB_OK;
}

int
grb_select_state(grb_ctx* c, grb_path_stats* current, uint64_t* curr_path, uint32_t* ids_inserted)
{
  cudaSetDevice(c->device);
  if (!c->sel_init) {
    return c->fail(GRB_ERR_STATE, "grb_select_state before grb_select_reads");
  }
  GrbSelState st;
  GRB_CUDA(c, cudaMemcpyAsync(&st, c->d_state, sizeof st, cudaMemcpyDeviceToHost, c->stream));
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (current) {
    *current = st.cur;
  }
  if (curr_path) {
    *curr_path = st.curr_path;
  }
  if (ids_inserted) {
    *ids_inserted = st.ids_inserted;
  }
  return GRB_OK;
}

int
grb_commit_profile(grb_ctx* c, uint64_t* out10)
{
  cudaSetDevice(c->device);
  if (!c->sel_init) {
    return c->fail(GRB_ERR_STATE, "grb_commit_profile before grb_select_reads");
  }
  GrbSelState st;
  GRB_CUDA(c, cudaMemcpyAsync(&st, c->d_state, sizeof st, cudaMemcpyDeviceToHost, c->stream));
  GRB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 10; ++i) {
    out10[i] = st.prof[i];
  }
  return GRB_OK;
}

int
grb_query_read(grb_ctx* c, uint64_t read_idx, uint32_t* best_id, uint32_t* best_count,
               uint32_t* n_cand, uint32_t* cand_ids, uint32_t* cand_counts, uint32_t cand_cap,
               uint64_t* counters)
{
  cudaSetDevice(c->device);
  if (!c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_query_read before grb_finalize_bitvector");
  }
  if (read_idx >= c->n_reads || (c->h_flags[read_idx] & 4u)) {
    return c->fail(GRB_ERR_ARG, "grb_query_read: bad read index or read with non-ACGT bases");
  }
  int rc = sel_prepare(c, c->h_len[read_idx]);
  if (rc != GRB_OK) {
    return rc;
  }
  cudaStream_t s = c->stream;
  const uint64_t tiles = c->h_len[read_idx] / c->p.tile_length;
  DevBuf<GrbSelState> tmp; // throw-away state so the loop counters are untouched
  GRB_CUDA(c, tmp.reserve(1, 0, s));
  GRB_CUDA(c, cudaMemsetAsync(tmp.p, 0, sizeof(GrbSelState), s));
  const unsigned qgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, 1024));
  k_query<512><<<qgrid, 512, c->query_smem, s>>>(c->reads_dev(), c->d_seed, c->filt, c->prm, c->sc,
                                                 tmp.p, read_idx);
  c->launches += 1;
  GrbSelState st;
  GRB_CUDA(c, cudaMemcpyAsync(&st, tmp.p, sizeof st, cudaMemcpyDeviceToHost, s));
  std::vector<uint32_t> ci(tiles * c->prm.cand_cap), cc(tiles * c->prm.cand_cap), nc(tiles);
  if (tiles) {
    GRB_CUDA(c, cudaMemcpyAsync(best_id, c->sc.best_id, tiles * 4, cudaMemcpyDeviceToHost, s));
    GRB_CUDA(c, cudaMemcpyAsync(best_count, c->sc.best_count, tiles * 4, cudaMemcpyDeviceToHost, s));
    GRB_CUDA(c, cudaMemcpyAsync(nc.data(), c->sc.n_cand, tiles * 4, cudaMemcpyDeviceToHost, s));
    GRB_CUDA(c, cudaMemcpyAsync(ci.data(), c->sc.cand_id, ci.size() * 4, cudaMemcpyDeviceToHost, s));
    GRB_CUDA(c, cudaMemcpyAsync(cc.data(), c->sc.cand_cnt, cc.size() * 4, cudaMemcpyDeviceToHost, s));
  }
  GRB_CUDA(c, cudaStreamSynchronize(s));
  GRB_CUDA(c, cudaGetLastError());
  for (uint64_t t = 0; t < tiles; ++t) {
    // order candidates like the reference's sorted list: count descending, then id ascending
    std::vector<std::pair<uint32_t, uint32_t>> v;
    for (uint32_t j = 0; j < nc[t]; ++j) {
      v.emplace_back(ci[t * c->prm.cand_cap + j], cc[t * c->prm.cand_cap + j]);
    }
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) {
      return a.second != b.second ? a.second > b.second : a.first < b.first;
    });
    n_cand[t] = nc[t];
    for (uint32_t j = 0; j < v.size() && j < cand_cap; ++j) {
      cand_ids[t * cand_cap + j] = v[j].first;
      cand_counts[t * cand_cap + j] = v[j].second;
    }
  }
  if (counters) {
    counters[0] += st.cur.queries;
    counters[1] += st.cur.hits;
    counters[2] += st.cur.misses;
  }
  return GRB_OK;
}

int
grb_insert_tiles(grb_ctx* c, uint64_t read_idx, uint32_t tile_start, uint32_t tile_end, uint32_t id)
{
  cudaSetDevice(c->device);
  if (!c->finalized) {
    return c->fail(GRB_ERR_STATE, "grb_insert_tiles before grb_finalize_bitvector");
  }
  if (read_idx >= c->n_reads || (c->h_flags[read_idx] & 4u)) {
    return c->fail(GRB_ERR_ARG, "grb_insert_tiles: bad read index or read with non-ACGT bases");
  }
  const uint64_t tiles = c->h_len[read_idx] / c->p.tile_length;
  if (tile_start >= tile_end || tile_end > tiles) {
    return c->fail(GRB_ERR_ARG, "grb_insert_tiles: empty or out-of-range tile range");
  }
  int rc = sel_prepare(c, c->h_len[read_idx]);
  if (rc != GRB_OK) {
    return rc;
  }
  cudaStream_t s = c->stream;
  const uint64_t h = c->h_seed.h;
  DevBuf<GrbSelState> tmp;
  GRB_CUDA(c, tmp.reserve(1, 0, s));
  GRB_CUDA(c, cudaMemsetAsync(tmp.p, 0, sizeof(GrbSelState), s));
  const unsigned qgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, 1024));
  k_query<512><<<qgrid, 512, c->query_smem, s>>>(c->reads_dev(), c->d_seed, c->filt, c->prm, c->sc,
                                                 tmp.p, read_idx); // fills the rank stash
  GrbReadPlan plan{};
  plan.verdict = GRB_UNTRIMMED;
  plan.trim_start = tile_start;
  plan.trim_end = tile_end - 1;
  plan.first_id = id;
  plan.n_blocks = 1;
  GRB_CUDA(c, cudaMemcpyAsync(c->sc.plan, &plan, sizeof plan, cudaMemcpyHostToDevice, s));
  GrbSelParams q = c->prm;
  q.block_size = tiles + 1; // the whole range is ONE insert call
  const uint64_t n = (uint64_t)(tile_end - tile_start) * c->tile_frames * h;
  const uint64_t tab = next_pow2(2 * n);
  DevBuf<uint64_t> key, mask;
  GRB_CUDA(c, key.reserve(tab, 0, s));
  GRB_CUDA(c, mask.reserve(tab, 0, s));
  GRB_CUDA(c, cudaMemsetAsync(key.p, 0xFF, tab * 8, s));
  GRB_CUDA(c, cudaMemsetAsync(mask.p, 0, tab * 8, s));
  GrbSelScratch sc = c->sc;
  sc.tab_key = key.p;
  sc.tab_mask = mask.p;
  k_insert_collect<<<grid_for(n, 256, c->sm_count * 4), 256, 0, s>>>(
    c->reads_dev(), q, sc, sc.stash, tmp.p, read_idx, 0, (uint32_t)tab);
  k_insert_apply<<<grid_for(tab, 256, c->sm_count * 4), 256, 0, s>>>(
    c->filt, sc, tmp.p, read_idx, 0, (uint32_t)tab);
  c->launches += 3;
  GRB_CUDA(c, cudaStreamSynchronize(s));
  GRB_CUDA(c, cudaGetLastError());
  return GRB_OK;
}

} // extern "C"
