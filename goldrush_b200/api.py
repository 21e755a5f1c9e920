"""ctypes mirror of include/goldrush_b200.h (same names, same argument meaning, errors -> GrbError)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "_lib", "libgoldrush_b200.so")


class GrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"goldrush_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """grb_params (opt:: globals, goldrush_path/opt.hpp:9-38)."""
    _fields_ = [
        ("assigned_max", C.c_uint64), ("unassigned_min", C.c_uint64), ("tile_length", C.c_uint64),
        ("block_size", C.c_uint64), ("hash_universe", C.c_uint64), ("genome_size", C.c_uint64),
        ("kmer_size", C.c_uint64), ("weight", C.c_uint64), ("min_length", C.c_uint64),
        ("hash_num", C.c_uint64), ("occupancy", C.c_double), ("ratio", C.c_double),
        ("max_paths", C.c_uint64), ("threshold", C.c_uint64), ("phred_min", C.c_uint32),
        ("phred_delta", C.c_uint32), ("silver_path", C.c_int32), ("device", C.c_int32),
        ("seeds", C.POINTER(C.c_char_p)),
    ]


class ReadMeta(C.Structure):
    _fields_ = [
        ("hdr_off", C.c_uint64), ("seq_off", C.c_uint64), ("qual_off", C.c_uint64),
        ("hdr_len", C.c_uint32), ("len", C.c_uint32), ("phred_first_half_sum", C.c_double),
        ("phred_total_sum", C.c_double), ("non_acgt", C.c_uint32), ("qual_len", C.c_uint32),
        ("name_hash", C.c_uint64),
    ]


class Decision(C.Structure):
    _fields_ = [
        ("verdict", C.c_uint8), ("pad", C.c_uint8 * 3), ("path", C.c_uint32),
        ("trim_start", C.c_uint32), ("trim_end", C.c_uint32), ("num_tiles", C.c_uint32),
        ("num_assigned", C.c_uint32),
    ]


class PathStats(C.Structure):
    _fields_ = [
        ("valid_reads", C.c_uint64), ("total_tiles", C.c_uint64), ("assigned_tiles", C.c_uint64),
        ("unassigned_tiles", C.c_uint64), ("queries", C.c_uint64), ("hits", C.c_uint64),
        ("misses", C.c_uint64), ("num_reads_in_path", C.c_uint64), ("inserted_bases", C.c_uint64),
        ("rollover_read", C.c_uint64), ("phred_sum_in_path", C.c_double),
    ]


class ProbeBenchResult(C.Structure):
    _fields_ = [("query_ms", C.c_double), ("insert_ms", C.c_double), ("probes", C.c_uint64),
                ("pop", C.c_uint64), ("filter_bits", C.c_uint64), ("footprint_bytes", C.c_uint64),
                ("checksum", C.c_uint64), ("keys_filled", C.c_uint64), ("probes_missed", C.c_uint64),
                ("line_query_ms", C.c_double), ("line_bytes", C.c_uint64)]


class RunOptions(C.Structure):
    _fields_ = [
        ("params", Params), ("seed_preset", C.c_char_p), ("prefix", C.c_char_p),
        ("filter_file", C.c_char_p), ("input_path", C.c_char_p), ("ntcard", C.c_int32),
        ("verbose", C.c_int32), ("debug", C.c_int32), ("write_outputs", C.c_int32),
        ("quiet", C.c_int32), ("jobs", C.c_int32), ("fastq_offset", C.c_uint64),
        ("fastq_total", C.c_uint64),
    ]


class RunResult(C.Structure):
    _fields_ = [
        ("num_reads", C.c_uint64), ("num_passed_reads", C.c_uint64), ("bases_pass1", C.c_uint64),
        ("reads_visited", C.c_uint64), ("bases_pass2", C.c_uint64), ("reads_selected", C.c_uint64),
        ("bases_selected", C.c_uint64), ("filter_bits", C.c_uint64), ("pop", C.c_uint64),
        ("phred_min", C.c_uint32), ("paths", C.c_uint32), ("ms_ingest", C.c_double),
        ("ms_pass1", C.c_double), ("ms_rank", C.c_double), ("ms_pass2", C.c_double),
        ("ms_wall", C.c_double), ("launches", C.c_uint64), ("out_digest", C.c_uint64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class PolishParams(C.Structure):
    _fields_ = [("hash_num", C.c_uint32), ("n_k", C.c_uint32), ("k_values", C.POINTER(C.c_uint32)),
                ("cbf_bytes", C.c_uint64), ("bf_bytes", C.c_uint64)]


class HostMsg(C.Structure):
    _fields_ = [("peer", C.c_int32), ("pad", C.c_int32), ("ptr", C.c_void_p), ("bytes", C.c_uint64)]


class SynthParams(C.Structure):
    _fields_ = [
        ("genome_len", C.c_uint64), ("seed", C.c_uint64), ("coverage", C.c_double),
        ("read_len", C.c_uint32), ("n50", C.c_uint32), ("sub_rate", C.c_double),
        ("ins_rate", C.c_double), ("del_rate", C.c_double), ("qmin", C.c_uint32),
        ("qmax", C.c_uint32),
    ]


READ_META_DTYPE = np.dtype([("hdr_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8"),
                            ("hdr_len", "<u4"), ("len", "<u4"), ("phred_first_half_sum", "<f8"),
                            ("phred_total_sum", "<f8"), ("non_acgt", "<u4"), ("qual_len", "<u4"),
                            ("name_hash", "<u8")])
DECISION_DTYPE = np.dtype([("verdict", "u1"), ("pad", "u1", (3,)), ("path", "<u4"),
                           ("trim_start", "<u4"), ("trim_end", "<u4"), ("num_tiles", "<u4"),
                           ("num_assigned", "<u4")])

VERDICTS = {0: "not_visited", 1: "skipped", 2: "untrimmed", 3: "trimmed", 4: "assigned"}
READ_PASS1, READ_PASS2 = 1, 2
KERNEL_CLASSES = ["fill", "rank", "query", "decide", "insert", "smooth", "dedupe", "commit", "gather"]

_lib = None


def lib():
    """Loads libgoldrush_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise GrbError(-2, f"{path} is missing: run `make lib` (or __graft_entry__.build())")
    L = C.CDLL(path)
    u64, u32, i32, sz, dbl, vp = C.c_uint64, C.c_uint32, C.c_int, C.c_size_t, C.c_double, C.c_void_p
    P = C.POINTER
    sigs = {
        "grb_params_default": (None, [P(Params)]),
        "grb_make_seed_pattern": (i32, [C.c_char_p, C.c_uint, C.c_uint, C.c_uint, P(C.c_char_p)]),
        "grb_calc_optimal_size": (u64, [u64, C.c_uint, dbl]),
        "grb_default_hash_universe": (u64, [u64, u64, u64]),
        "grb_phred_finalize": (None, [dbl, dbl, u64, P(u32), P(u32)]),
        "grb_phred_finalize_batch": (None, [vp, u64, vp, vp]),
        "grb_query_sharded": (i32, [vp]),
        "grb_create": (i32, [P(Params), P(vp)]),
        "grb_destroy": (None, [vp]),
        "grb_last_error": (C.c_char_p, [vp]),
        "grb_launch_count": (u64, [vp]),
        "grb_cached_memory_bytes": (u64, []),
        "grb_release_cached_memory": (None, []),
        "grb_reads_reserve": (i32, [vp, u64]),
        "grb_host_pin": (i32, [vp, sz, P(sz)]),
        "grb_host_unpin": (None, [vp, sz]),
        "grb_reads_ingest_fastq": (i32, [vp, vp, sz, i32, P(sz)]),
        "grb_reads_readahead": (i32, [vp, vp, sz]),
        "grb_reads_count": (u64, [vp]),
        "grb_reads_set_origin": (i32, [vp, u64]),
        "grb_reads_allgather": (i32, [vp]),
        "grb_reads_own_range": (i32, [vp, P(u64), P(u64)]),
        "grb_comm_allgather_host": (i32, [vp, vp, u64, vp, u64, P(u64)]),
        "grb_comm_exchange_host": (i32, [vp, vp, u32, vp, u32]),
        "grb_reads_get_meta": (i32, [vp, u64, u64, P(ReadMeta)]),
        "grb_reads_set_flags": (i32, [vp, u64, u64, vp]),
        "grb_reads_clear": (None, [vp]),
        "grb_phred_sums": (i32, [vp, C.c_char_p, sz, P(dbl), P(dbl)]),
        "grb_estimate_cardinality": (i32, [vp, u64, P(u64), P(u64)]),
        "grb_filter_alloc": (i32, [vp, u64]),
        "grb_build_bitvector": (i32, [vp]),
        "grb_build_bitvector_range": (i32, [vp, u64, u64]),
        "grb_finalize_bitvector": (i32, [vp, P(u64)]),
        "grb_reset_ids": (i32, [vp]),
        "grb_select_reads": (i32, [vp, u64, u64, P(Decision), P(PathStats), u32, P(u32), P(i32)]),
        "grb_select_state": (i32, [vp, P(PathStats), P(u64), P(u32)]),
        "grb_hash_sequence": (i32, [vp, C.c_char_p, sz, vp]),
        "grb_copy_bitvector": (i32, [vp, vp]),
        "grb_load_bitvector": (i32, [vp, vp]),
        "grb_rank": (i32, [vp, vp, sz, vp, vp]),
        "grb_get_ids": (i32, [vp, vp, sz, vp, vp]),
        "grb_set_ids": (i32, [vp, vp, sz, vp, vp]),
        "grb_query_read": (i32, [vp, u64, vp, vp, vp, vp, vp, u32, vp]),
        "grb_insert_tiles": (i32, [vp, u64, u32, u32, u32]),
        "grb_comm_unique_id": (i32, [vp]),
        "grb_comm_init": (i32, [vp, i32, i32, i32]),
        "grb_comm_destroy": (None, []),
        "grb_comm_info": (i32, [vp, P(i32), P(i32)]),
        "grb_bitvector_or_reduce": (i32, [vp]),
        "grb_bitvector_device": (i32, [vp, P(vp), P(u64)]),
        "grb_or_words": (i32, [vp, vp, vp, u64]),
        "grb_sync": (i32, [vp]),
        "grb_last_device_ms": (dbl, [vp]),
        "grb_stream": (i32, [vp, P(vp)]),
        "grb_profile_enable": (i32, [vp, i32]),
        "grb_kernel_time": (i32, [vp, i32, P(dbl), P(u64)]),
        "grb_commit_profile": (i32, [vp, P(u64)]),
        "grb_probe_bench": (i32, [vp, u64, dbl, u32, u64, u64, i32, P(ProbeBenchResult)]),
        "grb_run_path": (i32, [P(RunOptions), vp, sz, P(RunResult), C.c_char_p, sz]),
        "grb_run_two_stage": (i32, [P(RunOptions), P(RunOptions), vp, sz, P(RunResult), P(RunResult),
                                    C.c_char_p, sz]),
        "grb_synth_num_reads": (u64, [P(SynthParams)]),
        "grb_synth_fastq": (vp, [P(SynthParams), u64, u64, P(u64)]),
        "grb_free_host": (None, [vp]),
        "grb_test_group_hash_host": (i32, [P(C.c_char_p), u32, C.c_char_p, sz, vp]),
        "grb_test_next_record_start": (sz, [C.c_char_p, sz, sz]),
        "grb_abi_sizes": (None, [vp]),
        "grb_polish_kmer_threshold": (i32, [u64]),
        "grb_polish_plan_target": (i32, [u64, dbl, u32, P(C.c_char_p), vp, vp, vp, P(u32), P(C.c_int32)]),
        "grb_polish_fill_batches": (i32, [vp, P(PolishParams), u32, vp, C.c_char_p, vp, vp, vp]),
        "grb_test_polish_fill_host": (i32, [P(PolishParams), u32, vp, C.c_char_p, vp, vp, vp]),
        "grb_test_polish_fill_host_grouped": (i32, [P(PolishParams), u32, vp, C.c_char_p, vp, vp, vp, P(u64), P(u64)]),
        "grb_polish_index_build": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, sz]),
        "grb_polish_inputs_open": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, dbl, P(vp), C.c_char_p, sz]),
        "grb_polish_inputs_close": (None, [vp]),
        "grb_polish_inputs_mappings": (C.c_int64, [vp, C.c_char_p, C.c_char_p, sz]),
        "grb_polish_serve_batches": (i32, [vp, vp, P(PolishParams), dbl, u32, vp, P(C.c_char_p), vp, C.c_char_p, sz]),
        "grb_test_polish_serve_batches_host": (i32, [vp, P(PolishParams), dbl, u32, vp, P(C.c_char_p), vp, P(u64),
                                                     P(u64), C.c_char_p, sz]),
        "grb_test_plan_silver_parts": (i32, [vp, u32, C.c_int32, vp, vp]),
        "grb_test_decide_host": (i32, [u32, vp, vp, vp, vp, vp, u32, u64, u64, u64, u64, u64, u64,
                                       P(u32), vp, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._grb_symbols = sorted(sigs)
    _lib = L
    return L


# ---- host-side scalar helpers -------------------------------------------------------------------
def make_seed_pattern(preset, k, weight, h):
    """spaced_seeds.cpp:7-68"""
    bufs = [C.create_string_buffer(k + h + 2) for _ in range(h)]
    arr = (C.c_char_p * h)(*[C.cast(b, C.c_char_p) for b in bufs])
    rc = lib().grb_make_seed_pattern((preset or "").encode(), k, weight, h, arr)
    if rc:
        raise GrbError(rc, "cannot design a spaced seed for this k / w")
    return [b.value.decode() for b in bufs]


def calc_optimal_size(entries, hash_num, occupancy):
    return lib().grb_calc_optimal_size(entries, hash_num, occupancy)


def default_hash_universe(weight, genome_size, hash_num):
    return lib().grb_default_hash_universe(weight, genome_size, hash_num)


def phred_finalize(first_half_sum, total_sum, n):
    a, d = C.c_uint32(), C.c_uint32()
    lib().grb_phred_finalize(first_half_sum, total_sum, n, C.byref(a), C.byref(d))
    return a.value, d.value


# ---- multi-GPU: the process-wide NCCL communicator of the library (include/goldrush_b200.h) ------
def comm_unique_id() -> bytes:
    """Rank 0: a fresh NCCL unique id (128 bytes) to hand to every other rank."""
    buf = C.create_string_buffer(128)
    rc = lib().grb_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc:
        raise GrbError(rc, (lib().grb_last_error(None) or b"").decode())
    return buf.raw


def comm_init(ident: bytes, rank: int, world: int, device: int):
    """Every rank: joins the communicator; Engines / run_path created afterwards on `device` shard
    pass 1 and the speculative query across the ranks."""
    if len(ident) != 128:
        raise GrbError(-1, "comm_init: the NCCL unique id is 128 bytes")
    buf = C.create_string_buffer(ident, 128)
    rc = lib().grb_comm_init(C.cast(buf, C.c_void_p), rank, world, device)
    if rc:
        raise GrbError(rc, (lib().grb_last_error(None) or b"").decode())


def comm_destroy():
    lib().grb_comm_destroy()


def comm_info(engine=None):
    r, w = C.c_int(), C.c_int()
    lib().grb_comm_info(engine._h if engine is not None else None, C.byref(r), C.byref(w))
    return r.value, w.value


def default_params(**kw):
    p = Params()
    lib().grb_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One grb_ctx.  Method names follow the C ABI without the grb_ prefix."""

    def __init__(self, seeds, device=0, **params):
        self._L = lib()
        self.seeds = list(seeds)
        p = default_params(**params)
        p.hash_num = len(self.seeds)
        p.kmer_size = params.get("kmer_size", len(self.seeds[0]))
        p.device = device
        self._seed_arr = (C.c_char_p * len(self.seeds))(*[s.encode() for s in self.seeds])
        p.seeds = C.cast(self._seed_arr, C.POINTER(C.c_char_p))
        self.params = p
        h = C.c_void_p()
        rc = self._L.grb_create(C.byref(p), C.byref(h))
        if rc:
            raise GrbError(rc, self._L.grb_last_error(None).decode())
        self._h = h
        self.h = len(self.seeds)
        self.k = len(self.seeds[0])

    def close(self):
        if getattr(self, "_h", None):
            self._L.grb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc:
            raise GrbError(rc, self._L.grb_last_error(self._h).decode())

    # K1
    def reads_ingest_fastq(self, data, final=True, nbytes=None):
        """data: bytes, or a raw host address (int) with nbytes."""
        used = C.c_size_t()
        n = len(data) if nbytes is None else nbytes
        self._chk(self._L.grb_reads_ingest_fastq(self._h, data, n, int(final), C.byref(used)))
        return used.value

    def reads_readahead(self, base_ptr, total):
        """Consecutive ingest calls will walk the host range [base_ptr, base_ptr + total)."""
        self._chk(self._L.grb_reads_readahead(self._h, base_ptr, total))

    def reads_meta_array(self):
        """All read metadata as one numpy structured array (fields of grb_read_meta)."""
        n = self.reads_count()
        arr = np.zeros(max(1, n), dtype=READ_META_DTYPE)
        if n:
            self._chk(self._L.grb_reads_get_meta(self._h, 0, n,
                                                 C.cast(_ptr(arr), C.POINTER(ReadMeta))))
        return arr[:n]

    def reads_count(self):
        return self._L.grb_reads_count(self._h)

    def polish_fill_batches(self, params, batches, out=None):
        """grb_polish_fill_batches: Bloom filters [batch][k index][bf_bytes] of the batches' reads
        (out: optional preallocated uint8 array, e.g. page-locked with host_pin)."""
        seqs, off, thr, first = _polish_inputs(batches)
        if out is None:
            out = np.zeros(max(1, len(batches) * params.n_k * params.bf_bytes), dtype=np.uint8)
        self._chk(self._L.grb_polish_fill_batches(self._h, C.byref(params), len(batches), _ptr(first), seqs,
                                                  _ptr(off), _ptr(thr), _ptr(out)))
        return out.reshape(len(batches), params.n_k, params.bf_bytes)

    def polish_serve_batches(self, inputs, params, subsample_max_per_10kbp, batches, out=None):
        """grb_polish_serve_batches: batches = lists of target ids; Bloom filters [batch][k index][bf_bytes]."""
        first, ids = _polish_target_lists(batches)
        if out is None:
            out = np.zeros(max(1, len(batches) * params.n_k * params.bf_bytes), dtype=np.uint8)
        err = C.create_string_buffer(1024)
        rc = self._L.grb_polish_serve_batches(self._h, inputs._h, C.byref(params), float(subsample_max_per_10kbp),
                                              len(batches), _ptr(first), ids, _ptr(out), err, len(err))
        if rc:
            raise GrbError(rc, err.value.decode(errors="replace"))
        return out.reshape(len(batches), params.n_k, params.bf_bytes)

    def reads_set_origin(self, byte_offset):
        self._chk(self._L.grb_reads_set_origin(self._h, int(byte_offset)))

    def reads_allgather(self):
        """Several GPUs: every rank ingested its own slice; afterwards every store holds all reads."""
        self._chk(self._L.grb_reads_allgather(self._h))

    def reads_own_range(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._chk(self._L.grb_reads_own_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def reads_get_meta(self, first=0, count=None):
        if count is None:
            count = self.reads_count() - first
        arr = (ReadMeta * max(1, count))()
        if count:
            self._chk(self._L.grb_reads_get_meta(self._h, first, count, arr))
        return list(arr)[:count]

    def reads_set_flags(self, flags, first=0):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        self._chk(self._L.grb_reads_set_flags(self._h, first, len(f), _ptr(f)))

    def reads_clear(self):
        self._L.grb_reads_clear(self._h)

    def phred_sums(self, qual: bytes):
        a, b = C.c_double(), C.c_double()
        self._chk(self._L.grb_phred_sums(self._h, qual, len(qual), C.byref(a), C.byref(b)))
        return a.value, b.value

    # K5
    def estimate_cardinality(self, input_bytes):
        per = (C.c_uint64 * self.h)()
        tot = C.c_uint64()
        self._chk(self._L.grb_estimate_cardinality(self._h, input_bytes, per, C.byref(tot)))
        return list(per), tot.value

    # filter
    def filter_alloc(self, bits):
        self._chk(self._L.grb_filter_alloc(self._h, bits))
        self.filter_bits = bits

    def build_bitvector(self, first=None, count=None):
        if first is None:
            self._chk(self._L.grb_build_bitvector(self._h))
        else:
            self._chk(self._L.grb_build_bitvector_range(self._h, first, count))

    def bitvector_or_reduce(self):
        self._chk(self._L.grb_bitvector_or_reduce(self._h))

    def probe_bench(self, filter_bits, fill, h, n_probes=1 << 28, seed=42, reps=3):
        """grb_probe_bench: query / insert probe throughput at one filter size (cfg5)."""
        out = ProbeBenchResult()
        self._chk(self._L.grb_probe_bench(self._h, filter_bits, fill, h, n_probes, seed, reps,
                                          C.byref(out)))
        return out

    def finalize_bitvector(self):
        pop = C.c_uint64()
        self._chk(self._L.grb_finalize_bitvector(self._h, C.byref(pop)))
        self.pop = pop.value
        return pop.value

    def reset_ids(self):
        self._chk(self._L.grb_reset_ids(self._h))

    def copy_bitvector(self):
        w = np.zeros((self.filter_bits + 63) // 64, dtype=np.uint64)
        self._chk(self._L.grb_copy_bitvector(self._h, _ptr(w)))
        return w

    def load_bitvector(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        self._chk(self._L.grb_load_bitvector(self._h, _ptr(w)))

    def rank(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        r = np.zeros(len(pos), dtype=np.uint64)
        b = np.zeros(len(pos), dtype=np.uint8)
        self._chk(self._L.grb_rank(self._h, _ptr(pos), len(pos), _ptr(r), _ptr(b)))
        return r, b

    def get_ids(self, rank):
        rank = np.ascontiguousarray(rank, dtype=np.uint64)
        i = np.zeros(len(rank), dtype=np.uint32)
        c = np.zeros(len(rank), dtype=np.uint32)
        self._chk(self._L.grb_get_ids(self._h, _ptr(rank), len(rank), _ptr(i), _ptr(c)))
        return i, c

    def set_ids(self, rank, ids, counts):
        rank = np.ascontiguousarray(rank, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        self._chk(self._L.grb_set_ids(self._h, _ptr(rank), len(rank), _ptr(ids), _ptr(counts)))

    def hash_sequence(self, seq: bytes):
        frames = len(seq) - self.k + 1
        out = np.zeros(max(0, frames) * self.h, dtype=np.uint64)
        self._chk(self._L.grb_hash_sequence(self._h, seq, len(seq), _ptr(out)))
        return out.reshape(frames, self.h)

    # selection loop
    def query_read(self, read_idx, num_tiles, cand_cap=64):
        bi = np.zeros(max(1, num_tiles), dtype=np.uint32)
        bc = np.zeros(max(1, num_tiles), dtype=np.uint32)
        nc = np.zeros(max(1, num_tiles), dtype=np.uint32)
        ci = np.zeros(max(1, num_tiles) * cand_cap, dtype=np.uint32)
        cc = np.zeros(max(1, num_tiles) * cand_cap, dtype=np.uint32)
        cnt = np.zeros(3, dtype=np.uint64)
        self._chk(self._L.grb_query_read(self._h, read_idx, _ptr(bi), _ptr(bc), _ptr(nc), _ptr(ci),
                                         _ptr(cc), cand_cap, _ptr(cnt)))
        return (bi[:num_tiles], bc[:num_tiles], nc[:num_tiles],
                ci.reshape(-1, cand_cap)[:num_tiles], cc.reshape(-1, cand_cap)[:num_tiles], cnt)

    def insert_tiles(self, read_idx, tile_start, tile_end, id_):
        self._chk(self._L.grb_insert_tiles(self._h, read_idx, tile_start, tile_end, id_))

    def select_reads(self, first=0, count=None, stats_cap=64):
        if count is None:
            count = self.reads_count() - first
        dec = (Decision * max(1, count))()
        st = (PathStats * stats_cap)()
        ns = C.c_uint32()
        fin = C.c_int()
        self._chk(self._L.grb_select_reads(self._h, first, count, dec, st, stats_cap, C.byref(ns),
                                           C.byref(fin)))
        return list(dec)[:count], list(st)[:ns.value], bool(fin.value)

    def select_reads_array(self, first=0, count=None, stats_cap=64):
        """select_reads with the decisions as one numpy structured array."""
        if count is None:
            count = self.reads_count() - first
        dec = np.zeros(max(1, count), dtype=DECISION_DTYPE)
        st = (PathStats * stats_cap)()
        ns = C.c_uint32()
        fin = C.c_int()
        self._chk(self._L.grb_select_reads(self._h, first, count,
                                           C.cast(_ptr(dec), C.POINTER(Decision)), st, stats_cap,
                                           C.byref(ns), C.byref(fin)))
        return dec[:count], list(st)[:ns.value], bool(fin.value)

    def select_state(self):
        st = PathStats()
        cp = C.c_uint64()
        ids = C.c_uint32()
        self._chk(self._L.grb_select_state(self._h, C.byref(st), C.byref(cp), C.byref(ids)))
        return st, cp.value, ids.value

    # plumbing
    def bitvector_device(self):
        p = C.c_void_p()
        n = C.c_uint64()
        self._chk(self._L.grb_bitvector_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def or_words(self, dst_ptr, src_ptr, n_words):
        self._chk(self._L.grb_or_words(self._h, dst_ptr, src_ptr, n_words))

    def sync(self):
        self._chk(self._L.grb_sync(self._h))

    def last_device_ms(self):
        return self._L.grb_last_device_ms(self._h)

    def launch_count(self):
        return self._L.grb_launch_count(self._h)

    def stream(self):
        p = C.c_void_p()
        self._chk(self._L.grb_stream(self._h, C.byref(p)))
        return p.value

    def profile_enable(self, on=True):
        self._chk(self._L.grb_profile_enable(self._h, int(on)))

    def commit_profile(self):
        out = (C.c_uint64 * 10)()
        self._chk(self._L.grb_commit_profile(self._h, out))
        names = ["walk_cyc", "scans_with_a_changed_plan", "revalidate_cyc", "wait_and_scan_cyc",
                 "final_cyc", "conflict_frames", "reads", "iterations", "reads_inserting", "batches"]
        return dict(zip(names, [int(x) for x in out]))

    def kernel_time(self, kclass):
        """(milliseconds, launches) of one grb_kernel_class since profile_enable()."""
        ms, n = C.c_double(), C.c_uint64()
        self._chk(self._L.grb_kernel_time(self._h, KERNEL_CLASSES.index(kclass), C.byref(ms),
                                          C.byref(n)))
        return ms.value, n.value


def _run_options(input_path, prefix, seed_preset, filter_file, ntcard, verbose, write_outputs,
                 quiet, device, params):
    o = RunOptions()
    params = dict(params)
    o.fastq_offset = int(params.pop("fastq_offset", 0))  # slice mode (several GPUs), see the header
    o.fastq_total = int(params.pop("fastq_total", 0))
    o.params = default_params(**params)
    o.params.device = device
    o.seed_preset = (seed_preset or "").encode()
    o.prefix = prefix.encode()
    o.filter_file = filter_file.encode() if filter_file else None
    o.input_path = input_path.encode() if input_path else None
    o.ntcard = int(ntcard)
    o.verbose = int(verbose)
    o.write_outputs = int(write_outputs)
    o.quiet = int(quiet)
    return o


def run_path(fastq=None, input_path=None, prefix="goldrush_out", seed_preset="",
             filter_file=None, ntcard=False, verbose=False, write_outputs=True, quiet=True,
             device=0, nbytes=None, **params):
    """grb_run_path: the whole GoldRush-Path stage (goldrush_path.cpp main()).
    fastq: bytes, or a raw host address (int) with nbytes; None = read input_path."""
    L = lib()
    o = _run_options(input_path, prefix, seed_preset, filter_file, ntcard, verbose, write_outputs,
                     quiet, device, params)
    res = RunResult()
    err = C.create_string_buffer(1024)
    if nbytes is None:
        nbytes = len(fastq) if fastq is not None else 0
    rc = L.grb_run_path(C.byref(o), fastq, nbytes, C.byref(res), err, len(err))
    if rc:
        raise GrbError(rc, err.value.decode())
    return res


def run_two_stage(fastq, silver: dict, golden: dict, input_path=None, nbytes=None, device=0,
                  write_outputs=True, quiet=True):
    """grb_run_two_stage: the silver run and the golden run on its concatenated paths
    (bin/goldrush:240-260) in one call.  `silver` / `golden` are run_path keyword dicts
    (prefix, seed_preset, verbose and the opt:: parameters)."""
    L = lib()

    def opts(kw):
        kw = dict(kw)
        return _run_options(kw.pop("input_path", input_path), kw.pop("prefix", "goldrush_out"),
                            kw.pop("seed_preset", ""), kw.pop("filter_file", None),
                            kw.pop("ntcard", False), kw.pop("verbose", False),
                            kw.pop("write_outputs", write_outputs), kw.pop("quiet", quiet), device, kw)

    so, go = opts(silver), opts(golden)
    rs, rg = RunResult(), RunResult()
    err = C.create_string_buffer(1024)
    if nbytes is None:
        nbytes = len(fastq) if fastq is not None else 0
    rc = L.grb_run_two_stage(C.byref(so), C.byref(go), fastq, nbytes, C.byref(rs), C.byref(rg), err,
                             len(err))
    if rc:
        raise GrbError(rc, err.value.decode())
    return rs, rg


# ---- synthetic reads ------------------------------------------------------------------------------
def synth_params(genome_len, coverage, read_len, seed, err=0.01, n50=20000, qmin=12, qmax=30):
    return SynthParams(genome_len, seed, coverage, read_len, n50, err, err, err, qmin, qmax)


def synth_num_reads(sp):
    return lib().grb_synth_num_reads(C.byref(sp))


def synth_fastq_raw(sp, first=0, count=None):
    """(host address, bytes) of a malloc'ed FASTQ buffer; release with free_host()."""
    L = lib()
    if count is None:
        count = L.grb_synth_num_reads(C.byref(sp)) - first
    n = C.c_uint64()
    p = L.grb_synth_fastq(C.byref(sp), first, count, C.byref(n))
    if not p:
        raise MemoryError("grb_synth_fastq")
    return p, n.value


def free_host(p):
    lib().grb_free_host(p)


# ---- (f4) GoldPolish targeted Bloom filters ----
def polish_params(k_values, hash_num=4, cbf_bytes=10 << 20, bf_bytes=512 << 10):
    ks = np.ascontiguousarray(k_values, dtype=np.uint32)
    p = PolishParams(hash_num, len(ks), ks.ctypes.data_as(C.POINTER(C.c_uint32)), cbf_bytes, bf_bytes)
    p._keep = ks
    return p


def polish_plan_target(target_len, subsample_max_per_10kbp, ids, phred_avg, lens):
    """grb_polish_plan_target: (order, n_used, kmer_threshold) of one target's mapped reads."""
    n = len(ids)
    arr = (C.c_char_p * n)(*[i.encode() for i in ids])
    ph = np.ascontiguousarray(phred_avg, dtype=np.float64)
    ln = np.ascontiguousarray(lens, dtype=np.uint64)
    order = np.zeros(max(1, n), dtype=np.uint32)
    used, thr = C.c_uint32(), C.c_int32()
    rc = lib().grb_polish_plan_target(int(target_len), float(subsample_max_per_10kbp), n, arr, _ptr(ph),
                                      _ptr(ln), _ptr(order), C.byref(used), C.byref(thr))
    if rc:
        raise GrbError(rc, "grb_polish_plan_target")
    return order[:n], used.value, thr.value


def _polish_inputs(batches):
    """batches: list of lists of (sequence bytes, threshold) in insertion order."""
    seqs, off, thr, first = [], [0], [], [0]
    for b in batches:
        for s, t in b:
            seqs.append(s)
            off.append(off[-1] + len(s))
            thr.append(t)
        first.append(len(thr))
    return (b"".join(seqs), np.array(off, dtype=np.uint64), np.array(thr + [0], dtype=np.uint32),
            np.array(first, dtype=np.uint64))


def polish_fill_host_grouped(params, batches):
    """The warp kernel's algorithm emulated on the host: (filters, groups, groups replayed in lane order)."""
    seqs, off, thr, first = _polish_inputs(batches)
    out = np.zeros(max(1, len(batches) * params.n_k * params.bf_bytes), dtype=np.uint8)
    g, o = C.c_uint64(), C.c_uint64()
    rc = lib().grb_test_polish_fill_host_grouped(C.byref(params), len(batches), _ptr(first), seqs, _ptr(off),
                                                 _ptr(thr), _ptr(out), C.byref(g), C.byref(o))
    if rc:
        raise GrbError(rc, "grb_test_polish_fill_host_grouped")
    return out.reshape(len(batches), params.n_k, params.bf_bytes), g.value, o.value


def polish_fill_host(params, batches):
    seqs, off, thr, first = _polish_inputs(batches)
    out = np.zeros(max(1, len(batches) * params.n_k * params.bf_bytes), dtype=np.uint8)
    rc = lib().grb_test_polish_fill_host(C.byref(params), len(batches), _ptr(first), seqs, _ptr(off),
                                         _ptr(thr), _ptr(out))
    if rc:
        raise GrbError(rc, "grb_test_polish_fill_host")
    return out.reshape(len(batches), params.n_k, params.bf_bytes)


def polish_index_build(seqs_path, index_path):
    """grb_polish_index_build: what goldpolish-index writes for a FASTA / FASTQ file."""
    err = C.create_string_buffer(1024)
    rc = lib().grb_polish_index_build(os.fsencode(seqs_path), os.fsencode(index_path), err, len(err))
    if rc:
        raise GrbError(rc, err.value.decode(errors="replace"))


def _polish_target_lists(batches):
    first = np.zeros(len(batches) + 1, dtype=np.uint64)
    flat = []
    for i, b in enumerate(batches):
        flat += [t.encode() if isinstance(t, str) else t for t in b]
        first[i + 1] = len(flat)
    return first, (C.c_char_p * max(1, len(flat)))(*flat)


class PolishInputs:
    """grb_polish_inputs: the two sequence indexes and the mappings goldpolish-targeted-bfs loads."""

    def __init__(self, target_index, mappings, mapped_seqs, mapped_index, mx_max_per_10kbp):
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = lib().grb_polish_inputs_open(os.fsencode(target_index), os.fsencode(mappings), os.fsencode(mapped_seqs),
                                          os.fsencode(mapped_index), float(mx_max_per_10kbp), C.byref(h), err,
                                          len(err))
        if rc:
            raise GrbError(rc, err.value.decode(errors="replace"))
        self._h = h

    def close(self):
        if self._h:
            lib().grb_polish_inputs_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def mappings(self, target_id):
        """AllMappings::get_mappings: ids of the reads kept for the target, in load order."""
        buf = C.create_string_buffer(1 << 20)
        n = lib().grb_polish_inputs_mappings(self._h, target_id.encode(), buf, len(buf))
        ids = buf.value.decode().split("\n")[:-1]
        assert len(ids) == n, (len(ids), n)
        return ids

    def serve_batches_host(self, params, subsample_max_per_10kbp, batches):
        """Test hook: (filters, reads inserted, bases inserted) with the job code run on the host."""
        first, ids = _polish_target_lists(batches)
        out = np.zeros(max(1, len(batches) * params.n_k * params.bf_bytes), dtype=np.uint8)
        nr, nb = C.c_uint64(), C.c_uint64()
        err = C.create_string_buffer(1024)
        rc = lib().grb_test_polish_serve_batches_host(self._h, C.byref(params), float(subsample_max_per_10kbp),
                                                      len(batches), _ptr(first), ids, _ptr(out), C.byref(nr),
                                                      C.byref(nb), err, len(err))
        if rc:
            raise GrbError(rc, err.value.decode(errors="replace"))
        return out.reshape(len(batches), params.n_k, params.bf_bytes), nr.value, nb.value


def host_pin(ptr, nbytes):
    """grb_host_pin: page-locks [ptr, ptr + nbytes) in 1 GiB pieces; returns the bytes now pinned."""
    got = C.c_size_t()
    lib().grb_host_pin(ptr, nbytes, C.byref(got))
    return got.value


def host_unpin(ptr, pinned_bytes):
    lib().grb_host_unpin(ptr, pinned_bytes)


def synth_fastq(sp, first=0, count=None) -> bytes:
    L = lib()
    if count is None:
        count = L.grb_synth_num_reads(C.byref(sp)) - first
    n = C.c_uint64()
    p = L.grb_synth_fastq(C.byref(sp), first, count, C.byref(n))
    if not p:
        raise MemoryError("grb_synth_fastq")
    try:
        return C.string_at(p, n.value)
    finally:
        L.grb_free_host(p)
