"""CPU: the C-ABI library loads and exports every symbol the header declares, the host-side scalar
helpers agree with the oracle (and with KATs taken from the reference), the device decision code
(compiled for the host through the grb_test_decide_host hook) agrees with the oracle's smoothing on
random tile vectors, and the engine fails loudly without a CUDA device."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

import goldrush_b200 as grb
import oracle_util as ou
import parity_util as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "goldrush_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(grb_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 35
    L = grb.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in goldrush_b200.h but not exported"
    assert declared == set(L._grb_symbols), "ctypes mirror out of sync with the header"


def test_ctypes_mirror_has_the_sizes_of_the_c_structs():
    """goldrush_b200/api.py mirrors include/goldrush_b200.h by hand: every struct that crosses the ABI
    must have the size (and the numpy dtypes the itemsize) the library was compiled with."""
    out = (C.c_uint64 * 9)()
    grb.lib().grb_abi_sizes(out)
    A = grb.api
    mirrors = [A.Params, A.ReadMeta, A.Decision, A.PathStats, A.ProbeBenchResult, A.RunOptions,
               A.RunResult, A.SynthParams, A.HostMsg]
    assert [int(x) for x in out] == [C.sizeof(m) for m in mirrors]
    assert A.READ_META_DTYPE.itemsize == C.sizeof(A.ReadMeta)
    assert A.DECISION_DTYPE.itemsize == C.sizeof(A.Decision)


def test_seed_pattern_kats():
    # strings printed by the reference's make_seed_pattern in this image (glibc rand)
    assert grb.make_seed_pattern("", 22, 16, 3) == [
        "1111011100110011101111", "11110111001010011101111", "111101110010010011101111"]
    assert grb.make_seed_pattern("1011011110110111101101", 22, 16, 3) == [
        "1011011110110111101101", "10110111101010111101101", "101101111010010111101101"]
    for k, w, h in [(20, 12, 2), (24, 18, 4), (22, 16, 1), (32, 20, 5), (21, 10, 2)]:
        assert grb.make_seed_pattern("", k, w, h) == ou.make_seed_pattern("", k, w, h)


def test_filter_sizing_kats():
    # m_filterSize printed by the reference: G=1e6 -> 28473728, cfg sizes of 8(d)
    assert grb.default_hash_universe(16, 1000000, 3) == 3000000
    assert grb.calc_optimal_size(3000000, 1, 0.1) == 28473728
    assert grb.calc_optimal_size(grb.default_hash_universe(16, 5000000, 3), 1, 0.1) == 142368384
    assert grb.calc_optimal_size(grb.default_hash_universe(16, 100000000, 3), 1, 0.1) == 2847366528
    assert grb.calc_optimal_size(grb.default_hash_universe(16, 3000000000, 3), 1, 0.1) == 61146729472
    L = ou.lib()
    rng = random.Random(5)
    for _ in range(200):
        g = rng.randrange(1, 4 * 10 ** 9)
        w = rng.choice([10, 12, 14, 16, 18])
        h = rng.randrange(1, 6)
        occ = rng.choice([0.05, 0.1, 0.15, 0.2, 0.5])
        hu = grb.default_hash_universe(w, g, h)
        assert hu == L.grbo_default_hash_universe(w, g, h)
        assert grb.calc_optimal_size(hu, 1, occ) == L.grbo_calc_optimal_size(hu, 1, occ)


def test_phred_finalize_matches_oracle():
    rng = np.random.default_rng(3)
    for n in [1, 2, 3, 10, 101, 5000]:
        for _ in range(20):
            q = bytes((rng.integers(2, 41, size=n) + 33).astype(np.uint8))
            avg, delta, first, total = ou.calc_phred_average(q)
            assert grb.phred_finalize(first, total, n) == (avg, delta)


def _random_votes(rng, n, ids_pool):
    best_id, best_count, cands = [], [], []
    for _ in range(n):
        k = rng.choice([0, 0, 1, 1, 2, 3])
        chosen = rng.sample(ids_pool, min(k, len(ids_pool)))
        cl = sorted(((i, rng.choice([3, 4, 8, 11, 12, 30, 200])) for i in chosen), key=lambda t: t[0])
        if cl and rng.random() < 0.85:
            top = max(c for _, c in cl)
            bid = min(i for i, c in cl if c == top)
            best_id.append(bid)
            best_count.append(top)
        else:
            cl = []
            best_id.append(rng.choice(ids_pool + [0, 0]))
            best_count.append(rng.choice([0, 1, 2]) if best_id[-1] else 0)
            if best_count[-1] == 0:
                best_id[-1] = 0
        cands.append(cl)
    return best_id, best_count, cands


@pytest.mark.parametrize("seed", range(6))
def test_device_decision_code_matches_oracle(seed):
    L = grb.lib()
    rng = random.Random(seed)
    for it in range(1500):
        n = rng.choice([0, 1, 2, 3, 4, 5, 8, 14, 15, 16, 20, 25, 40, 90])
        pool = rng.choice([[1, 2], [5, 6, 7], [1, 2, 3, 9, 10, 50], [0xFFFFFFFF, 1, 2],
                           [3, 4, 5, 6, 7, 8, 9, 10, 11]])
        thr = rng.choice([0, 2, 5, 10])
        best_id, best_count, cands = _random_votes(rng, n, pool)
        cap = 8
        bi = np.array(best_id + [0], dtype=np.uint32)
        bc = np.array(best_count + [0], dtype=np.uint32)
        nc = np.array([len(c) for c in cands] + [0], dtype=np.uint32)
        ci = np.zeros((n + 1) * cap, dtype=np.uint32)
        cc = np.zeros((n + 1) * cap, dtype=np.uint32)
        for t, cl in enumerate(cands):
            order = list(cl)
            rng.shuffle(order)  # the device list is unordered
            for j, (a, b) in enumerate(order):
                ci[t * cap + j], cc[t * cap + j] = a, b
        T, B = 100, rng.choice([1, 2, 3, 10])
        u, a = rng.choice([0, 2, 5]), rng.choice([0, 1, 3])
        read_len = n * T + rng.randrange(0, T)
        ids_ins = C.c_uint32(rng.randrange(0, 1000))
        ids0 = ids_ins.value
        out_ids = np.zeros(n + 1, dtype=np.uint32)
        out_as = np.zeros(n + 1, dtype=np.uint8)
        plan = np.zeros(9, dtype=np.uint32)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        L.grb_test_decide_host(n, p(bi), p(bc), p(nc), p(ci), p(cc), cap, thr, read_len, T, B, u, a,
                               C.byref(ids_ins), p(out_ids), p(out_as), p(plan))
        # oracle: threshold + smoothing, then the same decision restated in Python from
        # goldrush_path.cpp:960-1053
        o_ids, o_as, o_na = ou.smooth_tiles(best_id, [0] * n, cands, thr)
        assert list(out_ids[:n]) == list(o_ids), (seed, it)
        assert list(out_as[:n]) == list(o_as), (seed, it)
        assert plan[6] == o_na
        n_un = n - o_na
        exp_ids = ids0
        if n_un >= u and o_na <= a:
            verdict, ts, te = 2, 0, max(n - 1, 0)
            exp_ids = (exp_ids + 1 + read_len // (T * B)) & 0xFFFFFFFF
        elif o_na == n:
            verdict, ts, te = 4, 0, 0
        else:
            ls, le = ou.find_longest_stretch(list(o_as))
            good, ts, te = ou.eval_flanks(ls, le, list(o_ids))
            if good:
                verdict = 3
                exp_ids = (exp_ids + 1 + (te - ts) // B) & 0xFFFFFFFF
            else:
                verdict, ts, te = 4, 0, 0
        assert plan[0] == verdict, (seed, it)
        assert ids_ins.value == exp_ids
        if verdict in (2, 3):
            assert (plan[1], plan[2]) == (ts, te), (seed, it)
            assert plan[3] == (ids0 + 1) & 0xFFFFFFFF


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    seeds = grb.make_seed_pattern("1011011110110111101101", 22, 16, 3)
    with pytest.raises(grb.GrbError) as e:
        grb.Engine(seeds, genome_size=1000000, weight=16)
    assert e.value.code == -2
    with pytest.raises(grb.GrbError):
        grb.run_path(b"@r\nACGT\n+\nIIII\n", kmer_size=22, weight=16, genome_size=1000,
                     seed_preset="1011011110110111101101")


def test_odd_kmer_size_is_refused_before_any_device_work():
    """An odd -k yields seeds of span k - 1 (spaced_seeds.cpp:28,58-60); the reference aborts on
    MIBloomFilter.hpp:180 after pass 1 (tests/test_oracle_vs_ref.py), the engine says no at once."""
    seeds = grb.make_seed_pattern("10110111101101111011010", 23, 16, 3)
    assert [len(s) for s in seeds] == [22, 23, 24]
    with pytest.raises(grb.GrbError) as e:
        grb.Engine(seeds, genome_size=1000000, weight=16, kmer_size=23)
    assert e.value.code == -1 and "MIBloomFilter.hpp:180" in str(e.value)


def test_no_product_source_touches_the_oracle():
    """Nothing under goldrush_b200/ or include/ may include, link or execute oracle/ (DESIGN.md 1)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for base in ("goldrush_b200", "include"):
        for d, _, files in os.walk(os.path.join(root, base)):
            for fn in files:
                if fn.endswith((".so", ".pyc")):
                    continue
                with open(os.path.join(d, fn), errors="replace") as f:
                    txt = f.read()
                if re.search(r'#include\s+"[^"]*oracle|libgrb_oracle|oracle/_(build|ref)|grbo_|'
                             r"goldrush-path-(ref|oracle)|import\s+oracle", txt):
                    bad.append(os.path.join(d, fn))
    assert not bad, bad
    with open(os.path.join(root, "Makefile")) as f:
        lib_rule = f.read().split("$(LIB):")[1].split("\n\n")[0]
    assert "oracle" not in lib_rule


def test_meson_files_name_the_sources_the_makefile_builds():
    """meson is not installed in the build image, so the meson.build files cannot be run here; what
    can be checked is that they describe the same build: every file they name exists, and every
    kernel header / host source the Makefile compiles is named."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ("meson.build", "goldrush_b200/meson.build", "goldrush_path/meson.build"):
        assert os.path.exists(os.path.join(root, rel)), rel
    with open(os.path.join(root, "goldrush_b200", "meson.build")) as f:
        lib_meson = f.read()
    named = set(re.findall(r"'((?:csrc|host|\.\./include)/[^']+)'", lib_meson))
    for n in named:
        assert os.path.exists(os.path.join(root, "goldrush_b200", n)), n
    csrc = {"csrc/" + f for f in os.listdir(os.path.join(root, "goldrush_b200", "csrc"))}
    assert csrc <= named, csrc - named
    with open(os.path.join(root, "Makefile")) as f:
        mk = f.read()
    host_lib = set(re.findall(r"goldrush_b200/(host/\w+\.cpp)", mk.split("HOST_LIB_SRC :=")[1].split("\n")[0]))
    assert host_lib and host_lib <= named, host_lib - named
    assert "arch=compute_100a,code=sm_100a" in lib_meson and "arch=compute_100a,code=sm_100a" in mk
    with open(os.path.join(root, "goldrush_path", "meson.build")) as f:
        exe = f.read()
    assert "executable('goldrush-path'" in exe and "install : true" in exe
    for n in re.findall(r"'\.\./(goldrush_b200/host/[^']+)'", exe):
        assert os.path.exists(os.path.join(root, n)), n


def test_record_boundary_search_cuts_only_between_records():
    """Several GPUs: each rank ingests bytes [cut(r), cut(r + 1)) of the FASTQ.  A cut must fall on a
    record start even when quality lines begin with '@' or '+' (both are legal Phred characters)."""
    rnd = random.Random(5)
    recs = []
    for i in range(300):
        n = rnd.randint(1, 60)
        seq = "".join(rnd.choice("ACGT") for _ in range(n))
        first = rnd.choice("@+I5")  # qualities that look like a header or a separator
        qual = first + "".join(rnd.choice("@+#5IJ") for _ in range(n - 1))
        recs.append(f"@r{i} c\n{seq}\n+\n{qual}\n")
    data = "".join(recs).encode()
    starts, pos = set(), 0
    for r in recs:
        starts.add(pos)
        pos += len(r)
    L = grb.lib()
    for frm in range(0, len(data), 7):
        got = L.grb_test_next_record_start(data, len(data), frm)
        want = min((s for s in starts if s >= frm), default=len(data))
        assert got == want, (frm, got, want)
    # every world size: the shares tile the input and each starts on a record
    for world in (2, 3, 8):
        cuts = [L.grb_test_next_record_start(data, len(data), len(data) * r // world) for r in range(world)]
        cuts.append(len(data))
        assert cuts[0] == 0 and cuts == sorted(cuts) and all(c in starts or c == len(data) for c in cuts)


def test_silver_parts_reach_the_golden_ranks_in_order_and_in_balance():
    """grb_run_two_stage, slice mode: every (path, rank) part of the silver output goes whole to one
    golden-stage rank; receiving ranks never decrease along the joined stream (so each rank's share is
    one consecutive slice of it), every byte is delivered once, and with equal shares per path the
    ranks end up within one part of total / ranks."""
    L = grb.lib()
    rnd = random.Random(9)
    for n_paths, ranks in ((5, 8), (2, 2), (12, 4), (1, 8), (3, 3)):
        sizes = np.array([[rnd.randint(900, 1100) for _ in range(n_paths)] for _ in range(ranks)],
                         dtype=np.uint64)
        to = np.zeros(n_paths * ranks, dtype=np.int32)
        got = np.zeros(ranks, dtype=np.uint64)
        ok = L.grb_test_plan_silver_parts(sizes.ctypes.data, n_paths, ranks, to.ctypes.data, got.ctypes.data)
        assert ok == 1
        assert (np.diff(to) >= 0).all() and to[0] == 0 and to[-1] == ranks - 1
        assert int(got.sum()) == int(sizes.sum())
        # the sequence is path-major in shell-glob order (_1, _10, _11, _12, _2, ...), rank-minor
        order = sorted(range(n_paths), key=lambda q: f"{q + 1}.fq")
        seq = [int(sizes[r][q]) for q in order for r in range(ranks)]
        per = [sum(b for b, t in zip(seq, to) if t == g) for g in range(ranks)]
        assert per == [int(x) for x in got]
        assert max(per) - min(per) <= 2 * 1100
    # one rank holds everything and there is a single path: nothing to hand to the others
    sizes = np.array([[5000], [0], [0]], dtype=np.uint64)
    to = np.zeros(3, dtype=np.int32)
    got = np.zeros(3, dtype=np.uint64)
    assert L.grb_test_plan_silver_parts(sizes.ctypes.data, 1, 3, to.ctypes.data, got.ctypes.data) == 0


def test_synth_generator_is_deterministic_and_thread_independent():
    sp = grb.api.synth_params(50000, 3.0, 2000, 77)
    a = grb.synth_fastq(sp)
    os.environ["OMP_NUM_THREADS"] = "1"
    b = grb.synth_fastq(sp, 0, grb.synth_num_reads(sp))
    assert a == b
    n = grb.synth_num_reads(sp)
    assert a.count(b"\n") == 4 * n
    # generating a sub-range gives the same records
    part = grb.synth_fastq(sp, 5, 3)
    assert part in a


@pytest.mark.parametrize("k,w,h,preset", [(22, 16, 3, "1011011110110111101101"), (22, 16, 1, "1011011110110111101101"),
                                          (20, 12, 2, ""), (24, 18, 4, ""), (32, 20, 5, "")])
def test_grouped_hash_code_matches_oracle_on_host(k, w, h, preset):
    """The hash formulation the query / fill kernels use (csrc/nthash.cuh: grouped half-hash tables,
    grb_lo64, grb_group_half, grb_combine), compiled for the host, against the oracle's SeedNtHash
    restatement: every frame, every pattern, stale tail included."""
    import ctypes as C
    seeds = grb.make_seed_pattern(preset, k, w, h)
    assert seeds == ou.make_seed_pattern(preset, k, w, h)
    rng = np.random.default_rng(k * 100 + h)
    for n in (k + h - 1, k + h + 5, 97, 1000, 5003):
        seq = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)])
        want = ou.hash_sequence(seq, seeds)
        out = np.zeros((n - k + 1) * h, dtype=np.uint64)
        arr = (C.c_char_p * h)(*[s.encode() for s in seeds])
        rc = grb.lib().grb_test_group_hash_host(arr, h, seq, n, out.ctypes.data)
        assert rc == 0
        assert np.array_equal(out.reshape(-1, h), want), (n, k, h)


@pytest.mark.skipif(not os.path.exists(pu.REF), reason="oracle/_ref not built here")
def test_command_line_refusals_match_the_reference_binary():
    """process_options (opt.cpp:90-217) needs no GPU: --help and every refusal that is decided before
    the input is opened give the same exit code, the same stdout and the same stderr as the
    reference's own opt.cpp (getopt's own diagnostics name the executable, so those compare by exit
    code only)."""
    import subprocess
    seed = "1011011110110111101101"
    lines = [
        ["--help"],
        [],
        ["-i", "x.fq", "-w", "16", "-g", "1000"],                        # span 0
        ["-i", "x.fq", "-k", "22", "-g", "1000"],                        # weight 0
        ["-i", "x.fq", "-k", "22", "-w", "16"],                          # genome size 0
        ["-i", "x.fq", "-k", "20", "-w", "16", "-g", "1000", "-s", seed],  # preset longer than k
        ["-i", "x.fq", "-k", "22", "-w", "14", "-g", "1000", "-s", seed],  # preset weight != w
        ["-i", "x.fq", "-k", "22", "-w", "16", "-g", "0"],
    ]
    for args in lines:
        r = subprocess.run([pu.REF] + args, capture_output=True)
        g = subprocess.run([pu.PRODUCT] + args, capture_output=True)
        assert (g.returncode, g.stdout, g.stderr) == (r.returncode, r.stdout, r.stderr), args
    for args in (["-z"], ["--no_such_option"], ["-k"]):
        r = subprocess.run([pu.REF] + args, capture_output=True)
        g = subprocess.run([pu.PRODUCT] + args, capture_output=True)
        assert g.returncode == r.returncode == 1, args
