"""(f4) GoldPolish targeted Bloom filters (subprojects/goldpolish/src/
goldpolish_targeted_bfs.cpp serve_batch + utils.cpp fill_bfs).

CPU: the port against the reference's own fill_bfs (oracle/_ref/libgoldpolish_ref.so, when built
here) and against the committed fixtures; the product's host-side planning against the oracle; the
product's job code (csrc/polish_core.h, one source for host and device) run on the host against the
port.  GPU: the kernel against the port.  The btllib semantics behind all of them are recalled, not
read (oracle/shim_polish/btllib/*.hpp): parity is unpinned at that boundary."""
import ctypes as C
import hashlib
import json
import os
import random

import numpy as np
import pytest

import goldrush_b200 as grb
import polish_util as pu

with open(os.path.join(pu.ROOT, "tests", "golden", "polish.json")) as f:
    GOLDEN = json.load(f)


@pytest.mark.parametrize("name", sorted(pu.CASES))
def test_port_reproduces_the_reference_fixture(name):
    batches, ks, h, cbf, bf = pu.case_batches(name)
    p = pu.port_fill(batches, ks, h, cbf, bf)
    assert hashlib.md5(p.tobytes()).hexdigest() == GOLDEN[name]["md5"]
    assert np.unpackbits(p, axis=2).sum(axis=2).tolist() == GOLDEN[name]["set_bits"]
    assert p.any()


@pytest.mark.skipif(not os.path.exists(pu.REF_SO), reason="oracle/_ref/libgoldpolish_ref.so not built here")
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_port_equals_reference_fill_bfs(seed):
    rnd = random.Random(seed)
    ks = sorted(rnd.sample(range(16, 70), 3), reverse=True)
    batches = pu.make_batches(seed, n_batches=2, reads_per_batch=20, genome_len=12000, max_len=3000)
    args = (batches, ks, rnd.choice([1, 3, 4]), 1 << rnd.randint(12, 16), 1 << rnd.randint(9, 13))
    assert (pu.ref_fill(*args) == pu.port_fill(*args)).all()


def test_kmer_threshold_and_target_plan_match_the_oracle():
    """serve_batch before any hashing (goldpolish_targeted_bfs.cpp:43-51,88-127): the threshold
    formula, the subsampling cap, and the (Phred as size_t desc, id asc) order with its ties."""
    L = grb.lib()
    O = C.CDLL(pu.PORT_SO)
    O.grbo_polish_kmer_threshold.argtypes = [C.c_uint64]
    for bases in [0, 1, 1_000_000, 1_563_000, 1_570_000, 20_000_000, 39_400_000, 39_500_000, 10**9, 10**12]:
        assert L.grb_polish_kmer_threshold(bases) == O.grbo_polish_kmer_threshold(bases)
    assert L.grb_polish_kmer_threshold(0) == 5 and L.grb_polish_kmer_threshold(10**12) == 13
    O.grbo_polish_plan_target.restype = C.c_uint32
    O.grbo_polish_plan_target.argtypes = [C.c_uint64, C.c_double, C.c_uint32, C.POINTER(C.c_char_p),
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    rnd = random.Random(4)
    for it in range(60):
        n = rnd.randint(0, 40)
        ids = [f"read{rnd.randint(0, 60)}.{i}" if rnd.random() < 0.8 else f"r{i}" for i in range(n)]
        phred = [rnd.choice([7.2, 7.9, 12.0, 12.99, 13.0, 20.5]) + rnd.random() * 0.01 for _ in range(n)]
        lens = [rnd.randint(500, 60000) for _ in range(n)]
        tlen = rnd.choice([1000, 9999, 10000, 25000, 400000])
        sub = rnd.choice([0.5, 1.0, 3.5, 40.0])
        order, used, thr = grb.api.polish_plan_target(tlen, sub, ids, phred, lens)
        arr = (C.c_char_p * max(1, n))(*[i.encode() for i in ids])
        ph, ln = np.array(phred + [0.0]), np.array(lens + [0], dtype=np.uint64)
        o_order = np.zeros(max(1, n), dtype=np.uint32)
        o_thr = C.c_int32()
        o_used = O.grbo_polish_plan_target(tlen, sub, n, arr, ph.ctypes.data, ln.ctypes.data,
                                           o_order.ctypes.data, C.byref(o_thr))
        assert used == o_used == min(n, int(tlen * sub / 10000.0))
        assert thr == o_thr.value
        assert order.tolist() == o_order[:n].tolist()


@pytest.mark.parametrize("name", sorted(pu.CASES))
def test_job_code_on_the_host_matches_the_port(name):
    """csrc/polish_core.h compiled for the host (rolling hash, restart after N, conservative
    update, per-k threshold) against the port's from-scratch restatement, bit for bit."""
    batches, ks, h, cbf, bf = pu.case_batches(name)
    got = grb.api.polish_fill_host(grb.api.polish_params(ks, h, cbf, bf), batches)
    assert hashlib.md5(got.tobytes()).hexdigest() == GOLDEN[name]["md5"]


@pytest.mark.parametrize("name", sorted(pu.CASES))
def test_warp_algorithm_on_the_host_matches_the_port(name):
    """The warp kernel's algorithm emulated lane by lane (2-bit packed segments, table-driven hashes,
    32 k-mers applied at once -- all counts read before any write -- unless two lanes share a
    counter): the same filters as the sequential port, i.e. conflict-free groups do commute."""
    batches, ks, h, cbf, bf = pu.case_batches(name)
    got, groups, ordered = grb.api.polish_fill_host_grouped(grb.api.polish_params(ks, h, cbf, bf), batches)
    assert hashlib.md5(got.tobytes()).hexdigest() == GOLDEN[name]["md5"]
    assert groups > 0
    if name == "low_complexity":
        assert ordered > 50  # repeats put identical k-mers into one group
    if name == "reference_sizes":
        assert ordered < groups // 50  # chance collisions are rare at 10 MiB


def test_threshold_below_four_is_refused():
    p = grb.api.polish_params([24], 4, 1 << 12, 1 << 10)
    with pytest.raises(grb.GrbError):
        grb.api.polish_fill_host(p, [[(b"ACGT" * 20, 3)]])  # utils.cpp:105-107


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(pu.CASES))
def test_device_bloom_filters_match_the_reference_fixture(name):
    batches, ks, h, cbf, bf = pu.case_batches(name)
    seeds = grb.make_seed_pattern("1011011110110111101101", 22, 16, 3)
    with grb.Engine(seeds, genome_size=1000000, weight=16) as e:
        got = e.polish_fill_batches(grb.api.polish_params(ks, h, cbf, bf), batches)
        assert e.launch_count() > 0
    assert hashlib.md5(got.tobytes()).hexdigest() == GOLDEN[name]["md5"]


@pytest.mark.gpu
def test_device_runs_many_batches_in_waves():
    """2 200 (batch, k) jobs with small filters: the call is cut into three waves (1024, 1024, 152
    jobs), the Bloom filters of one travelling back under the kernel of the next through two
    alternating device buffers; exercises the wave loop, the buffer reuse and the job indexing."""
    batches = pu.make_batches(31, n_batches=1100, reads_per_batch=4, genome_len=4000, max_len=600)
    ks, h, cbf, bf = [31, 25], 3, 1 << 13, 1 << 10
    want = pu.port_fill(batches, ks, h, cbf, bf)
    seeds = grb.make_seed_pattern("1011011110110111101101", 22, 16, 3)
    with grb.Engine(seeds, genome_size=1000000, weight=16) as e:
        got = e.polish_fill_batches(grb.api.polish_params(ks, h, cbf, bf), batches)
    assert (got == want).all() and want.any()
