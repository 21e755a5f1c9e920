#!/usr/bin/env python
"""(f4) throughput of the GoldPolish targeted-Bloom-filter builder: k-mer inserts per second of
grb_polish_fill_batches on the GPU (the reference's filter sizes: 10 MiB counting filter + 512 KiB
Bloom filter per batch and k, hash_num 4) against the reference's own fill_bfs
(oracle/_ref/libgoldpolish_ref.so) on one host core, on the same batches.  One JSON line.
usage: python tools/polish_bench.py [n_batches] [reads_per_batch] [read_len]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldrush_b200 as grb  # noqa: E402
import polish_util as pu  # noqa: E402

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rpb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rlen = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
ks, h, cbf, bf = [32, 28, 24], 4, 10 << 20, 512 << 10
rng = np.random.default_rng(5)
batches = []
for b in range(n_batches):  # each batch: reads over its own 20 kbp region at ~8x, threshold 6
    g = rng.integers(0, 4, 20000, dtype=np.uint8)
    reads = []
    for i in range(rpb):
        st = int(rng.integers(0, 20000 - rlen))
        r = g[st:st + rlen].copy()
        m = rng.random(rlen) < 0.02
        r[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
        reads.append((bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[r]), 6))
    batches.append(reads)
kmers = sum(max(0, len(s) - k + 1) for b in batches for s, _ in b for k in ks)
seeds = grb.make_seed_pattern("1011011110110111101101", 22, 16, 3)
params = grb.api.polish_params(ks, h, cbf, bf)
out = np.zeros(n_batches * len(ks) * bf, dtype=np.uint8)
pinned = grb.api.host_pin(out.ctypes.data, out.nbytes)  # the caller's result buffer, page-locked
with grb.Engine(seeds, genome_size=1000000, weight=16) as e:
    e.polish_fill_batches(params, batches[:8])  # warm-up
    t0 = time.time()
    got = e.polish_fill_batches(params, batches, out=out)
    t_gpu = time.time() - t0
    dev_ms = grb.lib().grb_last_device_ms(e._h)
grb.api.host_unpin(out.ctypes.data, pinned)
sample = batches[:4]
t0 = time.time()
want = pu.ref_fill(sample, ks, h, cbf, bf) if os.path.exists(pu.REF_SO) else pu.port_fill(sample, ks, h, cbf, bf)
t_cpu = time.time() - t0
k_sample = sum(max(0, len(s) - k + 1) for b in sample for s, _ in b for k in ks)
print(json.dumps({
    "what": "GoldPolish targeted Bloom filters (f4), k-mer inserts per second",
    "batches": n_batches, "reads_per_batch": rpb, "read_len": rlen, "k_values": ks, "hash_num": h,
    "cbf_bytes": cbf, "bf_bytes": bf, "kmer_inserts": kmers, "result_buffer_pinned_bytes": int(pinned),
    "mode": os.environ.get("GRB_POLISH", "warp"),
    "gpu_s_wall": round(t_gpu, 3), "gpu_device_ms": round(dev_ms, 1),
    "gpu_gkmers_per_s": round(kmers / (dev_ms * 1e-3) / 1e9, 3),
    "cpu_reference_one_core_mkmers_per_s": round(k_sample / t_cpu / 1e6, 2),
    "cpu_sample": f"{len(sample)} batches through the reference's own fill_bfs, one thread",
    "equal_on_sample": bool((got[:len(sample)] == want).all())}))
