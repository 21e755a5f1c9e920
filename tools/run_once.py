#!/usr/bin/env python
"""One grb_run_path call on a bench workload (no output files), for ncu captures and debug dumps.
usage: tools/run_once.py [cfg2] [n_calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import goldrush_b200 as grb  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = bench.WORKLOADS[name]
sp = grb.api.synth_params(w["genome"], w["cov"], w["read_len"], w["seed"])
ptr, n = grb.synth_fastq_raw(sp)
for i in range(calls):
    t0 = time.time()
    res = grb.run_path(ptr, nbytes=n, input_path="(memory)", write_outputs=False, quiet=True,
                       seed_preset=bench.SEED22, genome_size=w["genome"], phred_min=w["phred_min"],
                       **bench.PARAMS)
    print(f"call {i}: {time.time() - t0:.3f} s, pass2 {res.ms_pass2:.1f} ms, digest {res.out_digest}, "
          f"selected {res.reads_selected}, launches {res.launches}", flush=True)
grb.free_host(ptr)
