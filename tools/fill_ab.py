#!/usr/bin/env python
"""A/B of the pass-1 fill kernels on the bench workload: direct atomicOr over the whole bit vector
(k_fill_bits) against the L2-partitioned fill (k_fill_part + k_fill_apply, 1024- and 512-thread
CTAs).  Prints one JSON object: device ms per variant (best of 3) and whether the three bit vectors
are identical.  usage: python tools/fill_ab.py [cfg2]"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import goldrush_b200 as grb  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    w = bench.WORKLOADS[wl]
    P = bench.PARAMS
    sp = grb.api.synth_params(w["genome"], w["cov"], w["read_len"], w["seed"])
    fq_ptr, fq_len = grb.synth_fastq_raw(sp)
    seeds = grb.make_seed_pattern(bench.SEED22, P["kmer_size"], P["weight"], P["hash_num"])
    eng = grb.Engine(seeds, device=0, genome_size=w["genome"],
                     **{k: v for k, v in P.items() if k not in ("kmer_size", "hash_num")})
    off = 0
    while off < fq_len:
        n = min(1 << 30, fq_len - off)
        used = eng.reads_ingest_fastq(fq_ptr + off, final=(off + n == fq_len), nbytes=n)
        if used == 0:
            break
        off += used
    flags, meta, phred_min = bench.host_flags(eng, grb, w["phred_min"], np)
    eng.reads_set_flags(flags)
    bits = grb.calc_optimal_size(grb.default_hash_universe(P["weight"], w["genome"], P["hash_num"]),
                                 1, P["occupancy"])
    bases = int(meta["len"][flags & 1 != 0].sum())
    out = {"workload": wl, "filter_bits": int(bits), "bases_pass1": bases,
           "probes": bases * P["hash_num"], "variants": {}}
    variants = {"direct": {"GRB_FILL": "direct"}}
    for ps in (23, 24, 25, 26, 27):
        variants[f"part_pshift{ps}"] = {"GRB_FILL": "part", "GRB_FILL_PSHIFT": str(ps)}
        variants[f"part_pshift{ps}_bs1024"] = {"GRB_FILL": "part", "GRB_FILL_PSHIFT": str(ps),
                                              "GRB_FILL_BS": "1024"}
    only = os.environ.get("FILL_AB_ONLY")
    if only:
        variants = {k: v for k, v in variants.items() if k in only.split(",")}
    digests = set()
    for name, env in variants.items():
        for k in ("GRB_FILL", "GRB_FILL_BS", "GRB_FILL_PSHIFT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ms = []
        for _ in range(3):
            eng.filter_alloc(bits)
            eng.build_bitvector()
            ms.append(eng.last_device_ms())
        dig = hashlib.md5(eng.copy_bitvector().tobytes()).hexdigest()
        digests.add(dig)
        best = min(ms)
        out["variants"][name] = {"ms": ms, "best_ms": best, "gprobes_per_s": out["probes"] / best / 1e6,
                                 "gbp_per_s": bases / best / 1e6, "md5": dig}
    out["identical"] = len(digests) == 1
    print(json.dumps(out))


if __name__ == "__main__":
    main()
