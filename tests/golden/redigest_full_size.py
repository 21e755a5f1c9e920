#!/usr/bin/env python
"""Recomputes the `out_digest` fields of tests/golden/full_size.json after a change of the digest
function (oracle/grb_digest.cpp = grb_run_result.out_digest): re-runs the PORT on each config,
checks that every output file still has the md5 the reference's own sources produced (the md5s in
full_size.json are not touched), and stores the new digests of those very bytes.
    python tests/golden/redigest_full_size.py cfg1 cfg2 [--work DIR]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_full_size as m  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--work", default="/tmp/grb_full_size")
    a = ap.parse_args()
    os.makedirs(a.work, exist_ok=True)
    with open(m.OUT) as f:
        res = json.load(f)
    for cfg in a.configs:
        c = m.CONFIGS[cfg]
        fq = os.path.join(a.work, cfg + ".fq")
        if not os.path.exists(fq):
            m.subprocess.check_call([m.pu.SYNTH, "-G", str(c["genome"]), "-c", str(c["cov"]), "-l",
                                     str(c["read_len"]), "-s", str(c["seed"]), "-o", fq])
        port = m.one_assembly(m.pu.ORACLE, "port", cfg, fq, a.work, os.cpu_count() or 1)
        for st in ("silver", "golden"):
            assert port[st]["files"] == res[cfg][st]["files"], (cfg, st, "md5s differ from the reference's")
            assert port[st]["records"] == res[cfg][st]["records"]
            res[cfg][st]["out_digest"] = port[st]["out_digest"]
        with open(m.OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
        print(cfg, "ok", res[cfg]["silver"]["out_digest"], res[cfg]["golden"]["out_digest"], file=sys.stderr)
        for fn in os.listdir(a.work):
            if fn.startswith(cfg + "."):
                os.remove(os.path.join(a.work, fn))


if __name__ == "__main__":
    main()
