#!/usr/bin/env python
"""Full-size reference outputs for the BASELINE.json configs (tests/golden/full_size.json).

For each config: the synthetic read set of SURVEY.md 8(d) (build/grb-synth), then both launches of
one assembly exactly as bin/goldrush:240-260 issues them -- the --silver_path run, `cat` of its
<p>_N.fq files in shell-glob order, the golden run on that file -- with
  * oracle/_ref/goldrush-path-ref   (the reference's own sources, compiled unmodified), and
  * oracle/_build/goldrush-path-oracle (the port),
and the two must agree byte for byte.  Committed per stage: md5 of every output file, the
order-sensitive record digest (oracle/_build/grb-digest = grb_run_result.out_digest), the
--verbose counters, and the reference's own phase timers (seconds on the authoring container's 8
cores).  cfg3 (1 Gbp) and cfg4 (3 Gbp) cannot be produced in the authoring container: at 1 Gbp the
reference's m_data + m_counts alone are 8 bytes x up to 6.4e9 set bits = 51 GB next to a 3.3 GB
bit vector, its rank structure and the 60 GB input, on a box with 62 GB of RAM (the attempt was
killed by the kernel's OOM handler; the port, which also holds the input in memory, even earlier).
Parity at those sizes rests on size-independent properties instead (tests/test_gpu_parity.py,
bench.py: batch engine == one-read-at-a-time engine, any batch size, any number of GPUs).

Run in the authoring container only (needs /root/reference for oracle/_ref):
    python tests/golden/make_full_size.py cfg1 cfg2 [--work DIR] [--keep]
cfg1: ~1.5 min, cfg2: ~45 min (reference 25 + port 15).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu  # noqa: E402

DIGEST = os.path.join(ROOT, "oracle", "_build", "grb-digest")
OUT = os.path.join(ROOT, "tests", "golden", "full_size.json")
S = pu.golden_cases.SEED22
COMMON = ["-k", "22", "-w", "16", "-s", S, "-h", "3", "-t", "1000", "-b", "10", "-u", "5", "-a", "1",
          "-o", "0.1", "-x", "10", "-d", "5", "-r", "0.9", "--verbose"]
# genome, coverage, read length (0 = log-normal, N50 20 kbp), seed, -P, -g
CONFIGS = {
    "cfg1": dict(genome=5_000_000, cov=25, read_len=20000, seed=1001, phred_min=0, g="5e6"),
    "cfg2": dict(genome=100_000_000, cov=30, read_len=25000, seed=1002, phred_min=20, g="1e8"),
}


def run_stage(binary, args, inp, prefix, jobs):
    t0 = time.time()
    p = subprocess.run([binary] + COMMON + args + ["-j", str(jobs), "-i", inp, "-p", prefix],
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        raise SystemExit(f"{binary} failed ({p.returncode}): {p.stderr[-2000:]}")
    return p.stderr, wall


def outputs_of(prefix):
    d, base = os.path.dirname(prefix), os.path.basename(prefix)
    outs = []
    for f in os.listdir(d):
        m = re.fullmatch(re.escape(base) + r"_(\d+)\.fq", f)
        if m:
            outs.append((int(m.group(1)), os.path.join(d, f)))
        elif f == base + ".fa":
            outs.append((0, os.path.join(d, f)))
    return [o[1] for o in sorted(outs)]


def describe(outs, err, wall):
    dg = subprocess.check_output([DIGEST] + outs, text=True).split()
    files = []
    for o in outs:
        md5 = subprocess.check_output(["md5sum", o], text=True).split()[0]
        files.append({"suffix": re.search(r"(_\d+\.fq|\.fa)$", o).group(1), "md5": md5,
                      "bytes": os.path.getsize(o)})
    return {"out_digest": int(dg[0]), "records": int(dg[1]), "bytes": int(dg[2]), "files": files,
            "stats": pu.parse_stats(err),
            "phase_s": [float(x) for x in re.findall(r"^in ([0-9.]+)\s*$", err, re.M)],
            "wall_s": round(wall, 1)}


def one_assembly(binary, tag, cfg, fq, work, jobs):
    c = CONFIGS[cfg]
    P = ["-P", str(c["phred_min"]), "-g", c["g"]]
    sp = os.path.join(work, f"{cfg}.{tag}.silver")
    err, wall = run_stage(binary, ["-M", "5", "-m", "20000", "--silver_path"] + P, fq, sp, jobs)
    s_outs = outputs_of(sp)
    silver = describe(s_outs, err, wall)
    cat = sp + "_all.fastq"
    with open(cat, "wb") as f:  # `cat $(p1)_*.fq` (bin/goldrush:250-251): shell-glob order
        for o in sorted(s_outs, key=lambda o: os.path.basename(o).encode()):
            with open(o, "rb") as g:
                while True:
                    b = g.read(1 << 26)
                    if not b:
                        break
                    f.write(b)
    gp = os.path.join(work, f"{cfg}.{tag}.golden")
    err, wall = run_stage(binary, ["-m", "0"] + P, cat, gp, jobs)
    golden = describe(outputs_of(gp), err, wall)
    return {"silver": silver, "golden": golden}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=sorted(CONFIGS))
    ap.add_argument("--work", default="/tmp/grb_full_size")
    ap.add_argument("--port-only", action="store_true")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.work, exist_ok=True)
    res = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            res = json.load(f)
    for cfg in a.configs:
        c = CONFIGS[cfg]
        fq = os.path.join(a.work, cfg + ".fq")
        if not os.path.exists(fq):
            subprocess.check_call([pu.SYNTH, "-G", str(c["genome"]), "-c", str(c["cov"]), "-l",
                                   str(c["read_len"]), "-s", str(c["seed"]), "-o", fq])
        port = one_assembly(pu.ORACLE, "port", cfg, fq, a.work, a.jobs)
        entry = {"synth": {k: c[k] for k in ("genome", "cov", "read_len", "seed")},
                 "args_common": COMMON, "phred_min": c["phred_min"], "g": c["g"],
                 "input_bytes": os.path.getsize(fq), "cores": a.jobs}
        if a.port_only:
            entry.update(by="port (oracle/grb_oracle.cpp; byte-identical to the reference's own sources "
                            "on cfg1 and cfg2 in full, see those entries)", **port)
        else:
            ref = one_assembly(pu.REF, "ref", cfg, fq, a.work, a.jobs)
            for st in ("silver", "golden"):
                for key in ("out_digest", "records", "bytes", "files", "stats"):
                    assert ref[st][key] == port[st][key], (cfg, st, key, ref[st][key], port[st][key])
            entry.update(by="reference sources (oracle/_ref), port checked byte-identical", **ref)
            entry["port_phase_s"] = {st: port[st]["phase_s"] for st in ("silver", "golden")}
            entry["port_wall_s"] = {st: port[st]["wall_s"] for st in ("silver", "golden")}
        res[cfg] = entry
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
        print(cfg, "ok", file=sys.stderr)
        if not a.keep:
            for f in os.listdir(a.work):
                if f.startswith(cfg + "."):
                    os.remove(os.path.join(a.work, f))


if __name__ == "__main__":
    main()
