"""Helpers shared by the parity tests and tests/golden/make_golden.py."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases as golden_cases  # noqa: E402

SYNTH = os.path.join(ROOT, "build", "grb-synth")
REF = os.path.join(ROOT, "oracle", "_ref", "goldrush-path-ref")
ORACLE = os.path.join(ROOT, "oracle", "_build", "goldrush-path-oracle")
PRODUCT = os.path.join(ROOT, "build", "goldrush-path")
GOLDEN_JSON = os.path.join(ROOT, "tests", "golden", "cases.json")


def ensure_built():
    """Build the host tools / oracle if missing (cheap; no CUDA)."""
    if not (os.path.exists(SYNTH) and os.path.exists(ORACLE)):
        subprocess.check_call(["make", "-s", "-C", ROOT, "host-tools", "oracle"])


def case_by_name(name):
    for c in golden_cases.CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


def make_input(case, workdir, produce_outputs=None):
    """Materialise the FASTQ (and filter list) of a case under workdir; returns (path, extra_args).

    `produce_outputs(case, workdir)` is called for cases that consume another case's outputs
    (golden run on concatenated silver paths, bin/goldrush:249-251)."""
    path = os.path.join(workdir, case["name"] + ".fq")
    extra = []
    if "from_case" in case:
        src = case_by_name(case["from_case"])
        outs = produce_outputs(src, workdir)
        # `cat $(p1)_*.fq` (bin/goldrush:250-251): shell glob order = lexicographic (C collation)
        outs = sorted(outs, key=lambda o: os.path.basename(o).encode())
        with open(path, "wb") as f:
            for o in outs:
                with open(o, "rb") as g:
                    f.write(g.read())
        return path, extra
    subprocess.check_call([SYNTH] + case["synth"] + ["-o", path], stderr=subprocess.DEVNULL)
    if case.get("post"):
        with open(path, "rb") as f:
            data = f.read()
        data = golden_cases.POST[case["post"]](data)
        with open(path, "wb") as f:
            f.write(data)
    if case.get("filter_every"):
        names = []
        with open(path, "rb") as f:
            for i, line in enumerate(f):
                if i % 4 == 0 and (i // 4) % case["filter_every"] == 0:
                    names.append(line[1:].split()[0])
        fl = os.path.join(workdir, case["name"] + ".filter.txt")
        with open(fl, "wb") as f:
            f.write(b"\n".join(names) + b"\n")
        extra = ["-f", fl]
    return path, extra


def run_cli(binary, case, inp, extra, workdir, tag, jobs=4, env=None):
    """Runs a goldrush-path-compatible binary; returns (exit code, sorted output paths, stderr)."""
    prefix = os.path.join(workdir, f"{case['name']}.{tag}")
    for f in os.listdir(workdir):
        if f.startswith(os.path.basename(prefix) + "_") or f == os.path.basename(prefix) + ".fa":
            os.remove(os.path.join(workdir, f))
    cmd = [binary] + case["args"] + extra + ["-j", str(jobs), "-i", inp, "-p", prefix]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    outs = []
    base = os.path.basename(prefix)
    for f in os.listdir(workdir):
        m = re.fullmatch(re.escape(base) + r"_(\d+)\.fq", f)
        if m:
            outs.append((int(m.group(1)), os.path.join(workdir, f)))
        elif f == base + ".fa":
            outs.append((0, os.path.join(workdir, f)))
    outs.sort()
    return p.returncode, [o[1] for o in outs], p.stderr.decode(errors="replace")


STAT_RE = re.compile(r"^(Visited|Saw:|Assigned:|Unassigned:|Total queries:|Total hits:|"
                     r"Total misses:|Num reads:|Average Phred:|m_filterSize:|num_passed_reads:|"
                     r"Minimum phred score calculated with median:|Total expected entries for seed patterns:)\s*(\d+)", re.M)


def parse_stats(stderr):
    """The counters the reference prints with --verbose (goldrush_path.cpp:126-154,308-325)."""
    return [[m.group(1), int(m.group(2))] for m in STAT_RE.finditer(stderr)]


def digest_outputs(paths):
    out = []
    for p in paths:
        with open(p, "rb") as f:
            b = f.read()
        name = os.path.basename(p)
        suffix = re.search(r"(_\d+\.fq|\.fa)$", name).group(1)
        out.append({"suffix": suffix, "md5": golden_cases.md5(b), "bytes": len(b)})
    return out


def load_golden():
    with open(GOLDEN_JSON) as f:
        return json.load(f)
