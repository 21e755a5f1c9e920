"""Seeded parameter fuzz over the goldrush-path option space (k, weight, h, tile, block, smoothing
and Phred options, silver / golden mode, fixed and log-normal read lengths).  The same case list is
run three ways: reference sources == oracle port on CPU (test_oracle_vs_ref.py) and drop-in
executable == oracle on the GPU (test_gpu_parity.py).  Nothing here is a committed expectation:
both sides are computed at test time from the seeded input."""
import numpy as np

import parity_util as pu

N_CASES = 16


def _case(i):
    rng = np.random.default_rng(7000 + i)
    k = int(rng.choice([16, 18, 20, 22, 24, 26, 28, 32]))
    half = int(rng.integers(max(3, k // 4), k // 2))  # ones per half, below k/2 so the design loop ends fast
    h = int(rng.integers(1, 5))
    t = int(rng.choice([200, 300, 500, 1000]))
    G = int(rng.integers(80, 250)) * 1000
    cov = int(rng.integers(8, 15))
    lognormal = bool(rng.random() < 0.35)
    L = 0 if lognormal else int(rng.integers(12, 40)) * t // 2
    n50 = int(rng.integers(10, 30)) * t // 2
    silver = bool(rng.random() < 0.7)
    m = int((n50 if lognormal else L) * rng.choice([0.5, 0.8, 1.0])) if silver else 0
    args = ["-k", str(k), "-w", str(2 * half), "-h", str(h), "-t", str(t),
            "-b", str(int(rng.integers(2, 11))), "-u", str(int(rng.integers(3, 7))),
            "-a", str(int(rng.integers(1, 3))), "-o", str(rng.choice([0.05, 0.1, 0.2])),
            "-x", str(int(rng.integers(5, 13))), "-d", str(int(rng.integers(3, 9))),
            "-P", str(int(rng.choice([0, 12, 15, 18]))), "-g", str(G), "-m", str(m), "--verbose"]
    if silver:
        args += ["-r", str(rng.choice([0.7, 0.8, 0.9])), "-M", str(int(rng.choice([1, 2, 3, 5]))),
                 "--silver_path"]
    return dict(name=f"fuzz{i:02d}", synth=pu.golden_cases.synth_args(G, cov, L, 7100 + i, n50=n50),
                args=args)


CASES = [_case(i) for i in range(N_CASES)]
# -P 0 with more than 50 000 long-enough reads: the median is taken over the first 50 000 (the
# reference's counter stops one past the cap per thread, goldrush_path.cpp:91-103; the reference arm
# of the CPU test runs with -j 2); the late reads are all Q40 so that a median over everything would differ
CASES.append(dict(name="median_over_cap", synth=pu.golden_cases.synth_args(150000, 80, 220, 77, q="8,38"),
                  post="late_reads_q40",
                  args=["-k", "16", "-w", "10", "-h", "2", "-t", "100", "-b", "2", "-u", "1", "-a", "1", "-o",
                        "0.1", "-x", "5", "-d", "9", "-P", "0", "-g", "150000", "-r", "0.9", "-M", "2", "-m",
                        "200", "--silver_path", "--verbose"]))


# ---- degenerate inputs (golden mode, -m 0: nothing is filtered by length) ----
EDGE_ARGS = ["-k", "22", "-w", "16", "-h", "3", "-t", "500", "-b", "3", "-u", "2", "-a", "1", "-o", "0.1", "-x",
             "8", "-d", "5", "-P", "15", "-g", "30000", "-m", "0", "--verbose"]


def edge_inputs():
    """name -> FASTQ text.  An empty file and a file holding one newline are "not FASTQ" (exit 1,
    goldrush_path.cpp:247-250); a read shorter than the seed span ends the run in SeedNtHash (exit 1,
    after the output file was opened); reads of less than one tile and of exactly one tile are
    processed."""
    import random
    rnd = random.Random(3)

    def seq(n):
        return "".join(rnd.choice("ACGT") for _ in range(n))

    def rec(name, s):
        return f"@{name}\n{s}\n+\n{'I' * len(s)}\n"

    g = seq(30000)
    base = "".join(rec(f"r{i}", g[st:st + 3000]) for i, st in enumerate(range(0, 27000, 700)))
    last = rec("r_last", g[100:3100])
    return {
        "empty_file": "",
        "one_newline": "\n",
        "read_shorter_than_k": base + rec("tiny", seq(10)) + last,
        "read_shorter_than_a_tile": base + rec("sub", seq(450)) + last,
        "read_of_exactly_one_tile": base + rec("one", g[5000:5500]) + last,
        "no_final_newline": (base + last)[:-1],
    }


EDGE_EXIT = {"empty_file": 1, "one_newline": 1, "read_shorter_than_k": 1}
