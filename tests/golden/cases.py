"""Parity cases shared by tests/golden/make_golden.py (which runs the REFERENCE on them, in the
authoring container) and by the tests (which run the oracle and the CUDA engine on them).

Every input is produced by the deterministic generator in goldrush_b200/host/synth.cpp
(`build/grb-synth`), optionally post-processed by a pure function below, so only digests need to
be committed.  `args` is the goldrush-path command line without -i / -p / -j.
"""
import hashlib

SEED22 = "1011011110110111101101"  # bin/goldrush:73


def synth_args(G, cov, length, seed, err=0.01, n50=20000, q="12,30", max_reads=0):
    a = ["-G", str(G), "-c", str(cov), "-l", str(length), "-s", str(seed), "-e", str(err),
         "-n", str(n50), "-q", q]
    if max_reads:
        a += ["-N", str(max_reads)]
    return a


def mutate_n_and_case(data: bytes) -> bytes:
    """Every 7th record gets an N, every 5th is lower-cased (exercises the ACGTacgt filter,
    goldrush_path.cpp:293-301, and SeqReader case folding)."""
    lines = data.split(b"\n")
    out = []
    rec = 0
    for i in range(0, len(lines) - 3, 4):
        h, s, p, q = lines[i:i + 4]
        if rec % 7 == 3:
            s = s[:len(s) // 2] + b"N" + s[len(s) // 2 + 1:]
        if rec % 5 == 1:
            s = s.lower()
        out += [h, s, p, q]
        rec += 1
    return b"\n".join(out) + b"\n"


def ragged_tail(data: bytes) -> bytes:
    """Drops the final newline and appends two short records (shorter than a tile, and shorter
    than the seed span) to exercise the length filter and ragged input."""
    data = data.rstrip(b"\n")
    data += b"\n@tiny1 x\nACGTACGTAC\n+\nIIIIIIIIII\n@tiny2\nACG\n+\nIII"
    return data


CASES = [
    # name, synth params, post-process, args
    dict(name="silver_default",
         synth=synth_args(1000000, 12, 20000, 11),
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "1000", "-u", "5", "-a", "1",
               "-o", "0.1", "-x", "10", "-b", "10", "-d", "5", "-P", "0", "-g", "1e6", "-r", "0.9",
               "-M", "2", "-m", "20000", "--silver_path", "--verbose"]),
    dict(name="golden_default", from_case="silver_default",
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "1000", "-u", "5", "-a", "1",
               "-o", "0.1", "-x", "10", "-b", "10", "-d", "5", "-P", "0", "-g", "1e6", "-m", "0",
               "--verbose"]),
    dict(name="small_tiles_random_seed",
         synth=synth_args(100000, 30, 4000, 12, err=0.005),
         args=["-k", "20", "-w", "12", "-h", "2", "-t", "200", "-u", "3", "-a", "1", "-o", "0.1",
               "-x", "5", "-b", "3", "-d", "3", "-P", "15", "-g", "1e5", "-r", "0.8", "-M", "4",
               "-m", "4000", "--silver_path", "--verbose"]),
    dict(name="lognormal_h1",
         synth=synth_args(150000, 20, 0, 13, n50=8000),
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "1", "-t", "500", "-u", "5", "-a", "1",
               "-o", "0.1", "-x", "10", "-b", "10", "-d", "5", "-P", "0", "-g", "150000", "-r",
               "0.5", "-M", "2", "-m", "5000", "--silver_path", "--verbose"]),
    dict(name="golden_lognormal_all_lengths",
         synth=synth_args(120000, 15, 0, 14, n50=6000, err=0.003),
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "400", "-u", "4", "-a", "2",
               "-o", "0.1", "-x", "8", "-b", "4", "-d", "5", "-P", "12", "-g", "120000", "-m",
               "1000", "--verbose"]),
    dict(name="n_and_lowercase", synth=synth_args(100000, 25, 5000, 15), post="mutate_n_and_case",
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "250", "-u", "5", "-a", "1",
               "-o", "0.1", "-x", "10", "-b", "5", "-d", "5", "-P", "0", "-g", "1e5", "-r", "0.9",
               "-M", "2", "-m", "5000", "--silver_path", "--verbose"]),
    dict(name="ragged_tail_h4", synth=synth_args(80000, 20, 3000, 16), post="ragged_tail",
         args=["-k", "24", "-w", "18", "-h", "4", "-t", "300", "-u", "2", "-a", "1", "-o", "0.2",
               "-x", "6", "-b", "2", "-d", "4", "-P", "14", "-g", "8e4", "-r", "0.7", "-M", "3",
               "-m", "3000", "--silver_path", "--verbose"]),
    dict(name="filter_list", synth=synth_args(100000, 25, 5000, 17), filter_every=3,
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "250", "-u", "5", "-a", "1",
               "-o", "0.1", "-x", "10", "-b", "10", "-d", "5", "-P", "16", "-g", "1e5", "-r",
               "0.9", "-M", "2", "-m", "5000", "--silver_path", "--verbose"]),
    dict(name="hash_universe_override_block1",
         synth=synth_args(60000, 30, 2500, 18, err=0.002),
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "100", "-u", "5", "-a", "1",
               "-o", "0.15", "-x", "4", "-b", "1", "-d", "5", "-P", "13", "-g", "6e4", "-H",
               "400000", "-r", "0.9", "-M", "5", "-m", "2500", "--silver_path", "--verbose"]),
    # more than 9 silver paths: `cat $(p1)_*.fq` (bin/goldrush:250-251) joins them in the shell's
    # lexicographic order (_1, _10, _11, _12, _2, ...), which is the golden run's read order
    dict(name="silver_m12", synth=synth_args(60000, 40, 3000, 21, err=0.004),
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "250", "-u", "4", "-a", "1",
               "-o", "0.1", "-x", "8", "-b", "3", "-d", "5", "-P", "0", "-g", "6e4", "-r", "0.5",
               "-M", "12", "-m", "3000", "--silver_path", "--verbose"]),
    dict(name="golden_m12", from_case="silver_m12",
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "250", "-u", "4", "-a", "1",
               "-o", "0.1", "-x", "8", "-b", "3", "-d", "5", "-P", "0", "-g", "6e4", "-m", "0",
               "--verbose"]),
    dict(name="ntcard_sizing", synth=synth_args(100000, 25, 5000, 15), post="mutate_n_and_case",
         args=["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "250", "-u", "5", "-a", "1",
               "-o", "0.1", "-x", "10", "-b", "5", "-d", "5", "-P", "0", "-g", "1e5", "-r", "0.9",
               "-M", "2", "-m", "5000", "--silver_path", "--verbose", "--ntcard"]),
]

# odd -k: make_seed_pattern builds seeds of span k - 1 (two halves of k / 2 positions,
# spaced_seeds.cpp:28,58-60) and the reference aborts on MIBloomFilter.hpp:180
# (assert(m_sseeds[0].size() == kmerSize); meson's default build keeps asserts) once pass 1 is
# done: exit by SIGABRT, no record written.  Not a fixture: checked by test_oracle_vs_ref.py
# (reference and oracle abort alike) and test_host_logic.py (the engine refuses the option).
ODD_K_CASE = dict(name="odd_k_aborts", synth=synth_args(60000, 10, 4000, 19),
                  args=["-k", "23", "-w", "16", "-s", SEED22 + "0", "-h", "3", "-t", "250", "-u", "5",
                        "-a", "1", "-o", "0.1", "-x", "10", "-b", "5", "-d", "5", "-P", "0", "-g",
                        "6e4", "-r", "0.9", "-M", "3", "-m", "4000", "--silver_path", "--verbose"])

def late_reads_q40(data: bytes) -> bytes:
    """Every read after the 50 000th gets quality 'I' throughout: the -P 0 median must not see them
    (MEDIAN_SAMPLES_NEEDED, goldrush_path.cpp:38,93-96), the filters of pass 1 and 2 must."""
    lines = data.split(b"\n")
    for i in range(50000 * 4 + 3, len(lines) - 1, 4):
        lines[i] = b"I" * len(lines[i])
    return b"\n".join(lines)


POST = {"mutate_n_and_case": mutate_n_and_case, "ragged_tail": ragged_tail, "late_reads_q40": late_reads_q40}


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()
