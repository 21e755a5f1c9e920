"""Shared by tests/test_polish.py and tests/golden/make_polish_golden.py: deterministic inputs for the
GoldPolish targeted-Bloom-filter builder and the ctypes bindings of its two CPU checkers
(oracle/_ref/libgoldpolish_ref.so = the reference's own fill_bfs; oracle/_build/libgrb_oracle.so =
the port).  Test infrastructure only."""
import ctypes as C
import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgoldpolish_ref.so")
PORT_SO = os.path.join(ROOT, "oracle", "_build", "libgrb_oracle.so")
FILL_ARGS = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p, C.c_uint, C.c_size_t,
             C.c_size_t, C.c_void_p]


def make_batches(seed, n_batches=3, reads_per_batch=25, genome_len=30000, max_len=5000):
    """Batches of reads sampled from one genome (so that k-mers recur and cross the thresholds),
    with substitutions, a few N and lower-case characters, short reads (shorter than k) and an
    empty one; each read carries its target's k-mer threshold (5..8)."""
    rnd = random.Random(seed)
    genome = "".join(rnd.choice("ACGT") for _ in range(genome_len))
    batches = []
    for b in range(n_batches):
        reads = []
        thr = rnd.choice([5, 6, 7, 8])
        for i in range(reads_per_batch):
            if i % 9 == 4:
                thr = rnd.choice([5, 6, 7, 8])  # next target of the batch
            ln = rnd.choice([0, 7, 19, 33]) if i % 11 == 5 else rnd.randint(40, max_len)
            st = rnd.randint(0, genome_len - ln) if ln < genome_len else 0
            r = list(genome[st:st + ln])
            for j in range(len(r)):
                x = rnd.random()
                if x < 0.02:
                    r[j] = rnd.choice("ACGT")
                elif x < 0.0215:
                    r[j] = "N"
                elif x < 0.03:
                    r[j] = r[j].lower()
            reads.append(("".join(r).encode(), thr))
        batches.append(reads)
    return batches


def flat(batch):
    seqs = b"".join(s for s, _ in batch)
    off = np.zeros(len(batch) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s, _ in batch])
    thr = np.array([t for _, t in batch] + [0], dtype=np.uint32)
    return seqs, off, thr


def cpu_fill(so_path, fn_name, batches, k_values, hash_num, cbf_bytes, bf_bytes):
    lib = C.CDLL(so_path)
    fn = getattr(lib, fn_name)
    fn.argtypes = FILL_ARGS
    ks = np.ascontiguousarray(k_values, dtype=np.uint32)
    out = np.zeros((len(batches), len(ks), bf_bytes), dtype=np.uint8)
    for b, batch in enumerate(batches):
        seqs, off, thr = flat(batch)
        o = np.zeros(len(ks) * bf_bytes, dtype=np.uint8)
        rc = fn(seqs, off.ctypes.data, thr.ctypes.data, len(batch), hash_num, ks.ctypes.data, len(ks),
                cbf_bytes, bf_bytes, o.ctypes.data)
        assert rc == 0
        out[b] = o.reshape(len(ks), bf_bytes)
    return out


def ref_fill(*a):
    return cpu_fill(REF_SO, "grbp_ref_fill", *a)


def port_fill(*a):
    return cpu_fill(PORT_SO, "grbo_polish_fill", *a)


CASES = {
    # name: (seed, batches, reads per batch, k values, hash_num, cbf_bytes, bf_bytes)
    "small_filters_collide": (11, 3, 25, [32, 28, 24], 4, 1 << 15, 1 << 14),
    "reference_sizes": (12, 2, 120, [40, 32, 24], 4, 10 << 20, 512 << 10),
    "two_hashes_one_k": (13, 4, 12, [20], 2, 1 << 16, 1 << 12),
    # homopolymers and short tandem repeats: identical k-mers side by side, i.e. groups of 32 k-mers
    # that share counters and must be applied in order; k up to the table limit of 64
    "low_complexity": (14, 3, 20, [64, 33, 17], 4, 1 << 16, 1 << 13),
}


def add_repeats(batches, seed):
    rnd = random.Random(seed)
    out = []
    for batch in batches:
        nb = []
        for s, t in batch:
            s = bytearray(s)
            for _ in range(3):
                if len(s) > 400:
                    at = rnd.randint(0, len(s) - 300)
                    unit = bytes(rnd.choice(b"ACGT") for _ in range(rnd.choice([1, 1, 2, 3, 7])))
                    n = rnd.randint(60, 250)
                    s[at:at + n] = (unit * n)[:n]
            nb.append((bytes(s), t))
        out.append(nb)
    return out


def case_batches(name):
    seed, nb, rpb, ks, h, cbf, bf = CASES[name]
    batches = make_batches(seed, nb, rpb)
    if name == "low_complexity":
        batches = add_repeats(batches, seed)
    return batches, ks, h, cbf, bf
