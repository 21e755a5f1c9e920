"""Deterministic input files for the input side of the GoldPolish targeted-Bloom-filter builder
(target FASTA, mapped-read FASTQ, mappings in ntLink / PAF / SAM form, batches of target ids) and the
ctypes binding of its checker: the reference's own SeqIndex, AllMappings and serve_batch
(oracle/_ref/libgoldpolish_ref.so, compiled unmodified; regular files where the reference has named
pipes).  Test infrastructure only."""
import ctypes as C
import hashlib
import os
import random
import threading

import numpy as np

import polish_util as pu

SCENARIOS = {
    # name: (seed, mappings format, mx_max per 10 kbp, subsample max per 10 kbp, k values, cbf, bf)
    "ntlink_filtered": (41, "ntlink", 8.0, 5.0, [32, 28, 24], 1 << 16, 1 << 14),
    "ntlink_loose_reference_sizes": (42, "ntlink", 500.0, 40.0, [32, 24], 10 << 20, 512 << 10),
    "paf": (43, "paf", 8.0, 6.0, [40, 20], 1 << 16, 1 << 13),
    "sam": (44, "sam", 8.0, 3.5, [25], 1 << 15, 1 << 13),
}


def _mutate(rnd, s):
    r = list(s)
    for j in range(len(r)):
        x = rnd.random()
        if x < 0.01:
            r[j] = rnd.choice("ACGT")
        elif x < 0.0105:
            r[j] = "N"
    return "".join(r)


def make_files(name, workdir):
    """Writes targets.fa, reads.fq, mappings.<ext> under workdir; returns a dict of paths + batches."""
    seed, fmt, mx_max, subsample, ks, cbf, bf = SCENARIOS[name]
    rnd = random.Random(seed)
    d = os.path.join(workdir, name)
    os.makedirs(d, exist_ok=True)
    targets = []
    lengths = [1500, 6000, 9000, 14000, 21000, 25000]
    rnd.shuffle(lengths)
    for i, ln in enumerate(lengths):
        targets.append((f"t{i}", "".join(rnd.choice("ACGT") for _ in range(ln))))
    with open(os.path.join(d, "targets.fa"), "w") as f:
        for tid, s in targets:
            f.write(f">{tid} len={len(s)}\n{s}\n")
        f.write(f">t0 a second record under a used id\n{'ACGT' * 10}\n")
    reads, lines = [], []
    for ti, (tid, s) in enumerate(targets):
        n = rnd.choice([3, 12, 25, 40])
        for _ in range(n):
            ln = rnd.randint(300, min(4000, len(s)))
            st = rnd.randint(0, len(s) - ln)
            rid = f"r{len(reads)}"
            base_q = rnd.randint(5, 35)
            qual = "".join(chr(33 + max(2, min(40, base_q + rnd.randint(-2, 2)))) for _ in range(ln))
            reads.append((rid, _mutate(rnd, s[st:st + ln]), qual))
            lines.append((rid, tid, rnd.choice([0, 1, 2, 5, 9, 14, 22, 29, 30, 31, 40]), ln, len(s), st))
            if rnd.random() < 0.15:  # the same read on a second target
                other = targets[(ti + 1) % len(targets)]
                lines.append((rid, other[0], rnd.randint(1, 35), ln, len(other[1]), 0))
            if rnd.random() < 0.1:  # the same pair again, another minimizer count (first one counts)
                lines.append((rid, tid, rnd.randint(1, 35), ln, len(s), st))
    for _ in range(5):  # reads that map nowhere, and mappings to a target outside the index
        rid = f"r{len(reads)}"
        ln = rnd.randint(300, 900)
        reads.append((rid, "".join(rnd.choice("ACGT") for _ in range(ln)), "I" * ln))
        lines.append((rid, "t_unknown", 20, ln, 1000, 0))
    rnd.shuffle(lines)
    with open(os.path.join(d, "reads.fq"), "w") as f:
        for j, (rid, s, q) in enumerate(reads):
            sep = ["", " strand=+", "\tcomment with a tab"][j % 3]
            f.write(f"@{rid}{sep}\n{s}\n+\n{q}\n")
        rid, s, q = reads[0]
        f.write(f"@{rid} again\n{s[:50]}\n+\n{q[:50]}\n")
    ext = {"ntlink": "tsv", "paf": "paf", "sam": "sam"}[fmt]
    mp = os.path.join(d, "mappings." + ext)
    with open(mp, "w") as f:
        if fmt == "sam":
            f.write("@HD\tVN:1.6\tSO:unsorted\n")
            for tid, s in targets:
                f.write(f"@SQ\tSN:{tid}\tLN:{len(s)}\n")
        for n, (rid, tid, mx, ln, tl, st) in enumerate(lines):
            if fmt == "ntlink":
                f.write(f"{rid}\t{tid}\t{mx}\n")
            elif fmt == "paf":
                f.write(f"{rid}\t{ln}\t0\t{ln}\t+\t{tid}\t{tl}\t{st}\t{st + ln}\t{ln - 9}\t{ln}\t60\ttp:A:P\n")
                if n == 7:
                    f.write("@a line taken for a header\n\nr2\t100\t0\n")  # short line: keeps the target before it
            else:
                f.write(f"{rid}\t0\t{tid}\t{st + 1}\t60\t{ln}M\t*\t0\t0\t*\t*\n")
                if n == 5:
                    f.write("r3\t4\n")
    batches = [["t0", "t1"], ["t2"], ["t3", "t4", "t5"], []]
    return dict(dir=d, targets=os.path.join(d, "targets.fa"), reads=os.path.join(d, "reads.fq"), mappings=mp,
                batches=batches, mx_max=mx_max, subsample=subsample, ks=ks, cbf=cbf, bf=bf,
                target_ids=[t for t, _ in targets])


def sorted_lines_md5(path):
    with open(path, "rb") as f:
        lines = sorted(f.read().splitlines())
    return hashlib.md5(b"\n".join(lines)).hexdigest(), len(lines)


def ref_lib():
    lib = C.CDLL(pu.REF_SO)
    lib.grbp_ref_index_build.argtypes = [C.c_char_p, C.c_char_p]
    lib.grbp_ref_serve_batches.argtypes = [C.c_char_p] * 5 + [C.c_double, C.c_double, C.c_uint, C.c_void_p,
                                                              C.c_uint, C.c_size_t, C.c_size_t, C.c_char_p, C.c_uint]
    return lib


def ref_index(seqs_path, index_path):
    assert ref_lib().grbp_ref_index_build(seqs_path.encode(), index_path.encode()) == 0


def ref_serve(sc, target_index, mapped_index, hash_num=4):
    """The reference's serve_batch over the scenario's batches: Bloom filters [batch][k][bf_bytes]."""
    work = os.path.join(sc["dir"], "ref_work")
    os.makedirs(work, exist_ok=True)
    for b, ids in enumerate(sc["batches"]):
        with open(os.path.join(work, f"b{b}-target_ids_input"), "w") as f:
            f.write("".join(t + "\n" for t in ids) + "x\n")
    ks = np.ascontiguousarray(sc["ks"], dtype=np.uint32)
    # SeqIndex::get_seq keeps its file descriptor in a thread_local static (seqindex.hpp:66-81): a thread
    # can read from one sequence file only, so every call gets a thread of its own
    rc = []
    th = threading.Thread(target=lambda: rc.append(ref_lib().grbp_ref_serve_batches(
        sc["targets"].encode(), target_index.encode(), sc["mappings"].encode(), sc["reads"].encode(),
        mapped_index.encode(), sc["mx_max"], sc["subsample"], hash_num, ks.ctypes.data, len(ks), sc["cbf"],
        sc["bf"], work.encode(), len(sc["batches"]))))
    th.start()
    th.join()
    assert rc == [0]
    out = np.zeros((len(sc["batches"]), len(ks), sc["bf"]), dtype=np.uint8)
    for b in range(len(sc["batches"])):
        for i, k in enumerate(sc["ks"]):
            out[b, i] = np.fromfile(os.path.join(work, f"b{b}-k{k}.bf"), dtype=np.uint8)
    return out


def ref_mappings(sc, target_index, target_id):
    lib = C.CDLL(pu.REF_SO)
    lib.grbp_ref_mappings.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_char_p, C.c_char_p,
                                      C.c_size_t]
    lib.grbp_ref_mappings.restype = C.c_long
    buf = C.create_string_buffer(1 << 20)
    n = lib.grbp_ref_mappings(sc["targets"].encode(), target_index.encode(), sc["mappings"].encode(), sc["mx_max"],
                              target_id.encode(), buf, len(buf))
    ids = buf.value.decode().split("\n")[:-1] if n else []
    assert len(ids) == n
    return ids
