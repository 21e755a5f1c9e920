"""(f4) the input side of the GoldPolish targeted-Bloom-filter builder: sequence index, mappings
(ntLink / PAF / SAM), and serve_batch's per-target planning and sequence fetch, against the
reference's own SeqIndex / AllMappings / serve_batch (oracle/_ref/libgoldpolish_ref.so, compiled
unmodified) and against the fixtures it wrote (tests/golden/polish_inputs.json).  The CPU tests run
the host side with the job code compiled for the host; the GPU tests run the same call with the
kernels."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import goldrush_b200 as grb
import polish_inputs_util as piu
import polish_util as pu

with open(os.path.join(pu.ROOT, "tests", "golden", "polish_inputs.json")) as f:
    GOLDEN = json.load(f)
HAVE_REF = os.path.exists(pu.REF_SO)
NAMES = sorted(piu.SCENARIOS)


def _prepared(name, workdir):
    sc = piu.make_files(name, str(workdir))
    ti, ri = os.path.join(sc["dir"], "targets.idx"), os.path.join(sc["dir"], "reads.idx")
    grb.api.polish_index_build(sc["targets"], ti)
    grb.api.polish_index_build(sc["reads"], ri)
    return sc, ti, ri


def _open(sc, ti, ri):
    return grb.api.PolishInputs(ti, sc["mappings"], sc["reads"], ri, sc["mx_max"])


@pytest.mark.parametrize("name", NAMES)
def test_index_files_hold_the_lines_the_reference_writes(name, workdir):
    """goldpolish-index: same lines (the reference writes them in hash-table order, so as a set):
    FASTA and FASTQ, ids cut at the first blank / tab, a repeated id keeps its first record,
    Phred average over all but the last quality character, %g formatting."""
    sc, ti, ri = _prepared(name, workdir)
    assert list(piu.sorted_lines_md5(ti)) == GOLDEN[name]["target_index"]
    assert list(piu.sorted_lines_md5(ri)) == GOLDEN[name]["mapped_index"]
    if HAVE_REF:
        for seqs, mine in ((sc["targets"], ti), (sc["reads"], ri)):
            piu.ref_index(seqs, mine + ".ref")
            assert piu.sorted_lines_md5(mine + ".ref") == piu.sorted_lines_md5(mine)


@pytest.mark.parametrize("name", NAMES)
def test_kept_mappings_equal_the_references(name, workdir):
    """AllMappings after loading (+ the minimizer filter for ntLink input): per target the same read
    ids in the same order; targets outside the index and unknown ids give nothing."""
    sc, ti, ri = _prepared(name, workdir)
    with _open(sc, ti, ri) as pin:
        for t, want in GOLDEN[name]["mappings"].items():
            assert pin.mappings(t) == want, t
        assert pin.mappings("never_seen") == []
        if HAVE_REF:
            for t in sc["target_ids"]:
                assert pin.mappings(t) == piu.ref_mappings(sc, ti, t)


@pytest.mark.parametrize("name", NAMES)
def test_serve_batches_on_the_host_equals_the_references_serve_batch(name, workdir):
    sc, ti, ri = _prepared(name, workdir)
    params = grb.api.polish_params(sc["ks"], 4, sc["cbf"], sc["bf"])
    with _open(sc, ti, ri) as pin:
        got, n_reads, n_bases = pin.serve_batches_host(params, sc["subsample"], sc["batches"])
    g = GOLDEN[name]
    assert np.unpackbits(got, axis=2).sum(axis=2).tolist() == g["set_bits"]
    assert hashlib.md5(got.tobytes()).hexdigest() == g["bfs_md5"]
    assert n_reads > 0 and n_bases > n_reads
    if HAVE_REF:
        assert (got == piu.ref_serve(sc, ti, ri)).all()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built here")
@pytest.mark.parametrize("seed", range(8))
def test_minimizer_filter_fuzz_against_the_reference(seed, workdir):
    """AllMappings::filter (mappings.cpp:226-320) on hand-written index files: targets short and long,
    few and many mappings, minimizer counts clustered so that all three branches (everything fits /
    too many even at the top threshold / bisection) are taken."""
    rnd = random.Random(900 + seed)
    d = os.path.join(str(workdir), f"filter{seed}")
    os.makedirs(d, exist_ok=True)
    tidx, midx, mp = (os.path.join(d, n) for n in ("t.idx", "m.idx", "map.tsv"))
    seqs = os.path.join(d, "none.fq")
    open(seqs, "w").close()
    open(midx, "w").close()
    targets = [(f"c{i}", rnd.choice([400, 3000, 12000, 60000, 250000])) for i in range(12)]
    with open(tidx, "w") as f:
        for t, ln in targets:
            f.write(f"{t}\t{rnd.randint(0, 10**6)}\t{ln}\t0\n")
    with open(mp, "w") as f:
        rows = []
        for t, ln in targets:
            style = rnd.choice(["low", "high", "spread"])
            for j in range(rnd.choice([1, 4, 30, 120])):
                mx = {"low": rnd.randint(0, 6), "high": rnd.randint(28, 45), "spread": rnd.randint(0, 45)}[style]
                rows.append(f"q{j}_{t} {t} {mx}\n")
        rnd.shuffle(rows)
        f.writelines(rows)
    mx_max = rnd.choice([0.5, 3.0, 10.0, 40.0])
    sc = dict(targets=seqs, mappings=mp, mx_max=mx_max)
    kept = 0
    with grb.api.PolishInputs(tidx, mp, seqs, midx, mx_max) as pin:
        for t, _ in targets:
            mine = pin.mappings(t)
            assert mine == piu.ref_mappings(sc, tidx, t), (t, mx_max)
            kept += len(mine)
    assert kept > 0


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built here")
@pytest.mark.parametrize("seed", range(6))
def test_index_build_fuzz_against_the_reference(seed, workdir):
    """goldpolish-index on odd but legal files: ids followed by blanks, tabs or nothing, descriptions
    holding '@' and '>', quality lines that begin with '@' or '+', one-character reads (the Phred
    average is taken over all but the last character, over the one character when there is only one),
    repeated ids, a last line without newline; FASTQ and two-line FASTA."""
    rnd = random.Random(600 + seed)
    d = os.path.join(str(workdir), f"idx{seed}")
    os.makedirs(d, exist_ok=True)
    fastq = seed % 2 == 0
    path = os.path.join(d, "s.fq" if fastq else "s.fa")
    recs = []
    for i in range(rnd.randint(20, 60)):
        rid = rnd.choice([f"s{i}", f"s{i}", f"s{rnd.randint(0, i)}", f"x|{i}.1", f"{i}"])
        sep = rnd.choice(["", " ", "  two blanks", "\tafter a tab", " desc @with >signs", "\t", " \tmix"])
        ln = rnd.choice([1, 2, 3, 50, 400])
        s = "".join(rnd.choice("ACGTNacgt") for _ in range(ln))
        if fastq:
            q = "".join(chr(rnd.randint(33, 74)) for _ in range(ln))
            if rnd.random() < 0.3:
                q = rnd.choice("@+") + q[1:]
            recs.append(f"@{rid}{sep}\n{s}\n+{rnd.choice(['', rid])}\n{q}\n")
        else:
            recs.append(f">{rid}{sep}\n{s}\n")
    text = "".join(recs)
    if rnd.random() < 0.5:
        text = text[:-1]
    with open(path, "w") as f:
        f.write(text)
    mine, ref = os.path.join(d, "mine.idx"), os.path.join(d, "ref.idx")
    grb.api.polish_index_build(path, mine)
    piu.ref_index(path, ref)
    with open(mine, "rb") as f, open(ref, "rb") as g:
        assert sorted(f.read().splitlines()) == sorted(g.read().splitlines())


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built here")
@pytest.mark.parametrize("fmt,seed", [("paf", 0), ("paf", 1), ("sam", 0), ("sam", 1)])
def test_alignment_loader_fuzz_against_the_reference(fmt, seed, workdir):
    """load_paf / load_sam (mappings.cpp:109-224) on ragged files: header lines, blank lines, lines too
    short to hold the target column (they inherit the target of the line before), repeated pairs,
    targets outside the index, blanks and tabs mixed as separators."""
    rnd = random.Random(700 + seed + (10 if fmt == "sam" else 0))
    d = os.path.join(str(workdir), f"aln_{fmt}{seed}")
    os.makedirs(d, exist_ok=True)
    tidx, midx, seqs = (os.path.join(d, n) for n in ("t.idx", "m.idx", "none.fq"))
    open(seqs, "w").close()
    open(midx, "w").close()
    targets = [f"c{i}" for i in range(8)]
    with open(tidx, "w") as f:
        for t in targets:
            f.write(f"{t}\t0\t{rnd.randint(500, 90000)}\t0\n")
    mp = os.path.join(d, "aln." + fmt)
    with open(mp, "w") as f:
        for n in range(300):
            q, t = f"q{rnd.randint(0, 60)}", rnd.choice(targets + ["other", "c1"])
            x = rnd.random()
            sep = rnd.choice(["\t", " ", "\t\t"])
            if x < 0.05:
                f.write("@PG\tID:x\n")
            elif x < 0.1:
                f.write("\n")
            elif x < 0.2:
                f.write(sep.join([q, "16"]) + "\n")
            elif fmt == "paf":
                f.write(sep.join([q, "900", "0", "900", "+", t, "5000", "1", "901", "880", "900", "60"]) + "\n")
            else:
                f.write(sep.join([q, "0", t, "7", "60", "900M", "*", "0", "0", "*", "*"]) + "\n")
    sc = dict(targets=seqs, mappings=mp, mx_max=8.0)
    kept = 0
    with grb.api.PolishInputs(tidx, mp, seqs, midx, 8.0) as pin:
        for t in targets + ["other"]:
            mine = pin.mappings(t)
            assert mine == piu.ref_mappings(sc, tidx, t), t
            kept += len(mine)
    assert kept > 50


def test_input_errors_are_reported_not_fatal(workdir):
    """Where the reference dies (uncaught std::out_of_range from .at(), exit(1) from check_error), the
    library returns an error with a message."""
    sc, ti, ri = _prepared("sam", workdir)
    params = grb.api.polish_params(sc["ks"], 4, sc["cbf"], sc["bf"])
    with _open(sc, ti, ri) as pin:
        with pytest.raises(grb.GrbError, match="target id not in the target index"):
            pin.serve_batches_host(params, sc["subsample"], [["t0", "no_such_target"]])
    with pytest.raises(grb.GrbError, match="BAM"):
        grb.api.PolishInputs(ti, os.path.join(sc["dir"], "x.bam"), sc["reads"], ri, 8.0)
    with pytest.raises(grb.GrbError, match="cannot read"):
        grb.api.PolishInputs(ti, os.path.join(sc["dir"], "missing.paf"), sc["reads"], ri, 8.0)
    tsv = os.path.join(sc["dir"], "one.tsv")
    with open(tsv, "w") as f:
        f.write("r0 t0 5\n")
    with pytest.raises(grb.GrbError, match="not positive"):
        grb.api.PolishInputs(ti, tsv, sc["reads"], ri, 0.0)
    with open(tsv, "w") as f:
        f.write("r0 t0 five\n")
    with pytest.raises(grb.GrbError, match="not a minimizer count"):
        grb.api.PolishInputs(ti, tsv, sc["reads"], ri, 8.0)
    # a mapped read that the mapped-sequence index does not know
    short = os.path.join(sc["dir"], "short.idx")
    with open(ri) as f, open(short, "w") as g:
        g.writelines(f.readlines()[:3])
    with grb.api.PolishInputs(ti, sc["mappings"], sc["reads"], short, sc["mx_max"]) as pin:
        with pytest.raises(grb.GrbError, match="mapped read not in the mapped-sequence index"):
            pin.serve_batches_host(params, sc["subsample"], sc["batches"])
    with pytest.raises(grb.GrbError, match="cannot read"):
        grb.api.polish_index_build(os.path.join(sc["dir"], "absent.fa"), os.path.join(sc["dir"], "a.idx"))


def test_goldpolish_index_executable_is_a_drop_in(workdir):
    """build/goldpolish-index: the reference tool's two arguments and its refusal of anything else
    (goldpolish_index.cpp:6-9); the index it writes holds the reference's lines."""
    import subprocess
    exe = os.path.join(pu.ROOT, "build", "goldpolish-index")
    sc = piu.make_files("paf", str(workdir))
    out = os.path.join(sc["dir"], "cli.idx")
    subprocess.check_call([exe, sc["reads"], out])
    assert list(piu.sorted_lines_md5(out)) == GOLDEN["paf"]["mapped_index"]
    p = subprocess.run([exe, sc["reads"]], capture_output=True)
    assert p.returncode == 1 and p.stderr == b"Wrong args.\n"
    p = subprocess.run([exe, os.path.join(sc["dir"], "absent.fq"), out], capture_output=True)
    assert p.returncode == 1 and b"cannot read" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_serve_batches_on_the_device_equals_the_references_serve_batch(name, workdir):
    """grb_polish_serve_batches: index + mappings + planning + sequence fetch on the host, filters by
    the kernels; every batch's Bloom filters byte for byte what the reference's serve_batch saved."""
    sc, ti, ri = _prepared(name, workdir)
    params = grb.api.polish_params(sc["ks"], 4, sc["cbf"], sc["bf"])
    seeds = grb.make_seed_pattern("1011011110110111101101", 22, 16, 3)
    with _open(sc, ti, ri) as pin, grb.Engine(seeds, genome_size=1000000, weight=16) as e:
        got = e.polish_serve_batches(pin, params, sc["subsample"], sc["batches"])
        with pytest.raises(grb.GrbError, match="target id not in the target index"):
            e.polish_serve_batches(pin, params, sc["subsample"], [["nope"]])
    assert hashlib.md5(got.tobytes()).hexdigest() == GOLDEN[name]["bfs_md5"]
    if HAVE_REF:
        assert (got == piu.ref_serve(sc, ti, ri)).all()
