"""GPU parity tests: every kernel family of the engine against the CPU oracle on the same seeded
inputs, through the C ABI, bit-exact (all of this path is integer / byte / index work; the two
Phred sums are IEEE doubles added in the reference's order and must match to the last bit)."""
import os

import numpy as np
import pytest

import fuzz_cases
import goldrush_b200 as grb
import oracle_util as ou
import parity_util as pu

pytestmark = pytest.mark.gpu

SEED22 = "1011011110110111101101"
GOLDEN = pu.load_golden()


def _rand_seq(rng, n):
    return bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)])


def _fastq(records):
    return b"".join(b"@" + n + b"\n" + s + b"\n+\n" + q + b"\n" for n, s, q in records)


def _records(rng, n_reads, lo, hi, name=b"r"):
    out = []
    for i in range(n_reads):
        L = int(rng.integers(lo, hi))
        q = bytes((rng.integers(2, 41, size=L) + 33).astype(np.uint8))
        out.append((name + str(i).encode() + b" c=" + str(i).encode(), _rand_seq(rng, L), q))
    return out


@pytest.mark.parametrize("k,w,h,preset", [(22, 16, 3, SEED22), (22, 16, 1, SEED22),
                                          (20, 12, 2, ""), (24, 18, 4, ""), (32, 20, 5, "")])
def test_hash_sequence_matches_oracle(k, w, h, preset):
    seeds = grb.make_seed_pattern(preset, k, w, h)
    rng = np.random.default_rng(k * 100 + h)
    with grb.Engine(seeds, genome_size=1000000, weight=w) as e:
        for n in [len(seeds[-1]), len(seeds[-1]) + 1, 64, 65, 1000, 1021, 4097]:
            s = _rand_seq(rng, n)
            got = e.hash_sequence(s)
            exp = ou.hash_sequence(s, seeds)
            assert got.shape == exp.shape
            assert np.array_equal(got, exp), (n, k, h)
        low = _rand_seq(rng, 500)
        assert np.array_equal(e.hash_sequence(low.lower()), ou.hash_sequence(low, seeds))


def test_phred_sums_bit_exact():
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    rng = np.random.default_rng(9)
    with grb.Engine(seeds, genome_size=1000000, weight=16) as e:
        for n in [1, 2, 3, 7, 100, 101, 25000]:
            q = bytes((rng.integers(0, 60, size=n) + 33).astype(np.uint8))
            first, total = e.phred_sums(q)
            avg, delta, f0, t0 = ou.calc_phred_average(q)
            assert (first, total) == (f0, t0), n
            assert grb.phred_finalize(first, total, n) == (avg, delta)


def test_ingest_decodes_ragged_fastq():
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    rng = np.random.default_rng(21)
    recs = _records(rng, 300, 30, 3000)
    # N, lower case, CRLF, a last record without newline
    recs[3] = (recs[3][0], recs[3][1][:10] + b"N" + recs[3][1][11:], recs[3][2])
    recs[5] = (recs[5][0], recs[5][1].lower(), recs[5][2])
    recs[7] = (recs[7][0], recs[7][1][:5] + b"n" + recs[7][1][6:], recs[7][2])
    data = _fastq(recs)
    data = data.replace(b"@r9 c=9\n", b"@r9 c=9\r\n", 1)[:-1]
    with grb.Engine(seeds, genome_size=1000000, weight=16) as e:
        # feed in three chunks that split records at arbitrary bytes
        cuts = [0, len(data) // 3 + 7, 2 * len(data) // 3 + 1, len(data)]
        pos = 0
        for i in range(3):
            chunk = data[pos:cuts[i + 1]]
            used = e.reads_ingest_fastq(chunk, final=(i == 2))
            pos += used
        assert pos == len(data)
        assert e.reads_count() == len(recs)
        meta = e.reads_get_meta()
        for m, (name, seq, qual) in zip(meta, recs):
            assert m.len == len(seq) and m.qual_len == len(qual)
            assert data[m.seq_off:m.seq_off + m.len] == seq
            assert data[m.qual_off:m.qual_off + m.qual_len] == qual
            assert data[m.hdr_off:m.hdr_off + m.hdr_len] == name
            assert m.non_acgt == (1 if (b"N" in seq or b"n" in seq) else 0)
            avg, delta, f0, t0 = ou.calc_phred_average(qual)
            assert (m.phred_first_half_sum, m.phred_total_sum) == (f0, t0)


@pytest.mark.parametrize("input_bytes", [0, 60_000_000_000])
def test_cardinality_estimate_matches_oracle_for_both_sample_widths(input_bytes, workdir):
    """K5 (ntcard.hpp:81-154,180-188,248-274): the estimate per seed pattern and in total, with the
    sample width the input size selects -- sBits = 7 below 50 GB (the file's own size here) and
    sBits = 11 from 50 GB up, which no test file reaches, so the size is passed in on both sides."""
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    sp = grb.api.synth_params(400000, 12.0, 5000, 57)
    data = grb.synth_fastq(sp)
    fq = os.path.join(workdir, "ntcard_widths.fq")
    with open(fq, "wb") as f:
        f.write(data)
    L = ou.lib()
    per = (ou.C.c_uint64 * 3)()
    arr = (ou.C.c_char_p * 3)(*[s.encode() for s in seeds])
    want_total = L.grbo_ntcard_sized(fq.encode(), arr, 3, input_bytes, per)
    with grb.Engine(seeds, genome_size=400000, weight=16) as e:
        e.reads_ingest_fastq(data)
        got_per, got_total = e.estimate_cardinality(input_bytes or len(data))
    assert got_per == list(per) and got_total == want_total and want_total > 0


def test_ingest_readahead_equals_plain_ingest():
    """grb_reads_readahead: chunks copied ahead on the second stream (and re-aligned on the device)
    decode to exactly the same read store as chunks copied inside the call."""
    import ctypes
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    sp = grb.api.synth_params(150000, 20.0, 4000, 91)
    data = grb.synth_fastq(sp)
    buf = ctypes.create_string_buffer(data, len(data))
    base = ctypes.addressof(buf)
    chunk = 200003  # odd size: the device copy of a read-ahead chunk is never 16-byte aligned
    out = []
    for ahead in (False, True):
        with grb.Engine(seeds, genome_size=150000, weight=16) as e:
            if ahead:
                e.reads_readahead(base, len(data))
            off = 0
            while off < len(data):
                n = min(chunk, len(data) - off)
                used = e.reads_ingest_fastq(base + off, final=(off + n == len(data)), nbytes=n)
                assert used > 0
                off += used
            if ahead:
                e.reads_readahead(None, 0)
            meta = e.reads_meta_array()
            n_reads = e.reads_count()
            e.reads_set_flags(np.full(n_reads, 3, dtype=np.uint8))
            e.filter_alloc(2000000 + 64)
            e.build_bitvector()
            out.append((n_reads, meta.tobytes(), e.copy_bitvector().tobytes()))
    assert out[0][0] > 500 and out[0] == out[1]


def _build_pair(rng, n_reads, lo, hi, seeds, bits, **params):
    """Engine + oracle filter holding the same reads / bit vector."""
    recs = _records(rng, n_reads, lo, hi)
    e = grb.Engine(seeds, genome_size=1000000, weight=16, **params)
    e.reads_ingest_fastq(_fastq(recs))
    e.reads_set_flags(np.full(n_reads, 3, dtype=np.uint8))
    e.filter_alloc(bits)
    e.build_bitvector()
    f = ou.Filter(bits, len(seeds))
    for _, s, _ in recs:
        f.insert_bv(ou.hash_sequence(s, seeds))
    return e, f, recs


@pytest.mark.parametrize("h,bits", [(3, 1000000 + 64), (1, 333376), (4, 5000000 + 128)])
def test_bitvector_and_rank_match_oracle(h, bits):
    seeds = grb.make_seed_pattern(SEED22, 22, 16, h)
    rng = np.random.default_rng(100 + h)
    e, f, recs = _build_pair(rng, 40, 30, 6000, seeds, bits)
    with e:
        assert np.array_equal(e.copy_bitvector(), f.words())
        pop = e.finalize_bitvector()
        assert pop == f.setup()
        pos = np.concatenate([rng.integers(0, bits, size=5000, dtype=np.uint64),
                              np.array([0, 1, 63, 64, 191, 192, 193, bits - 1], dtype=np.uint64)])
        r, b = e.rank(pos)
        for p, rr, bb in zip(pos, r, b):
            er, eb = f.rank(int(p))
            assert (rr, bb) == (er, eb), p


@pytest.mark.parametrize("h,bits,env", [
    (3, 1000000 + 64, {"GRB_FILL_PSHIFT": "16"}),                       # 16 partitions
    (3, 1000000 + 64, {"GRB_FILL_PSHIFT": "16", "GRB_FILL_BS": "1024"}),
    (1, 333376, {"GRB_FILL_PSHIFT": "12"}),                             # 82 partitions
    (4, 5000000 + 128, {"GRB_FILL_PSHIFT": "14"}),                      # 306 partitions
    (4, 5000000 + 128, {"GRB_FILL_PSHIFT": "12", "GRB_FILL_MAXPART": "100"}),  # folded to 77
    (3, 1000000 + 64, {"GRB_FILL_PSHIFT": "16", "GRB_FILL_CAP": "300"}),  # lists overflow
    (2, 700032, {}),                                                    # one partition
])
def test_partitioned_fill_equals_oracle(h, bits, env, monkeypatch):
    """Pass 1 through the L2-partitioned fill (k_fill_part + k_fill_apply), forced on small filters:
    the same bit vector as MIBFConstructSupport::insertBV, including the list-overflow fallback."""
    monkeypatch.setenv("GRB_FILL", "part")
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    seeds = grb.make_seed_pattern(SEED22, 22, 16, h)
    rng = np.random.default_rng(300 + h)
    e, f, recs = _build_pair(rng, 60, 30, 9000, seeds, bits)
    with e:
        assert np.array_equal(e.copy_bitvector(), f.words())
        assert e.finalize_bitvector() == f.setup()


def test_probe_microbenchmark_builds_the_filter_it_claims():
    """grb_probe_bench (cfg5): the synthetic filter reaches the requested share of set bits, every
    probe of the timed launches meets a set bit (the checksum carries no miss marker), and half of
    the slots hold an ID."""
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    with grb.Engine(seeds, genome_size=1000000, weight=16) as e:
        for h in (1, 3, 5):
            r = e.probe_bench(50_000_000 + 64, 0.3, h, n_probes=1 << 20, reps=2)
            assert abs(r.pop / r.filter_bits - 0.3) < 0.005
            assert r.probes == ((1 << 20) // h) * h and r.query_ms > 0 and r.insert_ms > 0
            assert r.checksum > 0 and r.probes_missed == 0
            assert r.footprint_bytes == ((r.filter_bits + 191) // 192) * 32 + (r.pop + 1) * 8  # 8-byte {id, count} slots


def _tile_hashes(seq, t, T, k, seeds):
    tile = seq[t * T:t * T + T + k - 1]
    return ou.hash_sequence(tile, seeds)


@pytest.mark.parametrize("h,T,B", [(3, 1000, 10), (2, 200, 3), (4, 300, 1)])
def test_query_vote_and_insert_match_oracle(h, T, B):
    seeds = grb.make_seed_pattern(SEED22, 22, 16, h)
    k = 22
    rng = np.random.default_rng(7 * h + T)
    bits = 3000000 + 64
    e, f, recs = _build_pair(rng, 12, 4 * T, 12 * T + 50, seeds, bits, tile_length=T, block_size=B)
    with e:
        pop = e.finalize_bitvector()
        assert pop == f.setup()
        next_id = 1
        for step in range(3):
            for ri, (_, seq, _) in enumerate(recs):
                nt = len(seq) // T
                bi, bc, nc, ci, cc, cnt = e.query_read(ri, nt, cand_cap=64)
                exp_cnt = np.zeros(3, dtype=np.uint64)
                for t in range(nt):
                    hv = _tile_hashes(seq, t, T, k, seeds)
                    obi, obc, on, oci, occ, oc = f.query_tile(hv, 64)
                    exp_cnt += oc
                    assert (bi[t], bc[t], nc[t]) == (obi, obc, on), (step, ri, t)
                    assert list(ci[t][:on]) == list(oci) and list(cc[t][:on]) == list(occ)
                assert np.array_equal(cnt, exp_cnt)
                # insert a tile range as ONE call, ids chosen to exercise the reservoir rule
                a = int(rng.integers(0, nt))
                b = int(rng.integers(a + 1, nt + 1))
                e.insert_tiles(ri, a, b, next_id)
                flat = np.concatenate([_tile_hashes(seq, t, T, k, seeds).ravel() for t in range(a, b)])
                f.insert_mibf(flat, next_id)
                next_id += int(rng.integers(1, 3))
            ranks = np.arange(pop, dtype=np.uint64)
            ids, counts = e.get_ids(ranks)
            for r in rng.integers(0, pop, size=20000):
                assert (ids[r], counts[r]) == f.get(int(r)), (step, r)
            nz = np.flatnonzero(counts)
            for r in nz[:20000]:
                assert (ids[r], counts[r]) == f.get(int(r))
        # saturation bit is preserved by setData and masked by the query
        some = np.flatnonzero(ids)[:50].astype(np.uint64)
        e.set_ids(some, ids[some] | np.uint32(0x80000000), counts[some])
        for r in some:
            f.set(int(r), int(ids[r]) | 0x80000000, int(counts[r]))
        seq = recs[0][1]
        nt = len(seq) // T
        bi, bc, nc, ci, cc, cnt = e.query_read(0, nt, cand_cap=64)
        for t in range(nt):
            obi, obc, on, oci, occ, oc = f.query_tile(_tile_hashes(seq, t, T, k, seeds), 64)
            assert (bi[t], bc[t], nc[t]) == (obi, obc, on)
        e.insert_tiles(0, 0, nt, 4242)
        f.insert_mibf(np.concatenate([_tile_hashes(seq, t, T, k, seeds).ravel() for t in range(nt)]),
                      4242)
        ids2, counts2 = e.get_ids(some)
        for r, i2, c2 in zip(some, ids2, counts2):
            assert (i2, c2) == f.get(int(r))
        e.reset_ids()
        ids3, counts3 = e.get_ids(ranks)
        assert not ids3.any() and not counts3.any()


def _args_to_params(args):
    """goldrush-path command line -> grb_run_path keyword arguments."""
    m = {"-k": ("kmer_size", int), "-w": ("weight", int), "-h": ("hash_num", int),
         "-t": ("tile_length", int), "-u": ("unassigned_min", int), "-a": ("assigned_max", int),
         "-o": ("occupancy", float), "-x": ("threshold", int), "-b": ("block_size", int),
         "-d": ("phred_delta", int), "-P": ("phred_min", int), "-r": ("ratio", float),
         "-M": ("max_paths", int), "-m": ("min_length", int), "-H": ("hash_universe", int),
         "-g": ("genome_size", lambda v: int(float(v)))}
    kw, i = {}, 0
    while i < len(args):
        a = args[i]
        if a == "-s":
            kw["seed_preset"] = args[i + 1]
            i += 2
        elif a == "-f":
            kw["filter_file"] = args[i + 1]
            i += 2
        elif a in m:
            kw[m[a][0]] = m[a][1](args[i + 1])
            i += 2
        elif a == "--silver_path":
            kw["silver_path"] = 1
            i += 1
        elif a == "--ntcard":
            kw["ntcard"] = True
            i += 1
        elif a == "--verbose":
            kw["verbose"] = True
            i += 1
        else:
            raise ValueError(a)
    return kw


_product_outputs = {}


def _product_outputs_for(case, workdir):
    """Runs the engine through grb_run_path with the FASTQ in a HOST buffer."""
    if case["name"] not in _product_outputs:
        inp, extra = pu.make_input(case, workdir, _product_outputs_for)
        with open(inp, "rb") as f:
            data = f.read()
        prefix = os.path.join(workdir, case["name"] + ".gpu")
        for fn in os.listdir(workdir):
            if fn.startswith(os.path.basename(prefix)):
                os.remove(os.path.join(workdir, fn))
        kw = _args_to_params(case["args"] + extra)
        res = grb.run_path(data, input_path=inp, prefix=prefix, quiet=True, **kw)
        outs = sorted((os.path.join(workdir, fn) for fn in os.listdir(workdir)
                       if fn.startswith(os.path.basename(prefix))),
                      key=lambda p: (len(p), p))
        _product_outputs[case["name"]] = (outs, res, inp)
    return _product_outputs[case["name"]][0]


@pytest.mark.parametrize("name", [c["name"] for c in pu.golden_cases.CASES])
def test_selected_reads_match_reference_fixture(name, workdir):
    """Silver / golden path files byte-identical to what the reference wrote for the same input."""
    case = pu.case_by_name(name)
    outs = _product_outputs_for(case, workdir)
    g = GOLDEN[name]
    assert pu.digest_outputs(outs) == g["outputs"]
    res = _product_outputs[name][1]
    stats = dict((k, v) for k, v in g["stats"] if k in ("m_filterSize:", "num_passed_reads:"))
    assert res.filter_bits == stats["m_filterSize:"]
    assert res.num_passed_reads == stats["num_passed_reads:"]


@pytest.mark.parametrize("name", ["silver_default", "golden_lognormal_all_lengths", "ntcard_sizing",
                                  "filter_list"])
def test_cli_binary_matches_reference_fixture(name, workdir):
    """The drop-in executable: same files AND the same --verbose counters on stderr."""
    case = pu.case_by_name(name)
    inp, extra = pu.make_input(case, workdir, _product_outputs_for)
    rc, outs, err = pu.run_cli(pu.PRODUCT, case, inp, extra, workdir, "cli", jobs=4)
    g = GOLDEN[name]
    assert rc == g["exit_code"], err[-2000:]
    assert pu.digest_outputs(outs) == g["outputs"]
    assert pu.parse_stats(err) == g["stats"]


@pytest.mark.parametrize("case", fuzz_cases.CASES, ids=[c["name"] for c in fuzz_cases.CASES])
def test_parameter_fuzz_cli_equals_oracle(case, workdir):
    """Seeded fuzz over the option space (tests/fuzz_cases.py: k 16-32, any weight, h 1-4, tile
    200-1000, block / smoothing / Phred options, silver and golden mode, fixed and log-normal
    lengths): the drop-in executable against the CPU oracle run here on the same input — files by
    md5, exit code, --verbose counters.  The CPU suite checks the oracle against the reference's
    own sources on this very list."""
    inp, extra = pu.make_input(case, workdir)
    rc_o, outs_o, err_o = pu.run_cli(pu.ORACLE, case, inp, extra, workdir, "ora", jobs=4)
    rc_g, outs_g, err_g = pu.run_cli(pu.PRODUCT, case, inp, extra, workdir, "gpu", jobs=4)
    assert rc_g == rc_o == 0, err_g[-2000:]
    assert outs_o and pu.digest_outputs(outs_g) == pu.digest_outputs(outs_o)
    assert pu.parse_stats(err_g) == pu.parse_stats(err_o)


@pytest.mark.parametrize("name", sorted(fuzz_cases.edge_inputs()))
def test_degenerate_inputs_cli_equals_oracle(name, workdir):
    """Empty and one-newline files, reads shorter than k / than a tile / of exactly one tile, a last
    record without newline: the drop-in executable exits as the oracle does (which the CPU suite
    holds against the reference's own sources) and writes the same files and counters."""
    inp = os.path.join(workdir, name + ".fq")
    with open(inp, "w") as f:
        f.write(fuzz_cases.edge_inputs()[name])
    case = dict(name="edge_" + name, args=fuzz_cases.EDGE_ARGS)
    rc_o, outs_o, err_o = pu.run_cli(pu.ORACLE, case, inp, [], workdir, "ora", jobs=2)
    rc_g, outs_g, err_g = pu.run_cli(pu.PRODUCT, case, inp, [], workdir, "gpu", jobs=2)
    assert rc_g == rc_o == fuzz_cases.EDGE_EXIT.get(name, 0), err_g[-2000:]
    assert pu.digest_outputs(outs_g) == pu.digest_outputs(outs_o)
    assert pu.parse_stats(err_g) == pu.parse_stats(err_o)


@pytest.mark.parametrize("pair", [("silver_default", "golden_default"), ("silver_m12", "golden_m12")])
def test_two_stage_call_equals_two_reference_runs(pair, workdir):
    """grb_run_two_stage (bin/goldrush:240-260 in one call): the silver files of the silver case
    AND the golden path the reference wrote from their concatenation, without the files travelling
    through a second process.  With 12 paths the concatenation order is the shell glob's
    (_1, _10, _11, _12, _2, ...), not the numeric one."""
    silver, golden = pu.case_by_name(pair[0]), pu.case_by_name(pair[1])
    inp, extra = pu.make_input(silver, workdir, _product_outputs_for)
    with open(inp, "rb") as f:
        data = f.read()
    ps, pg = os.path.join(workdir, "two.silver"), os.path.join(workdir, "two.golden")
    for fn in os.listdir(workdir):
        if fn.startswith("two."):
            os.remove(os.path.join(workdir, fn))
    rs, rg = grb.run_two_stage(data, dict(prefix=ps, **_args_to_params(silver["args"])),
                               dict(prefix=pg, **_args_to_params(golden["args"])), input_path=inp)
    s_outs = sorted((os.path.join(workdir, fn) for fn in os.listdir(workdir)
                     if fn.startswith("two.silver")), key=lambda p: (len(p), p))
    g_outs = [os.path.join(workdir, "two.golden.fa")]
    assert pu.digest_outputs(s_outs) == GOLDEN[pair[0]]["outputs"]
    assert pu.digest_outputs(g_outs) == GOLDEN[pair[1]]["outputs"]
    assert rs.reads_selected > 0 and rg.reads_selected > 0
    assert rg.num_reads == rs.reads_selected  # the second stage read exactly the silver records


def test_cli_error_paths(workdir):
    import subprocess
    fa = os.path.join(workdir, "x.fa")
    with open(fa, "w") as f:
        f.write(">a\nACGT\n")
    base = [pu.PRODUCT, "-k", "22", "-w", "16", "-s", SEED22, "-g", "1000"]
    p = subprocess.run(base + ["-i", fa], capture_output=True)
    assert p.returncode == 1 and b"Gold Path requires fastq format" in p.stderr
    p = subprocess.run([pu.PRODUCT, "-w", "16", "-g", "1000", "-i", fa], capture_output=True)
    assert p.returncode == 1 and b"span of spaced seed cannot be 0" in p.stderr
    p = subprocess.run(base[:-2] + ["-i", fa], capture_output=True)
    assert p.returncode == 1 and b"genome size cannot be 0" in p.stderr
    fq = os.path.join(workdir, "short.fq")
    with open(fq, "w") as f:
        f.write("@a\n" + "ACGT" * 100 + "\n+\n" + "I" * 400 + "\n")
    p = subprocess.run(base + ["-i", fq, "-m", "20000"], capture_output=True)
    assert p.returncode == 1 and b"no reads passed" in p.stderr


def test_split_selection_calls_equal_one_call():
    """grb_select_reads keeps its loop state between calls: two halves == one pass."""
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    sp = grb.api.synth_params(200000, 12.0, 5000, 31)
    data = grb.synth_fastq(sp)
    outs = []
    for split in (False, True):
        with grb.Engine(seeds, genome_size=200000, weight=16, tile_length=250, min_length=5000,
                        silver_path=1, max_paths=3, ratio=0.9) as e:
            e.reads_ingest_fastq(data)
            n = e.reads_count()
            e.reads_set_flags(np.full(n, 3, dtype=np.uint8))
            e.filter_alloc(grb.calc_optimal_size(grb.default_hash_universe(16, 200000, 3), 1, 0.1))
            e.build_bitvector()
            e.finalize_bitvector()
            if split:
                d1, s1, f1 = e.select_reads(0, n // 2)
                d2, s2, f2 = e.select_reads(n // 2, n - n // 2)
                dec, st = d1 + d2, s1 + s2
            else:
                dec, st, fin = e.select_reads()
            outs.append(([(d.verdict, d.path, d.trim_start, d.trim_end) for d in dec],
                         [(s.valid_reads, s.queries, s.hits, s.misses, s.rollover_read) for s in st]))
    assert outs[0] == outs[1]
    assert any(v[0] in (2, 3) for v in outs[0][0])


def _select_all(data, env, **params):
    """Decisions + rollover snapshots + final counters of one full selection under `env`."""
    seeds = grb.make_seed_pattern(SEED22, 22, 16, params.pop("hash_num", 3))
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        with grb.Engine(seeds, weight=16, **params) as e:
            e.reads_ingest_fastq(data)
            n = e.reads_count()
            e.reads_set_flags(np.full(n, 3, dtype=np.uint8))
            e.filter_alloc(grb.calc_optimal_size(
                grb.default_hash_universe(16, params["genome_size"], len(seeds)), 1, 0.1))
            e.build_bitvector()
            pop = e.finalize_bitvector()
            dec, st, fin = e.select_reads()
            cur, path, ids = e.select_state()
            ranks = np.arange(min(pop, 2000000), dtype=np.uint64)
            slot_ids, slot_counts = e.get_ids(ranks)
            return ([(d.verdict, d.path, d.trim_start, d.trim_end, d.num_tiles, d.num_assigned)
                     for d in dec],
                    [(s.valid_reads, s.total_tiles, s.assigned_tiles, s.queries, s.hits, s.misses,
                      s.num_reads_in_path, s.inserted_bases, s.rollover_read) for s in st],
                    (cur.valid_reads, cur.queries, cur.hits, cur.misses, path, ids, fin),
                    slot_ids.tobytes(), slot_counts.tobytes())
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("shape", [
    dict(genome=300000, cov=14.0, read_len=6000, seed=41, tile_length=300, block_size=4,
         max_paths=3, ratio=0.9, threshold=10, hash_num=3),
    dict(genome=150000, cov=25.0, read_len=0, seed=42, tile_length=200, block_size=1,
         max_paths=2, ratio=0.8, threshold=5, hash_num=2),
    dict(genome=200000, cov=20.0, read_len=9000, seed=43, tile_length=500, block_size=10,
         max_paths=1, ratio=0.9, threshold=10, hash_num=3, silver=0),
])
def test_batch_engine_equals_serial_engine(shape):
    """The speculative batch + ordered commit must reproduce the one-read-at-a-time loop exactly:
    same decisions, same per-path counters (queries / hits / misses included), same final ID and
    count slots, for every batch size (1 = no speculation at all)."""
    sp = grb.api.synth_params(shape["genome"], shape["cov"], shape["read_len"], shape["seed"],
                              n50=7000)
    data = grb.synth_fastq(sp)
    params = dict(genome_size=shape["genome"], tile_length=shape["tile_length"],
                  block_size=shape["block_size"], min_length=3 * shape["tile_length"],
                  silver_path=shape.get("silver", 1), max_paths=shape["max_paths"],
                  ratio=shape["ratio"], threshold=shape["threshold"], hash_num=shape["hash_num"])
    ref = _select_all(data, {"GRB_ENGINE": "serial"}, **dict(params))
    assert any(d[0] in (2, 3) for d in ref[0]) and any(d[0] == 4 for d in ref[0])
    for eng, b in (("batch", "1"), ("batch", "3"), ("batch", "32"), ("batch", "128"),
                   ("batch", "512")):
        got = _select_all(data, {"GRB_ENGINE": eng, "GRB_BATCH_READS": b}, **dict(params))
        assert got[0] == ref[0], b
        assert got[1] == ref[1], b
        assert got[2] == ref[2], b
        assert got[3] == ref[3] and got[4] == ref[4], b


# ---- BASELINE.json sizes ---------------------------------------------------------------------------
CFG2 = dict(genome=100_000_000, cov=30.0, read_len=25000, seed=1002)  # configs[1], SURVEY.md 8(d)
CFG2_PARAMS = dict(kmer_size=22, weight=16, hash_num=3, tile_length=1000, block_size=10,
                   unassigned_min=5, assigned_max=1, occupancy=0.1, threshold=10, phred_delta=5,
                   ratio=0.9, max_paths=5, min_length=20000, silver_path=1, phred_min=20)


def test_cfg2_shaped_sample_matches_cpu_oracle(workdir):
    """Reads of the bench workload's shape (25 kbp, tile 1000, k 22 / w 16 / h 3, -P 20): the first
    1500 reads of the cfg2 read set against the CPU oracle, byte for byte, with the genome size set
    so that all five silver paths close inside the sample (rollovers + the exit(0) at path M + 1)."""
    import subprocess
    sp = grb.api.synth_params(CFG2["genome"], CFG2["cov"], CFG2["read_len"], CFG2["seed"])
    data = grb.synth_fastq(sp, 0, 1500)
    fq = os.path.join(workdir, "cfg2_sample.fq")
    with open(fq, "wb") as f:
        f.write(data)
    params = dict(CFG2_PARAMS, genome_size=2_000_000)
    res = grb.run_path(data, input_path=fq, prefix=os.path.join(workdir, "cfg2s.gpu"), quiet=True,
                       seed_preset=SEED22, **params)
    args = ["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "1000", "-b", "10", "-u", "5", "-a",
            "1", "-o", "0.1", "-x", "10", "-d", "5", "-r", "0.9", "-M", "5", "-m", "20000", "-P", "20",
            "-g", "2000000", "--silver_path"]
    subprocess.check_call([pu.ORACLE] + args + ["-j", "8", "-i", fq, "-p",
                                                os.path.join(workdir, "cfg2s.cpu")],
                          stderr=subprocess.DEVNULL)
    n_files = 0
    for i in range(1, 6):
        g, c = (os.path.join(workdir, f"cfg2s.{t}_{i}.fq") for t in ("gpu", "cpu"))
        assert os.path.exists(g) == os.path.exists(c), i
        if os.path.exists(c):
            n_files += 1
            with open(g, "rb") as fg, open(c, "rb") as fc:
                assert pu.golden_cases.md5(fg.read()) == pu.golden_cases.md5(fc.read()), i
    assert n_files >= 2 and res.paths >= 2 and res.reads_selected > 0


def test_cfg2_full_size_is_independent_of_batching_and_fill_path(monkeypatch):
    """The whole cfg2 read set (120 000 reads, 3.0 Gbp): how the work is cut must not show in the
    result.  320-read batches + L2-partitioned fill + 16 384-read slices (defaults) against 96-read
    batches + direct atomic fill + one slice: same output digest, same counts, same filter."""
    sp = grb.api.synth_params(CFG2["genome"], CFG2["cov"], CFG2["read_len"], CFG2["seed"])
    ptr, n = grb.synth_fastq_raw(sp)
    try:
        kw = dict(nbytes=n, input_path="(memory)", seed_preset=SEED22, write_outputs=False,
                  quiet=True, genome_size=CFG2["genome"], **CFG2_PARAMS)
        a = grb.run_path(ptr, **kw)
        monkeypatch.setenv("GRB_BATCH_READS", "96")
        monkeypatch.setenv("GRB_FILL", "direct")
        monkeypatch.setenv("GRB_SLICE_READS", "0")
        b = grb.run_path(ptr, **kw)
    finally:
        grb.free_host(ptr)
    key = lambda r: (r.out_digest, r.num_reads, r.num_passed_reads, r.reads_visited, r.bases_pass2,
                     r.reads_selected, r.bases_selected, r.filter_bits, r.pop, r.paths)
    assert key(a) == key(b)
    assert a.num_reads == 120000 and a.reads_selected > 20000 and a.paths in (5, 6)
    assert a.pop < a.filter_bits and a.bases_pass2 == a.reads_visited * 25000


# ---- full-size reference digests (tests/golden/full_size.json, made by make_full_size.py) ---------
def _full_size():
    import json
    with open(os.path.join(pu.ROOT, "tests", "golden", "full_size.json")) as f:
        return json.load(f)


def _stats_dict(stats, which):
    """Last value of each --verbose counter of parse_stats (the final path's running totals)."""
    out = {}
    for k, v in stats:
        out[k] = v
    return out[which]


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_full_size_config_matches_reference_digests(cfg):
    """BASELINE.json configs[0] and configs[1] at their own size: both launches of one assembly
    (silver run, golden run on the concatenated silver paths, bin/goldrush:240-260) through
    grb_run_two_stage on a host FASTQ buffer, against what the reference's own sources wrote for the
    same input: record digest of all output files, record and byte counts, filter size, and the
    number of reads that passed pass 1."""
    fs = _full_size()[cfg]
    s = fs["synth"]
    sp = grb.api.synth_params(s["genome"], float(s["cov"]), s["read_len"], s["seed"])
    ptr, n = grb.synth_fastq_raw(sp)
    assert n == fs["input_bytes"], "generator drifted from the fixture"
    common = dict(kmer_size=22, weight=16, hash_num=3, tile_length=1000, block_size=10,
                  unassigned_min=5, assigned_max=1, occupancy=0.1, threshold=10, phred_delta=5,
                  ratio=0.9, genome_size=s["genome"], phred_min=fs["phred_min"], seed_preset=SEED22)
    # page-locked in 1 GiB pieces (grb_host_pin): the ingest copies of cfg2's 6 GB cross five piece
    # boundaries, and one cudaMemcpyAsync must not span two registrations
    pinned = grb.api.host_pin(ptr, n)
    assert pinned == n
    try:
        rs, rg = grb.run_two_stage(ptr, dict(common, max_paths=5, min_length=20000, silver_path=1),
                                   dict(common, min_length=0, silver_path=0), nbytes=n,
                                   input_path="(memory)", write_outputs=False)
    finally:
        grb.api.host_unpin(ptr, pinned)
        grb.free_host(ptr)
    for res, ref in ((rs, fs["silver"]), (rg, fs["golden"])):
        assert res.out_digest == ref["out_digest"]
        assert res.reads_selected == ref["records"]
        assert res.filter_bits == _stats_dict(ref["stats"], "m_filterSize:")
        assert res.num_passed_reads == _stats_dict(ref["stats"], "num_passed_reads:")
    assert rg.num_reads == rs.reads_selected


def test_reads_of_hundreds_of_tiles_match_cpu_oracle(workdir):
    """Reads far longer than 160 tiles (the count matrix of the smoothing passes leaves shared
    memory, the human-scale config's 200 kbp reads at tile 1000 do the same): log-normal lengths with
    N50 20 kbp at tile length 100, i.e. 200 tiles typical and up to 2000, against the CPU oracle byte
    for byte."""
    import subprocess
    sp = grb.api.synth_params(400000, 12.0, 0, 91, n50=20000)
    data = grb.synth_fastq(sp)
    fq = os.path.join(workdir, "longreads.fq")
    with open(fq, "wb") as f:
        f.write(data)
    params = dict(kmer_size=22, weight=16, hash_num=3, tile_length=100, block_size=10,
                  unassigned_min=5, assigned_max=1, occupancy=0.1, threshold=10, phred_delta=5,
                  ratio=0.9, max_paths=3, min_length=2000, silver_path=1, phred_min=15,
                  genome_size=400000)
    res = grb.run_path(data, input_path=fq, prefix=os.path.join(workdir, "long.gpu"), quiet=True,
                       seed_preset=SEED22, **params)
    args = ["-k", "22", "-w", "16", "-s", SEED22, "-h", "3", "-t", "100", "-b", "10", "-u", "5", "-a",
            "1", "-o", "0.1", "-x", "10", "-d", "5", "-r", "0.9", "-M", "3", "-m", "2000", "-P", "15",
            "-g", "400000", "--silver_path"]
    subprocess.check_call([pu.ORACLE] + args + ["-j", "8", "-i", fq, "-p",
                                                os.path.join(workdir, "long.cpu")],
                          stderr=subprocess.DEVNULL)
    n_files = 0
    for i in range(1, 4):
        g, c = (os.path.join(workdir, f"long.{t}_{i}.fq") for t in ("gpu", "cpu"))
        assert os.path.exists(g) == os.path.exists(c), i
        if os.path.exists(c):
            n_files += 1
            with open(g, "rb") as fg, open(c, "rb") as fc:
                assert pu.golden_cases.md5(fg.read()) == pu.golden_cases.md5(fc.read()), i
    assert n_files >= 1 and res.reads_selected > 0
