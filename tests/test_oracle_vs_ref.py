"""CPU: the oracle restatement (oracle/grb_oracle.cpp) against the REFERENCE binary
(oracle/_ref/goldrush-path-ref = the reference's own goldrush_path/*.cpp compiled unmodified by
oracle/Makefile) on inputs that are NOT in tests/golden/: other seeds, k/w/h, tile and block sizes.
The binary is built in the authoring container and travels with the snapshot; where it is absent
the committed fixtures (test_oracle_golden.py) are the pin and this module is skipped."""
import os

import pytest

import fuzz_cases
import parity_util as pu

pytestmark = pytest.mark.skipif(not os.path.exists(pu.REF), reason="oracle/_ref not built here")

S = pu.golden_cases.synth_args
CASES = [
    dict(name="xr_silver_h2", synth=S(150000, 10, 5000, 31),
         args=["-k", "20", "-w", "12", "-h", "2", "-t", "400", "-b", "3", "-u", "4", "-a", "1", "-o",
               "0.1", "-x", "8", "-d", "5", "-P", "0", "-g", "150000", "-r", "0.9", "-M", "3", "-m",
               "5000", "--silver_path", "--verbose"]),
    dict(name="xr_golden_h4", synth=S(120000, 8, 6000, 32),
         args=["-k", "24", "-w", "16", "-h", "4", "-t", "500", "-b", "5", "-u", "5", "-a", "1", "-o",
               "0.15", "-x", "10", "-d", "6", "-P", "15", "-g", "120000", "-m", "0", "--verbose"]),
    dict(name="xr_lognormal", synth=S(200000, 9, 0, 33, n50=7000),
         args=["-k", "22", "-w", "16", "-s", pu.golden_cases.SEED22, "-h", "3", "-t", "500", "-b",
               "4", "-u", "5", "-a", "1", "-o", "0.1", "-x", "10", "-d", "5", "-P", "0", "-g",
               "200000", "-r", "0.8", "-M", "2", "-m", "4000", "--silver_path", "--verbose"]),
]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_equals_reference_binary(case, workdir):
    inp, extra = pu.make_input(case, workdir)
    rc_r, outs_r, err_r = pu.run_cli(pu.REF, case, inp, extra, workdir, "ref", jobs=2)
    rc_o, outs_o, err_o = pu.run_cli(pu.ORACLE, case, inp, extra, workdir, "ora", jobs=4)
    assert rc_r == rc_o, (err_r[-400:], err_o[-400:])
    assert outs_r, "reference wrote nothing: " + err_r[-400:]
    assert pu.digest_outputs(outs_r) == pu.digest_outputs(outs_o)
    assert pu.parse_stats(err_r) == pu.parse_stats(err_o)


@pytest.mark.parametrize("case", fuzz_cases.CASES, ids=[c["name"] for c in fuzz_cases.CASES])
def test_parameter_fuzz_oracle_equals_reference_binary(case, workdir):
    """Seeded fuzz over k / weight / h / tile / smoothing / Phred options and both modes
    (tests/fuzz_cases.py): files, exit code and --verbose counters of the port equal the reference
    sources'.  The GPU suite runs the drop-in executable over the same list."""
    inp, extra = pu.make_input(case, workdir)
    rc_r, outs_r, err_r = pu.run_cli(pu.REF, case, inp, extra, workdir, "ref", jobs=2)
    rc_o, outs_o, err_o = pu.run_cli(pu.ORACLE, case, inp, extra, workdir, "ora", jobs=4)
    assert rc_r == rc_o == 0, (err_r[-400:], err_o[-400:])
    assert outs_r and sum(d["bytes"] for d in pu.digest_outputs(outs_r)) > 0
    assert pu.digest_outputs(outs_r) == pu.digest_outputs(outs_o)
    assert pu.parse_stats(err_r) == pu.parse_stats(err_o)


@pytest.mark.parametrize("name", sorted(fuzz_cases.edge_inputs()))
def test_degenerate_inputs_oracle_equals_reference_binary(name, workdir):
    """Empty and one-newline files, reads shorter than k / than a tile / of exactly one tile, a last
    record without newline: same exit code, same files, same counters."""
    inp = os.path.join(workdir, name + ".fq")
    with open(inp, "w") as f:
        f.write(fuzz_cases.edge_inputs()[name])
    case = dict(name="edge_" + name, args=fuzz_cases.EDGE_ARGS)
    rc_r, outs_r, err_r = pu.run_cli(pu.REF, case, inp, [], workdir, "ref", jobs=2)
    rc_o, outs_o, err_o = pu.run_cli(pu.ORACLE, case, inp, [], workdir, "ora", jobs=2)
    assert rc_r == rc_o == fuzz_cases.EDGE_EXIT.get(name, 0), (err_r[-300:], err_o[-300:])
    assert pu.digest_outputs(outs_r) == pu.digest_outputs(outs_o)
    assert pu.parse_stats(err_r) == pu.parse_stats(err_o)
    if rc_r == 0:
        assert outs_r and os.path.getsize(outs_r[0]) > 0


def test_odd_k_aborts_in_reference_and_oracle(workdir):
    """An odd -k gives seeds of span k - 1 and the reference dies on the assertion in the filter's
    constructor (MIBloomFilter.hpp:180) after pass 1; the oracle mirrors that, the engine refuses
    the option up front (test_host_logic.py)."""
    import signal
    case = pu.golden_cases.ODD_K_CASE
    inp, extra = pu.make_input(case, workdir)
    rc_r, outs_r, err_r = pu.run_cli(pu.REF, case, inp, extra, workdir, "ref", jobs=2)
    rc_o, outs_o, err_o = pu.run_cli(pu.ORACLE, case, inp, extra, workdir, "ora", jobs=2)
    assert rc_r == -signal.SIGABRT and rc_o == -signal.SIGABRT, (rc_r, rc_o)
    assert "m_sseeds[0].size() == kmerSize" in err_r and "m_sseeds[0].size() == kmerSize" in err_o
    for o in outs_r + outs_o:  # the first output file is opened before the abort and stays empty
        assert os.path.getsize(o) == 0
