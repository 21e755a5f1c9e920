"""CPU: the oracle reproduces every golden fixture (outputs of the REFERENCE run in the authoring
container, tests/golden/make_golden.py), and the synthetic generator still produces the inputs the
fixtures were made from."""
import pytest

import parity_util as pu

GOLDEN = pu.load_golden()
_outputs = {}


def _oracle_outputs(case, workdir):
    if case["name"] not in _outputs:
        inp, extra = pu.make_input(case, workdir, _oracle_outputs)
        rc, outs, err = pu.run_cli(pu.ORACLE, case, inp, extra, workdir, "ora", jobs=8)
        _outputs[case["name"]] = (rc, outs, err, inp)
    return _outputs[case["name"]][1]


@pytest.mark.parametrize("name", [c["name"] for c in pu.golden_cases.CASES])
def test_oracle_matches_reference_fixture(name, workdir):
    case = pu.case_by_name(name)
    _oracle_outputs(case, workdir)
    rc, outs, err, inp = _outputs[name]
    g = GOLDEN[name]
    with open(inp, "rb") as f:
        assert pu.golden_cases.md5(f.read()) == g["input_md5"], "generator drifted from the fixture"
    assert rc == g["exit_code"]
    assert pu.digest_outputs(outs) == g["outputs"]
    assert pu.parse_stats(err) == g["stats"]
