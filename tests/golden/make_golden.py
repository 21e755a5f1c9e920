#!/usr/bin/env python
"""Generates tests/golden/cases.json by running the REFERENCE (oracle/_ref/goldrush-path-ref: the
unmodified sources of /root/reference/goldrush_path compiled against oracle/shim) on every case in
cases.py, and cross-checks the CPU oracle against it on the spot.

Run in the authoring container only (needs /root/reference to build oracle/_ref):
    make -C oracle all && make host-tools && python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parity_util as pu  # noqa: E402


def main():
    work = tempfile.mkdtemp(prefix="grb_golden_")
    produced = {}

    def ref_outputs(case, workdir):
        if case["name"] not in produced:
            run(case)
        return produced[case["name"]]

    results = {}

    def run(case):
        inp, extra = pu.make_input(case, work, ref_outputs)
        with open(inp, "rb") as f:
            in_md5 = pu.golden_cases.md5(f.read())
        rc, outs, err = pu.run_cli(pu.REF, case, inp, extra, work, "ref", jobs=2)
        rc1, outs1, err1 = pu.run_cli(pu.REF, case, inp, extra, work, "ref1", jobs=1)
        rco, outso, erro = pu.run_cli(pu.ORACLE, case, inp, extra, work, "ora", jobs=8)
        d, d1, do = pu.digest_outputs(outs), pu.digest_outputs(outs1), pu.digest_outputs(outso)
        assert rc == rc1 == rco, (case["name"], rc, rc1, rco, err[-500:], erro[-500:])
        assert d == d1, ("reference output depends on -j", case["name"])
        assert d == do, ("oracle differs from reference", case["name"], d, do)
        st, sto = pu.parse_stats(err), pu.parse_stats(erro)
        assert st == sto, ("oracle counters differ", case["name"], st, sto)
        produced[case["name"]] = outs
        results[case["name"]] = {"input_md5": in_md5, "exit_code": rc, "outputs": d, "stats": st}
        print(case["name"], "ok:", [(o["suffix"], o["bytes"]) for o in d], file=sys.stderr)

    for case in pu.golden_cases.CASES:
        if case["name"] not in results:
            run(case)
    with open(pu.GOLDEN_JSON, "w") as f:
        json.dump(results, f, indent=1, sort_keys=True)
    print("wrote", pu.GOLDEN_JSON)


if __name__ == "__main__":
    main()
