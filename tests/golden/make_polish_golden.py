#!/usr/bin/env python
"""tests/golden/polish.json: md5 of the Bloom filters the REFERENCE's own fill_bfs
(oracle/_ref/libgoldpolish_ref.so: subprojects/goldpolish/src/utils.cpp compiled unmodified against
the stand-in btllib headers of oracle/shim_polish) produces for the cases of tests/polish_util.py,
cross-checked against the port on the spot.  Authoring container only:
    make -C oracle all && python tests/golden/make_polish_golden.py"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polish_util as pu  # noqa: E402


def main():
    out = {}
    for name in pu.CASES:
        batches, ks, h, cbf, bf = pu.case_batches(name)
        r = pu.ref_fill(batches, ks, h, cbf, bf)
        p = pu.port_fill(batches, ks, h, cbf, bf)
        assert (r == p).all(), name
        out[name] = {"md5": hashlib.md5(r.tobytes()).hexdigest(),
                     "set_bits": [[int(x) for x in row] for row in
                                  __import__("numpy").unpackbits(r, axis=2).sum(axis=2)]}
        print(name, out[name]["md5"], out[name]["set_bits"], file=sys.stderr)
    with open(os.path.join(pu.ROOT, "tests", "golden", "polish.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
