#!/usr/bin/env python
"""tests/golden/polish_inputs.json: what the REFERENCE's own SeqIndex, AllMappings and serve_batch
(oracle/_ref/libgoldpolish_ref.so: subprojects/goldpolish/src/{seqindex,mappings,utils,
goldpolish_targeted_bfs}.cpp compiled unmodified against oracle/shim_polish) produce for the scenarios
of tests/polish_inputs_util.py: md5 of the sorted index lines, the kept mappings per target, md5 and set
bits of every batch's Bloom filters.  Authoring container only:
    make -C oracle all && python tests/golden/make_polish_inputs_golden.py"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polish_inputs_util as piu  # noqa: E402
import polish_util as pu  # noqa: E402


def main():
    out = {}
    with tempfile.TemporaryDirectory() as wd:
        for name in piu.SCENARIOS:
            sc = piu.make_files(name, wd)
            ti, ri = os.path.join(sc["dir"], "targets.ref.idx"), os.path.join(sc["dir"], "reads.ref.idx")
            piu.ref_index(sc["targets"], ti)
            piu.ref_index(sc["reads"], ri)
            bfs = piu.ref_serve(sc, ti, ri)
            out[name] = {
                "target_index": piu.sorted_lines_md5(ti), "mapped_index": piu.sorted_lines_md5(ri),
                "mappings": {t: piu.ref_mappings(sc, ti, t) for t in sc["target_ids"] + ["t_unknown"]},
                "bfs_md5": hashlib.md5(bfs.tobytes()).hexdigest(),
                "set_bits": [[int(x) for x in row] for row in np.unpackbits(bfs, axis=2).sum(axis=2)],
            }
            print(name, out[name]["bfs_md5"], out[name]["set_bits"], file=sys.stderr)
    with open(os.path.join(pu.ROOT, "tests", "golden", "polish_inputs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
