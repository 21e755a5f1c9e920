"""One rank of the multi-GPU parity run (launched by tests/test_gpu_multi.py through torchrun, one
process per GPU).  Every rank runs the same golden cases through grb_run_path with the library's NCCL
communicator in place (pass 1 sharded + OR-reduced, each batch's query sharded + all-gathered,
commit replicated) and writes the digests of ITS OWN output files; the test compares every rank's
digests with the reference fixtures.  Also checks, at the engine level, that the OR-reduced bit
vector and the decisions equal those of an unsharded context on the same GPU."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import goldrush_b200 as grb  # noqa: E402
from goldrush_b200 import multi  # noqa: E402
import parity_util as pu  # noqa: E402
from test_gpu_parity import _args_to_params, SEED22  # noqa: E402

CASES = ["silver_default", "small_tiles_random_seed", "golden_lognormal_all_lengths", "ragged_tail_h4"]


def main():
    out_dir = sys.argv[1]
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # only carries the NCCL unique id and the barriers
    result = {"rank": rank, "world": world, "cases": {}}

    # ---- unsharded reference on this GPU, before the communicator exists ----
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    sp = grb.api.synth_params(300000, 12.0, 5000, 77)
    data = grb.synth_fastq(sp)
    kw = dict(genome_size=300000, weight=16, tile_length=250, min_length=5000, silver_path=1,
              max_paths=3, ratio=0.9, device=local)

    def engine_run():
        with grb.Engine(seeds, **kw) as e:
            e.profile_enable(True)
            e.reads_ingest_fastq(data)
            n = e.reads_count()
            e.reads_set_flags(np.full(n, 3, dtype=np.uint8))
            e.filter_alloc(grb.calc_optimal_size(grb.default_hash_universe(16, 300000, 3), 1, 0.1))
            e.build_bitvector()
            bits = e.copy_bitvector()
            pop = e.finalize_bitvector()
            dec, stats, fin = e.select_reads_array()
            return bits, pop, dec.copy(), grb.api.comm_info(e), e.kernel_time("gather")[1]

    bits1, pop1, dec1, info1, _ = engine_run()
    assert info1 == (0, 1)

    assert multi.init_comm(local) == (rank, world)
    assert grb.api.comm_info() == (rank, world)
    # default: pass 1 sharded + OR-reduced, pass 2 replicated
    bitsw, popw, decw, infow, n_gather0 = engine_run()
    assert infow == (rank, world)
    # GRB_SHARD_QUERY=1: each batch's speculative query sharded over tiles + all-gathered as well
    os.environ["GRB_SHARD_QUERY"] = "1"
    bitsq, popq, decq, _, n_gather1 = engine_run()
    assert n_gather1 > n_gather0 >= 0
    result["engine"] = {"bits_equal": bool(np.array_equal(bits1, bitsw) and np.array_equal(bits1, bitsq)),
                        "pop_equal": pop1 == popw == popq,
                        "dec_equal": bool(np.array_equal(dec1.view(np.uint8), decw.view(np.uint8)) and
                                          np.array_equal(dec1.view(np.uint8), decq.view(np.uint8))),
                        "selected": int(((decw["verdict"] == 2) | (decw["verdict"] == 3)).sum())}
    multi.assert_replicas_agree(decw.view(np.uint8))
    multi.assert_replicas_agree(decq.view(np.uint8))

    # ---- golden cases through the whole-stage call ----
    work = os.path.join(out_dir, f"rank{rank}")
    os.makedirs(work, exist_ok=True)
    produced = {}

    def outputs_for(case, workdir):
        if case["name"] not in produced:
            inp, extra = pu.make_input(case, workdir, outputs_for)
            with open(inp, "rb") as f:
                fq = f.read()
            prefix = os.path.join(workdir, case["name"] + ".gpu")
            res = grb.run_path(fq, input_path=inp, prefix=prefix, quiet=True, device=local,
                               **_args_to_params(case["args"] + extra))
            outs = sorted((os.path.join(workdir, fn) for fn in os.listdir(workdir)
                           if fn.startswith(os.path.basename(prefix))), key=lambda p: (len(p), p))
            produced[case["name"]] = (outs, res)
        return produced[case["name"]][0]

    for i, name in enumerate(CASES):
        os.environ["GRB_SHARD_QUERY"] = "1" if i % 2 == 0 else "0"
        outs = outputs_for(pu.case_by_name(name), work)
        res = produced[name][1]
        result["cases"][name] = {"outputs": pu.digest_outputs(outs), "filter_bits": res.filter_bits,
                                 "num_passed_reads": res.num_passed_reads, "launches": res.launches}
        dist.barrier()

    # ---- both launches of one assembly in one call, every rank holding the whole input ----
    os.environ.pop("GRB_SHARD_QUERY", None)
    silver, golden = pu.case_by_name("silver_default"), pu.case_by_name("golden_default")
    inp, _ = pu.make_input(silver, work, outputs_for)
    with open(inp, "rb") as f:
        fq = f.read()
    ps, pg = os.path.join(work, "two.silver"), os.path.join(work, "two.golden")
    rs, rg = grb.run_two_stage(fq, dict(prefix=ps, **_args_to_params(silver["args"])),
                               dict(prefix=pg, **_args_to_params(golden["args"])), input_path=inp,
                               device=local)
    s_outs = sorted((os.path.join(work, fn) for fn in os.listdir(work) if fn.startswith("two.silver")),
                    key=lambda p: (len(p), p))
    result["two_stage"] = {"silver": pu.digest_outputs(s_outs),
                           "golden": pu.digest_outputs([pg + ".fa"])}
    dist.barrier()

    # ---- slice mode: this rank is handed only its own records of the input (cut at a record
    # boundary); no output files, the record digest is put together from the ranks' parts ----
    starts = [0]
    pos, line = 0, 0
    while True:
        nl = fq.find(b"\n", pos)
        if nl < 0:
            break
        pos = nl + 1
        line += 1
        if line % 4 == 0 and pos < len(fq):
            starts.append(pos)
    n_rec = len(starts)
    cut = [starts[n_rec * r // world] if r < world else len(fq) for r in range(world + 1)]
    mine = fq[cut[rank]:cut[rank + 1]]
    kw_s = dict(_args_to_params(silver["args"]), fastq_offset=cut[rank], fastq_total=len(fq))
    kw_g = _args_to_params(golden["args"])
    r1 = grb.run_path(mine, input_path="(slice)", write_outputs=False, device=local, **kw_s)
    ss, sg = grb.run_two_stage(mine, kw_s, kw_g, input_path="(slice)", device=local,
                               write_outputs=False)
    result["slice"] = {"bytes": len(mine), "digest": r1.out_digest, "expect": rs.out_digest,
                       "selected": r1.reads_selected, "expect_selected": rs.reads_selected,
                       "two_silver": ss.out_digest, "two_golden": sg.out_digest,
                       "expect_golden": rg.out_digest, "golden_reads": sg.num_reads,
                       "expect_golden_reads": rg.num_reads}
    dist.barrier()

    grb.api.comm_destroy()
    with open(os.path.join(out_dir, f"result{rank}.json"), "w") as f:
        json.dump(result, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
