import os, sys, time, subprocess, hashlib
sys.path.insert(0, "/root/repo")
import bench, goldrush_b200 as grb
wname = sys.argv[1]
w = bench.WORKLOADS[wname]
sp = grb.api.synth_params(w["genome"], w["cov"], w["read_len"], w["seed"])
ptr, n = grb.synth_fastq_raw(sp)
out = {}
for mode in ("batch", "serial"):
    os.environ["GRB_ENGINE"] = mode
    t = time.time()
    res = grb.run_path(ptr, nbytes=n, input_path="(memory)", seed_preset=bench.SEED22, write_outputs=False,
                       quiet=True, genome_size=w["genome"], phred_min=w["phred_min"], **bench.PARAMS)
    out[mode] = (res.out_digest, res.reads_visited, res.reads_selected, res.bases_selected)
    print(mode, out[mode], f"pass2 {res.ms_pass2:.1f} ms wall {time.time()-t:.2f}s", flush=True)
assert out["batch"] == out["serial"]
print("batch == serial on", wname)
