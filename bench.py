#!/usr/bin/env python
"""bench.py — GoldRush-Path Gbp/s hashed+queried on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A step is one complete GoldRush-Path silver run over the synthetic read set of the workload:
bit-vector fill (K2+K4a) -> rank build (K4b) -> ordered selection loop (K2+K3+decide+K4c).
`value` times that with the decoded reads already resident in HBM (CUDA events on the engine's
stream); `e2e` times the public whole-stage call grb_run_path() on a pinned HOST FASTQ buffer
(H2D of the FASTQ, K1 decode, the step above, D2H of the decisions, host-side record assembly).
Both divide the bases that were hashed AND queried (sum of num_tiles * tile_length over the reads
the selection loop visited) by the time.  The reference arm (`--impl reference`) runs the
reference's own sources (oracle/_ref/goldrush-path-ref, built unmodified against stand-in
third-party headers) on a bounded sample of the same read set with all host cores and reports
the same quantity from the reference's own phase timers.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED22 = "1011011110110111101101"
# genome seed / shape per config; default GoldRush-Path parameters (bin/goldrush:61-78)
WORKLOADS = {
    "cfg1": dict(genome=5_000_000, cov=25.0, read_len=20000, seed=1001, phred_min=0,
                 desc="synthetic 5 Mbp genome, 25x 20 kbp reads"),
    "cfg2": dict(genome=100_000_000, cov=30.0, read_len=25000, seed=1002, phred_min=20,
                 desc="synthetic 100 Mbp genome, 30x 25 kbp reads"),
    "tiny": dict(genome=1_000_000, cov=12.0, read_len=20000, seed=7, phred_min=0,
                 desc="synthetic 1 Mbp genome, 12x 20 kbp reads (debug)"),
}
PARAMS = dict(kmer_size=22, weight=16, hash_num=3, tile_length=1000, block_size=10,
              unassigned_min=5, assigned_max=1, occupancy=0.1, threshold=10, phred_delta=5,
              ratio=0.9, max_paths=5, min_length=20000, silver_path=1)
REF_SAMPLE_READS = 1200  # bounded sample for the CPU arms (about 10-20 s of host time)


def cli_args(w, phred_min):
    p = PARAMS
    return ["-k", str(p["kmer_size"]), "-w", str(p["weight"]), "-s", SEED22, "-h", str(p["hash_num"]),
            "-t", str(p["tile_length"]), "-b", str(p["block_size"]), "-u", str(p["unassigned_min"]),
            "-a", str(p["assigned_max"]), "-o", str(p["occupancy"]), "-x", str(p["threshold"]),
            "-d", str(p["phred_delta"]), "-r", str(p["ratio"]), "-M", str(p["max_paths"]),
            "-m", str(p["min_length"]), "-P", str(phred_min), "-g", str(w["genome"]),
            "--silver_path", "--verbose"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_flags(eng, grb, phred_min, np):
    """Per-read pass-1 / pass-2 flags from the device Phred sums: the reference's filters
    (goldrush_path.cpp:261-301, 907-932) with the glibc-exact final log10 on the host."""
    meta = eng.reads_meta_array()
    n = len(meta)
    L = grb.lib()
    avg = np.zeros(n, dtype=np.uint32)
    delta = np.zeros(n, dtype=np.uint32)
    a, d = C.c_uint32(), C.c_uint32()
    fh, tot, ql = meta["phred_first_half_sum"], meta["phred_total_sum"], meta["qual_len"]
    for i in range(n):
        L.grb_phred_finalize(float(fh[i]), float(tot[i]), int(ql[i]), C.byref(a), C.byref(d))
        avg[i] = a.value
        delta[i] = d.value
    long_enough = meta["len"] >= PARAMS["min_length"]
    if phred_min == 0:  # calc_min_phred_threshold (goldrush_path.cpp:79-107), first 50000 in file order
        s = np.sort(avg[long_enough][:50000])[::-1]
        scores = np.zeros(50000, dtype=np.uint32)
        scores[:len(s)] = s
        phred_min = max(10, int(scores[min(len(s), 50000) // 2]))
    ok = long_enough & (avg >= phred_min) & (delta < PARAMS["phred_delta"]) & (meta["non_acgt"] == 0)
    flags = (ok.astype(np.uint8) * 1) | (ok.astype(np.uint8) * 2)
    return flags, meta, phred_min


def run_reference_sample(fastq_path, w, phred_min, jobs):
    """One run of the reference's own sources on a FASTQ file; returns the reference's phase
    timers and its 'Saw: N tiles' counter."""
    ref = os.path.join(ROOT, "oracle", "_ref", "goldrush-path-ref")
    kind = "reference"
    if not os.path.exists(ref):
        ref = os.path.join(ROOT, "oracle", "_build", "goldrush-path-oracle")
        kind = "port"
    out = tempfile.mkdtemp(prefix="grb_ref_")
    t0 = time.time()
    p = subprocess.run([ref] + cli_args(w, phred_min) + ["-j", str(jobs), "-i", fastq_path, "-p",
                                                         os.path.join(out, "ref")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    err = p.stderr
    tiles = [int(x) for x in re.findall(r"^Saw: (\d+) tiles", err, re.M)]
    secs = [float(x) for x in re.findall(r"^in ([0-9.]+)\s*$", err, re.M)]
    for f in os.listdir(out):
        os.remove(os.path.join(out, f))
    os.rmdir(out)
    if p.returncode != 0 or not tiles:
        raise RuntimeError("reference run failed: " + err[-2000:])
    # the reference exits inside silver_path_check after the last path (no final timer line)
    phase_s = sum(secs) if len(secs) >= 2 else wall
    return dict(kind=kind, bases=tiles[-1] * PARAMS["tile_length"], phase_s=phase_s, wall_s=wall)


def write_sample(grb, w, n_reads):
    sp = grb.api.synth_params(w["genome"], w["cov"], w["read_len"], w["seed"])
    n = min(n_reads, grb.synth_num_reads(sp))
    data = grb.synth_fastq(sp, 0, n)
    fd, path = tempfile.mkstemp(prefix="grb_sample_", suffix=".fq",
                                dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    with os.fdopen(fd, "wb") as f:
        f.write(data)
    return path, n


def bench_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import goldrush_b200 as grb  # only the host-side synthetic read generator is used here
    cores = os.cpu_count() or 1
    path, n = write_sample(grb, w, REF_SAMPLE_READS)
    phred_min = w["phred_min"]
    try:
        for _ in range(args.warmup):
            run_reference_sample(path, w, phred_min, cores)
        t, bases, kind = 0.0, 0, "reference"
        for _ in range(args.steps):
            r = run_reference_sample(path, w, phred_min, cores)
            t += r["phase_s"]
            bases += r["bases"]
            kind = r["kind"]
    finally:
        os.remove(path)
    v = bases / t / 1e9
    sample = f"first {n} reads of the {args.workload} read set, unchanged parameters, -j {cores}"
    line = {
        "impl": "reference", "metric": "GoldRush-Path Gbp/s hashed+queried", "value": v,
        "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}, default GoldRush-Path params "
                               f"(k=22 w=16 h=3 t=1000 b=10 x=10 o=0.1 -M 5 --silver_path)",
                   "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Gbp/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_ours(args, w):
    import numpy as np
    import torch
    import torch.distributed as dist
    import goldrush_b200 as grb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the one JSON line must be the only thing on stdout: NCCL prints its version banner there
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- synthetic reads (host, pinned) ----
    sp = grb.api.synth_params(w["genome"], w["cov"], w["read_len"], w["seed"])
    t_s = time.time()
    fq_ptr, fq_len = grb.synth_fastq_raw(sp)
    t_synth = time.time() - t_s
    cudart = torch.cuda.cudart()
    pinned = int(cudart.cudaHostRegister(fq_ptr, fq_len, 0)) == 0

    if world > 1:
        from goldrush_b200 import multi
        multi.init_comm(local)
    seeds = grb.make_seed_pattern(SEED22, PARAMS["kmer_size"], PARAMS["weight"], PARAMS["hash_num"])
    eng = grb.Engine(seeds, device=local, genome_size=w["genome"],
                     **{k: v for k, v in PARAMS.items() if k not in ("kmer_size", "hash_num")})
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))

    # ---- K1 once: reads resident in HBM for the `value` measurement ----
    off, chunk = 0, 1 << 30
    while off < fq_len:
        n = min(chunk, fq_len - off)
        used = eng.reads_ingest_fastq(fq_ptr + off, final=(off + n == fq_len), nbytes=n)
        if used == 0:
            break
        off += used
    flags, meta, phred_min = host_flags(eng, grb, w["phred_min"], np)
    eng.reads_set_flags(flags)
    n_reads = len(meta)
    hash_universe = grb.default_hash_universe(PARAMS["weight"], w["genome"], PARAMS["hash_num"])
    filter_bits = grb.calc_optimal_size(hash_universe, 1, PARAMS["occupancy"])
    bases_pass1 = int(meta["len"][flags & 1 != 0].sum())

    # multi-GPU: the engine was created after multi.init_comm, so the library itself
    # shards pass 1 (+ OR-reduce) and each batch's speculative query (+ all-gather) over NCCL; the
    # ordered commit is replicated.  Same call sequence at every N.
    def one_step():
        eng.filter_alloc(filter_bits)
        eng.build_bitvector()
        pop = eng.finalize_bitvector()
        dec, stats, fin = eng.select_reads_array()
        return pop, dec

    def timed(fn, n_iter):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for _ in range(n_iter):
            out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    timed(one_step, args.warmup)
    eng.profile_enable(True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t0 = time.time()
    ms_total, (pop, dec) = timed(one_step, args.steps)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    launches = eng.launch_count() - launches0
    ktime = {k: eng.kernel_time(k) for k in grb.api.KERNEL_CLASSES}
    commit_prof = eng.commit_profile()
    eng.profile_enable(False)

    if world > 1:  # the commit is replicated: every rank must hold the same decisions
        multi.assert_replicas_agree(dec.view(np.uint8))
    visited = dec["verdict"] >= 2
    bases_pass2 = int((dec["num_tiles"][visited].astype(np.int64) * PARAMS["tile_length"]).sum())
    reads_visited = int(visited.sum())
    selected = (dec["verdict"] == 2) | (dec["verdict"] == 3)
    ms_step = ms_total / args.steps
    value = bases_pass2 / (ms_step * 1e-3) / 1e9

    # roofline of the dominant kernel (k_query): 64 algorithmic bytes per probe = one 32-byte filter
    # block (bit words + running rank) + one 32-byte sector holding the {id,count} slot
    st, _, _ = eng.select_state()
    frames = 0
    T, k = PARAMS["tile_length"], PARAMS["kmer_size"]
    lens = meta["len"][visited].astype(np.int64)
    nt = lens // T
    last_len = np.minimum(T + k - 1, lens - (nt - 1) * T)
    frames = int(((nt - 1) * T + last_len - k + 1).sum())
    probes_per_step = frames * PARAMS["hash_num"]
    q_ms, q_n = ktime["query"]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = probes_per_step * args.steps * 64 / (q_ms * 1e-3) / 1e9 if q_ms > 0 else 0.0
    # second denominator: what random 32-byte sector gathers reach on this GPU over a footprint of
    # the same order (tools/sector_roofline.cu, measured on B200, profiles/sector_roofline_r01.json);
    # DRAM traffic of one k2_query launch from the committed `ncu --set full` capture
    sector_peak, traffic = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_inputs.json")) as f:
            ri = json.load(f)
        sector_peak = float(ri["random_sector_gather_gbs"])
        # measured DRAM bytes per probe of the committed capture x the probes of one launch here
        traffic = (ri["k2_query_dram_bytes_per_probe"] * probes_per_step * args.steps / max(1, q_n)
                   if args.workload == "cfg2" else None)
    except (OSError, KeyError, ValueError):
        pass
    # third denominator: the filter's own probe sequence without hashing / voting, measured over
    # footprints by tools/probe_bench.py (cfg5, profiles/probe_bench_r01.jsonl)
    probe_peak = None
    try:
        footprint_gb = ((filter_bits + 191) // 192 * 32 + (int(pop) + 1) * 16) / 1e9
        rows = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "probe_bench_r01.jsonl"))]
        rows = [r for r in rows if r.get("h") == PARAMS["hash_num"] and "query_gprobes_per_s" in r]
        near = min(rows, key=lambda r: abs(r["footprint_gb"] - footprint_gb))
        probe_peak = {"gprobes_per_s": near["query_gprobes_per_s"], "at_footprint_gb": near["footprint_gb"],
                      "footprint_gb": round(footprint_gb, 2)}
    except (OSError, ValueError, KeyError):
        pass
    gprobes = probes_per_step * args.steps / (q_ms * 1e-3) / 1e9 if q_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k2_query", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "bytes_per_probe": 64, "probes_per_step": probes_per_step,
                "launches_per_step": q_n // max(1, args.steps),
                "avg_launch_us": 1e3 * q_ms / max(1, q_n), "traffic": traffic,
                "random_sector_peak": sector_peak,
                "frac_of_random_sector_peak": achieved / sector_peak if sector_peak else None,
                "gprobes_per_s": gprobes, "probe_microbench": probe_peak,
                "frac_of_probe_microbench": (gprobes / probe_peak["gprobes_per_s"]) if probe_peak else None}
    del eng

    # ---- e2e: the public whole-stage call on the pinned host FASTQ ----
    e2e_steps = max(1, min(args.steps, 3))
    res = None
    # one untimed call first: CUDA module load, and the library's device-allocation cache takes
    # over the blocks of the engine above (a resident service pays cudaMalloc once, not per run)
    if args.warmup > 0:
        grb.run_path(fq_ptr, nbytes=fq_len, input_path="(memory)", seed_preset=SEED22,
                     write_outputs=False, quiet=True, device=local, genome_size=w["genome"],
                     phred_min=w["phred_min"], **PARAMS)
    barrier()
    torch.cuda.synchronize()
    t_e0 = time.time()
    for _ in range(e2e_steps):
        res = grb.run_path(fq_ptr, nbytes=fq_len, input_path="(memory)", seed_preset=SEED22,
                           write_outputs=False, quiet=True, device=local, genome_size=w["genome"],
                           phred_min=w["phred_min"], **PARAMS)
    torch.cuda.synchronize()
    t_e = torch.tensor([(time.time() - t_e0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    assert res.bases_pass2 == bases_pass2, (res.bases_pass2, bases_pass2)
    e2e = {"value": res.bases_pass2 / e2e_s / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": fq_len,
           "d2h_bytes_per_step": int(n_reads * (24 + 56)), "s_per_step": e2e_s,
           "phases_ms": {"ingest": res.ms_ingest, "pass1": res.ms_pass1, "rank": res.ms_rank,
                         "pass2": res.ms_pass2, "wall": res.ms_wall},
           "pinned": pinned}
    if pinned:
        cudart.cudaHostUnregister(fq_ptr)
    grb.free_host(fq_ptr)

    if rank == 0:
        cpu = None
        if world == 1 and not os.environ.get("GRB_BENCH_SKIP_CPU"):  # A/B runs skip the CPU leg
            cores = os.cpu_count() or 1
            path, n = write_sample(grb, w, REF_SAMPLE_READS)
            try:
                r = run_reference_sample(path, w, w["phred_min"], cores)
            finally:
                os.remove(path)
            cpu = {"value": r["bases"] / r["phase_s"] / 1e9, "unit": "Gbp/s", "cores": cores,
                   "kind": r["kind"],
                   "sample": f"first {n} reads of the {args.workload} read set, unchanged "
                             f"parameters, -j {cores}; {r['phase_s']:.1f} s of reference phase timers"}
        line = {
            "metric": "GoldRush-Path Gbp/s hashed+queried", "value": value, "unit": "Gbp/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}, default GoldRush-Path params "
                                   f"(k=22 w=16 h=3 t=1000 b=10 x=10 o=0.1 -M 5 --silver_path, "
                                   f"-P {phred_min})",
                       "reads": n_reads, "reads_visited": reads_visited,
                       "reads_selected": int(selected.sum()), "bases_pass1": bases_pass1,
                       "bases_pass2": bases_pass2, "filter_bits": int(filter_bits), "pop": int(pop),
                       "l2": "inputs larger than L2 (filter blocks + ID slots + packed reads)",
                       "parallelism": ("one GPU" if world == 1 else
                                       f"filter replicated on {world} GPUs; pass-1 reads and each "
                                       f"batch's query tiles sharded, NCCL OR-reduce / all-gather; "
                                       f"ordered commit replicated (decisions checked equal)"),
                       "synth_s": round(t_synth, 1)},
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in ktime.items()},
            "commit_profile_last_step": commit_prof,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        sys.stdout.flush()
        os.dup2(json_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        grb.api.comm_destroy()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    # torchrun exports OMP_NUM_THREADS=1; the synthetic read generator and the reference arm are
    # OpenMP code: give every rank its share of the cores instead (before libgomp is loaded)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    if args.impl == "reference":
        bench_reference(args, w)
    else:
        bench_ours(args, w)


if __name__ == "__main__":
    main()
