#!/usr/bin/env python
"""cfg5 (BASELINE.json configs[4]): miBF probe microbenchmark — query and insert probe throughput
against filter footprint (1 .. 150 GB of filter blocks + 8-byte ID slots) and seed patterns h = 1..5,
through the C ABI (grb_probe_bench), 2^28 probes per launch, best of 3.  One JSON object per line.

usage: python tools/probe_bench.py [--footprints 1,2,4,...,150] [--h 1,2,3,4,5] [--fill 0.47]

fill 0.47 is the share of set bits the cfg2 run ends pass 1 with (pop / filter_bits)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import goldrush_b200 as grb  # noqa: E402

SEED22 = "1011011110110111101101"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--footprints", default="1,2,4,8,16,32,64,96,128,150")
    ap.add_argument("--h", default="1,2,3,4,5")
    ap.add_argument("--fill", type=float, default=0.47)
    ap.add_argument("--probes", type=int, default=1 << 28)
    args = ap.parse_args()
    seeds = grb.make_seed_pattern(SEED22, 22, 16, 3)
    for gb in [float(x) for x in args.footprints.split(",")]:
        # footprint = bits / 6 (32-byte blocks of 192 bits) + 8 B per set bit ({id, count} slot)
        bits = int(gb * 1e9 / (1.0 / 6.0 + 8.0 * args.fill))
        bits += 64 - bits % 64
        for h in [int(x) for x in args.h.split(",")]:
            with grb.Engine(seeds, genome_size=1000000, weight=16) as e:
                try:
                    r = e.probe_bench(bits, args.fill, h, n_probes=args.probes)
                except grb.GrbError as err:
                    print(json.dumps({"footprint_gb_target": gb, "h": h, "error": str(err)}), flush=True)
                    continue
            q = r.probes / (r.query_ms * 1e-3)
            i = r.probes / (r.insert_ms * 1e-3)
            print(json.dumps({
                "footprint_gb_target": gb, "footprint_gb": round(r.footprint_bytes / 1e9, 3), "h": h,
                "filter_bits": r.filter_bits, "pop": r.pop, "fill": round(r.pop / r.filter_bits, 4),
                "probes": r.probes, "query_ms": round(r.query_ms, 4), "insert_ms": round(r.insert_ms, 4),
                "query_gprobes_per_s": round(q / 1e9, 3), "insert_gprobes_per_s": round(i / 1e9, 3),
                "query_algorithmic_gbs": round(q * 64 / 1e9, 1),    # 64 B per probe
                "insert_algorithmic_gbs": round(i * 96 / 1e9, 1),   # 32 B block + 32 B slot read + 32 B write
                # DRAM accesses per second: a query probe is 2 random sector reads, an insert probe 2
                # reads + 1 write-back (tools/sector_roofline.cu: ~36 G random accesses/s on this GPU)
                "query_gaccess_per_s": round(2 * q / 1e9, 2), "insert_gaccess_per_s": round(3 * i / 1e9, 2),
                "line_query_gprobes_per_s": round(r.probes / (r.line_query_ms * 1e-3) / 1e9, 3)
                if r.line_query_ms > 0 else None,
                "line_gb": round(r.line_bytes / 1e9, 3),
                "checksum": r.checksum, "probes_missed": r.probes_missed}), flush=True)
            grb.lib().grb_release_cached_memory()


if __name__ == "__main__":
    main()
