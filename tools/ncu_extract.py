#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): per captured launch the duration, DRAM
bytes, DRAM throughput %, registers, grid -- written as profiles/<tag>_ncu_full.json -- and the
per-launch DRAM traffic of the two streaming kernels merged into profiles/traffic.json, which
bench.py reports as roofline.traffic.

    python tools/ncu_extract.py gpurun_out/prof_r1d.ncu-rep r1d --n 268435456 --mvec 10 --gpus 1
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12,
         "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("tag")
    ap.add_argument("--n", type=int, default=1 << 28)
    ap.add_argument("--mvec", type=int, default=10)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "")}
        for metric, key in WANT.items():
            if metric in idx:
                val = float(r[idx[metric]].replace(",", ""))
                val *= SCALE.get(units[idx[metric]], 1.0)
                rec[key] = val
        rec["dram_bytes"] = rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
        rec["dram_tbs"] = rec["dram_bytes"] / rec["duration"] / 1e12
        out.append(rec)
    dst = os.path.join(ROOT, "profiles", "%s_ncu_full.json" % a.tag)
    with open(dst, "w") as fh:
        json.dump({"report": os.path.basename(a.report), "n": a.n, "mvec": a.mvec, "gpus": a.gpus,
                   "note": "ncu --set full --clock-control none; per-launch values (cold cache, serialised replays)",
                   "launches": out}, fh, indent=1)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = {}
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh)
    key = "n%d_m%d_g%d" % (a.n, a.mvec, a.gpus)
    ent = traffic.setdefault(key, {})
    for name in ("nka_pass_a", "nka_pass_b"):
        vals = [r["dram_bytes"] for r in out if r["kernel"].startswith(name)]
        if vals:
            ent[name.replace("nka_", "")] = sum(vals) / len(vals)
    ent["source"] = os.path.basename(dst)
    with open(tpath, "w") as fh:
        json.dump(traffic, fh, indent=1)
    for r in out:
        print("%-22s %.3f ms  dram %.2f GB (R %.2f W %.2f)  %.2f TB/s  %.1f%% of peak  regs %d grid %d x %d"
              % (r["kernel"], r["duration"] * 1e3, r["dram_bytes"] / 1e9, r.get("dram_read", 0) / 1e9,
                 r.get("dram_write", 0) / 1e9, r["dram_tbs"], r.get("dram_pct_of_peak", 0), r.get("registers", 0),
                 r.get("grid", 0), r.get("block", 0)))


if __name__ == "__main__":
    main()
