#!/usr/bin/env python
"""bench.py -- accel_update throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is ONE accel_update in steady state (subspace full, one eviction per
call) on synthetic fp64 vectors, n = 2^28, mvec = 10, vtol = 0.01
(BASELINE.json configs[2], the configuration the metric is quoted on).  With
N > 1 (launched by torchrun, one process per GPU) every vector is split into N
contiguous slabs (strong scaling: n is fixed) and the only exchange is one
66-double NCCL all-reduce per update.

Prints one JSON line (rank 0).  `value` is device-resident throughput;
`e2e` is the same metric through the reference-facing C entry point
nka_accel_update(NKA, double*) with HOST (pinned) buffers, so each step pays
the host->device and device->host copy of f.  `cpu_baseline` / `--impl
reference` time the reference's own serial C implementation (oracle/_ref,
compiled from /root/reference; else the oracle port) on a bounded slice of the
same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_FULL = 1 << 28
MVEC = 10
VTOL = 0.01


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL, help="global vector length (default 2^28)")
    ap.add_argument("--mvec", type=int, default=MVEC)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def algorithmic_bytes(n: int, m: int) -> int:
    """reads M w + M v + f, writes f_out + new w + new v."""
    return (2 * m + 4) * n * 8


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax.append(mx)
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                for nm, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# the reference arm / cpu baseline: the reference's own serial C accel_update
# --------------------------------------------------------------------------
def cpu_reference_run(n_full: int, mvec: int, steps: int, warmup: int, budget_s: float):
    """Time the reference's C accel_update (1 thread: the reference is serial) on a slice of
    the workload small enough to finish in ~budget_s; returns updates/s extrapolated
    linearly to n_full, with a description of the sample."""
    import numpy as np
    from oracle import api
    calls = steps + warmup + mvec + 2
    # ~1.0 s per update at n = 2^24, mvec = 10 on this class of host
    per_elem = 1.0 / (1 << 24) * (2 * mvec + 4) / 24.0
    n_s = 1 << 16
    while n_s * 2 <= min(n_full, 1 << 24) and (n_s * 2) * per_elem * calls <= budget_s:
        n_s *= 2
    kind = "reference" if api.ref_lib() is not None else "port"
    acc = api.RefNKA(n_s, mvec, VTOL) if kind == "reference" else api.OracleNKA(n_s, mvec, VTOL)
    rng = np.random.default_rng(1234)
    pool = [rng.uniform(-0.5, 0.5, n_s) for _ in range(mvec + 3)]
    k = 0
    for _ in range(mvec + 2 + warmup):
        acc.accel_update(pool[k % len(pool)])
        k += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        acc.accel_update(pool[k % len(pool)])
        k += 1
    dt = time.perf_counter() - t0
    nvec = acc.num_vec()
    acc.close()
    ups_sample = steps / dt
    value = ups_sample * (n_s / n_full)
    sample = ("%s C accel_update (gcc -O3, serial, default dp) timed at n=2^%d, mvec=%d, %d steady-state calls "
              "(num_vec=%d); updates/s scaled linearly by n_sample/n to n=2^%d"
              % ("reference src-C" if kind == "reference" else "oracle port of the reference",
                 n_s.bit_length() - 1, mvec, steps, nvec, n_full.bit_length() - 1))
    return {"value": value, "unit": "updates/s", "cores": 1, "kind": kind, "sample": sample,
            "host_cores": os.cpu_count(), "ms_per_update_sample": 1e3 * dt / steps, "n_sample": n_s}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(args.n, args.mvec, args.steps, args.warmup, budget_s=90.0)
    line = {
        "impl": "reference",
        "metric": "accel_update/sec", "value": res["value"], "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_local=None),
        "hbm_gbs": res["value"] * algorithmic_bytes(args.n, args.mvec) / 1e9,
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_local, comm_mode=None):
    cfg = {
        "workload": "synthetic accel_update microbench (BASELINE.json configs[2]): n=2^%d fp64, mvec=%d, vtol=%g, "
                    "steady state (subspace full, one eviction per call), f_t i.i.d. uniform(-0.5,0.5)"
                    % (args.n.bit_length() - 1, args.mvec, VTOL),
        "n": args.n, "mvec": args.mvec, "vtol": VTOL,
        "l2": "inputs larger than L2: every column is %.0f MiB per GPU, no flush needed"
              % ((n_local or args.n) * 8 / 2 ** 20),
        "parallelism": ("row slabs, %d GPU(s), %s" % (args.gpus, {
            "peer": "partial dot products summed across ranks inside pass A through NVLink peer memory "
                    "(no collective launch)",
            "nccl": "one 66-double NCCL all-reduce per update"}.get(comm_mode, "one small sum-allreduce per update")))
                       if args.gpus > 1 else "single GPU",
    }
    if n_local is not None:
        cfg["n_local"] = n_local
    return cfg


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d processes (one per GPU)" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nka_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from nka_b200 import NKA
    from nka_b200.distributed import distributed_nka

    n, m = args.n, args.mvec
    stream = torch.cuda.Stream()
    if world > 1:
        acc, lo, hi = distributed_nka(n, m, VTOL, device=local_rank, stream=stream.cuda_stream)
    else:
        acc, lo, hi = NKA(n, m, VTOL, device=local_rank, stream=stream.cuda_stream), 0, n
    n_local = hi - lo

    # synthetic inputs, resident in HBM: a pool of mvec+3 independent vectors used round-robin.
    # accel_update overwrites f with f + (a tiny projection), so a revisited buffer is again an
    # i.i.d.-like vector unrelated to the subspace, which by then has evicted everything built from it.
    pool_n = m + 3
    gen = torch.Generator(device="cuda").manual_seed(1000 + rank)
    pool = [torch.rand(n_local, dtype=torch.float64, device="cuda", generator=gen) - 0.5 for _ in range(pool_n)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k = 0
    with torch.cuda.stream(stream):
        for _ in range(m + 2):                      # reach steady state (not counted as warm-up)
            acc.accel_update(pool[k % pool_n]); k += 1
        for _ in range(max(args.warmup, 0)):
            acc.accel_update(pool[k % pool_n]); k += 1
    barrier()
    assert acc.num_vec() == m, "not in steady state: num_vec=%d" % acc.num_vec()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    acc.timing_enable(True)
    acc.timing_reset()
    launches0 = acc.launch_count()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            acc.accel_update(pool[k % pool_n]); k += 1
        ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = acc.launch_count() - launches0
    kt = acc.timing_read()
    acc.timing_enable(False)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    nvec_end = acc.num_vec()
    st = acc.state()
    geom = acc.launch_geometry()

    # ---- e2e: the reference-facing entry point with host buffers ----
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(n_local, dtype=torch.float64, pin_memory=True) for _ in range(3)]
        for hb, src in zip(host, pool):
            hb.copy_(src)
        torch.cuda.synchronize()
        from nka_b200 import _lib
        lib = _lib.load()
        h = acc._handle()
        j = 0
        for _ in range(2):
            lib.nka_accel_update(h, host[j % 3].data_ptr()); j += 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            lib.nka_accel_update(h, host[j % 3].data_ptr()); j += 1     # H2D + kernels + D2H, synchronous
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": args.e2e_steps / dt, "unit": "updates/s",
               "h2d_bytes_per_step": n_local * 8 * world, "d2h_bytes_per_step": n_local * 8 * world,
               "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "api": "nka_accel_update(NKA, double* host_f) -- include/nonlinear_krylov_accelerator.h, pinned host f"}
        del host

    if rank == 0:
        sampler.stop()

    if rank == 0:
        peak, peak_kind = measured_peak()
        ups = args.steps / (elapsed_ms * 1e-3)
        algo = algorithmic_bytes(n, m)
        gbs = ups * algo / 1e9
        # dominant kernel = pass B (f + M Z-column reads, 3 column writes).  Its share of the
        # algorithmic bytes: the M "v" columns + the three writes (f_out, new w, new v) = (M+3) n 8;
        # pass A's share: the M "w" columns + f = (M+1) n 8 (it streams exactly that: lazy last column).
        kb = kt["pass_b"]
        ka = kt["pass_a"]
        algo_b = (m + 3) * n_local * 8
        algo_a = (m + 1) * n_local * 8
        ms_b = kb["ms"] / max(kb["count"], 1)
        ms_a = ka["ms"] / max(ka["count"], 1)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as fh:
                    tj = json.load(fh)
                key = "n%d_m%d_g%d" % (n, m, world)
                traffic = tj.get(key, {}).get("pass_b")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": "nka_pass_b<%d,2>" % m,
                    "achieved": algo_b / (ms_b * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": algo_b / (ms_b * 1e-3) / 1e9 / peak, "traffic": traffic,
                    "peak_kind": "of " + peak_kind, "algorithmic_bytes_per_launch": algo_b,
                    "avg_launch_ms": ms_b, "launches_timed": kb["count"]}
        line = {
            "metric": "accel_update/sec", "value": ups, "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n_local, acc.comm_mode()),
            "comm_mode": acc.comm_mode(),
            "hbm_gbs": gbs, "roofline_frac_update": gbs / (peak * world),
            "roofline_update": {"algorithmic_bytes": algo, "formula": "(2M+4)*n*8", "achieved_gbs": gbs,
                                "peak_gbs": peak * world, "frac": gbs / (peak * world),
                                "peak_kind": "of " + peak_kind},
            "clocks": sampler.summary(t_wall0, t_wall1),
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "kernels": {
                "pass_a": {"avg_ms": ms_a, "algorithmic_gbs": algo_a / (ms_a * 1e-3) / 1e9,
                           "frac": algo_a / (ms_a * 1e-3) / 1e9 / peak, "count": ka["count"]},
                "pass_b": {"avg_ms": ms_b, "algorithmic_gbs": algo_b / (ms_b * 1e-3) / 1e9,
                           "frac": algo_b / (ms_b * 1e-3) / 1e9 / peak, "count": kb["count"]},
                "state": {"avg_ms": kt["state"]["ms"] / max(kt["state"]["count"], 1)},
                "materialise": {"avg_ms": kt["materialise"]["ms"] / max(kt["materialise"]["count"], 1)},
                "allreduce": {"avg_ms": kt["allreduce"]["ms"] / max(kt["allreduce"]["count"], 1)},
                "geometry": geom,
            },
            "num_vec": nvec_end, "state_error": st["error"],
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_run(n, m, steps=5, warmup=0, budget_s=20.0)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    acc.delete()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
