"""Two-GPU parity (-m gpu; skipped on a single-GPU box): every rank owns a row slab, the only
exchange is the sum of the partial dot products -- fused into pass A through peer memory
("peer", the default on one NVLink domain) or one NCCL all-reduce ("nccl", NKA_PEER_REDUCE=0)
-- and the joined result equals the serial reference on the full vector with identical
decisions on both ranks."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, mode, q):
    os.environ["NKA_PEER_REDUCE"] = "1" if mode == "peer" else "0"
    os.environ["NKA_PEER_TIMEOUT_S"] = "30"
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenarios as S
    from nka_b200.distributed import distributed_nka
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, mvec, vtol, mk = S.SCENARIOS[name]
        if mode == "dp":
            # the reference's own mechanism for a parallel run (src-F08-vector/README.md:16-22): every
            # process passes its portion of the vector and a dot product that sums over the processes
            from nka_b200 import NKA
            from nka_b200.distributed import slab_bounds
            from nka_b200.nka import DP_FUNC
            gloo = dist.new_group(backend="gloo")
            lo, hi = slab_bounds(n, world, rank)

            def dp(m, x, y):
                t = torch.tensor([x[0] * y[0]], dtype=torch.float64)
                dist.all_reduce(t, group=gloo)
                return float(t[0])
            cb = DP_FUNC(dp)
            acc = NKA(hi - lo, mvec, vtol, device=rank)
            acc.set_dot_prod(cb)
        else:
            acc, lo, hi = distributed_nka(n, mvec, vtol, device=rank)
        outs, nvec, decisions = [], [], []
        comm_mode = acc.comm_mode()
        for op in mk():
            if op[0] == "update":
                d = torch.from_numpy(np.ascontiguousarray(op[1][lo:hi])).cuda()
                acc.accel_update(d)
                outs.append(d.cpu().numpy())
                st = acc.state()
                decisions.append((st["ndrop_last"], st["relaxed_last"], st["evicted_last"], st["error"]))
            elif op[0] == "relax":
                acc.relax()
            else:
                acc.restart()
            nvec.append(acc.num_vec())
        q.put((rank, {"outs": outs, "nvec": nvec, "decisions": decisions, "bounds": (lo, hi),
                      "comm_mode": comm_mode}))
        acc.delete()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["iid_n1000_m10", "picard_n500_m5_v2", "mixed_n257_m5", "relax_restart_n96_m4",
                                  "n3_m5_rankdef"])
@pytest.mark.parametrize("mode", ["peer", "nccl", "dp"])
def test_two_gpu_slabs_match_serial_oracle(name, mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenarios as S
    from oracle import api
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, mvec, vtol, mk = S.SCENARIOS[name]
    ops = mk()
    inputs = [op[1] for op in ops if op[0] == "update"]
    serial, nv_ref = S.run_ops(api.OracleNKA(n, mvec, vtol, dotmode=0), ops)
    arbiter, _ = S.run_ops(api.OracleNKA(n, mvec, vtol, dotmode=1), ops)
    scales, tols = S.tolerances(serial, arbiter, inputs)
    assert got[0]["comm_mode"] == got[1]["comm_mode"] == ("single" if mode == "dp" else mode)
    assert got[0]["nvec"] == got[1]["nvec"] == nv_ref
    assert got[0]["decisions"] == got[1]["decisions"]
    assert all(d[3] == 0 for d in got[0]["decisions"])
    for t in range(len(inputs)):
        joined = np.concatenate([got[0]["outs"][t], got[1]["outs"][t]])
        assert np.linalg.norm(joined - arbiter[t]) / scales[t] <= tols[t], (name, t)


# ---------------------------------------------------------------------------------------------
# the example on row slabs (BASELINE.json configs[3] in miniature): halo rows for the residual,
# global norm, slab-pipelined exact-order SSOR, accelerator over all slabs
# ---------------------------------------------------------------------------------------------
def _example_worker(rank, world, port, case, q):
    os.environ["NKA_PEER_TIMEOUT_S"] = "30"
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from nka_b200.example import distributed_system, Solver, FIELD_U, FIELD_R, FIELD_Z
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = {}
        if case["kind"] == "kernels":
            nx, ny = case["nx"], case["ny"]
            sy = distributed_system(0.02, nx, ny, scaling=1, device=rank)
            rng = np.random.default_rng(42)
            u = rng.uniform(0.0, 0.3, (ny, nx))
            z = rng.uniform(-0.01, 0.01, (ny, nx))
            sy.set(FIELD_U, u[sy.k0:sy.k1])
            sy.set(FIELD_Z, z[sy.k0:sy.k1])
            out["rnorm"] = sy.residual(subtract_z=True)
            out["u"] = sy.get(FIELD_U)
            out["r"] = sy.get(FIELD_R)
            sy.pc_ssor(case["nsweep"], 1.4)
            out["z"] = sy.get(FIELD_Z)
            sy.pc_ssor(1, 1.4)                      # a second application: tags keep counting
            out["z2"] = sy.get(FIELD_Z)
            out["rows"] = (sy.k0, sy.k1)
            sy.delete()
        elif case["kind"] == "steps":
            # the Picard loop one kernel at a time, every field of every iteration saved for the parent
            from nka_b200 import _lib
            lib = _lib.load()
            nx, ny = case["nx"], case["ny"]
            sy = distributed_system(0.02, nx, ny, scaling=1, device=rank)
            so = Solver(sy, nsweep=2, omega=1.4, mvec=case["mvec"], vtol=0.01)
            save = lambda tag, it, a: np.save(os.path.join(case["dir"], "%s_%d_r%d.npy" % (tag, it, rank)), a)
            out["rnorm"] = [sy.residual(subtract_z=False)]
            save("r", 0, sy.get(FIELD_R))
            out["nvec"] = []
            for it in range(1, case["iters"] + 1):
                sy.pc_ssor(2, 1.4)
                save("zssor", it, sy.get(FIELD_Z))
                lib.nka_accel_update_dev(so.accel._handle(), sy.field_ptr(FIELD_Z))
                out["nvec"].append(so.accel.num_vec())
                save("zacc", it, sy.get(FIELD_Z))
                out["rnorm"].append(sy.residual(subtract_z=True))
                save("u", it, sy.get(FIELD_U))
                save("r", it, sy.get(FIELD_R))
            out["rows"] = (sy.k0, sy.k1)
            out["mode"] = so.accel.comm_mode()
            so.delete(); sy.delete()
        else:
            sy = distributed_system(0.02, case["nx"], case["ny"], scaling=case["scaling"], device=rank)
            so = Solver(sy, nsweep=2, omega=1.4, mvec=case["mvec"], vtol=0.01)
            res = so.solve(maxitr=case.get("maxitr"), record_nvec=case["mvec"] > 0)
            out = {"iters": res["iters"], "rnorm": res["rnorm"], "nvec": res.get("nvec"),
                   "mode": so.accel.comm_mode() if so.accel else None}
            so.delete(); sy.delete()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def _run_example(case, world=2):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_example_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=case.get("timeout", 240)) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


@pytest.mark.parametrize("nx,ny,nsweep", [(70, 64, 2), (33, 9, 1), (200, 131, 3)])
def test_slab_residual_and_ssor_bit_identical_to_serial_loops(nx, ny, nsweep):
    """Two row slabs: u - z, residual and the SSOR sweeps equal the CPU loops on the whole grid bit
    for bit (the sweep crosses the slab boundary strip by strip through NVLink peer memory)."""
    from oracle import api
    got = _run_example({"kind": "kernels", "nx": nx, "ny": ny, "nsweep": nsweep})
    rng = np.random.default_rng(42)
    u = rng.uniform(0.0, 0.3, (ny, nx))
    z = rng.uniform(-0.01, 0.01, (ny, nx))
    orc = api.OracleSystem(nx, ny, 0.02, 1)
    unew = u - z
    pad = np.zeros((ny + 2, nx + 2)); pad[1:-1, 1:-1] = unew
    r = orc.residual(pad).reshape(ny, nx)
    zz = orc.pc_ssor(nsweep, 1.4, r.ravel().copy()).reshape(ny, nx)
    zz2 = orc.pc_ssor(1, 1.4, r.ravel().copy()).reshape(ny, nx)
    assert got[0]["rows"][1] == got[1]["rows"][0]
    for key, want in (("u", unew), ("r", r), ("z", zz), ("z2", zz2)):
        joined = np.concatenate([got[0][key], got[1][key]], axis=0)
        assert np.array_equal(joined, want), key
    assert got[0]["rnorm"] == got[1]["rnorm"]
    assert abs(got[0]["rnorm"] - orc.norm2(r.ravel())) <= 1e-13 * got[0]["rnorm"]


def test_slab_example_reproduces_golden_tables():
    """The reference's pinned run (src-F95/reference_output) on two GPUs: 26 accelerated iterations,
    every printed line equal; 367 unaccelerated."""
    from oracle import api
    with open(os.path.join(ROOT, "tests", "golden", "example_c_f95.txt")) as fh:
        acc_txt, unacc_txt = fh.read().split("UNACCELERATED SOLVE")
    lines = lambda text: [ln for ln in text.splitlines() if ":" in ln[:4] and ln[:3].strip().isdigit()]
    got = _run_example({"kind": "solve", "nx": 50, "ny": 50, "scaling": 0, "mvec": 5})
    assert got[0]["iters"] == got[1]["iters"] == 26
    assert got[0]["mode"] == "peer"
    assert np.array_equal(got[0]["rnorm"], got[1]["rnorm"])
    assert api.format_table(got[0]["rnorm"]) == lines(acc_txt)
    assert list(got[0]["nvec"]) == [0, 1, 2, 3, 4, 5] + [5] * 20
    got = _run_example({"kind": "solve", "nx": 50, "ny": 50, "scaling": 0, "mvec": 0})
    assert got[0]["iters"] == 367
    assert api.format_table(got[0]["rnorm"]) == lines(unacc_txt)


def test_all_gpus_8192_picard_steps_against_cpu_oracle(tmp_path_factory):
    """BASELINE.json configs[3] at a size the CPU oracle still checks: 8192 x 8192 on every GPU of
    the box (row slabs), three Picard iterations taken one kernel at a time.  Each device kernel is
    compared with the CPU loops ON THE DEVICE'S OWN INPUT, so the comparison is exact where the
    reference is order-exact: residual and update_system (np.array_equal on r and u), the
    slab-pipelined SSOR sweeps across all slab boundaries (np.array_equal on z), accel_update with
    the cross-rank sum fused into pass A (against the long-double arbiter replaying the device's
    f-sequence: 1e-12, or twice the reference's own serial-vs-long-double spread where that is
    larger; identical num_vec), norms to 1e-9."""
    import shutil
    import torch
    from oracle import api
    from parity_log import record_parity
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(world, 8)
    nx = ny = 8192
    iters, mvec = 3, 5
    d = "/dev/shm/nka_steps_%d" % os.getpid()
    os.makedirs(d, exist_ok=True)
    try:
        got = _run_example({"kind": "steps", "nx": nx, "ny": ny, "iters": iters, "mvec": mvec, "dir": d,
                            "timeout": 900}, world=world)
        join = lambda tag, it: np.concatenate([np.load(os.path.join(d, "%s_%d_r%d.npy" % (tag, it, r)), mmap_mode="r")
                                               for r in range(world)], axis=0)
        assert all(got[r]["mode"] == "peer" for r in range(world))
        assert all(got[r]["rnorm"] == got[0]["rnorm"] and got[r]["nvec"] == got[0]["nvec"] for r in range(world))
        assert [got[r]["rows"] for r in range(world)] == [((ny * r) // world, (ny * (r + 1)) // world) for r in range(world)]
        orc = api.OracleSystem(nx, ny, 0.02, 1)
        arb = api.OracleNKA(nx * ny, mvec, 0.01, dotmode=1)
        ser = api.OracleNKA(nx * ny, mvec, 0.01, dotmode=0)      # the unmodified reference's arithmetic
        spread = 0.0
        pad = np.zeros((ny + 2, nx + 2))
        r_cpu = orc.residual(pad).reshape(ny, nx)
        assert np.array_equal(join("r", 0), r_cpu)
        assert abs(got[0]["rnorm"][0] - orc.norm2(r_cpu.ravel())) <= 1e-9 * got[0]["rnorm"][0]
        u_prev = np.zeros((ny, nx))
        for it in range(1, iters + 1):
            r_dev = join("r", it - 1)
            z_cpu = orc.pc_ssor(2, 1.4, np.ascontiguousarray(r_dev).ravel()).reshape(ny, nx)
            z_dev = join("zssor", it)
            assert np.array_equal(z_dev, z_cpu), ("ssor", it)
            want = np.ascontiguousarray(z_dev).ravel().copy()
            arb.accel_update(want)
            wser = np.ascontiguousarray(z_dev).ravel().copy()
            ser.accel_update(wser)
            zacc = np.ascontiguousarray(join("zacc", it))
            # the example's corrections are nearly parallel (ill-conditioned Gram matrix): the bar is
            # 1e-12 or the reference's own serial-vs-long-double spread on this very sequence, as in
            # tests/scenarios.py: tolerances (here with factor 2: the well-conditioned end of that family)
            spread = max(spread, np.linalg.norm(wser - want) / np.linalg.norm(want))
            err = np.linalg.norm(zacc.ravel() - want) / np.linalg.norm(want)
            record_parity("example_8192_all_gpus_it%d" % it, n=nx * ny, mvec=mvec, vtol=0.01, ranks=world,
                          err_vs_arbiter=float(err), reference_serial_vs_arbiter=float(spread),
                          tol_used=float(max(1e-12, 2 * spread)))
            assert err <= max(1e-12, 2 * spread), ("accel", it, err, spread)
            assert got[0]["nvec"][it - 1] == arb.num_vec()
            u_dev = join("u", it)
            assert np.array_equal(u_dev, u_prev - zacc), ("u", it)
            pad[1:-1, 1:-1] = u_dev
            r_cpu = orc.residual(pad).reshape(ny, nx)
            assert np.array_equal(join("r", it), r_cpu), ("residual", it)
            assert abs(got[0]["rnorm"][it] - orc.norm2(r_cpu.ravel())) <= 1e-9 * got[0]["rnorm"][it]
            u_prev = np.ascontiguousarray(u_dev)
    finally:
        shutil.rmtree(d, ignore_errors=True)
