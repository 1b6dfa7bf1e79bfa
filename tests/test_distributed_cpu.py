"""N > 1 host logic on CPU (gloo, world_size 2).

The CUDA path cannot run here, so this checks what surrounds it: the slab partition, the
id exchange over torch.distributed, and -- with the REFERENCE's own C accelerator driven
through its `dp` hook by a gloo all-reduce -- that "each rank updates its slab, dot products
are summed over ranks" reproduces the serial full-vector update exactly as the reference
promises (src-F08-vector/README.md:16-22).  That is the contract libnka_b200's NCCL path
implements on the GPUs (checked for real by tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_partition():
    from nka_b200.distributed import slab_bounds
    for n in (0, 1, 2, 3, 7, 16, 1001, 1 << 20, (1 << 28) + 5):
        for world in (1, 2, 3, 4, 8):
            edges = [slab_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
            assert all(lo % 2 == 0 for lo, _ in edges)
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 2 or n < 2 * world
    with pytest.raises(ValueError):
        slab_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, mvec, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nka_b200.distributed import exchange_unique_id, slab_bounds
    from oracle import api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = exchange_unique_id()
        lo, hi = slab_bounds(n, world, rank)
        lib = api.ref_lib()
        have_ref = lib is not None
        flags = [None] * world
        dist.all_gather_object(flags, have_ref)
        results = {"uid": uid, "bounds": (lo, hi)}
        if all(flags):
            DP = C.CFUNCTYPE(C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))

            def global_dot(length, x, y):
                a = np.ctypeslib.as_array(x, shape=(length,)) if length else np.zeros(0)
                b = np.ctypeslib.as_array(y, shape=(length,)) if length else np.zeros(0)
                t = torch.tensor([float(np.dot(a, b))], dtype=torch.float64)
                dist.all_reduce(t)
                return float(t.item())

            cb = DP(global_dot)
            h = lib.nka_init(hi - lo, mvec, 0.01, C.cast(cb, C.c_void_p))
            rng = np.random.default_rng(77)
            outs, nvec = [], []
            for t in range(mvec + 5):
                full = rng.uniform(-0.5, 0.5, n) * 0.8 ** t
                mine = np.ascontiguousarray(full[lo:hi])
                lib.nka_accel_update(h, mine.ctypes.data_as(C.POINTER(C.c_double)))
                outs.append(mine)
                nvec.append(lib.nka_num_vec(h))
            lib.nka_delete(h)
            results["outs"] = outs
            results["nvec"] = nvec
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def test_two_rank_slab_update_equals_serial_update():
    import torch.multiprocessing as mp
    from oracle import api
    n, mvec, world = 1001, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, mvec, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the 128-byte id reached both ranks unchanged
    assert got[0]["uid"] == got[1]["uid"] and len(got[0]["uid"]) == 128
    assert got[0]["bounds"][1] == got[1]["bounds"][0]
    if "outs" not in got[0]:
        pytest.skip("oracle/_ref not built in this tree")
    # serial reference on the full vectors (long-double dots: summation order differs across the split)
    ser = api.OracleNKA(n, mvec, 0.01, dotmode=1)
    rng = np.random.default_rng(77)
    for t in range(mvec + 5):
        full = rng.uniform(-0.5, 0.5, n) * 0.8 ** t
        ser.accel_update(full)
        joined = np.concatenate([got[0]["outs"][t], got[1]["outs"][t]])
        assert np.linalg.norm(joined - full) <= 1e-12 * np.linalg.norm(full), t
        assert got[0]["nvec"][t] == got[1]["nvec"][t] == ser.num_vec()


def test_row_slabs_tile_the_grid():
    """nka_b200.example.row_slab: contiguous, balanced, complete (the slab example's partition)."""
    from nka_b200.example import row_slab
    for ny in (9, 50, 64, 131, 4096, 32768):
        for world in (1, 2, 3, 4, 8):
            if ny // world < 3:
                continue
            bounds = [row_slab(ny, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == ny
            assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in bounds]
            assert min(sizes) >= 3 and max(sizes) - min(sizes) <= 1
