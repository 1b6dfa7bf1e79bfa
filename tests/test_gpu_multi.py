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
@pytest.mark.parametrize("mode", ["peer", "nccl"])
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
    assert got[0]["comm_mode"] == got[1]["comm_mode"] == mode
    assert got[0]["nvec"] == got[1]["nvec"] == nv_ref
    assert got[0]["decisions"] == got[1]["decisions"]
    assert all(d[3] == 0 for d in got[0]["decisions"])
    for t in range(len(inputs)):
        joined = np.concatenate([got[0]["outs"][t], got[1]["outs"][t]])
        assert np.linalg.norm(joined - arbiter[t]) / scales[t] <= tols[t], (name, t)
