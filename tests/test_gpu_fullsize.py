"""Full-size checks (-m gpu) at BASELINE.json's n = 2^28, mvec = 10, through
size-independent properties, since no CPU oracle finishes there in seconds.

1. Tiling: if every input is a length-p vector repeated n/p times (n/p a power
   of 4), all dot products scale by exactly n/p, the norm by its exact square
   root, and the Cholesky/solve see the same numbers; so the n-length result
   must be the tiled p-length ORACLE result to rounding.  This is a genuine
   comparison with the reference algorithm at full size.
2. Scaling by a power of two is exact in binary floating point: the update of
   (2^k f_t) must be bit-identical to 2^k times the update of (f_t).
"""
import numpy as np
import pytest

from oracle import api

pytestmark = pytest.mark.gpu


def _free_gib():
    import torch
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


def test_full_size_tiled_inputs_match_oracle():
    import torch
    from nka_b200 import NKA
    n, mvec, p = 1 << 28, 10, 1 << 16          # n/p = 4096 = 4^6
    if _free_gib() < 60:
        pytest.skip("needs ~50 GiB of device memory")
    rng = np.random.default_rng(7)
    orc = api.OracleNKA(p, mvec, 0.01, dotmode=1)
    acc = NKA(n, mvec, 0.01)
    f = torch.empty(n, dtype=torch.float64, device="cuda")
    for t in range(mvec + 4):
        small = rng.uniform(-0.5, 0.5, p)
        want = small.copy()
        orc.accel_update(want)
        f.view(n // p, p).copy_(torch.from_numpy(small).cuda().unsqueeze(0).expand(n // p, p))
        acc.accel_update(f)
        tiles = f.view(n // p, p)
        # every tile is the same vector ...
        assert bool((tiles[0:1] == tiles).all())
        # ... and equals the oracle's p-length correction
        got = tiles[0].cpu().numpy()
        assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want), t
        assert acc.num_vec() == orc.num_vec()
    assert acc.defined()
    acc.delete()


def test_stress_config5_full_size_tiled():
    """BASELINE.json configs[4]: ill-conditioned sequence forcing vtol drops, the s == 0 relax
    guard, relax() and restart(), at n = 2^26.  Inputs are a p = 4096 stress sequence tiled
    n/p = 4^7 times, so the oracle at length p decides what must happen at length n: identical
    num_vec / drop / relax / eviction decisions on every call (14 drops, margins >= 4e-5), and
    every tile equal to the oracle's correction."""
    import torch
    import scenarios as S
    from nka_b200 import NKA
    n, p, mvec, vtol = 1 << 26, 4096, 8, 0.2
    if _free_gib() < 12:
        pytest.skip("needs ~10 GiB of device memory")
    ops = S.mixed_stress(p, 30, 23)
    inputs = [op[1] for op in ops if op[0] == "update"]
    serial, _ = S.run_ops(api.OracleNKA(p, mvec, vtol, dotmode=0), ops)
    arbiter, _ = S.run_ops(api.OracleNKA(p, mvec, vtol, dotmode=1), ops)
    scales, tols = S.tolerances(serial, arbiter, inputs)
    orc = api.OracleNKA(p, mvec, vtol, dotmode=1)
    acc = NKA(n, mvec, vtol)
    f = torch.empty(n, dtype=torch.float64, device="cuda")
    it, ndrops, nrelaxed = 0, 0, 0
    for op in ops:
        if op[0] == "update":
            want = op[1].copy()
            orc.accel_update(want)
            f.view(n // p, p).copy_(torch.from_numpy(op[1]).cuda().unsqueeze(0).expand(n // p, p))
            acc.accel_update(f)
            st = acc.state()
            assert st["error"] == 0
            assert (st["ndrop_last"], bool(st["relaxed_last"]), bool(st["evicted_last"])) == \
                   (orc.ndrop_last(), orc.relaxed_last(), orc.evicted_last()), it
            ndrops += st["ndrop_last"]; nrelaxed += st["relaxed_last"]
            tiles = f.view(n // p, p)
            assert bool((tiles[0:1] == tiles).all())
            got = tiles[0].cpu().numpy()
            assert np.linalg.norm(got - arbiter[it]) / scales[it] <= tols[it], it
            it += 1
        elif op[0] == "relax":
            orc.relax(); acc.relax()
        else:
            orc.restart(); acc.restart()
        assert acc.num_vec() == orc.num_vec(), it
    assert ndrops >= 10 and nrelaxed >= 1        # the sequence really exercised the drop / guard paths
    assert acc.defined()
    acc.delete()


def test_power_of_two_scaling_is_bit_exact():
    import torch
    from nka_b200 import NKA
    n, mvec = 1 << 24, 5
    outs = []
    for scale in (1.0, 2.0 ** 7):
        g = torch.Generator(device="cuda").manual_seed(11)
        acc = NKA(n, mvec, 0.01)
        run = []
        for t in range(mvec + 3):
            f = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5) * scale
            acc.accel_update(f)
            run.append(f / scale)
        outs.append(run)
        acc.delete()
    for a, b in zip(*outs):
        assert bool((a == b).all())
