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
from parity_log import record_parity

pytestmark = pytest.mark.gpu


def _arbiter(n, mvec, vtol):
    """The reference itself with a long-double dot product injected through its documented dp hook
    (src-C/nonlinear_krylov_accelerator.c:227-231) when the compiled reference travelled with the
    tree (oracle/_ref), else the port in the same mode (bit-identical: tests/test_oracle.py)."""
    if api.ref_lib() is not None and (mvec + 1) * n < 2 ** 31:
        return api.RefNKA(n, mvec, vtol, long_double_dp=True), "reference src-C + long-double dp hook"
    return api.OracleNKA(n, mvec, vtol, dotmode=1), "oracle port, long-double dots"


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


def test_iid_2p24_nonperiodic_against_reference_long_double():
    """n = 2^24 (the 4096^2 example's length), mvec = 10, i.i.d. NON-periodic inputs: every element
    of every correction is compared with the reference run on the same 2^24 values (no tiling, so a
    swapped / re-read tile or a chunk-offset error cannot hide).  Strict 1e-12."""
    import torch
    from nka_b200 import NKA
    n, mvec = 1 << 24, 10
    rng = np.random.default_rng(2024)
    orc, kind = _arbiter(n, mvec, 0.01)
    acc = NKA(n, mvec, 0.01)
    worst = 0.0
    for t in range(mvec + 4):
        f = rng.random(n) - 0.5
        want = f.copy()
        orc.accel_update(want)
        d = torch.from_numpy(f).cuda()
        acc.accel_update(d)
        got = d.cpu().numpy()
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        worst = max(worst, err)
        assert err <= 1e-12, (t, err)
        assert acc.num_vec() == orc.num_vec()
    assert acc.defined()
    acc.delete()
    record_parity("fullsize_iid_n2p24_m10", n=n, mvec=mvec, vtol=0.01, calls=mvec + 4, drops=0,
                  err_vs_arbiter=worst, tol_used=1e-12, arbiter=kind, inputs="non-periodic i.i.d.")


def test_host_chunked_path_2p24_nonperiodic_elementwise():
    """The 16-chunk host-pointer path on non-periodic data: the result of nka_accel_update(host f)
    equals the device-pointer path's on the same inputs at every element, to reduction-order
    rounding (the chunked sweep folds its partial rows in a different order)."""
    import torch
    from nka_b200 import NKA
    n, mvec = (1 << 24) + 10, 6
    g = torch.Generator(device="cuda").manual_seed(5)
    a, b = NKA(n, mvec, 0.01), NKA(n, mvec, 0.01)
    host = np.empty(n)                                    # pageable, as a reference caller's malloc
    for t in range(mvec + 3):
        f = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
        host[:] = f.cpu().numpy()
        a.accel_update(f)
        b.accel_update(host)
        got = torch.from_numpy(host).cuda()
        assert float((got - f).abs().max()) <= 1e-13, t
        assert a.num_vec() == b.num_vec()
    a.delete(); b.delete()


def test_stress_config5_2p26_nonperiodic_against_reference():
    """BASELINE.json configs[4] as SURVEY.md 8(d) states it: the ill-conditioned family (collinear
    drift + an exact repeat + relax() + restart()) at n = 2^26 with NON-periodic inputs, compared
    with the reference itself at that length (11 * 2^26 < 2^31 fits its int arithmetic): identical
    num_vec / drop / relax decisions after every call, corrections against the long-double arbiter."""
    import torch
    import scenarios as S
    from nka_b200 import NKA
    n, mvec, vtol = 1 << 26, 8, 0.2
    if _free_gib() < 12:
        pytest.skip("needs ~10 GiB of device memory")
    ops = S.mixed_stress(n, 18, 23)
    orc, kind = _arbiter(n, mvec, vtol)
    acc = NKA(n, mvec, vtol)
    it, worst, nv_seq, ndrops, nrelaxed = 0, 0.0, [], 0, 0
    for op in ops:
        if op[0] == "update":
            want = op[1].copy()
            orc.accel_update(want)
            d = torch.from_numpy(op[1]).cuda()
            acc.accel_update(d)
            got = d.cpu().numpy()
            st = acc.state()
            assert st["error"] == 0
            ndrops += st["ndrop_last"]; nrelaxed += st["relaxed_last"]
            scale = max(np.linalg.norm(want), np.linalg.norm(op[1]))
            err = np.linalg.norm(got - want) / scale
            worst = max(worst, err)
            assert err <= 1e-12, (it, err)
            it += 1
        elif op[0] == "relax":
            orc.relax(); acc.relax()
        else:
            orc.restart(); acc.restart()
        assert acc.num_vec() == orc.num_vec(), it
        nv_seq.append(acc.num_vec())
    assert acc.defined()
    assert ndrops >= 6 and nrelaxed >= 1         # the sequence really exercised the drop / guard paths
    acc.delete()
    record_parity("fullsize_stress_n2p26_m8", n=n, mvec=mvec, vtol=vtol, calls=it, num_vec=nv_seq, drops=ndrops,
                  err_vs_arbiter=worst, tol_used=1e-12, arbiter=kind, inputs="non-periodic stress family")


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
