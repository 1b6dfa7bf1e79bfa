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
