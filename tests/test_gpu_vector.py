"""gpu_vector operations and the vector flavour of accel_update (-m gpu), against numpy
restatements of src-F08-vector/grid_vector_type.F90:108-197 and the oracle."""
import numpy as np
import pytest

import scenarios as S
from oracle import api

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 3, 1000, 4097, 1 << 20])
def test_vector_ops_match_grid_vector_expressions(n):
    from nka_b200.vector import GpuVector
    rng = np.random.default_rng(n + 1)
    hx, hy, hz = (rng.uniform(-1, 1, n) for _ in range(3))
    x, y, z = GpuVector(n), GpuVector(n), GpuVector(n)
    x.set(hx); y.set(hy); z.set(hz)
    a, b, c = 0.75, -1.25, 0.5

    z.update(a, x); hz = a * hx + hz                      # update1_  :121-129
    assert np.array_equal(z.get(), hz) or np.allclose(z.get(), hz, rtol=0, atol=2e-16 * 4)
    z.update(a, x, b); hz = a * hx + b * hz               # update2_  :132-140
    np.testing.assert_allclose(z.get(), hz, rtol=0, atol=1e-15)   # fma vs two roundings
    z.update(a, x, b, y); hz = a * hx + b * hy + hz       # update3_  :143-154
    np.testing.assert_allclose(z.get(), hz, rtol=0, atol=1e-15)
    z.update(a, x, b, y, c); hz = a * hx + b * hy + c * hz  # update4_ :157-168
    np.testing.assert_allclose(z.get(), hz, rtol=0, atol=1e-15)
    before = z.get()                                      # (host hz now differs from z by fma roundings)
    z.scale(-2.0); hz = -2.0 * hz                         # scale     :114-118
    assert np.array_equal(z.get(), -2.0 * before)
    np.testing.assert_allclose(z.get(), hz, rtol=0, atol=4e-15)
    # zero-coefficient short cuts of the base class (vector_class.F90:176,189,203-206)
    before = z.get()
    z.update(0.0, x); assert np.array_equal(z.get(), before)
    z.update(0.0, x, 3.0); assert np.array_equal(z.get(), 3.0 * before)
    # reductions
    if n:
        assert abs(x.dot(y) - float(np.dot(hx.astype(np.longdouble), hy.astype(np.longdouble)))) \
            <= 1e-13 * np.linalg.norm(hx) * np.linalg.norm(hy)
        assert abs(x.norm2() - np.linalg.norm(hx)) <= 1e-14 * np.linalg.norm(hx)
    else:
        assert x.dot(y) == 0.0 and x.norm2() == 0.0
    w = x.clone()
    assert np.array_equal(w.get(), hx)
    w.setval(1.5)
    assert np.array_equal(w.get(), np.full(n, 1.5)) and np.array_equal(x.get(), hx)
    w.copy(x)
    assert np.array_equal(w.get(), hx)
    clones = x.clone(3)
    assert len(clones) == 3 and all(np.array_equal(v.get(), hx) for v in clones)
    with pytest.raises(TypeError):
        x.copy(GpuVector(n + 1))


def test_vector_flavour_accel_update_matches_oracle():
    """init(vec, mvec) + accel_update(class(vector) f): src-F08-vector/nka_type.F90:175-188, 219-400."""
    from nka_b200.vector import GpuVector, VectorNKA
    n, mvec, vtol, mk = S.SCENARIOS["iid_n1000_m10"]
    proto = GpuVector(n)
    proto.setval(0.0)
    acc = VectorNKA().init(proto, mvec, vtol)
    orc = api.OracleNKA(n, mvec, vtol, flavour=1)
    f = proto.clone()
    for op in mk():
        want = op[1].copy()
        orc.accel_update(want)
        f.set(op[1])
        acc.accel_update(f)
        got = f.get()
        assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
        assert acc.num_vec() == orc.num_vec()
    assert acc.defined() and acc.max_vec() == mvec and acc.vec_tol() == vtol
    with pytest.raises(TypeError):
        acc.accel_update(np.zeros(n))
