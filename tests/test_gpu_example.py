"""The example's physics on the device (-m gpu), through the C-ABI of include/nka_example.h,
against the oracle's restatement of system_type / solver_type (oracle/nka_oracle.c, pinned to
the reference's golden reference_output files by tests/test_oracle.py).

Bar: residual, face coefficients and the SSOR preconditioner are BIT-IDENTICAL to the serial CPU
loops (same operand order, no fma, exact Gauss-Seidel order); the Picard tables reproduce the
reference's golden output line for line with the identical iteration count.
"""
import json
import os

import numpy as np
import pytest

from oracle import api

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _pad(u):
    ny, nx = u.shape
    up = np.zeros((ny + 2, nx + 2))
    up[1:-1, 1:-1] = u
    return up


def _random_u(nx, ny, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, 0.3, (ny, nx))          # a + u stays positive


@pytest.fixture(params=[1, 2, 3], ids=["sweep1", "sweep2", "sweep3"])
def ssor_kernel(request, monkeypatch):
    """All SSOR kernels (NKA_SSOR_KERNEL is read when a System is created): ex_ssor_sweep,
    ex_ssor_sweep2 (the dependent chain on a warp of its own) and ex_ssor_sweep3 (two columns and
    two independent chains per lane, 64-column strips)."""
    monkeypatch.setenv("NKA_SSOR_KERNEL", str(request.param))
    return request.param


def test_ssor_division_identical():
    """ex_ssor_sweep2 takes the reciprocal of the divisor off the dependent chain; the quotient
    must be the device's IEEE quotient in every bit, out-of-range operands included."""
    from nka_b200 import _lib
    lib = _lib.load()
    assert lib.nka_example_division_check(1 << 28, 12345, 0) == 0
    assert lib.nka_example_division_check(1 << 24, 99, 0) == 0


SHAPES = [(3, 3), (5, 4), (4, 9), (31, 17), (32, 32), (33, 70), (50, 50), (96, 40), (257, 129), (300, 300), (700, 64)]
# strip-edge cases of the 64-column strips of ex_ssor_sweep3: one column into a strip, one short of it, exactly full
SSOR_SHAPES = SHAPES + [(63, 5), (64, 40), (65, 33), (127, 12), (129, 70)]


@pytest.mark.parametrize("nx,ny", SHAPES)
@pytest.mark.parametrize("scaling", [0, 1])
def test_residual_and_coefficients_bit_identical(nx, ny, scaling):
    from nka_b200.example import System, FIELD_U, FIELD_R, FIELD_AXL, FIELD_AYD, FIELD_AC
    u = _random_u(nx, ny, nx * 1000 + ny)
    sy = System(0.02, nx, ny, scaling=scaling)
    sy.set(FIELD_U, u)
    assert np.array_equal(sy.get(FIELD_U), u)                   # natural <-> wavefront round trip
    rn = sy.residual()
    orc = api.OracleSystem(nx, ny, 0.02, scaling)
    r = orc.residual(_pad(u)).reshape(ny, nx)
    ax, ay, ac = orc.coefficients()
    assert np.array_equal(sy.get(FIELD_R), r)
    assert np.array_equal(sy.get(FIELD_AXL), ax[:, :nx])
    assert np.array_equal(sy.get(FIELD_AYD), ay[:ny, :])
    assert np.array_equal(sy.get(FIELD_AC), ac)
    assert abs(rn - orc.norm2(r)) <= 1e-14 * rn
    sy.delete()


@pytest.mark.parametrize("nx,ny", SSOR_SHAPES)
@pytest.mark.parametrize("nsweep", [1, 2, 3])
def test_pc_ssor_bit_identical_to_serial_gauss_seidel(nx, ny, nsweep, ssor_kernel):
    from nka_b200.example import System, FIELD_U, FIELD_R, FIELD_Z
    u = _random_u(nx, ny, nx * 77 + ny)
    sy = System(0.02, nx, ny, scaling=1)
    sy.set(FIELD_U, u)
    sy.residual()
    orc = api.OracleSystem(nx, ny, 0.02, 1)
    r = orc.residual(_pad(u))
    for rep in range(2):                                        # twice: the edge channels must come back clean
        sy.pc_ssor(nsweep, 1.4)
        z = sy.get(FIELD_Z)
        want = orc.pc_ssor(nsweep, 1.4, r).reshape(ny, nx)
        assert np.array_equal(z, want), (rep, float(np.abs(z - want).max()))
    assert np.array_equal(sy.get(FIELD_R), r.reshape(ny, nx))   # r itself is left alone
    sy.delete()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("nx,ny", [(9700, 7), (40000, 5), (9601, 6)])
def test_pc_ssor_more_strips_than_resident_ctas(nx, ny, ssor_kernel):
    """Grids wider than 32 x (resident CTAs): every CTA then walks several strips in turn (the
    32768-wide slab configuration does), so the hand-over barriers, mailbox and edge channels
    must come back idle after each strip.  9700 columns = 304 strips > the 296 resident CTAs of
    ex_ssor_sweep2 (and 152 strips of 64 > the 148 of ex_ssor_sweep3); 40000 columns = 1250 strips >
    ex_ssor_sweep's residency too; 9601 = 150 full strips of 64 and one column."""
    from nka_b200.example import System, FIELD_U, FIELD_Z
    u = _random_u(nx, ny, 4242)
    sy = System(0.02, nx, ny, scaling=1)
    sy.set(FIELD_U, u)
    sy.residual()
    orc = api.OracleSystem(nx, ny, 0.02, 1)
    r = orc.residual(_pad(u))
    for rep in range(2):
        sy.pc_ssor(2, 1.4)
        z = sy.get(FIELD_Z)
        want = orc.pc_ssor(2, 1.4, r).reshape(ny, nx)
        assert np.array_equal(z, want), (rep, float(np.abs(z - want).max()))
    sy.delete()


def test_update_then_residual_bit_identical():
    """u = u - r ; residual (src-F08/nka_example.F90:248-249) fused in one kernel."""
    from nka_b200.example import System, FIELD_U, FIELD_R, FIELD_Z
    nx, ny = 130, 75
    u = _random_u(nx, ny, 5)
    z = np.random.default_rng(6).uniform(-0.01, 0.01, (ny, nx))
    sy = System(0.02, nx, ny, scaling=0)
    sy.set(FIELD_U, u)
    sy.set(FIELD_Z, z)
    rn = sy.residual(subtract_z=True)
    orc = api.OracleSystem(nx, ny, 0.02, 0)
    r = orc.residual(_pad(u - z)).reshape(ny, nx)
    assert np.array_equal(sy.get(FIELD_U), u - z)
    assert np.array_equal(sy.get(FIELD_R), r)
    assert abs(rn - orc.norm2(r)) <= 1e-14 * rn
    sy.delete()


def _golden_table(part):
    with open(os.path.join(GOLD, "example_c_f95.txt")) as fh:
        acc, unacc = fh.read().split("UNACCELERATED SOLVE")
    text = acc if part == "accelerated" else unacc
    return [ln for ln in text.splitlines() if ":" in ln[:4] and ln[:3].strip().isdigit()]


def test_example_on_device_reproduces_golden_accelerated_table(ssor_kernel):
    """src-F95/reference_output:5-31 == src-C/reference_output: 26 iterations, every line equal;
    no vector is ever dropped (num_vec 0,1,2,3,4,5,5,...), as in the reference run."""
    from nka_b200.example import System, Solver
    sy = System(0.02, 50, 50, scaling=0)
    so = Solver(sy, nsweep=2, omega=1.4, mvec=5, vtol=0.01)
    out = so.solve(record_nvec=True)
    ref = api.example_solve(mvec=5, record=True)
    assert out["iters"] == 26 == ref["iters"]
    assert api.format_table(out["rnorm"]) == _golden_table("accelerated")
    assert list(out["nvec"]) == list(ref["nvec"])
    so.delete(); sy.delete()


def test_example_on_device_reproduces_golden_unaccelerated_table(ssor_kernel):
    """src-F95/reference_output:36-403: 367 iterations.  Without NKA every kernel on the path is
    bit-identical to the CPU code, so the norms agree to the last digits of the blocked sum."""
    from nka_b200.example import System, Solver
    sy = System(0.02, 50, 50, scaling=0)
    so = Solver(sy, nsweep=2, omega=1.4, mvec=0)
    out = so.solve()
    assert out["iters"] == 367
    assert api.format_table(out["rnorm"]) == _golden_table("unaccelerated")
    ref = api.example_solve(mvec=0)
    assert np.allclose(out["rnorm"], ref["rnorm"], rtol=1e-13, atol=0)
    so.delete(); sy.delete()


def test_example_on_device_f08_reference_output_lines():
    """src-F08/reference_output:7,16,25 (the three runs the F08 flavours ship)."""
    from nka_b200.example import System, Solver
    with open(os.path.join(GOLD, "example_f08.json")) as fh:
        runs = json.load(fh)["runs"]
    for run in runs:
        args = run["args"]
        nsweep = int(args[args.index("--sweeps") + 1]) if "--sweeps" in args else 2
        mvec = int(args[args.index("--nka-vec") + 1]) if "--nka-vec" in args else 0
        sy = System(0.02, 50, 50, scaling=1)
        so = Solver(sy, nsweep=nsweep, omega=1.4, mvec=mvec)
        out = so.solve()
        assert api.format_table(out["rnorm"])[-1] == run["last_line"], args
        so.delete(); sy.delete()


def test_example_rectangular_and_multi_cta_solve_matches_oracle_history(ssor_kernel):
    """A grid wider than one CTA's 256 columns, not a multiple of 32: same iteration count and
    residual history as the CPU oracle (NKA sums in another order: 1e-9 on the norms)."""
    from nka_b200.example import System, Solver
    nx, ny = 300, 173
    sy = System(0.05, nx, ny, scaling=1)
    so = Solver(sy, nsweep=2, omega=1.4, mvec=4, vtol=0.01)
    out = so.solve(maxitr=40, record_nvec=True)
    ref = api.example_solve(nx=nx, ny=ny, a=0.05, nsweep=2, omega=1.4, mvec=4, scaling=1, maxitr=40, record=True)
    assert out["iters"] == ref["iters"]
    assert np.allclose(out["rnorm"], ref["rnorm"], rtol=1e-9, atol=0)
    assert list(out["nvec"]) == list(ref["nvec"])
    so.delete(); sy.delete()


def test_full_size_4096_residual_and_ssor_bit_identical(ssor_kernel):
    """BASELINE.json configs[1]: the 4096 x 4096 grid.  One residual and one 2-sweep SSOR
    application against the serial CPU loops, bit for bit; then 3 accelerated Picard iterations
    against the oracle's history."""
    from nka_b200.example import System, Solver, FIELD_R, FIELD_Z
    nx = ny = 4096
    sy = System(0.02, nx, ny, scaling=1)
    rn0 = sy.residual()
    orc = api.OracleSystem(nx, ny, 0.02, 1)
    r = orc.residual(np.zeros((ny + 2, nx + 2)))
    assert np.array_equal(sy.get(FIELD_R).ravel(), r)
    sy.pc_ssor(2, 1.4)
    z = orc.pc_ssor(2, 1.4, r)
    assert np.array_equal(sy.get(FIELD_Z).ravel(), z)
    assert abs(rn0 - orc.norm2(r)) <= 1e-13 * rn0
    so = Solver(sy, nsweep=2, omega=1.4, mvec=5)
    out = so.solve(maxitr=3)
    ref = api.example_solve(nx=nx, ny=ny, nsweep=2, omega=1.4, mvec=5, scaling=1, maxitr=3)
    assert np.allclose(out["rnorm"], ref["rnorm"], rtol=1e-9, atol=0)
    so.delete(); sy.delete()
