"""Host-side mirror of the reference example's two modules over the C-ABI
(include/nka_example.h): `System` = system_type (init, residual, pc_ssor;
src-F08/nka_example.F90:67-181) and `Solver` = solver_type (init, solve; :187-258).

Everything runs on the device: the iterate, the residual, the SSOR preconditioner in the
reference's exact Gauss-Seidel order, and accel_update; the host only sees one residual
norm per Picard iteration.  Python is the binding; nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .nka import NKA

FIELD_U, FIELD_R, FIELD_Z, FIELD_AXL, FIELD_AYD, FIELD_AC = range(6)


class System:
    """type(system): the discrete nonlinear system -div((a+u) grad u) = q, device resident."""

    def __init__(self, a: float | None = None, nx: int | None = None, ny: int | None = None,
                 scaling: int = 1, device: int = -1, stream: int | None = None,
                 slab: tuple[int, int] | None = None):
        self._h = None
        self._lib = _lib.load()
        self.distributed = False
        if a is not None:
            self.init(a, nx, ny, scaling=scaling, device=device, stream=stream, slab=slab)

    def init(self, a: float, nx: int, ny: int, scaling: int = 1, device: int = -1, stream: int | None = None,
             slab: tuple[int, int] | None = None):
        """init(a, nx, ny): src-F08/nka_example.F90:86-101.  scaling=0 gives the F95/C flavour
        (q = hx*hy), scaling=1 the F08 flavours (q = 1).  slab=(k0, k1): this process holds rows
        [k0, k1) of the nx x ny grid (one process per GPU; call comm_init next); self.ny is then
        the LOCAL number of rows, self.ny_global the grid's."""
        if not a > 0.0:
            raise ValueError("a must be > 0")          # ASSERT(a > 0) :90
        if nx < 3 or ny < 3:
            raise ValueError("nx, ny must be >= 3")    # :91-92
        self.delete()
        k0, k1 = slab if slab is not None else (0, ny)
        if not (0 <= k0 and k1 <= ny and k1 - k0 >= 3):
            raise ValueError("a slab needs at least 3 rows of the grid")
        self._h = self._lib.nka_system_init_slab(nx, ny, k0, k1, a, scaling, device, stream)
        self.nx, self.ny, self.a, self.scaling = nx, k1 - k0, a, scaling
        self.ny_global, self.k0, self.k1 = ny, k0, k1
        self.stream = stream
        self.device = device
        self.distributed = False
        return self

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        """Collective: join the slabs (rank r directly above rank r-1) into one grid."""
        buf = C.create_string_buffer(unique_id, 128)
        rc = self._lib.nka_system_comm_init(self._handle(), nranks, rank, buf)
        if rc != 0:
            raise RuntimeError("nka_system_comm_init failed (%d)" % rc)
        self.distributed = nranks > 1

    def delete(self):
        if self._h:
            self._lib.nka_system_delete(self._h)
            self._h = None

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("system is not initialised")
        return self._h

    def size(self) -> int: return self._lib.nka_system_size(self._handle())

    def residual(self, subtract_z: bool = False) -> float:
        """residual(uext, r) of the current u (after u <- u - z if subtract_z); returns norm2(r)."""
        return self._lib.nka_system_residual(self._handle(), int(subtract_z))

    def pc_ssor(self, nsweep: int, omega: float) -> None:
        """pc_ssor(nsweep, omega, r): z <- SSOR(r), exact lexicographic order."""
        if nsweep < 1 or not omega > 0.0:
            raise ValueError("nsweep >= 1 and omega > 0 required")     # :156-157
        if self._lib.nka_system_pc_ssor(self._handle(), nsweep, omega) != 0:
            raise RuntimeError("an earlier SSOR sweep reported an internal error")

    # -- data access (natural order on the host, wavefront-major on the device) --
    def get(self, field: int) -> np.ndarray:
        out = np.empty((self.ny, self.nx))
        self._lib.nka_system_get_field(self._handle(), field, out.ctypes.data)
        return out

    def set(self, field: int, values) -> None:
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(self.ny, self.nx)
        self._lib.nka_system_set_field(self._handle(), field, v.ctypes.data)

    def field_ptr(self, field: int) -> int: return self._lib.nka_system_field(self._handle(), field)
    def index(self, j: int, k: int) -> int: return self._lib.nka_system_index(self._handle(), j, k)
    def launch_count(self) -> int: return self._lib.nka_system_launch_count(self._handle())
    def timing_enable(self, on: bool = True): self._lib.nka_system_timing_enable(self._handle(), int(on))

    def timing_read(self) -> dict:
        ms = (C.c_double * 2)()
        cnt = (C.c_ulonglong * 2)()
        self._lib.nka_system_timing_read(self._handle(), ms, cnt)
        return {"pc_ssor": {"ms": ms[0], "count": cnt[0]}, "residual": {"ms": ms[1], "count": cnt[1]}}


class Solver:
    """type(solver): Picard iteration with SSOR preconditioning and optional NKA acceleration."""

    MAXITR = 999          # src-F08/nka_example.F90:233
    TOL = 1.0e-6          # :234

    def __init__(self, sys: System | None = None, nsweep: int = 2, omega: float = 1.4, mvec: int = 0,
                 vtol: float = 0.01):
        self.accel = None
        if sys is not None:
            self.init(sys, nsweep, omega, mvec, vtol)

    def init(self, sys: System, nsweep: int, omega: float, mvec: int, vtol: float = 0.01):
        """init(sys, nsweep, omega, mvec): :208-224 (mvec = 0: unaccelerated)."""
        if nsweep <= 0 or not omega > 0.0 or mvec < 0:
            raise ValueError("nsweep > 0, omega > 0, mvec >= 0 required")      # :214-216
        self.sys, self.nsweep, self.omega = sys, nsweep, omega
        if self.accel is not None:
            self.accel.delete()
        self.accel = NKA(sys.size(), mvec, vtol, device=sys.device, stream=sys.stream) if mvec > 0 else None
        if self.accel is not None and sys.distributed:
            _lib.load().nka_comm_share_system(self.accel._handle(), sys._handle())   # dot products over all slabs
        return self

    def solve(self, maxitr: int | None = None, tol: float | None = None, record_nvec: bool = False) -> dict:
        """solve(uext): :226-256.  Returns iters, the residual norms and (optionally) num_vec per call."""
        maxitr = self.MAXITR if maxitr is None else maxitr
        tol = self.TOL if tol is None else tol
        rnorm = np.zeros(maxitr + 1)
        nvec = np.zeros(maxitr, dtype=np.int32) if (record_nvec and self.accel is not None) else None
        lib = _lib.load()
        it = lib.nka_example_solve(self.sys._handle(), self.accel._handle() if self.accel else None,
                                   self.nsweep, self.omega, maxitr, tol,
                                   rnorm.ctypes.data_as(C.POINTER(C.c_double)),
                                   nvec.ctypes.data_as(C.POINTER(C.c_int)) if nvec is not None else None)
        out = {"iters": it, "rnorm": rnorm[: it + 1].copy()}
        if nvec is not None:
            out["nvec"] = nvec[:it].copy()
        return out

    def delete(self):
        if self.accel is not None:
            self.accel.delete()
            self.accel = None


def row_slab(ny: int, world: int, rank: int) -> tuple[int, int]:
    """Balanced contiguous rows [k0, k1) of rank `rank`."""
    return (ny * rank) // world, (ny * (rank + 1)) // world


def distributed_system(a: float, nx: int, ny: int, scaling: int = 1, group=None, device: int = -1,
                       stream: int | None = None) -> System:
    """Collective (torch.distributed is only the plumbing that ships the NCCL id): every rank gets
    its row slab of the nx x ny grid, joined for halo rows, the global norm and the slab-pipelined
    SSOR sweep."""
    import torch.distributed as dist
    from .distributed import exchange_unique_id
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sy = System(a, nx, ny, scaling=scaling, device=device, stream=stream, slab=row_slab(ny, world, rank))
    sy.comm_init(world, rank, exchange_unique_id(group))
    return sy
