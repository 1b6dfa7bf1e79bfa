"""Build-time guard of the code-generation properties the measured numbers depend on (no GPU
needed: nvcc cross-compiles sm_100a here and cuobjdump reads the objects).  DESIGN.md section 4 and
section 10 state them; a compiler or flag change that breaks one shows up here, not as an
unexplained slowdown on the next GPU run.

  * the two streaming sweeps of accel_update use 16-byte global accesses and do not spill;
    pass B's 512-thread CTAs need <= 128 registers to launch at all;
  * the SSOR sweep's hot loops have no local-memory traffic and no memory fence, copy operands
    with LDGSTS (cp.async) and hand blocks over with named barriers.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def objdir():
    from nka_b200 import build
    build.build_library()
    return build.LIB_DIR


def _res_usage(obj):
    out = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = {"reg": int(m.group(2)), "stack": int(m.group(3)), "shared": int(m.group(4)),
                           "local": int(m.group(5))}
    return res


def _sass(obj, fun):
    return subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout


def _find(res, prefix):
    names = [n for n in res if n.startswith(prefix)]
    assert names, prefix
    return names


def test_streaming_sweeps_vectorised_and_spill_free(objdir):
    a = _res_usage(os.path.join(objdir, "nka_pass_a.o"))
    b = _res_usage(os.path.join(objdir, "nka_pass_b.o"))
    for res in (a, b):
        for name, r in res.items():
            if "nka_pass_" in name:
                assert r["local"] == 0, (name, r)            # no spills in any instantiation
    # pass B: 512-thread CTAs up to 12 streamed columns (65536 registers per SM / 512), 256 beyond
    for name in _find(b, "_Z10nka_pass_bILi"):
        nz = int(re.match(r"_Z10nka_pass_bILi(\d+)E", name).group(1))
        assert b[name]["reg"] <= (128 if nz < 13 else 255), (name, b[name])
    # the BASELINE configuration (mvec = 10, 16-byte accesses)
    fb = _find(b, "_Z10nka_pass_bILi10ELi2E")[0]
    sass = _sass(os.path.join(objdir, "nka_pass_b.o"), fb)
    assert "LDG.E.128" in sass and "STG.E.128" in sass
    fa = _find(a, "_Z10nka_pass_aILi10ELi2E")[0]
    assert "LDG.E.128" in _sass(os.path.join(objdir, "nka_pass_a.o"), fa)


def test_ssor_sweep_hot_loops(objdir):
    obj = os.path.join(objdir, "nka_example.o")
    res = _res_usage(obj)
    names = _find(res, "_Z14ex_ssor_sweep2ILi")
    assert len(names) == 8                                    # direction x trace x slabs
    for name in names:
        assert res[name]["local"] == 0 and res[name]["stack"] == 0, (name, res[name])
        # two CTAs of 128 threads per SM must fit the register file with room to spare
        assert res[name]["reg"] <= 168, (name, res[name])
    for name in names:
        if "Lb0ELb0E" not in name:                            # the one-GPU, trace-off kernels are the measured ones
            continue
        sass = _sass(obj, name)
        assert "STL" not in sass and "LDL" not in sass
        assert "MEMBAR" not in sass                           # barriers order the hand-over; a fence stalls on the cp.asyncs in flight
        assert "LDGSTS" in sass and "BAR.ARV" in sass and "SHFL" in sass
        assert "MUFU.RCP64H" in sass                          # the reciprocal half of the split division (copier warp)


def test_residual_strip_kernel_does_not_spill(objdir):
    """The residual walk keeps its prefetched operands in registers: a spilled prefetch value makes the warp wait
    for its loads at once (57 % of all stall samples in the 64-register build, profiles/r2ae_residual_strip_8192_ncu.txt)."""
    obj = os.path.join(objdir, "nka_example.o")
    res = _res_usage(obj)
    name = _find(res, "_Z24ex_residual_strip_kernel")[0]
    assert res[name]["stack"] == 0 and res[name]["local"] == 0, res[name]
    assert res[name]["reg"] <= 85, res[name]                  # three CTAs of 256 threads per SM
    sass = _sass(obj, name)
    assert "STL" not in sass and "LDL" not in sass
