"""CPU check of the device algorithm's scalar half.

tests/model/ compiles nka_b200/csrc/nka_state.h -- the exact code the state
kernel runs on the device -- for the host, with plain-loop stand-ins for the
streaming kernels.  Run against the oracle this pins the raw-chain storage
scheme, the materialisation rule, the drop/evict/relax decisions and the host
bookkeeping (pending flag, list-length bound) without a GPU.  The product
never uses this model; the -m gpu tests check the real kernels.
"""
import numpy as np
import pytest

import scenarios as S
from model import ModelNKA
from oracle import api


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_model_matches_oracle(name, lazy):
    n, mvec, vtol, mk = S.SCENARIOS[name]
    ops = mk()
    inputs = [op[1] for op in ops if op[0] == "update"]
    serial, _ = S.run_ops(api.OracleNKA(n, mvec, vtol, dotmode=0), ops)
    arbiter, _ = S.run_ops(api.OracleNKA(n, mvec, vtol, dotmode=1), ops)
    scales, tols = S.tolerances(serial, arbiter, inputs)
    orc = api.OracleNKA(n, mvec, vtol, dotmode=1)
    mod = ModelNKA(n, mvec, vtol, lazy=lazy)
    it = 0
    for op in ops:
        if op[0] == "update":
            fin = op[1]
            a, b = fin.copy(), fin.copy()
            orc.accel_update(a)
            mod.accel_update(b)
            assert (orc.ndrop_last(), orc.relaxed_last(), orc.evicted_last()) == \
                   (mod.ndrop_last(), bool(mod.relaxed_last()), bool(mod.evicted_last()))
            assert np.linalg.norm(a - b) / scales[it] <= tols[it], (name, it)
            it += 1
        elif op[0] == "relax":
            orc.relax(); mod.relax()
        else:
            orc.restart(); mod.restart()
        assert orc.num_vec() == mod.num_vec()
        assert mod.defined() and mod.error() == 0
        assert mod.host_pending() == mod.dev_pending()
        assert mod.list_len() <= mod.ub_len()
    assert mod.bound_violations() == 0


def test_model_materialises_only_on_breaks():
    """No vtol drop, no relax, no s == 0: the chain is never materialised."""
    n, mvec, vtol, mk = S.SCENARIOS["iid_n1000_m10"]
    mod = ModelNKA(n, mvec, vtol)
    S.run_ops(mod, mk())
    assert mod.mat_entries() == 0
    n, mvec, vtol, mk = S.SCENARIOS["relax_restart_n96_m4"]
    mod = ModelNKA(n, mvec, vtol)
    S.run_ops(mod, mk())
    assert mod.mat_entries() > 0


def test_lazy_last_column_fixup_is_exercised_and_bit_identical():
    """Skipping the doomed oldest column must not change a single bit: when a vtol drop (or the
    s == 0 guard) keeps that column after all, the fix-up supplies its two dot products."""
    for name in ("picard_n500_m5_v2", "repeats_n128_m4", "mixed_n257_m5"):
        n, mvec, vtol, mk = S.SCENARIOS[name]
        lazy, eager = ModelNKA(n, mvec, vtol, lazy=True), ModelNKA(n, mvec, vtol, lazy=False)
        a, nva = S.run_ops(lazy, mk())
        b, nvb = S.run_ops(eager, mk())
        assert lazy.fixups() > 0 and eager.fixups() == 0
        assert nva == nvb
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
