"""CPU tier: the C-ABI library loads and exports every symbol include/mpcqp_b200.h declares; default settings
and parameters read back; creating an engine without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpcqp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcqp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as G
    G.build()
    from intent_mpc_b200 import engine
    lib = engine.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mpcqp_b200.h but not exported"


def test_defaults_read_back():
    from intent_mpc_b200 import engine
    s = engine.default_settings()
    assert (s.rho, s.sigma, s.alpha, s.eps_abs, s.eps_rel) == (0.1, 1e-6, 1.6, 1e-3, 1e-3)
    assert (s.max_iter, s.scaling, s.check_termination, s.adaptive_rho_interval) == (4000, 10, 25, 25)
    p = engine.MpcParamsC()
    engine.load_library().mpcqp_default_mpc_params(C.byref(p))
    assert p.horizon == 30 and p.ts == 0.1 and p.max_acc == 20.0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from intent_mpc_b200 import engine
    with pytest.raises(engine.EngineError):
        engine.Engine(0)
