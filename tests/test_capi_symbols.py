"""CPU-only: the product library loads and exports every symbol include/piqp_b200.h declares."""
import os
import re


def test_library_exports_all_declared_symbols():
    import piqp_b200
    from piqp_b200._lib import SYMBOLS
    L = piqp_b200.lib()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "piqp_b200.h")).read()
    declared = set(re.findall(r"\b(b200(?:kkt|qp)?_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(SYMBOLS), declared ^ set(SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_settings_struct_defaults_match_reference():
    """settings.hpp:45-82"""
    import ctypes as C
    import piqp_b200
    s = piqp_b200.Settings()
    piqp_b200.lib().b200qp_set_default_settings_dense(C.byref(s))
    assert (s.rho_init, s.delta_init, s.eps_abs, s.eps_rel) == (1e-6, 1e-4, 1e-8, 1e-9)
    assert (s.max_iter, s.max_factor_retires, s.preconditioner_iter, s.tau) == (250, 10, 10, 0.99)
    assert s.iterative_refinement_max_iter == 10 and s.iterative_refinement_min_improvement_rate == 5.0
    assert s.kkt_solver == 0
    piqp_b200.lib().b200qp_set_default_settings_sparse(C.byref(s))
    assert s.kkt_solver == 1


def test_no_device_is_an_error_not_a_fallback():
    """without a GPU the product must fail loudly (no CPU path exists)"""
    import numpy as np
    import piqp_b200
    if piqp_b200.lib().b200_device_count() > 0:
        return
    try:
        piqp_b200.DenseKKT(np.eye(3))
    except RuntimeError as e:
        assert "failed" in str(e)
    else:
        raise AssertionError("expected a RuntimeError without a CUDA device")
