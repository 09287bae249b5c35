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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """the boundary is a C ABI: include/piqp_b200.h compiles as C99 (-pedantic) and a C program links against the library
    (examples/c_api_demo.c; without a device it prints the error and exits 0 -- no compute here)"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "piqp_b200.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    import piqp_b200
    piqp_b200.lib()
    libdir = os.path.join(root, "piqp_b200")
    exe = str(tmp_path / "c_api_demo")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c_api_demo.c"),
                    "-L", libdir, "-lpiqp_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    if piqp_b200.lib().b200_device_count() > 0:
        assert "status 1" in r.stdout and "0.42857" in r.stdout, r.stdout
    else:
        assert "no CUDA device" in r.stdout
