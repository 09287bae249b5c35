"""Reference-side bindings: include/piqp_b200_adapter.hpp (KKTSolverBase adapter), integration/piqp_batched.{h,c}
(piqp_setup_dense_batched & co over the reference's own C types) and integration/piqp_python_batched.cpp (pybind11
DenseSolverBatched).  CPU: everything compiles and links -- against the REAL reference headers where they are mounted, against the
minimal mocks in tests/adapter_mock otherwise.  GPU: the built artefacts run and reproduce the reference's known answers."""
import os
import subprocess
import sys
import sysconfig

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "adapter_mock")
OUT = os.path.join(MOCK, "build")
REF_C = "/root/reference/interfaces/c/include"
LINK = ["-L", os.path.join(ROOT, "piqp_b200"), "-lpiqp_b200", "-Wl,-rpath," + os.path.join(ROOT, "piqp_b200")]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, " ".join(cmd) + "\n" + r.stdout + r.stderr
    return r.stdout


def build_adapter_demo():
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "adapter_demo")
    _run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", MOCK, os.path.join(MOCK, "adapter_demo.cpp")] + LINK + ["-o", exe])
    return exe


def build_batched_demo():
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "batched_demo")
    _run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(MOCK, "c"), "-I", os.path.join(ROOT, "integration"), "-I", os.path.join(ROOT, "include"),
          os.path.join(ROOT, "integration", "piqp_batched.c"), os.path.join(MOCK, "c", "batched_demo.c")] + LINK + ["-lm", "-o", exe])
    return exe


def build_pybind_module():
    import pybind11
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "piqp_batched" + sysconfig.get_config_var("EXT_SUFFIX"))
    src = os.path.join(ROOT, "integration", "piqp_python_batched.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        _run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], "-I", os.path.join(ROOT, "include"), src] + LINK + ["-o", so])
    return so


def test_adapter_header_compiles_against_the_plugin_interface():
    import piqp_b200
    piqp_b200.lib()
    build_adapter_demo()


def test_batched_c_binding_compiles_with_mock_and_with_the_reference_headers():
    import piqp_b200
    piqp_b200.lib()
    build_batched_demo()
    if os.path.isdir(REF_C):      # build container: the reference's own piqp.h / piqp_typedef.h; the static asserts in piqp_batched.c prove the layouts
        _run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", REF_C, "-I", os.path.join(ROOT, "integration"), "-I", os.path.join(ROOT, "include"),
              "-c", os.path.join(ROOT, "integration", "piqp_batched.c"), "-o", os.path.join(OUT, "piqp_batched_ref.o")])


def test_pybind_module_builds():
    build_pybind_module()


@pytest.mark.gpu
def test_adapter_drives_all_backends_through_kkt_solver_base(b200):
    out = _run([build_adapter_demo()])
    assert "ADAPTER_OK" in out, out


@pytest.mark.gpu
def test_batched_c_binding_reproduces_known_answers(b200):
    out = _run([build_batched_demo()])
    assert "BATCHED_BINDING_OK" in out, out


@pytest.mark.gpu
def test_pybind_dense_solver_batched(b200):
    """dense/solver_test.cpp:30-101 through the pybind module: setup / solve / update / solve on a batch of two"""
    from helpers import simple_qp, simple_qp_update
    so = build_pybind_module()
    sys.path.insert(0, os.path.dirname(so))
    import piqp_batched
    q1 = simple_qp(); q2 = simple_qp_update(q1)
    st = lambda a, b: np.stack([a, b])
    s = piqp_batched.DenseSolverBatched()
    s.setup(st(q1["P"], q2["P"]), st(q1["c"], q2["c"]), st(q1["A"], q2["A"]), st(q1["b"], q2["b"]), st(q1["G"], q2["G"]),
            st(q1["h_l"], q2["h_l"]), st(q1["h_u"], q2["h_u"]), st(q1["x_l"], q2["x_l"]), st(q1["x_u"], q2["x_u"]))
    assert s.solve() == [1, 1]
    r = s.result
    assert np.allclose(r["x"][0], [0.4285714, 0.2142857], atol=1e-6) and np.allclose(r["x"][1], [0.2763157, 0.0921056], atol=1e-6)
    s.update(P=st(q2["P"], q2["P"]), A=st(q2["A"], q2["A"]), h_u=st(q2["h_u"], q2["h_u"]), x_u=st(q2["x_u"], q2["x_u"]))
    assert s.solve() == [1, 1]
    assert np.allclose(s.result["x"][0], [0.2763157, 0.0921056], atol=1e-6)
