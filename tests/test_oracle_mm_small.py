"""The oracle on real Maros-Meszaros problems (committed subset of the reference's fixtures, tests/golden/make_mm_small.py).
The reference's own test (tests/src/sparse/maros_meszaros_tests.cpp:20-36) asserts PIQP_SOLVED on every file with default
settings; the objective values the oracle reaches are the published Maros-Meszaros optima (e.g. DUAL1 3.50129662e-02,
HS118 664.82045, QADLITTL 4.80318859e+05, QISRAEL 2.53478378e+07, QSCAGR7 2.68659486e+07, VALUES -1.39662114)."""
import numpy as np
import pytest

from helpers import load_mm_small

PROBLEMS, GOLD = load_mm_small()
PUBLISHED = {"DUAL1": 3.50129662e-02, "DUAL2": 3.37336761e-02, "DUAL3": 1.35755839e-01, "DUAL4": 7.46090842e-01, "DUALC1": 6.15525083e+03,
             "HS118": 6.64820450e+02, "LOTSCHD": 2.39841589e+03, "QADLITTL": 4.80318859e+05, "QAFIRO": -1.59078179e+00, "QGROW7": -4.27987139e+07,
             "QISRAEL": 2.53478378e+07, "QPCBLEND": -7.84254092e-03, "QRECIPE": -2.66616000e+02, "QSC205": -5.81395349e-03, "QSCAGR7": 2.68659486e+07,
             "QSHARE1B": 7.20078318e+05, "QSHARE2B": 1.17036917e+04, "VALUES": -1.39662114e+00}


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_oracle_solves_mm_problem(oracle, name):
    s = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt"))
    s.setup(*PROBLEMS[name])
    assert s.solve() == 1 == GOLD[name]["status"]            # the reference's assertion
    r = s.result()
    assert r.info.iter == GOLD[name]["iter"]
    assert r.info.primal_obj == pytest.approx(GOLD[name]["primal_obj"], rel=1e-9, abs=1e-9)
    if name in PUBLISHED:                                    # published optimum (Maros & Meszaros 1999), solver tolerance 1e-8 rel
        assert r.info.primal_obj == pytest.approx(PUBLISHED[name], rel=2e-6, abs=1e-7)


@pytest.mark.parametrize("solver", ["sparse_ldlt_eq_cond", "sparse_ldlt_ineq_cond", "sparse_ldlt_cond"])
@pytest.mark.parametrize("name", ["DUALC1", "HS118", "LOTSCHD", "QAFIRO", "QPCBLEND", "QSC205"])
def test_oracle_condensed_modes_on_mm(oracle, name, solver):
    s = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver))
    s.setup(*PROBLEMS[name])
    assert s.solve() == 1
    assert s.result().info.primal_obj == pytest.approx(GOLD[name]["primal_obj"], rel=1e-6, abs=1e-7)


def test_oracle_solves_mid_size_mm_problems(oracle):
    """mid-size real problems (n_kkt 1 500 .. 25 000): SOLVED like the reference asserts; the oracle takes the product's
    fill-reducing ordering (host-only symbolic phase) because its own exact minimum degree is meant for small problems"""
    import scipy.sparse as sp
    from helpers import load_mm_mid
    from piqp_b200.backend import sparse_ldlt_symbolic
    probs, gold = load_mm_mid()
    for name in ("CVXQP3_M", "STCQP2", "AUG3DCQP", "DTOC3"):
        a = probs[name]
        perm = sparse_ldlt_symbolic(sp.triu(a[0]), a[2], a[4])["perm"]
        s = oracle.SparseSolver(oracle.default_settings(kkt_solver="sparse_ldlt"), kkt_perm=perm); s.setup(*a)
        assert s.solve() == 1 == gold[name]["status"]
        r = s.result()
        assert r.info.iter == gold[name]["iter"]
        assert r.info.primal_obj == pytest.approx(gold[name]["primal_obj"], rel=1e-9, abs=1e-9)


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_oracle_dense_solver_on_mm_problem(oracle, name):
    """tests/src/dense/maros_meszaros_tests.cpp:20-60: the reference runs its DENSE solver on every Maros-Meszaros problem with
    n <= 1000 and p + m <= 1000 and asserts PIQP_SOLVED; the oracle's dense backend on the committed subset, against the sparse
    backend's objective"""
    P, c, A, b, G, h_l, h_u, x_l, x_u = PROBLEMS[name]
    d = lambda M: None if M is None else np.asarray(M.todense())
    s = oracle.DenseSolver()
    s.setup(d(P), c, d(A), b, d(G), h_l, h_u, x_l, x_u)
    assert s.solve() == 1
    assert s.result().info.primal_obj == pytest.approx(GOLD[name]["primal_obj"], rel=1e-6, abs=1e-7)
