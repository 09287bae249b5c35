"""GPU: the device-resident batched IP loop + sparse LDL^T CUDA backends on real Maros-Meszaros problems (BASELINE config 5
family; committed subset of the reference's fixtures).  Bar: the reference's assertion (PIQP_SOLVED), the oracle's iteration
count, |dx|_inf <= 1e-8 max(1, |x|_inf)."""
import numpy as np
import pytest

from helpers import load_mm_small

pytestmark = pytest.mark.gpu
PROBLEMS, GOLD = load_mm_small()
# LP-like problems whose minimiser is not unique (a flat optimal face): two correct solvers agree on the objective to 1e-13 but
# land on different points of the face (measured |dx| up to 1.6e-2, profiles/r01c_mm_small_diag.txt) -> objective parity there
DEGENERATE = {"QADLITTL", "QAFIRO", "QSC205", "QSHARE1B", "QSHARE2B", "QGROW7"}
# Numerically chaotic problems: LDL^T without pivoting at delta = 1e-10 loses ~10 digits and the iteration path depends on the
# summation order.  The ORACLE itself needs 29..152 iterations on QRECIPE under 1e-15 data perturbations and hits max_iter on
# QBEACONF under random elimination orders; the multifrontal kernels hit max_iter on both, the level-scheduled (left-looking,
# the reference's summation order) kernels solve them.  Tested with that kernel family; iteration parity is not asserted.
CHAOTIC = {"QBEACONF", "QRECIPE"}


def _check(oracle, b200, name, solver, batch=2, iter_parity=True):
    args = PROBLEMS[name]
    o = oracle.SparseSolver(oracle.default_settings(kkt_solver=solver)); o.setup(*args)
    status = o.solve(); ro = o.result()
    s = b200.SparseSolverBatched(kkt_solver=solver)
    s.setup(batch, *args)
    infos = s.solve(); r = s.result()
    for k in range(batch):
        assert infos[k].status == status == 1, (name, infos[k].status, status)
        assert abs(infos[k].primal_obj - ro.info.primal_obj) <= 1e-8 * max(1.0, abs(ro.info.primal_obj)), name
        if name in CHAOTIC or not iter_parity:
            continue
        if name in DEGENERATE:      # flat optimal face: the iteration count moves by a few with the elimination order / rounding (QSHARE2B: 18..23 on the CPU)
            assert abs(infos[k].iter - ro.info.iter) <= max(3, ro.info.iter // 3), (name, infos[k].iter, ro.info.iter)
        else:
            assert infos[k].iter == ro.info.iter, (name, infos[k].iter, ro.info.iter)
            assert np.abs(r.x[k] - ro.x).max() <= 1e-8 * max(1.0, np.abs(ro.x).max()), name
    assert np.array_equal(r.x[0], r.x[batch - 1])


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_mm_problem_sparse_ldlt(oracle, b200, name, monkeypatch):
    monkeypatch.setenv("B200_LDLT_LEVELS", "1" if name in CHAOTIC else "0")
    _check(oracle, b200, name, "sparse_ldlt")
    assert GOLD[name]["status"] == 1


@pytest.mark.parametrize("solver", ["sparse_ldlt_eq_cond", "sparse_ldlt_ineq_cond", "sparse_ldlt_cond"])
@pytest.mark.parametrize("name", ["DUALC1", "HS118", "LOTSCHD", "QAFIRO", "QPCBLEND", "QSC205"])
def test_mm_problem_condensed_modes(oracle, b200, name, solver, monkeypatch):
    """condensed KKT matrices of LP-like problems carry delta^-1 A^T A with delta = 1e-10: the iteration count is sensitive to
    the summation order there (QAFIRO / sparse_ldlt_cond differs by a few iterations), so the bar is status + objective"""
    monkeypatch.setenv("B200_LDLT_LEVELS", "0")
    _check(oracle, b200, name, solver, iter_parity=False)


def _mid():
    from helpers import load_mm_mid
    return load_mm_mid()


@pytest.mark.parametrize("name", ["CVXQP1_M", "CVXQP2_M", "CVXQP3_M", "STCQP2", "CONT-050", "AUG3DCQP", "QSHIP08L", "LISWET1", "DTOC3", "STADAT1"])
def test_mid_size_mm_problem(b200, name, monkeypatch):
    """real problems with n_kkt 1 500 .. 25 000 (fronts up to 512 rows: HBM fronts in the CTA-per-QP schedule; STCQP2 takes the
    whole-GPU schedule): SOLVED, the oracle's objective (tests/golden/mm_mid_golden.json) and its iteration count"""
    monkeypatch.setenv("B200_LDLT_LEVELS", "0")
    probs, gold = _mid()
    s = b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
    s.setup(1, *probs[name])
    info = s.solve()[0]
    g = gold[name]
    assert info.status == 1 == g["status"], (name, info.status)
    assert abs(info.primal_obj - g["primal_obj"]) <= 1e-7 * max(1.0, abs(g["primal_obj"])), (name, info.primal_obj, g["primal_obj"])
    assert abs(info.iter - g["iter"]) <= max(2, g["iter"] // 10), (name, info.iter, g["iter"])


def test_stcqp2_both_schedules_agree(b200, monkeypatch):
    """STCQP2 (largest front 512 rows): the CTA-per-QP schedule and the whole-GPU schedule reach the same solution"""
    probs, _ = _mid()
    xs = []
    for wide in ("0", "1"):
        monkeypatch.setenv("B200_LDLT_WIDE", wide); monkeypatch.setenv("B200_LDLT_LEVELS", "0")
        s = b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
        s.setup(1, *probs["STCQP2"])
        info = s.solve()[0]
        assert info.status == 1
        xs.append((s.result().x[0].copy(), info.iter))
    assert xs[0][1] == xs[1][1]
    assert np.abs(xs[0][0] - xs[1][0]).max() <= 1e-8 * max(1.0, np.abs(xs[0][0]).max())


@pytest.mark.xfail(strict=False, reason="unverified on the GPU: the no-FMA elimination loop was written after the round's GPU budget was spent; the CPU study "
                                         "(oracle/experimental_multifrontal.hpp) predicts SOLVED")
@pytest.mark.parametrize("name", sorted(CHAOTIC))
def test_chaotic_mm_problem_with_multifrontal_kernels(oracle, b200, name, monkeypatch):
    """QBEACONF / QRECIPE through the default (multifrontal) kernels: with the fused multiply-add of the pivot update they ran into
    max_iter; the update now rounds multiply and subtract separately like the reference (sparse/ldlt.hpp:151-158)"""
    monkeypatch.setenv("B200_LDLT_LEVELS", "0")
    s = b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
    s.setup(1, *PROBLEMS[name])
    info = s.solve()[0]
    assert info.status == 1, (name, info.status, info.iter)
    assert abs(info.primal_obj - GOLD[name]["primal_obj"]) <= 1e-6 * max(1.0, abs(GOLD[name]["primal_obj"]))
