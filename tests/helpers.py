import numpy as np

INF = np.inf


def simple_qp():
    """tests/src/dense/solver_test.cpp:30-57 (first QP of SimpleQPWithUpdate)"""
    P = np.array([[6., 0], [0, 4]]); c = np.array([-1., -4])
    A = np.array([[1., -2]]); b = np.array([0.])
    G = np.array([[1., 0], [1, 0], [1, 0]]); h_l = np.array([-1, -INF, -2]); h_u = np.array([INF, 1, 2.])
    x_l = np.array([-INF, -1]); x_u = np.array([INF, 1.])
    return dict(P=P, c=c, A=A, b=b, G=G, h_l=h_l, h_u=h_u, x_l=x_l, x_u=x_u)


def simple_qp_update(q):
    """solver_test.cpp:74-77"""
    q = {k: v.copy() for k, v in q.items()}
    q["P"][0, 0] = 8; q["A"][0, 1] = -3; q["h_u"][0] = 2; q["x_u"][1] = 2
    return q


def primal_infeasible_qp():
    """solver_test.cpp:107-128"""
    return dict(P=np.array([[6., 0], [0, 4]]), c=np.array([-1., -4]), A=np.array([[1., -2]]), b=np.array([0.]),
                G=np.array([[1., 0], [0, 1], [-1, 0], [0, -1]]), h_l=None, h_u=np.array([0., 2, 1, -1]), x_l=None, x_u=None)


def dual_infeasible_qp():
    """solver_test.cpp:134-154"""
    return dict(P=np.zeros((2, 2)), c=np.array([-1., -1]), A=None, b=None, G=np.array([[-1., 0], [0, -1]]),
                h_l=None, h_u=np.array([0., 0]), x_l=None, x_u=None)


def ill_conditioned_qp():
    """solver_test.cpp:156-182"""
    P = np.diag([61, 2e9, 61, 2e9, 1000, 100.])
    A = np.array([[1, 0, 1, 0, 1, 0], [2.4, 0, -2.4, 0, 0, 1.]])
    x_l = np.array([-2e4, -0.3491, -2e4, -0.3491, -INF, -INF]); x_u = np.array([2e4, 0.3491, 2e4, 0.3491, INF, INF])
    return dict(P=P, c=np.zeros(6), A=A, b=np.zeros(2), G=None, h_l=None, h_u=None, x_l=x_l, x_u=x_u)


def inf_bounds_qp():
    """solver_test.cpp:347-377"""
    G = np.array([[1, 0, 0, 0], [1, 0, -1, 0], [-1, 0, -1, 0], [-1, 0, 0, 0], [-1, 0, 1, 0], [1, 0, 1, 0.]])
    return dict(P=np.eye(4), c=np.ones(4), A=None, b=None, G=G, h_l=None, h_u=np.array([1, 1, 1, 1, INF, INF]), x_l=None, x_u=None)


def setup_args(q):
    return (q["P"], q["c"], q.get("A"), q.get("b"), q.get("G"), q.get("h_l"), q.get("h_u"), q.get("x_l"), q.get("x_u"))


def kkt_residuals(q, r):
    """max KKT violation of a returned primal-dual point (stationarity, feasibility, complementarity)."""
    P = np.triu(q["P"]); P = P + P.T - np.diag(np.diag(P))
    x = r.x
    g = P @ x + q["c"]
    if q.get("A") is not None and len(r.y):
        g = g + q["A"].T @ r.y
    if q.get("G") is not None and len(r.z_l):
        g = g + q["G"].T @ (r.z_u - r.z_l)
    g = g + r.z_bu - r.z_bl
    res = [np.abs(g).max()]
    if q.get("A") is not None and len(r.y):
        res.append(np.abs(q["A"] @ x - q["b"]).max())
    if q.get("G") is not None and len(r.z_l):
        Gx = q["G"] @ x
        hl = q["h_l"] if q.get("h_l") is not None else np.full(len(Gx), -INF)
        hu = q["h_u"] if q.get("h_u") is not None else np.full(len(Gx), INF)
        res.append(max(0.0, (hl - Gx).max(), (Gx - hu).max()))
    xl = q["x_l"] if q.get("x_l") is not None else np.full(len(x), -INF)
    xu = q["x_u"] if q.get("x_u") is not None else np.full(len(x), INF)
    res.append(max(0.0, (xl - x).max(), (x - xu).max()))
    return max(res)


def load_scenario_mpc():
    """tests/golden/scenario_mpc_small.npz (+ golden JSON): the QP of the reference's documentation notebook."""
    import json
    import os
    import scipy.sparse as sp
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = np.load(os.path.join(here, "scenario_mpc_small.npz"))
    g = json.load(open(os.path.join(here, "scenario_mpc_small_golden.json")))
    n, p = int(d["n"]), int(d["p"])
    P = sp.csc_matrix((d["P_data"], d["P_indices"], d["P_indptr"]), shape=(n, n))
    A = sp.csc_matrix((d["A_data"], d["A_indices"], d["A_indptr"]), shape=(p, n))
    return dict(P=P, c=d["c"], A=A, b=d["b"], G=None, h_l=None, h_u=None, x_l=d["x_l"], x_u=d["x_u"]), g


def trace_as_printed(t):
    """oracle/b200 trace rows (rho, delta, mu, p_step, d_step, prim_res, dual_res, prim_obj, dual_obj, gap) ->
    the column order the reference prints (prim_obj dual_obj gap prim_res dual_res rho delta mu p_step d_step)"""
    return np.column_stack([t[:, 7], t[:, 8], t[:, 9], t[:, 5], t[:, 6], t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]])


def load_mm_mid():
    """mid-size real Maros-Meszaros problems (tests/golden/make_mm_mid.py)"""
    return load_mm_small("mm_mid")


def load_mm_small(stem="mm_small"):
    """the committed subset of the reference's Maros-Meszaros fixtures (tests/golden/make_mm_small.py):
    {name: (P, c, A, b, G, h_l, h_u, x_l, x_u)} as setup() arguments, plus the golden table"""
    import json
    import os
    import numpy as np
    import scipy.sparse as sp
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(here, stem + ".npz"))
    gold = json.load(open(os.path.join(here, stem + "_golden.json")))
    out = {}
    for g in gold["problems"]:
        nm = g["name"]
        n, p, m = (int(v) for v in z[nm + "/dims"])
        mat = lambda k, r: sp.csc_matrix((z["%s/%s_data" % (nm, k)], z["%s/%s_indices" % (nm, k)], z["%s/%s_indptr" % (nm, k)]), shape=(r, n))
        P, A, G = mat("P", n), mat("A", p), mat("G", m)
        v = lambda k: np.asarray(z["%s/%s" % (nm, k)], dtype=float)
        out[nm] = (P, v("c"), A if p else None, v("b") if p else None, G if m else None, v("h_l") if m else None, v("h_u") if m else None, v("x_l"), v("x_u"))
    return out, {g["name"]: g for g in gold["problems"]}
