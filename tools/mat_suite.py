"""The reference's `.mat` suites end to end through the batched sparse solver, with the reference's own status assertions
(SURVEY 8f rank 3):

  * 138 Maros-Meszaros QPs  -> PIQP_SOLVED                       (tests/src/sparse/maros_meszaros_tests.cpp:20-36)
  * 94 feasible Netlib LPs  -> PIQP_SOLVED                       (tests/src/sparse/netlib_lp_tests.cpp:23-38, infeasibility_threshold = 0.01)
  * 29 infeasible Netlib LPs -> PIQP_PRIMAL_/DUAL_INFEASIBLE     (netlib_lp_tests.cpp:40-55)

Fixture schema = include/piqp/utils/io_utils.hpp:59-94 (variables P, c, A, b, G, h_l, h_u, x_l, x_u of a MATLAB v5 file).

    python tools/mat_suite.py pack                      # build container: /root/reference/tests/data/**.mat -> tests/golden/_suites/*.npz
                                                        #   (33 MB, git-ignored; travels to the GPU box with the snapshot, where the reference is not mounted)
    python tools/mat_suite.py run [--suite mm|netlib_feas|netlib_infeas|all] [--max-kkt N] [--out gpurun_out/r02_mat_suite.json]

`run` loads a packed suite (or the .mat directory itself where it exists), solves every problem with `SparseSolverBatched`
(kkt_solver = sparse_ldlt, batch 1, default settings + the suite's overrides), compares with the asserted status and writes one JSON
(per problem: n, p, m, n_kkt, status, iter, primal_obj, setup / solve seconds; totals per suite).  Exit code 0 iff every status
matches.  The JSON is rewritten after every problem so that an interrupted run keeps what it has."""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/tests/data"
PACK = os.path.join(ROOT, "tests", "golden", "_suites")
SUITES = {
    "mm": dict(dir="maros_meszaros", expect=(1,), settings={}),
    "netlib_feas": dict(dir="netlib/data", expect=(1,), settings={"infeasibility_threshold": 0.01}),
    "netlib_infeas": dict(dir="netlib/infeas", expect=(-2, -3), settings={"infeasibility_threshold": 0.01}),
}


def load_mat(path):
    import scipy.io
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d = scipy.io.loadmat(path)
    g = lambda k: np.asarray(d[k], dtype=float).ravel()
    P, A, G = sp.csc_matrix(d["P"]), sp.csc_matrix(d["A"]), sp.csc_matrix(d["G"])
    for M in (P, A, G):
        M.sort_indices()
    return P, g("c"), A, g("b"), G, g("h_l"), g("h_u"), g("x_l"), g("x_u")


def pack():
    os.makedirs(PACK, exist_ok=True)
    for suite, cfg in SUITES.items():
        src = os.path.join(REF, cfg["dir"])
        arrays, names = {}, []
        for fn in sorted(os.listdir(src)):
            if not fn.endswith(".mat"):
                continue
            nm = fn[:-4]
            P, c, A, b, G, h_l, h_u, x_l, x_u = load_mat(os.path.join(src, fn))
            names.append(nm)
            for key, M in (("P", P), ("A", A), ("G", G)):
                arrays["%s/%s_indptr" % (nm, key)] = M.indptr.astype(np.int32)
                arrays["%s/%s_indices" % (nm, key)] = M.indices.astype(np.int32)
                arrays["%s/%s_data" % (nm, key)] = M.data.astype(np.float64)
            for key, v in (("c", c), ("b", b), ("h_l", h_l), ("h_u", h_u), ("x_l", x_l), ("x_u", x_u)):
                arrays["%s/%s" % (nm, key)] = v
            arrays["%s/dims" % nm] = np.array([P.shape[0], A.shape[0], G.shape[0]], dtype=np.int32)
        arrays["__names__"] = np.array(names)
        out = os.path.join(PACK, suite + ".npz")
        np.savez_compressed(out, **arrays)
        print("%-14s %3d problems -> %s (%.1f MB)" % (suite, len(names), out, os.path.getsize(out) / 1e6))


def problems(suite):
    """yields (name, setup-args) of a suite: from the reference's directory where it is mounted, else from the packed npz"""
    cfg = SUITES[suite]
    src = os.path.join(REF, cfg["dir"])
    if os.path.isdir(src):
        for fn in sorted(os.listdir(src)):
            if fn.endswith(".mat"):
                yield fn[:-4], load_mat(os.path.join(src, fn))
        return
    z = np.load(os.path.join(PACK, suite + ".npz"))
    for nm in [str(x) for x in z["__names__"]]:
        n, p, m = (int(v) for v in z[nm + "/dims"])
        mat = lambda k, r: sp.csc_matrix((z["%s/%s_data" % (nm, k)], z["%s/%s_indices" % (nm, k)], z["%s/%s_indptr" % (nm, k)]), shape=(r, n))
        v = lambda k: np.asarray(z["%s/%s" % (nm, k)], dtype=float)
        yield nm, (mat("P", n), v("c"), mat("A", p), v("b"), mat("G", m), v("h_l"), v("h_u"), v("x_l"), v("x_u"))


def run(suites, max_kkt, out, only=None):
    import piqp_b200
    assert piqp_b200.lib().b200_device_count() > 0, "mat_suite run needs a CUDA device"
    report = {"suites": {}, "problems": []}
    ok_all = True
    for suite in suites:
        cfg = SUITES[suite]
        tot = dict(n=0, match=0, skipped=[], mismatched=[], solve_s=0.0, setup_s=0.0)
        for nm, args in problems(suite):
            if only and nm not in only:
                continue
            P, c, A, b, G, h_l, h_u, x_l, x_u = args
            n, p, m = P.shape[0], A.shape[0], G.shape[0]
            if n + p + m > max_kkt:
                tot["skipped"].append(nm)
                continue
            s = piqp_b200.SparseSolverBatched(kkt_solver="sparse_ldlt")
            for k, v in cfg["settings"].items():
                setattr(s.settings, k, v)
            t0 = time.perf_counter()
            row = {"suite": suite, "name": nm, "n": n, "p": p, "m": m, "n_kkt": n + p + m}
            try:
                s.setup(1, P, c, A if p else None, b if p else None, G if m else None, h_l if m else None, h_u if m else None, x_l, x_u)
                t1 = time.perf_counter()
                info = s.solve()[0]
                t2 = time.perf_counter()
                row.update(status=int(info.status), iter=int(info.iter), primal_obj=float(info.primal_obj), setup_s=t1 - t0, solve_s=t2 - t1)
            except Exception as e:      # a problem the backend rejects is a mismatch, not a crash of the suite
                row.update(status=None, error=repr(e)[:300])
            del s
            row["match"] = row.get("status") in cfg["expect"]
            tot["n"] += 1; tot["match"] += bool(row["match"])
            tot["solve_s"] += row.get("solve_s", 0.0); tot["setup_s"] += row.get("setup_s", 0.0)
            if not row["match"]:
                tot["mismatched"].append((nm, row.get("status")))
            report["problems"].append(row)
            report["suites"][suite] = tot
            print("%-14s %-12s n_kkt %7d  status %s iter %s  setup %.2fs solve %.2fs %s" % (suite, nm, n + p + m, row.get("status"), row.get("iter"), row.get("setup_s", 0), row.get("solve_s", 0),
                                                                                             "" if row["match"] else "   <-- expected %s" % (cfg["expect"],)), flush=True)
            if out:
                json.dump(report, open(out, "w"), indent=1)
        ok_all &= tot["match"] == tot["n"]
        report["suites"][suite] = tot
    if out:
        json.dump(report, open(out, "w"), indent=1)
    print(json.dumps(report["suites"]))
    return 0 if ok_all else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["pack", "run"])
    ap.add_argument("--suite", default="all")
    ap.add_argument("--max-kkt", type=int, default=400000)
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.cmd == "pack":
        pack()
        return 0
    suites = list(SUITES) if a.suite == "all" else a.suite.split(",")
    return run(suites, a.max_kkt, a.out, set(a.only.split(",")) if a.only else None)


if __name__ == "__main__":
    sys.exit(main())
