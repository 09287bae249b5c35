"""Iteration-count parity table on the committed Maros-Meszaros fixtures (tests/golden/mm_small.npz, mm_mid.npz).

Per problem and KKT mode:  oracle (its own min-degree ordering) | oracle under the PRODUCT's permutation (kkt_perm = what
b200_sparse_ldlt_symbolic_mode returns, i.e. the ordering the CUDA path factorises in) | oracle under the product's permutation
with the serial multifrontal restatement of the product's summation order (ORACLE_MULTIFRONTAL=1) | and, when a GPU is
visible, the CUDA path: default kernels, forced HBM fronts (blocked elimination), level-scheduled kernels.
The oracle's own spread over elimination orders is what the named exceptions in tests/test_gpu_mm_small.py are justified with.

  python tools/mm_iter_parity.py [--set small|mid] [--modes sparse_ldlt,...] [--out profiles/r02_mm_iter_parity.json]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def oracle_run(args, solver, perm=None, seed_perm=None):
    from oracle import pyoracle
    o = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver=solver), kkt_perm=perm)
    o.setup(*args)
    st = o.solve()
    r = o.result()
    return {"status": int(st), "iter": int(r.info.iter), "obj": float(r.info.primal_obj)}


def gpu_run(b200, args, solver, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        s = b200.SparseSolverBatched(kkt_solver=solver)
        s.setup(1, *args)
        info = s.solve()[0]
        return {"status": int(info.status), "iter": int(info.iter), "obj": float(info.primal_obj)}
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="small")
    ap.add_argument("--modes", default="sparse_ldlt")
    ap.add_argument("--out", default=None)
    ap.add_argument("--names", default=None)
    ap.add_argument("--random-perms", type=int, default=0, help="also run the oracle under K random permutations (its own spread)")
    a = ap.parse_args()
    import numpy as np
    import piqp_b200
    from piqp_b200.backend import sparse_ldlt_symbolic, SparseKKT
    from helpers import load_mm_small
    probs, gold = load_mm_small("mm_small" if a.set == "small" else "mm_mid")
    have_gpu = piqp_b200.lib().b200_device_count() > 0
    names = sorted(probs) if not a.names else a.names.split(",")
    rows = []
    for solver in a.modes.split(","):
        mode = SparseKKT.MODES[solver]
        for nm in names:
            args = probs[nm]
            P, A, G = args[0], args[2], args[4]
            sym = sparse_ldlt_symbolic(P, A, G, mode=mode)
            row = {"name": nm, "solver": solver, "n_kkt": int(len(sym["perm"])), "largest_front": sym["largest_front"]}
            row["oracle_own"] = oracle_run(args, solver)
            row["oracle_prodperm"] = oracle_run(args, solver, perm=sym["perm"])
            if a.random_perms:
                rng = np.random.default_rng(7)
                row["oracle_random"] = [oracle_run(args, solver, perm=rng.permutation(len(sym["perm"])).astype(np.int32))["iter"] for _ in range(a.random_perms)]
            if have_gpu:
                row["gpu_default"] = gpu_run(piqp_b200, args, solver, {"B200_LDLT_LEVELS": "0"})
                row["gpu_hbm_fronts"] = gpu_run(piqp_b200, args, solver, {"B200_LDLT_LEVELS": "0", "B200_FRONT_SMEM_ROWS": "6"})
                row["gpu_levels"] = gpu_run(piqp_b200, args, solver, {"B200_LDLT_LEVELS": "1"})
            rows.append(row)
            f = lambda k: ("%4d%s" % (row[k]["iter"], "" if row[k]["status"] == 1 else "!")) if k in row else "   -"
            print("%-10s %-22s nk=%6d  own %s  prodperm %s  | gpu %s  hbm %s  lev %s  %s" % (
                nm, solver, row["n_kkt"], f("oracle_own"), f("oracle_prodperm"), f("gpu_default"), f("gpu_hbm_fronts"), f("gpu_levels"),
                row.get("oracle_random", "")), flush=True)
    if a.out:
        json.dump({"set": a.set, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
