"""diagnostic: CUDA vs oracle on the committed Maros-Meszaros subset (x error, objective difference, iterations)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import piqp_b200
from oracle import pyoracle
from helpers import load_mm_small
P, G = load_mm_small()
for name in sorted(P):
    o = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt")); o.setup(*P[name]); st = o.solve(); ro = o.result()
    s = piqp_b200.SparseSolverBatched(kkt_solver="sparse_ldlt"); s.setup(2, *P[name]); infos = s.solve(); r = s.result()
    ex = np.abs(r.x[0] - ro.x).max() / max(1.0, np.abs(ro.x).max())
    print("%-10s status %d/%d iter %d/%d  xerr %.2e  obj %.12g / %.12g  rel %.1e" % (name, infos[0].status, st, infos[0].iter, ro.info.iter, ex, infos[0].primal_obj, ro.info.primal_obj,
          abs(infos[0].primal_obj - ro.info.primal_obj) / max(1.0, abs(ro.info.primal_obj))))
