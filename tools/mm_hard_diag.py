"""diagnostic for the numerically hard Maros-Meszaros problems: where does the CUDA path leave the oracle's iteration path?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import piqp_b200
from oracle import pyoracle
from piqp_b200.backend import c_abi_vtable
from helpers import load_mm_small
P, G = load_mm_small()
vt = pyoracle.BackendVTable()
for k, v in c_abi_vtable().items():
    setattr(vt, k, v)
names = sys.argv[1:] or ["QBEACONF", "QRECIPE"]
for name in names:
    a = P[name]
    o = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt")); o.setup(*a); st = o.solve(); to = np.array(o.trace())
    print(name, "oracle:", st, o.result().info.iter)
    g = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt"), backend_vtable=vt); g.setup(*a); st = g.solve()
    print(name, "oracle IP loop + CUDA backend (C-ABI):", st, g.result().info.iter)
    for env in ({}, {"B200_LDLT_LEVELS": "1"}, {"B200_LDLT_NO_AMALG": "1"}):
        for k2, v2 in env.items():
            os.environ[k2] = v2
        for ks in ("sparse_ldlt", "sparse_ldlt_cond"):
            s = piqp_b200.SparseSolverBatched(kkt_solver=ks); s.settings.verbose = 2
            s.setup(1, *a); infos = s.solve()
            print(name, env, ks, "batched CUDA:", infos[0].status, infos[0].iter)
            if not env and ks == "sparse_ldlt":
                tg = np.array(s.trace(0))
                k = min(len(to), len(tg))
                dev = [i for i in range(k) if not np.allclose(tg[i, :7], to[i, :7], rtol=1e-3, atol=1e-12)]
                first = dev[0] if dev else k
                print("   first deviating iteration:", first)
                for i in range(max(0, first - 1), min(k, first + 3)):
                    print("   it %2d oracle" % i, ["%.3e" % v for v in to[i, :9]])
                    print("         cuda  ", ["%.3e" % v for v in tg[i, :9]])
        for k2 in env:
            del os.environ[k2]
