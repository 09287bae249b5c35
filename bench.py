#!/usr/bin/env python
"""bench.py -- measures the KKT factor+solve hot path (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload dense|multistage|sparse|sparse_c3] [--batch B] ...

Default workload (the one BASELINE.json's metric is quoted on, config 2): dense QPs n=1024, p=0, m=512, batch=256 per
GPU.  One "step" = one batched interior-point solve() of the per-GPU batch, i.e. ~10-15 passes of the hot path
(assemble + factorise + 2 KKT solves + residual mat-vecs).  `value` = algorithmic factor+solve GFLOP/s
over the whole step with inputs resident in HBM; `e2e` = the same metric through the public C-ABI with HOST buffers
(setup H2D + Ruiz + solve + D2H).  `--workload multistage` runs BASELINE config 4 (MPC N=100, nx=12, nu=4, 128 QPs per
GPU) through the block-tridiagonal-arrow backend; `--workload sparse` a batch of 148 random sparse QPs (n_kkt = 850) and
`--workload sparse_c3` BASELINE config 3 (ONE sparse QP, n_kkt = 20 000, whole-GPU schedule) through the supernodal
multifrontal LDL^T backend.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "dense", "multistage", "sparse", "sparse_c3", "mm_suite"],
                    help="all (default): the headline dense config 2 as the line's value + configs 4, 3 and 5 under config.workloads")
    ap.add_argument("--sub-steps", type=int, default=0, help="timed steps of the secondary workloads of --workload all (0 = auto)")
    ap.add_argument("--mm-max-kkt", type=int, default=30000)
    ap.add_argument("--mm-replicas", type=int, default=8)
    ap.add_argument("--density", type=float, default=0.01, help="sparse workload: density of P_utri, A, G")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (weak scaling); 0 = workload default")
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--p", type=int, default=0)
    ap.add_argument("--m", type=int, default=512)
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--nx", type=int, default=12)
    ap.add_argument("--nu", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=0, help="QPs in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    return a


def workload_args(a, name):
    """per-workload copy of the arguments with that workload's default batch / shape"""
    b = argparse.Namespace(**vars(a))
    b.workload = name
    if b.batch == 0 or a.workload == "all":
        b.batch = {"dense": 256, "multistage": 128, "sparse": 148, "sparse_c3": 1}.get(name, 1)      # sparse: one CTA per QP, one QP per SM; sparse_c3: one QP over the whole GPU
        if a.workload == "all" and name == "dense" and a.batch:
            b.batch = a.batch
    if name == "sparse" and (b.n, b.p, b.m) == (1024, 0, 512):      # sparse defaults (random patterns fill in heavily: n_kkt=850 -> nnz(L)=53k, 308 etree levels; 10-17 IP iterations)
        b.n, b.p, b.m = 500, 100, 250
    return b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        num = lambda s: s.replace(".", "", 1).isdigit()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and num(r[1])]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and num(r[2])]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------------
class DenseWorkload:
    name = "dense"

    def __init__(self, a):
        self.a = a
        self.n, self.p, self.m = a.n, a.p, a.m

    def describe(self, B):
        return "dense batched QP n=%d p=%d m=%d batch=%d per GPU (BASELINE config 2), full IP solve per step" % (self.n, self.p, self.m, B)

    def work(self):   # SURVEY.md 8(d), per call and instance
        n, p, m = float(self.n), float(self.p), float(self.m)
        return n * n * m + n ** 3 / 3.0, 2 * n * n + 4 * n * m + 4 * n * p

    def working_set_gb(self, B):
        return B * (3 * self.n * self.n + self.n * (self.m + self.p)) * 8 / 1e9

    def device_data(self, B, seed0, dev):
        from piqp_b200.synth import dense_batch_torch
        return dense_batch_torch(B, self.n, self.p, self.m, seed0=seed0, device=dev)

    def _pick(self, d, conv):
        g = lambda k: conv(d[k])
        p, m = self.p, self.m
        return (g("P"), g("c"), g("A") if p else None, g("b") if p else None, g("G") if m else None, g("h_l") if m else None,
                g("h_u") if m else None, g("x_l"), g("x_u"))

    def make_solver(self, local, data, on_host=False):
        import piqp_b200
        s = piqp_b200.DenseSolverBatched(device=local)
        s.setup(*self._pick(data, (lambda t: t.numpy()) if on_host else (lambda t: t)))
        return s

    def host_data(self, data):
        return {k: v.cpu().pin_memory() for k, v in data.items()}

    def h2d_bytes(self, host):
        keys = ["c", "x_l", "x_u"] + (["A", "b"] if self.p else []) + (["G", "h_l", "h_u"] if self.m else [])
        n, B = self.n, host["P"].shape[0]
        # P: only the upper triangle crosses PCIe, as trapezoids of 128 rows (b200qp_setup_dense, capi.cu: load_problem)
        p_elems = sum((n - r0) * min(128, n - r0) for r0 in range(0, n, 128)) if n >= 256 else n * n
        return sum(host[k].numel() * 8 for k in keys) + B * p_elems * 8

    def cpu_solvers(self, n_qp, seed0, native):
        from oracle import pyoracle
        from piqp_b200.synth import dense_strongly_convex_qp
        out = []
        for i in range(n_qp):
            q = dense_strongly_convex_qp(self.n, self.p, self.m, seed=seed0 + i)

            def make(q=q):
                s = pyoracle.DenseSolver(native=native)
                s.setup(q["P"], q["c"], q["A"] if self.p else None, q["b"] if self.p else None, q["G"] if self.m else None,
                        q["h_l"] if self.m else None, q["h_u"] if self.m else None, q["x_l"], q["x_u"])
                return s
            out.append(make)
        return out

    roofline_kernel = "gemm_nt_t64_bulk_kernel<EPI_ASSEMBLE,true> (K = P + diag + G^T Z^-1 G, DMMA m8n8k4 fp64, 128x64 tiles, operands by TMA bulk copies into an mbarrier ring, two CTAs per SM)"


class MultistageWorkload:
    name = "multistage"

    def __init__(self, a):
        self.a = a
        self.N, self.nx, self.nu = a.horizon, a.nx, a.nu
        self._work = None

    def describe(self, B):
        return ("sparse_multistage MPC horizon N=%d nx=%d nu=%d (n=%d, p=%d) batch=%d per GPU (BASELINE config 4), full IP solve per step"
                % (self.N, self.nx, self.nu, self.N * (self.nx + self.nu) + self.nx, self.N * self.nx, B))

    def work(self):
        return self._work

    def working_set_gb(self, B):
        return B * self.N * (3 * 16 * 16 + 2 * 12 * 16) * 8 * 3 / 1e9

    def device_data(self, B, seed0, dev):
        import torch
        from piqp_b200.synth import mpc_batch
        d = mpc_batch(B, N=self.N, nx=self.nx, nu=self.nu, seed0=seed0)
        self.pat = d
        return {k: torch.from_numpy(np_).to(dev) for k, np_ in (("Ax", d["Ax"]), ("c", d["c"]), ("b", d["b"]), ("x_l", d["x_l"]), ("x_u", d["x_u"]))}

    def make_solver(self, local, data, on_host=False):
        import piqp_b200
        s = piqp_b200.SparseSolverBatched(device=local, kkt_solver="sparse_multistage")
        conv = (lambda t: t.numpy()) if on_host else (lambda t: t)
        B = data["c"].shape[0]
        s.setup(B, self.pat["P"], conv(data["c"]), self.pat["A"], conv(data["b"]), None, None, None, conv(data["x_l"]), conv(data["x_u"]), Ax=conv(data["Ax"]))
        w = s.work()
        self._work = (w[0], w[2])
        self.bytes = (w[1], w[3])
        return s

    def host_data(self, data):
        return {k: v.cpu().pin_memory() for k, v in data.items()}

    def h2d_bytes(self, host):
        return sum(v.numel() * 8 for v in host.values()) + self.pat["P"].nnz * 8 * host["c"].shape[0]

    def cpu_solvers(self, n_qp, seed0, native):
        import scipy.sparse as sp
        from oracle import pyoracle
        from piqp_b200.synth import mpc_batch
        d = mpc_batch(n_qp, N=self.N, nx=self.nx, nu=self.nu, seed0=seed0)
        out = []
        for k in range(n_qp):
            A = sp.csc_matrix((d["Ax"][k], d["A"].indices, d["A"].indptr), shape=d["A"].shape)

            def make(k=k, A=A):
                s = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_multistage"), native=native)
                s.setup(d["P"], d["c"][k], A, d["b"][k], None, None, None, d["x_l"][k], d["x_u"][k])
                return s
            out.append(make)
        if self._work is None:
            import ctypes as C
            s0 = out[0]()
            f = s0._L.orc_multistage_factor_flops(C.c_void_p(s0._h))
            n = d["n"]
            self._work = (f, 2.0 * n * 16)     # solve flops only used on the reference arm when no GPU ran: coarse
        return out

    roofline_kernel = "partitioned block-tridiagonal Cholesky: msw_factor_chain_kernel on K runs per QP (one warp per run, fronts in registers) + msp_spike_kernel (DMMA) + reduced chain; one launch group = one factorisation of the batch"


class SparseWorkload(MultistageWorkload):
    """BASELINE config 3 / 5 family: random sparse QPs sharing one pattern, kkt_solver = sparse_ldlt (KKT_FULL, n_kkt = n+p+m)."""
    name = "sparse"

    def __init__(self, a):
        self.a = a
        self.n, self.p, self.m, self.density = a.n, a.p, a.m, a.density
        self._work = None

    def describe(self, B):
        return ("sparse random QP n=%d p=%d m=%d density=%.3g%% (n_kkt=%d), sparse_ldlt, shared pattern, batch=%d per GPU (BASELINE config 3/5 family), full IP solve per step"
                % (self.n, self.p, self.m, 100 * self.density, self.n + self.p + self.m, B))

    def working_set_gb(self, B):
        return B * (self.bytes[0] if getattr(self, "bytes", None) else 0.0) / 1e9

    _keys = ("Px", "Ax", "Gx", "c", "b", "h_l", "h_u", "x_l", "x_u")

    def device_data(self, B, seed0, dev):
        import torch
        from piqp_b200.synth import sparse_batch
        self.pat = sparse_batch(B, self.n, self.p, self.m, self.density, seed0=42)      # one pattern for every rank
        return {k: torch.from_numpy(self.pat[k]).to(dev) for k in self._keys}

    def make_solver(self, local, data, on_host=False):
        import piqp_b200
        s = piqp_b200.SparseSolverBatched(device=local, kkt_solver="sparse_ldlt")
        g = (lambda k: data[k].numpy()) if on_host else (lambda k: data[k])
        B = data["c"].shape[0]
        s.setup(B, self.pat["P"], g("c"), self.pat["A"], g("b"), self.pat["G"], g("h_l"), g("h_u"), g("x_l"), g("x_u"), Px=g("Px"), Ax=g("Ax"), Gx=g("Gx"))
        w = s.work()
        self._work = (w[0], w[2])
        self.bytes = (w[1], w[3])
        return s

    def h2d_bytes(self, host):
        return sum(v.numel() * 8 for v in host.values())

    def cpu_solvers(self, n_qp, seed0, native):
        import scipy.sparse as sp
        from oracle import pyoracle
        from piqp_b200.synth import sparse_batch
        d = sparse_batch(n_qp, self.n, self.p, self.m, self.density, seed0=42)
        out = []
        mk = lambda pat, v: sp.csc_matrix((v, pat.indices, pat.indptr), shape=pat.shape)
        # ordering: the pattern is shared, so the fill-reducing permutation is computed once, OUTSIDE the timed region, by
        # the product's host-only symbolic phase and handed to every oracle solver (the oracle's own exact minimum-degree
        # code is O(n^3)-ish and would misrepresent the reference's AMD; both arms then factor the same L pattern)
        import piqp_b200
        perm = piqp_b200.sparse_ldlt_symbolic(d["P"], d["A"], d["G"])["perm"]
        for k in range(n_qp):
            mats = (mk(d["P"], d["Px"][k]), mk(d["A"], d["Ax"][k]), mk(d["G"], d["Gx"][k]))

            def make(k=k, mats=mats):
                s = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt"), native=native, kkt_perm=perm)
                s.setup(mats[0], d["c"][k], mats[1], d["b"][k], mats[2], d["h_l"][k], d["h_u"][k], d["x_l"][k], d["x_u"][k])
                return s
            out.append(make)
        if self._work is None:
            nnzL, flops = out[0]().ldlt_stats()[:2]
            self._work = (float(flops), 4.0 * nnzL + self.n + self.p + self.m)
        return out

    roofline_kernel = "mf_factor_kernel (supernodal multifrontal sparse LDL^T, one CTA per QP, fronts in shared memory, left-looking blocked elimination of the fronts beyond it)"


class SparseC3Workload(SparseWorkload):
    """BASELINE config 3: ONE sparse random QP n=10 000, p=m=5 000, density 1 % (n_kkt = 20 000; the random pattern fills in to
    a ~10 000-row dense root front, 333 GFLOP per factorisation), kkt_solver = sparse_ldlt.  The backend switches to its
    whole-GPU schedule (sparse_wide.cuh): the roofline kernel is the DMMA trailing update of the blocked LDL^T of the root.
    CPU arm: the reference's scalar up-looking LDL^T needs minutes per factorisation at this size, so the bounded CPU sample
    is the SAME family at n=2 000 (n_kkt = 4 000, 2.4 GFLOP per factorisation), reported in the same size-normalised GFLOP/s."""
    name = "sparse_c3"
    CPU_N = 2000

    def __init__(self, a):
        super().__init__(a)
        if (a.n, a.p, a.m) == (1024, 0, 512):
            self.n, self.p, self.m = 10000, 5000, 5000
        self._cpu_work = None

    def describe(self, B):
        return ("sparse random QP n=%d p=%d m=%d density=%.3g%% (n_kkt=%d), sparse_ldlt, batch=%d per GPU (BASELINE config 3), full IP solve per step, whole-GPU supernodal schedule"
                % (self.n, self.p, self.m, 100 * self.density, self.n + self.p + self.m, B))

    def cpu_solvers(self, n_qp, seed0, native):
        small = SparseWorkload(self.a)
        small.n, small.p, small.m, small.density = self.CPU_N, self.CPU_N // 2, self.CPU_N // 2, self.density
        out = small.cpu_solvers(n_qp, seed0, native)
        self._cpu_work = small.work()
        return out

    def cpu_work(self):
        return self._cpu_work

    def cpu_factor_solve_sample(self, threads):
        """CPU arm of config 3 (oracle port of sparse::KKT<FULL> + sparse::LDLt): `threads` solvers of the same family at n = CPU_N,
        each timing setup-free  1 x update_scalings_and_factor + 2 x backend solve  (one IP iteration's worth of the hot path) under
        the product's permutation; size-normalised GFLOP/s.  One factorisation at the full n = 10 000 needs ~5 minutes on a core."""
        from concurrent.futures import ThreadPoolExecutor
        import numpy as np
        from oracle import pyoracle
        try:
            pyoracle.build(native=True); native = True
        except Exception:
            native = False
        makers = self.cpu_solvers(threads, 1042, native)
        with ThreadPoolExecutor(max_workers=threads) as ex:
            solvers = list(ex.map(lambda mk: mk(), makers))
        ff, sf = self._cpu_work
        n, p, m = solvers[0].dims[:3]
        rng = np.random.default_rng(1)
        x_reg, z_reg = rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 2.0, m)
        r = (rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m))

        def one(s):
            s.backend_factor(0.9, x_reg, z_reg); s.backend_solve(*r); s.backend_solve(*r)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, solvers))
        dt = time.perf_counter() - t0
        return {"value": threads * (ff + 2 * sf) / dt * 1e-9, "unit": "GFLOP/s", "cores": threads, "kind": "port", "seconds": dt,
                "sample": "%d oracle solvers (one per thread), each 1 factorisation + 2 backend solves of the same family at n=%d, p=m=%d under the product's permutation, %s build"
                          % (threads, self.CPU_N, self.CPU_N // 2, "-march=native" if native else "x86-64-v3")}

    def cpu_sample_note(self):
        return "same family at n=%d, p=m=%d (one factorisation at n=%d takes minutes on one core)" % (self.CPU_N, self.CPU_N // 2, self.n)

    roofline_kernel = "gemm_nt_t64_bulk_kernel<EPI_SUB,true> (far updates F22 -= L21 D L21^T of the two-level blocked LDL^T of the root front, DMMA m8n8k4 fp64, K = 256; window updates K = 64)"


def cpu_oracle_sample(wl, n_qp, threads, seed0=1042):
    """The CPU restatement of the reference (oracle, kind "port"): `n_qp` QPs of the workload's shape, one solver per host
    thread (the reference is single-threaded per solve).  Two clocks: solve() only (comparable with `value`) and
    setup() + solve() (comparable with the GPU arm's `e2e`, which pays setup + H2D + solve + D2H).
    Returns (gflops, qps, seconds, iters, native, e2e_gflops, e2e_qps)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    try:
        pyoracle.build(native=True)
        native = True
    except Exception:
        native = False
    makers = wl.cpu_solvers(n_qp, seed0, native)
    ts = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        solvers = list(ex.map(lambda mk: mk(), makers))  # setup(): ctypes releases the GIL inside the oracle
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda s: s.solve(), solvers))      # ctypes releases the GIL inside orc_solve
    t1 = time.perf_counter()
    dt = t1 - t0
    ff, sf = wl.cpu_work() if hasattr(wl, "cpu_work") else wl.work()
    flops, iters = 0.0, []
    for s in solvers:
        i = s.info()
        flops += i.n_factor * ff + i.n_backend_solve * sf
        iters.append(int(i.iter))
    return flops / dt * 1e-9, n_qp / dt, dt, iters, native, flops / (t1 - ts) * 1e-9, n_qp / (t1 - ts)


def blas_proxy_dense(n, p, m, iters, n_qp, threads):
    """The closest stand-in for the reference's Eigen::LLT + Eigen GEBP products that exists in this image: the same sequence of
    dense kernels per IP iteration through OpenBLAS (scipy.linalg.blas / lapack), one QP per host thread, BLAS itself
    single-threaded (the reference is single-threaded per solve):  dsyrk (G^T Z^-1 G, n x m) + dpotrf (n) + 2 x (2 dtrsv + the
    G / A mat-vecs).  Returns (GFLOP/s with the algorithmic flop count of SURVEY 8d, QP/s, seconds)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from scipy.linalg import blas, lapack
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    rng = np.random.default_rng(0)
    probs = []
    for _ in range(n_qp):
        G = np.asfortranarray(rng.standard_normal((m, n))) if m else None
        A = np.asfortranarray(rng.standard_normal((p, n))) if p else None
        P = np.asfortranarray(np.eye(n) * (n + m))
        probs.append((P, G, A, rng.uniform(0.5, 2.0, max(m, 1)), rng.standard_normal(n)))

    def one(q):
        P, G, A, z, r = q
        for _ in range(iters):
            K = P.copy(order="F")
            if G is not None:
                W = np.asfortranarray(G * np.sqrt(z[:m])[:, None])
                K = blas.dsyrk(1.0, W, beta=1.0, c=K, trans=1, lower=1, overwrite_c=1)
            if A is not None:
                K = blas.dsyrk(1.0, A, beta=1.0, c=K, trans=1, lower=1, overwrite_c=1)
            L, info = lapack.dpotrf(K, lower=1, overwrite_a=1)
            for _s in range(2):
                x = r.copy()
                if G is not None:
                    x = x + G.T @ (z[:m] * (G @ r)[:m])
                x = blas.dtrsv(L, x, lower=1, trans=0); x = blas.dtrsv(L, x, lower=1, trans=1)
                if G is not None:
                    _ = G @ x
                if A is not None:
                    _ = A @ x
        return 0

    def run():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, probs))
        return time.perf_counter() - t0
    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            run(); dt = run()
    else:
        run(); dt = run()
    fn, fp, fm = float(n), float(p), float(m)
    flops = n_qp * iters * ((fn * fn * fm + fn ** 3 / 3.0) + 2 * (2 * fn * fn + 4 * fn * fm + 4 * fn * fp))
    return flops / dt * 1e-9, n_qp / dt, dt


def blas_proxy_root_front(f, algorithmic_flops):
    """config 3: 99 % of the factorisation is the dense root front of the fill-in; a supernodal CPU code would run it as one
    LAPACK dpotrf with every host core (OpenBLAS threading on).  Returns (GFLOP/s on the algorithmic flops, seconds)."""
    import numpy as np
    from scipy.linalg import lapack
    rng = np.random.default_rng(0)
    A = rng.standard_normal((f, 64))
    K = np.asfortranarray(A @ A.T + np.eye(f) * f)
    t0 = time.perf_counter()
    lapack.dpotrf(K, lower=1, overwrite_a=1)
    dt = time.perf_counter() - t0
    return algorithmic_flops / dt * 1e-9, dt


def run_reference(args, wl, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The real PIQP cannot be built
    here (Eigen absent), so this times the oracle port; rank 0 only, bounded sample per step."""
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 32)
    n_qp = args.cpu_sample or (threads if wl.name in ("dense", "sparse_c3") else 8 * threads)
    vals, qpss, secs, e2ev, e2eq = [], [], [], [], []
    native = False
    for it in range(args.warmup + args.steps):
        g, q, dt, iters, native, eg, eq = cpu_oracle_sample(wl, n_qp, threads, seed0=1042 + 1000 * it)
        if it >= args.warmup:
            vals.append(g); qpss.append(q); secs.append(dt); e2ev.append(eg); e2eq.append(eq)
        if it == 0 and dt > 40:      # keep the whole run within a few minutes
            n_qp = max(1, int(n_qp * 20 / dt))
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "KKT factor+solve GFLOP/s fp64", "value": v, "unit": "GFLOP/s", "qps": sum(qpss) / len(qpss),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.describe(n_qp) + " -- CPU: %d QPs per step on %d host threads" % (n_qp, threads)
                               + ((" (" + wl.cpu_sample_note() + ")") if hasattr(wl, "cpu_sample_note") else "")},
        "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": min(threads, n_qp), "kind": "port",
                         "sample": "%d QPs per step%s, one oracle solver per thread, %s build"
                                   % (n_qp, (" of " + wl.cpu_sample_note()) if hasattr(wl, "cpu_sample_note") else "", "-march=native" if native else "x86-64-v3")},
        "e2e": {"value": sum(e2ev) / len(e2ev), "unit": "GFLOP/s", "qps": sum(e2eq) / len(e2eq), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "what": "setup() + solve() per QP on the host (the GPU arm's e2e also pays setup)"},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def measure(args, wl, ctx):
    """one workload on this rank's GPU: device-resident value, e2e, roofline, cpu_baseline -> the JSON line (rank 0) or None"""
    import ctypes as C
    import torch
    import piqp_b200
    rank, world, local, dev, dist = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"], ctx["dist"]
    B = args.batch

    # The one collective of the path: rank 0 broadcasts the problem descriptor (global batch, seed base) over NCCL/NVLink at
    # setup; every rank then owns a contiguous shard of the global batch (weak scaling: `--batch` instances per GPU).
    # No collective follows inside the IP loop.
    from piqp_b200.distributed import ShardedBatch
    sb = ShardedBatch(B * world, 42, dist=dist, device=dev)
    B, seed0 = sb.local_batch, sb.seed0
    assert (sb.lo, sb.hi) == (rank * B, (rank + 1) * B)
    data = wl.device_data(B, seed0 + rank * B, dev)
    torch.cuda.synchronize()
    solver = wl.make_solver(local, data)
    ff, sf = wl.work()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        solver.solve()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    L0 = piqp_b200.lib().b200_kernel_launch_count()
    keys = ("factor_calls", "backend_solves", "factor_ms", "solve_ms", "total_ms", "assemble_ms", "assemble_launches", "cholesky_ms", "cholesky_calls",
            "backend_solve_ms", "ip_iterations", "lockstep_iterations")
    agg = {k: 0 for k in keys}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        infos = solver.solve()
        st = solver.stats()
        for k in keys:
            agg[k] += getattr(st, k)
    barrier()
    wall = time.perf_counter() - t0
    L1 = piqp_b200.lib().b200_kernel_launch_count()
    clocks = sampler.stop() if sampler else None
    # The timed region above runs the product's default path (IP iterations replayed as CUDA graphs, no per-phase events).  The
    # per-bucket / per-kernel-class device times behind `buckets` and `roofline` come from a SEPARATE profiled pass of the same
    # solves (events between the phases => stepwise launches), outside the timed region.
    solver.set_profiling(True)
    prof_steps = max(1, min(args.steps, 3))
    prof = {k: 0 for k in keys}
    for _ in range(prof_steps):
        solver.solve()
        st = solver.stats()
        for k in keys:
            prof[k] += getattr(st, k)
    solver.set_profiling(False)
    for k in ("factor_ms", "solve_ms", "assemble_ms", "assemble_launches", "cholesky_ms", "cholesky_calls", "backend_solve_ms"):
        agg[k] = prof[k] * (args.steps / float(prof_steps))
    agg["profiled_total_ms_per_step"] = prof["total_ms"] / prof_steps
    # device time of the timed region: the library's own CUDA events on ITS stream (total_ms); max over ranks
    dev_ms = torch.tensor([agg["total_ms"]], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    statuses = [i.status for i in infos]
    flops_local = agg["factor_calls"] * ff + agg["backend_solves"] * sf
    tot = torch.tensor([flops_local, float(B * args.steps), float(agg["ip_iterations"])], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    step_ms = float(dev_ms.item()) / args.steps
    gflops = float(tot[0].item()) / (float(dev_ms.item()) * 1e-3) * 1e-9
    qps = float(tot[1].item()) / (float(dev_ms.item()) * 1e-3)

    # ---- e2e through the public C-ABI with HOST (pinned) buffers: setup (H2D + pack/gather + Ruiz) + solve + D2H of x ----
    e2e = None
    if not args.no_e2e:
        host = wl.host_data(data)
        n_x = solver.n
        hx = torch.empty((B, n_x), dtype=torch.float64).pin_memory()
        h2d = wl.h2d_bytes(host)
        d2h = hx.numel() * 8
        times, fl = [], []
        for rep in range(1 + max(2, min(args.steps, 3))):
            barrier()
            t0 = time.perf_counter()
            s2 = wl.make_solver(local, host, on_host=True)
            s2.solve()
            piqp_b200._lib.check(s2._L.b200qp_get_result(s2._h, C.cast(hx.data_ptr(), piqp_b200._lib.dp), *([None] * 9), 0), "get_result")
            dt = time.perf_counter() - t0
            st2 = s2.stats()
            if rep > 0:
                times.append(dt); fl.append(st2.factor_calls * ff + st2.backend_solves * sf)
            del s2
        print("e2e repetitions (s): %s" % ["%.4f" % t for t in times], file=sys.stderr)
        what = "b200qp_setup_*(host pinned buffers) + b200qp_solve + b200qp_get_result(x -> host), per GPU batch"
        single_times = list(times)
        if wl.name == "dense" and B >= 8:
            # The same public calls with the per-GPU batch cut into 4 sub-batches, one handle (= one CUDA stream) and one host
            # thread each.  The setups (H2D copies: 3.2 GB per GPU in total, ~60 ms over PCIe) run one after the other; the
            # interior-point solve of a sub-batch starts as soon as its own setup is done and overlaps the copies of the next ones.
            # Same inputs (slices of the same pinned buffers), same outputs (slices of hx), every QP solved.
            from concurrent.futures import ThreadPoolExecutor
            nchunk = max(2, int(os.environ.get("B200_E2E_CHUNKS", "4")))
            bounds = [(B * c // nchunk, B * (c + 1) // nchunk) for c in range(nchunk)]
            chunks = [{k: v[lo:hi] for k, v in host.items()} for lo, hi in bounds]
            turn = [threading.Event() for _ in range(nchunk + 1)]

            use_turns = os.environ.get("B200_E2E_TURNS", "0") == "1"      # 1: bench-side turns (setups one after the other), 0: the library's H2D turnstile

            def run_chunk(ci):
                torch.cuda.set_device(local)
                if use_turns:
                    turn[ci].wait()
                try:
                    sc = wl.make_solver(local, chunks[ci], on_host=True)
                finally:
                    turn[ci + 1].set()
                infos3 = sc.solve()
                lo, hi = bounds[ci]
                piqp_b200._lib.check(sc._L.b200qp_get_result(sc._h, C.cast(hx[lo:hi].data_ptr(), piqp_b200._lib.dp), *([None] * 9), 0), "get_result")
                st3 = sc.stats()
                ok = all(i.status == 1 for i in infos3)
                fl3 = st3.factor_calls * ff + st3.backend_solves * sf
                del sc
                return fl3, ok
            ptimes, pfl = [], []
            with ThreadPoolExecutor(max_workers=nchunk) as ex:
                for rep in range(1 + max(2, min(args.steps, 3))):
                    barrier()
                    for ev in turn:
                        ev.clear()
                    t0 = time.perf_counter()
                    futs = [ex.submit(run_chunk, ci) for ci in range(nchunk)]
                    turn[0].set()
                    res = [f.result() for f in futs]
                    dt = time.perf_counter() - t0
                    assert all(r[1] for r in res)
                    if rep > 0:
                        ptimes.append(dt); pfl.append(sum(r[0] for r in res))
            print("e2e repetitions, %d pipelined sub-batches (s): %s" % (nchunk, ["%.4f" % t for t in ptimes]), file=sys.stderr)
            if max(ptimes) < max(times):
                times, fl = ptimes, pfl
                what = ("b200qp_setup_dense(host pinned buffers) + b200qp_solve + b200qp_get_result(x -> host) on %d sub-batches of the per-GPU batch, one handle / "
                        "stream / host thread each; the library's H2D turnstile gives the copies of one sub-batch the whole link while the previous ones equilibrate and solve" % nchunk)
        tmax = torch.tensor([max(times)], dtype=torch.float64, device=dev)      # conservative: slowest repetition
        fsum = torch.tensor([sum(fl) / len(fl)], dtype=torch.float64, device=dev)
        tsingle = torch.tensor([max(single_times)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(fsum, op=dist.ReduceOp.SUM); dist.all_reduce(tsingle, op=dist.ReduceOp.MAX)
        e2e = {"value": float(fsum.item()) / float(tmax.item()) * 1e-9, "unit": "GFLOP/s", "qps": B * world / float(tmax.item()),
               "seconds_per_step": float(tmax.item()), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "single_handle_qps": B * world / float(tsingle.item()), "what": what}

    if rank != 0:
        return None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    if wl.name == "dense":
        # FP64 peak: MEASURED_PEAKS.json has only HBM and bf16 figures; calibrate the fp64 pipe with cuBLAS DGEMM here.
        a = torch.randn(6144, 6144, dtype=torch.float64, device=dev); bmat = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, bmat)
        best = 1e9
        for _ in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, bmat); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fp64_peak = 2 * 6144 ** 3 / (best * 1e-3) * 1e-12
        del a, bmat
        n, p, m = float(wl.n), float(wl.p), float(wl.m)
        fl_launch = n * n * m * (agg["factor_calls"] / max(1, agg["assemble_launches"]))
        ms_launch = agg["assemble_ms"] / max(1, agg["assemble_launches"])
        achieved = fl_launch / (ms_launch * 1e-3) * 1e-12 if ms_launch > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("gemm_nt_t64_bulk_kernel_assemble_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": wl.roofline_kernel, "bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak if fp64_peak else None,
                    "peak_source": "cuBLAS DGEMM 6144^3 measured in this run (MEASURED_PEAKS.json carries no fp64 figure; DMMA.8x8x4 probe: 37.1 TFLOP/s, profiles/dmma_probe_b200.txt)",
                    "flops_per_launch": fl_launch, "ms_per_launch": ms_launch, "traffic": traffic, "traffic_source": "profiles/traffic.json (one ncu --set full capture, not measured in this run)",
                    "cholesky_tflops": (agg["factor_calls"] * n ** 3 / 3.0) / (agg["cholesky_ms"] * 1e-3) * 1e-12 if agg["cholesky_ms"] else None,
                    "backend_solve_gbs": (agg["backend_solves"] * 8.0 * (n * n + 2 * n * m + 2 * n * p)) / (agg["backend_solve_ms"] * 1e-3) * 1e-9 if agg["backend_solve_ms"] else None,
                    "hbm_peak_gbs": hbm_peak}
    elif wl.name == "sparse_c3":
        # the factorisation is one dense-contraction-dominated pass (99 % of its flops are the trailing updates of the root front):
        # tensor-pipe bound; achieved = algorithmic factor flops (sum c_j^2 + 2 c_j, unpadded) / time of the factor kernels
        a = torch.randn(6144, 6144, dtype=torch.float64, device=dev); bmat = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, bmat)
        best = 1e9
        for _ in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, bmat); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fp64_peak = 2 * 6144 ** 3 / (best * 1e-3) * 1e-12
        del a, bmat
        fb, sb = wl.bytes
        ms_launch = agg["cholesky_ms"] / max(1, agg["cholesky_calls"])
        achieved = ff / (ms_launch * 1e-3) * 1e-12 if ms_launch > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("sparse_c3_factor_bytes_per_factorisation")
        except Exception:
            pass
        roofline = {"kernel": wl.roofline_kernel, "bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak if fp64_peak else None,
                    "peak_source": "cuBLAS DGEMM 6144^3 measured in this run (MEASURED_PEAKS.json carries no fp64 figure)",
                    "flops_per_launch": ff, "ms_per_launch": ms_launch, "traffic": traffic,
                    "note": "one 'launch' = one numeric factorisation (level-parallel leaf fronts, pull-form extend-add, ~157 panel + trailing-update launch pairs on the root front); time = CUDA events around the whole factorisation",
                    "backend_solve_gbs": (agg["backend_solves"] * sb) / (agg["backend_solve_ms"] * 1e-3) * 1e-9 if agg["backend_solve_ms"] else None,
                    "hbm_peak_gbs": hbm_peak}
    else:
        fb, sb = wl.bytes
        by_launch = fb * (agg["factor_calls"] / max(1, agg["cholesky_calls"]))
        ms_launch = agg["cholesky_ms"] / max(1, agg["cholesky_calls"])
        achieved = by_launch / (ms_launch * 1e-3) * 1e-9 if ms_launch > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                "msw_factor_chain_kernel_bytes_per_launch" if wl.name == "multistage" else "mf_factor_kernel_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": wl.roofline_kernel, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": hbm_src, "bytes_per_launch": by_launch, "ms_per_launch": ms_launch, "traffic": traffic, "traffic_source": "profiles/traffic.json (one ncu --set full capture, not measured in this run)",
                    "note": ("latency-bound chain of N dependent 16x16 block steps per QP; algorithmic bytes = blocks read + factor written and read once" if wl.name == "multistage"
                             else "one launch = one numeric factorisation of every QP of the batch; algorithmic bytes = 12 nnz(L) written + 12 nnz(KKT) read per QP (SURVEY 8d); latency-bound walk over the supernodes"),
                    "factor_gflops": (agg["factor_calls"] * ff) / (agg["cholesky_ms"] * 1e-3) * 1e-9 if agg["cholesky_ms"] else None,
                    "backend_solve_gbs": (agg["backend_solves"] * sb) / (agg["backend_solve_ms"] * 1e-3) * 1e-9 if agg["backend_solve_ms"] else None}

    cpu = None
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N = 1 only (the driver's scaling runs do not repeat it)
        threads = min(os.cpu_count() or 1, 32)
        if wl.name == "sparse_c3":
            cpu = wl.cpu_factor_solve_sample(min(threads, 8))
            f_root = int(solver.symbolic()["largest_front"]) if hasattr(solver, "symbolic") else wl.n
            bg, bdt = blas_proxy_root_front(f_root, ff)
            cpu["blas_proxy"] = {"value": bg, "unit": "GFLOP/s", "cores": os.cpu_count(), "seconds": bdt,
                                 "what": "LAPACK dpotrf (OpenBLAS, all host threads) of one dense %d x %d front = the root front that carries 99 %% of this factorisation, "
                                         "credited with the algorithmic factor flops of the FULL-SIZE problem; what a supernodal CPU code would reach, not what the reference's "
                                         "scalar up-looking LDL^T does (measured once in the build container: 329 s for this factorisation on one core = 1.0 GFLOP/s)" % (f_root, f_root)}
        else:
            n_qp = args.cpu_sample or (min(threads, 16) if wl.name == "dense" else 16 * threads)
            g, q, dt, iters, native, eg, eq = cpu_oracle_sample(wl, n_qp, min(threads, n_qp))
            cpu = {"value": g, "unit": "GFLOP/s", "qps": q, "cores": min(threads, n_qp), "kind": "port", "setup_plus_solve_gflops": eg, "setup_plus_solve_qps": eq,
                   "sample": "%d QPs of %s (seeds 1042..), one oracle solver per thread, %.1f s, iters %s, %s build"
                             % (n_qp, wl.cpu_sample_note() if hasattr(wl, "cpu_sample_note") else "the same shape", dt, sorted(set(iters)), "-march=native" if native else "x86-64-v3")}
            if wl.name == "dense":
                it = max(1, int(round(agg["ip_iterations"] / float(B * args.steps))))
                bg, bq, bdt = blas_proxy_dense(wl.n, wl.p, wl.m, it, min(threads, 16), min(threads, 16))
                cpu["blas_proxy"] = {"value": bg, "unit": "GFLOP/s", "qps": bq, "cores": min(threads, 16), "seconds": bdt,
                                     "what": "per IP iteration dsyrk + dpotrf + 2 x (2 dtrsv + mat-vecs) through OpenBLAS (scipy), %d iterations per QP, one QP per host thread, "
                                             "BLAS single-threaded: stand-in for the reference's Eigen::LLT / GEBP path (Eigen is not in this image)" % it}

    line = {
        "metric": "KKT factor+solve GFLOP/s fp64", "value": gflops, "unit": "GFLOP/s", "qps": qps,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.describe(B),
                   "parallelism": "batch sharded over %d GPU(s), one NCCL broadcast at setup, no collective inside the IP loop" % world,
                   "l2": "per-step working set %.2f GB per GPU vs 126 MB L2%s" % (wl.working_set_gb(B), "" if wl.working_set_gb(B) > 0.2 else " (fits: factor blocks are L2-resident by design)"),
                   "ip_iterations_per_qp": agg["ip_iterations"] / float(B * args.steps), "statuses_all_solved": all(s == 1 for s in statuses)},
        "buckets": {"factor_gflops": agg["factor_calls"] * ff / (agg["factor_ms"] * 1e-3) * 1e-9 if agg["factor_ms"] else None,
                    "solve_gflops": agg["backend_solves"] * sf / (agg["solve_ms"] * 1e-3) * 1e-9 if agg["solve_ms"] else None,
                    "factor_ms_per_step": agg["factor_ms"] / args.steps, "solve_ms_per_step": agg["solve_ms"] / args.steps,
                    "assemble_ms_per_step": agg["assemble_ms"] / args.steps, "cholesky_ms_per_step": agg["cholesky_ms"] / args.steps,
                    "backend_solve_ms_per_step": agg["backend_solve_ms"] / args.steps, "wall_s": wall},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(L1 - L0), "clocks": clocks,
    }
    return line


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    classes = {"dense": DenseWorkload, "multistage": MultistageWorkload, "sparse": SparseWorkload, "sparse_c3": SparseC3Workload}
    primary = "dense" if args.workload in ("all", "mm_suite") else args.workload
    if args.impl == "reference":
        a = workload_args(args, primary)
        run_reference(a, classes[primary](a), rank)
        return
    import torch
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout for the one JSON line (NCCL prints its version banner to stdout otherwise)
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    ctx = {"rank": rank, "world": world, "local": local, "dev": dev, "dist": dist}
    if args.workload == "mm_suite":
        from tools import mm_suite
        line = mm_suite.run(ctx, max_kkt=args.mm_max_kkt, replicas=args.mm_replicas)
    else:
        a = workload_args(args, primary)
        line = measure(a, classes[primary](a), ctx)
    if args.workload == "all":
        # BASELINE configs 4, 3 and 5 ride in the same line (config.workloads): same measurement contract per workload, fewer steps,
        # so that the driver's N = 1, 2, 4, 8 runs see config 4 at batch 128 per GPU (1024 over 8) and the config-5 suite sharded 1 -> 8
        subs = {}
        sub_steps = args.sub_steps or max(3, min(args.steps, 10))
        for key, name in (("multistage_c4", "multistage"), ("sparse_c3", "sparse_c3")):
            a = workload_args(args, name)
            a.steps, a.warmup = (sub_steps, 3) if name == "multistage" else (max(2, sub_steps // 3), 3)
            t0 = time.perf_counter()
            try:
                sub = measure(a, classes[name](a), ctx)
            except Exception as e:      # a secondary workload must not take the headline down with it
                sub = {"error": repr(e)} if rank == 0 else None
            if rank == 0:
                sub["bench_seconds"] = time.perf_counter() - t0
                subs[key] = sub
        t0 = time.perf_counter()
        try:
            from tools import mm_suite
            sub = mm_suite.run(ctx, max_kkt=args.mm_max_kkt, replicas=args.mm_replicas)
        except Exception as e:
            sub = {"error": repr(e)} if rank == 0 else None
        if rank == 0:
            sub["bench_seconds"] = time.perf_counter() - t0
            subs["mm_suite"] = sub
            line["config"]["workloads"] = subs
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
