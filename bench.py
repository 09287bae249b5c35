#!/usr/bin/env python
"""bench.py -- measures the KKT factor+solve hot path (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--n N --m M --p P]

One "step" = one batched interior-point solve() of `batch` dense QPs per GPU (BASELINE config 2:
n=1024, p=0, m=512, batch=256), i.e. ~10-15 passes of the hot path (assemble + Cholesky + 2 KKT solves +
residual mat-vecs).  `value` = algorithmic factor+solve GFLOP/s over the whole step with inputs
resident in HBM; `e2e` = the same metric through the public C-ABI with HOST buffers (setup H2D + solve + D2H).
Prints ONE JSON line on rank 0.
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
    ap.add_argument("--batch", type=int, default=256, help="instances per GPU (weak scaling)")
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--p", type=int, default=0)
    ap.add_argument("--m", type=int, default=512)
    ap.add_argument("--cpu-sample", type=int, default=0, help="QPs in the CPU-baseline sample (0 = one per core, max 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def factor_flops(n, p, m):   # per factor call and instance
    return float(n) * n * m + float(n) ** 3 / 3.0


def solve_flops(n, p, m):    # per backend solve call and instance
    return 2.0 * n * n + 4.0 * n * m + 4.0 * n * p


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
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_sample(n, p, m, n_qp, threads, seed0=1042):
    """The CPU restatement of the reference (oracle, kind "port") solving `n_qp` QPs of the same shape with one
    solver per host thread; returns (gflops, qps, seconds, iters)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    from piqp_b200.synth import dense_strongly_convex_qp
    try:
        pyoracle.build(native=True)
        native = True
    except Exception:
        native = False
    qs = [dense_strongly_convex_qp(n, p, m, seed=seed0 + i) for i in range(n_qp)]
    solvers = []
    for q in qs:
        s = pyoracle.DenseSolver(native=native)
        s.setup(q["P"], q["c"], q["A"] if p else None, q["b"] if p else None, q["G"] if m else None,
                q["h_l"] if m else None, q["h_u"] if m else None, q["x_l"], q["x_u"])
        solvers.append(s)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda s: s.solve(), solvers))      # ctypes releases the GIL inside orc_solve
    dt = time.perf_counter() - t0
    flops = 0.0
    iters = []
    for s in solvers:
        i = s.info()
        flops += i.n_factor * factor_flops(n, p, m) + i.n_backend_solve * solve_flops(n, p, m)
        iters.append(int(i.iter))
    return flops / dt * 1e-9, n_qp / dt, dt, iters, native


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The real PIQP cannot be
    built here (Eigen absent), so this runs the oracle port; rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 32)
    n_qp = args.cpu_sample or threads
    vals, qpss, secs = [], [], []
    for it in range(args.warmup + args.steps):
        g, q, dt, iters, native = cpu_oracle_sample(args.n, args.p, args.m, n_qp, threads, seed0=1042 + 1000 * it)
        if it >= args.warmup:
            vals.append(g); qpss.append(q); secs.append(dt)
        if it == 0 and dt > 40:      # keep the whole run within a few minutes
            n_qp = max(1, int(n_qp * 20 / dt))
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "KKT factor+solve GFLOP/s fp64", "value": v, "unit": "GFLOP/s", "qps": sum(qpss) / len(qpss),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "dense batched QP n=%d p=%d m=%d (BASELINE config 2 shape), %d QPs per step on host threads" % (args.n, args.p, args.m, n_qp)},
        "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": threads, "kind": "port",
                         "sample": "%d QPs per step, one oracle solver per thread, %s build" % (n_qp, "-march=native" if native else "x86-64-v3")},
        "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import piqp_b200
    from piqp_b200.synth import dense_batch_torch
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    n, p, m, B = args.n, args.p, args.m, args.batch

    # The one collective of the path: rank 0 broadcasts the problem descriptor (shape, seed base, settings vector)
    # over NCCL/NVLink at setup; every rank then owns the contiguous shard [rank*B, (rank+1)*B) of the global batch.
    desc = torch.tensor([n, p, m, B, 42], dtype=torch.int64, device=dev)
    if dist:
        dist.broadcast(desc, src=0)
    n, p, m, B, seed0 = [int(v) for v in desc.tolist()]
    data = dense_batch_torch(B, n, p, m, seed0=seed0 + rank * B, device=dev)
    torch.cuda.synchronize()

    solver = piqp_b200.DenseSolverBatched(device=local)
    arg = lambda k: data[k] if (k not in ("A", "b") or p) and (k not in ("G", "h_l", "h_u") or m) else None
    solver.setup(data["P"], data["c"], arg("A"), arg("b"), arg("G"), arg("h_l"), arg("h_u"), data["x_l"], data["x_u"])
    solver.set_profiling(True)

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
    agg = dict(factor_calls=0, backend_solves=0, factor_ms=0.0, solve_ms=0.0, total_ms=0.0, assemble_ms=0.0, assemble_launches=0,
               cholesky_ms=0.0, backend_solve_ms=0.0, iters=0, lockstep=0)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        infos = solver.solve()
        st = solver.stats()
        agg["factor_calls"] += st.factor_calls; agg["backend_solves"] += st.backend_solves
        agg["factor_ms"] += st.factor_ms; agg["solve_ms"] += st.solve_ms; agg["total_ms"] += st.total_ms
        agg["assemble_ms"] += st.assemble_ms; agg["assemble_launches"] += st.assemble_launches
        agg["cholesky_ms"] += st.cholesky_ms; agg["backend_solve_ms"] += st.backend_solve_ms
        agg["iters"] += st.ip_iterations; agg["lockstep"] += st.lockstep_iterations
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    # device time of the timed region: the library's own CUDA events on ITS stream (total_ms) -- torch events only see torch's stream
    dev_ms = torch.tensor([agg["total_ms"]], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    L1 = piqp_b200.lib().b200_kernel_launch_count()
    clocks = sampler.stop() if sampler else None
    statuses = [i.status for i in infos]
    flops_local = agg["factor_calls"] * factor_flops(n, p, m) + agg["backend_solves"] * solve_flops(n, p, m)
    tot = torch.tensor([flops_local, float(B * args.steps), float(agg["iters"])], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    step_ms = float(dev_ms.item()) / args.steps
    gflops = float(tot[0].item()) / (float(dev_ms.item()) * 1e-3) * 1e-9
    qps = float(tot[1].item()) / (float(dev_ms.item()) * 1e-3)

    # ---- e2e through the public C-ABI with HOST (pinned) buffers: setup (H2D + Ruiz) + solve + D2H of x ----
    e2e = None
    if not args.no_e2e:
        host = {k: v.cpu().pin_memory() for k, v in data.items()}
        hx = torch.empty((B, n), dtype=torch.float64).pin_memory()
        harg = lambda k: host[k].numpy() if (k not in ("A", "b") or p) and (k not in ("G", "h_l", "h_u") or m) else None
        h2d = sum(host[k].numel() * 8 for k in host if harg(k) is not None)
        d2h = hx.numel() * 8
        times, fl = [], []
        for rep in range(1 + max(2, min(args.steps, 3))):
            barrier()
            t0 = time.perf_counter()
            s2 = piqp_b200.DenseSolverBatched(device=local)
            s2.setup(host["P"].numpy(), host["c"].numpy(), harg("A"), harg("b"), harg("G"), harg("h_l"), harg("h_u"), host["x_l"].numpy(), host["x_u"].numpy())
            s2.solve()
            import ctypes as C
            piqp_b200._lib.check(s2._L.b200qp_get_result(s2._h, C.cast(hx.data_ptr(), piqp_b200._lib.dp), *([None] * 9), 0), "get_result")
            dt = time.perf_counter() - t0
            st2 = s2.stats()
            if rep > 0:
                times.append(dt); fl.append(st2.factor_calls * factor_flops(n, p, m) + st2.backend_solves * solve_flops(n, p, m))
            del s2
        tt = torch.tensor([max(times), sum(fl) / len(fl)], dtype=torch.float64, device=dev)   # conservative: slowest repetition
        tmax = tt[:1].clone(); fsum = tt[1:].clone()
        if dist:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
        e2e = {"value": float(fsum.item()) / float(tmax.item()) * 1e-9, "unit": "GFLOP/s", "qps": B * world / float(tmax.item()),
               "seconds_per_step": float(tmax.item()), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "what": "b200qp_setup_dense(host pinned buffers) + b200qp_solve + b200qp_get_result(x -> host), per GPU batch"}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the KKT assembly contraction (gemm_nt_tile_kernel<EPI_ASSEMBLE>, DMMA) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # FP64 peak: MEASURED_PEAKS.json has only HBM and bf16 figures; calibrate the fp64 pipe with cuBLAS DGEMM here.
    a = torch.randn(6144, 6144, dtype=torch.float64, device=dev); bmat = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, bmat)
    best = 1e9
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, bmat); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fp64_peak_tflops = 2 * 6144 ** 3 / (best * 1e-3) * 1e-12
    del a, bmat
    asm_flops_per_launch = float(n) * n * m * (agg["factor_calls"] / max(1, agg["assemble_launches"]))
    asm_ms = agg["assemble_ms"] / max(1, agg["assemble_launches"])
    achieved = asm_flops_per_launch / (asm_ms * 1e-3) * 1e-12 if asm_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("gemm_nt_tile_kernel_assemble_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "gemm_nt_tile_kernel<EPI_ASSEMBLE,true> (K = P + diag + G^T Z^-1 G, DMMA m8n8k4 fp64)",
                "bound": "tensor", "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp64_peak_tflops if fp64_peak_tflops else None,
                "peak_source": "cuBLAS DGEMM 6144^3 measured in this run (MEASURED_PEAKS.json carries no fp64 figure; hbm_gbs=%s)" % peaks.get("hbm_gbs"),
                "flops_per_launch": asm_flops_per_launch, "ms_per_launch": asm_ms, "traffic": traffic,
                "cholesky_tflops": (agg["factor_calls"] * float(n) ** 3 / 3.0) / (agg["cholesky_ms"] * 1e-3) * 1e-12 if agg["cholesky_ms"] else None,
                "backend_solve_gbs": (agg["backend_solves"] * 8.0 * (float(n) * n + 2.0 * n * m + 2.0 * n * p)) / (agg["backend_solve_ms"] * 1e-3) * 1e-9 if agg["backend_solve_ms"] else None,
                "hbm_peak_gbs": peaks.get("hbm_gbs")}

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = min(cores, 32)
        n_qp = args.cpu_sample or min(threads, 16)
        g, q, dt, iters, native = cpu_oracle_sample(n, p, m, n_qp, min(threads, n_qp))
        cpu = {"value": g, "unit": "GFLOP/s", "qps": q, "cores": min(threads, n_qp), "kind": "port",
               "sample": "%d QPs of the same shape (seeds 1042..), one oracle solver per thread, %.1f s, iters %s, %s build"
                         % (n_qp, dt, sorted(set(iters)), "-march=native" if native else "x86-64-v3")}

    line = {
        "metric": "KKT factor+solve GFLOP/s fp64", "value": gflops, "unit": "GFLOP/s", "qps": qps,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "dense batched QP n=%d p=%d m=%d batch=%d per GPU (BASELINE config 2), full IP solve per step" % (n, p, m, B),
                   "parallelism": "batch sharded over %d GPU(s), no collective inside the IP loop" % world,
                   "l2": "per-step working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (B * (3 * n * n + n * m) * 8 / 1e9),
                   "ip_iterations_per_qp": agg["iters"] / float(B * args.steps), "statuses_all_solved": all(s == 1 for s in statuses)},
        "buckets": {"factor_gflops": agg["factor_calls"] * factor_flops(n, p, m) / (agg["factor_ms"] * 1e-3) * 1e-9 if agg["factor_ms"] else None,
                    "solve_gflops": agg["backend_solves"] * solve_flops(n, p, m) / (agg["solve_ms"] * 1e-3) * 1e-9 if agg["solve_ms"] else None,
                    "factor_ms_per_step": agg["factor_ms"] / args.steps, "solve_ms_per_step": agg["solve_ms"] / args.steps,
                    "assemble_ms_per_step": agg["assemble_ms"] / args.steps, "cholesky_ms_per_step": agg["cholesky_ms"] / args.steps,
                    "backend_solve_ms_per_step": agg["backend_solve_ms"] / args.steps, "wall_s": wall},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(L1 - L0), "clocks": clocks,
    }
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
