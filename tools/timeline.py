"""Device timeline of one batched solve (B200_TIMELINE=1, include/piqp_b200.h: b200_timeline_dump).

    B200_TIMELINE=1 python tools/timeline.py --workload multistage [--batch 128] [--out gpurun_out/timeline_ms.txt]

There is no nsys in the image and ncu serialises launches (cold caches), so this is how the repo looks at a LIVE captured IP
iteration: every launch is followed by a one-thread kernel that stamps %globaltimer.  The script runs two solves of a bench.py
workload (the second one is reported), sorts the stamps by time, attributes to every kernel the time since the previous stamp
(= its duration + the launch gap in front of it + ~2 us of stamp) and prints (a) the totals per kernel over the last full IP iteration
that ran as a graph replay and (b) that iteration launch by launch."""
import argparse
import collections
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="multistage")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    assert os.environ.get("B200_TIMELINE"), "run with B200_TIMELINE=1"
    import torch
    import bench
    import piqp_b200
    argv, sys.argv = sys.argv, ["bench.py", "--workload", a.workload] + (["--batch", str(a.batch)] if a.batch else [])
    args = bench.workload_args(bench.parse(), a.workload)
    sys.argv = argv
    wl = {"dense": bench.DenseWorkload, "multistage": bench.MultistageWorkload, "sparse": bench.SparseWorkload, "sparse_c3": bench.SparseC3Workload}[a.workload](args)
    dev = torch.device("cuda:0")
    data = wl.device_data(args.batch, 1000, dev)
    s = wl.make_solver(0, data)
    s.solve()
    torch.cuda.synchronize()
    lib = piqp_b200.lib()
    tmp = a.out or "/tmp/b200_timeline.txt"
    lib.b200_timeline_dump(ctypes.c_char_p(tmp.encode()))      # drop setup + first solve
    t0 = time.perf_counter()
    s.solve()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n = lib.b200_timeline_dump(ctypes.c_char_p(tmp.encode()))
    rows = [l.rstrip("\n").split("\t") for l in open(tmp)]
    st = [(int(r[2]), r[1]) for r in rows]
    print("stamps %d, solve wall %.3f ms, iterations %s" % (n, wall * 1e3, [int(i.iter) for i in s.info()[:4]]))
    heads = [i for i, x in enumerate(st) if x[1] == "k_head"]
    print("k_head at %s" % heads[:40])
    if len(heads) >= 5:
        lo, hi = heads[2], heads[3]       # one full iteration between two convergence checks, a graph replay in the middle of the solve
        it = st[lo:hi + 1]
        agg = collections.OrderedDict()
        for prev, cur in zip(it[:-1], it[1:]):
            e = agg.setdefault(cur[1], [0, 0])
            e[0] += 1
            e[1] += cur[0] - prev[0]
        tot = sum(v[1] for v in agg.values())
        print("\none IP iteration (between the 3rd and 4th k_head stamps): %.1f us, %d launches" % (tot / 1e3, len(it) - 1))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("  %-44s x%-3d %8.1f us  %5.1f %%" % (k[:44], v[0], v[1] / 1e3, 100.0 * v[1] / tot))
        print("\nlaunch by launch:")
        for prev, cur in zip(it[:-1], it[1:]):
            print("  %-44s %7.1f us" % (cur[1][:44], (cur[0] - prev[0]) / 1e3))
    print("\nfirst to last stamp of the solve: %.1f us" % ((st[-1][0] - st[0][0]) / 1e3))
    agg = collections.OrderedDict()
    for prev, cur in zip(st[:-1], st[1:]):
        e = agg.setdefault(cur[1], [0, 0])
        e[0] += 1
        e[1] += cur[0] - prev[0]
    tot = sum(v[1] for v in agg.values())
    print("whole solve by kernel:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-44s x%-4d %9.1f us  %5.1f %%  (%.1f us each)" % (k[:44], v[0], v[1] / 1e3, 100.0 * v[1] / tot, v[1] / 1e3 / v[0]))


if __name__ == "__main__":
    main()
