"""Single-pair latency of BASELINE configs[0] (bundled pair 1->3, single level) through the C ABI: preprocessing of both
frames and the solve, per kernel variant.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edge_alignment_b200 as ea

z = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))
ctx = ea.Context(0)
fs = ea.FrameSet(ctx, ea.frame_params(), 2)
bgr = np.ascontiguousarray(z["bgr"][[0, 2]]); dep = np.ascontiguousarray(z["depth"][[0, 2]])
out = {}
def timeit(fn, n=30):
    for _ in range(5): fn()
    ctx.sync(); t = time.perf_counter()
    for _ in range(n): fn()
    ctx.sync(); return (time.perf_counter() - t) / n * 1e3
out["preprocess_both_frames_ms"] = timeit(lambda: fs.preprocess_host([0, 1], bgr, dep, ea.ROLE_BOTH))
for stride in (30, 1):
    for kern, name in ((-1, "task_graph"), (1, "cta_per_pair"), (2, "cluster2"), (4, "cluster4"), (8, "cluster8"), (0, "auto")):
        sp = ea.solve_params(point_stride=stride, cluster_size=kern)
        ms = timeit(lambda: ctx.solve_batch(fs, [0], fs, [1], None, sp))
        poses, S = ctx.solve_batch(fs, [0], fs, [1], None, sp)
        out["solve_stride%d_%s_ms" % (stride, name)] = ms
        out["iterations_stride%d" % stride] = S[0][0]["iterations"]
try:
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    K = tuple(float(v) for v in z["K"])
    xyz, _ = O.get_aX(z["bgr"][0], z["depth"][0], K); dt, _ = O.get_distance_transform(z["bgr"][2])
    for stride in (30, 1):
        t = time.perf_counter(); O.solve(xyz, dt, K, ea.IDENTITY, stride=stride); out["oracle_cpu_solve_stride%d_ms" % stride] = (time.perf_counter() - t) * 1e3
    t = time.perf_counter(); O.get_aX(z["bgr"][0], z["depth"][0], K); O.get_distance_transform(z["bgr"][2]); out["oracle_cpu_preprocess_ms"] = (time.perf_counter() - t) * 1e3
except Exception as e:
    out["oracle_error"] = str(e)
print(json.dumps(out))
