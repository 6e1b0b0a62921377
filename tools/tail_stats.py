"""GPU-box helper: how well can a one-pair-per-CTA schedule balance the bench workload?  For every tracker step it takes the
per-stream work of the step (sum over levels of evaluations x (residuals + a fixed per-evaluation cost)) and list-schedules it
on 148 workers (a) in the order the kernel uses (sorted by the PREVIOUS step's work), (b) sorted by the step's own work
(perfect prediction), and prints makespan / mean load for both."""
import heapq, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, synth
import edge_alignment_b200 as ea

S, T, NL = int(os.environ.get("STREAMS", "592")), 24, 3
bgr, dep, _ = synth.make_sequences(S, T, seed=1234, device="cuda:0")
ctx = ea.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
tr = ea.Tracker(ctx, ea.frame_params(n_levels=NL), ea.solve_params(point_stride=1, loss_type=ea.LOSS_HUBER, loss_scale=0.1), S, 10)
fb, fd = 640 * 480 * 3 * S, 640 * 480 * 2 * S
def sched(work, order, P=148):
    h = [0.0] * P; heapq.heapify(h)
    for i in order:
        t = heapq.heappop(h); heapq.heappush(h, t + work[i])
    return max(h)
prev = None
allw = []
EVAL_FIXED = 6000.0     # point-equivalents of the serial section per evaluation (~3.5 us at ~20 us per 30k-point evaluation)
for t in range(T):
    tr.step_device(bgr.data_ptr() + t * fb, dep.data_ptr() + t * fd)
    if t == 0: continue
    _, Ss = tr.poses()
    w = np.zeros(S)
    for i, s in enumerate(Ss): w[i // NL] += s.evaluations * (s.n_residuals + EVAL_FIXED)
    mean = w.sum() / 148
    perfect = sched(w, np.argsort(-w))
    used = sched(w, np.argsort(-prev) if prev is not None else np.arange(S))
    print("step %2d: mean load %.2e  max pair %.2f of mean  makespan/mean: previous-step order %.3f, perfect order %.3f   cv of pair work %.2f"
          % (t, mean, w.max() / mean, used / mean, perfect / mean, w.std() / w.mean()))
    prev = w
    allw.append(w)

np.save(os.path.join(ROOT, 'gpurun_out', 'pair_work_%d.npy' % S), np.array(allw))
