#!/usr/bin/env python
"""Throughput against batch size for one workload (device-resident inputs): separates fixed costs and the tail of the persistent
kernel (a few configurations need 50x the mean work) from the steady-state rate.  Also prints the work-trace percentiles."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from openrave_b200.checker import BATCH_ENV, BATCH_SELF, CudaCollisionChecker  # noqa: E402

import torch  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "wam_lab1_100k"
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [12500, 25000, 50000, 100000, 200000, 400000, 1000000]
env, robot, _, seed = bench.load_world(workload)
cc = CudaCollisionChecker(env, 0)
env.SetCollisionChecker(cc)
bq = cc.bind(robot)
nmax = max(sizes)
lo, hi = robot.GetActiveDOFLimits()
rng = np.random.default_rng(seed)
qh = np.array(lo) + rng.random((nmax, len(lo))) * (np.array(hi) - np.array(lo))
q = torch.from_numpy(qh).cuda()
v = torch.zeros(nmax, dtype=torch.int32, device="cuda"); fe = torch.zeros_like(v); fs = torch.zeros_like(v)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
bq.check_device(q.data_ptr(), min(nmax, 100000), BATCH_ENV | BATCH_SELF, v.data_ptr(), fe.data_ptr(), fs.data_ptr(), 0, st.cuda_stream)
torch.cuda.synchronize()
print("==", workload, os.environ.get("CC_TUNE_DEFER", ""))
for n in sizes:
    ms = []
    for rep in range(8):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        bq.check_device(q.data_ptr(), n, BATCH_ENV | BATCH_SELF, v.data_ptr(), fe.data_ptr(), fs.data_ptr(), 0, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        if rep >= 3:
            ms.append(e0.elapsed_time(e1))
    m = float(np.mean(ms))
    print("n=%8d  %8.3f ms  %7.2f M configs/s  (min %.3f ms)" % (n, m, n / m / 1e3, min(ms)))
bq.trace_work(True)
n = min(nmax, 200000)
vv, _, _, _ = bq.check(qh[:n], BATCH_ENV | BATCH_SELF)
w = bq.work_trace(n).astype(np.int64)
bq.trace_work(False)
print("work trace (%d configs): mean %.1f iterations, percentiles 50/90/99/99.9/max %s %d, share of configs > 64 iterations %.4f carrying %.3f of the work" % (
    n, w.mean(), [int(np.percentile(w, p)) for p in (50, 90, 99, 99.9)], int(w.max()), (w > 64).mean(), w[w > 64].sum() / max(1, w.sum())))
