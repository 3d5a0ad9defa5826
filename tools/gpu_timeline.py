#!/usr/bin/env python
"""Per-chunk timeline of one batched call (cc_scene_trace_timeline): proves (or refutes) the two-stream overlap the bench's
step time relies on.  Prints one JSON object: per chunk {lane, lane_free_ms, fk_end_ms, traversal_end_ms}, the step's
duration, the sum of the per-chunk [lane free, traversal end] intervals and the time both lanes were busy at once."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from openrave_b200.checker import BATCH_ENV, BATCH_SELF, CudaCollisionChecker  # noqa: E402

import torch  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "pr2_pr2test1_1m"
env, robot, n, seed = bench.load_world(workload)
cc = CudaCollisionChecker(env, 0)
env.SetCollisionChecker(cc)
bq = cc.bind(robot)
q = torch.from_numpy(bench.sample(robot, n, seed)).cuda()
v = torch.zeros(n, dtype=torch.int32, device="cuda"); fe = torch.zeros_like(v); fs = torch.zeros_like(v)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
for _ in range(4):
    bq.check_device(q.data_ptr(), n, BATCH_ENV | BATCH_SELF, v.data_ptr(), fe.data_ptr(), fs.data_ptr(), 0, st.cuda_stream)
torch.cuda.synchronize()
cc.TraceTimeline(True)
recs = []
for rep in range(3):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    bq.check_device(q.data_ptr(), n, BATCH_ENV | BATCH_SELF, v.data_ptr(), fe.data_ptr(), fs.data_ptr(), 0, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms, lane = cc.GetTimeline()
    step = e0.elapsed_time(e1)
    # time during which at least two lanes had a chunk in flight
    grid = np.linspace(0, step, 20001)
    busy = [np.zeros_like(grid, dtype=bool) for _ in range(max(2, max(lane) + 1))]
    for (a, b, c), l in zip(ms, lane):
        busy[l] |= (grid >= a) & (grid <= c)
    nbusy = np.sum(busy, axis=0)
    recs.append({"step_ms": step,
                 "chunks": [{"lane": int(l), "lane_free_ms": round(a, 4), "fk_end_ms": round(b, 4), "traversal_end_ms": round(c, 4)}
                            for (a, b, c), l in zip(ms, lane)],
                 "sum_of_chunk_intervals_ms": float(sum(c - a for a, b, c in ms)),
                 "sum_of_fk_intervals_ms": float(sum(b - a for a, b, c in ms)),
                 "lanes": len(busy),
                 "two_or_more_lanes_busy_ms": float((nbusy >= 2).mean() * step),
                 "any_lane_busy_ms": float((nbusy >= 1).mean() * step)})
print(json.dumps({"workload": workload, "n": n, "steps": recs}, indent=1))
