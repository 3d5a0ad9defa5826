#!/usr/bin/env python
"""Full-size parity sweeps: the CUDA path against the CPU oracle on BASELINE.json's own batches.

  python tests/parity_sweep.py --workload all --out-dir profiles --tag r02      # on the GPU box

For every collision workload of bench.py (C1 WAM/lab1 100k, C2 PR2/pr2test1 1M, C4 HIRO-NX 1M, C5 clutter — 20k of its batch
by default: the reference builds 1.9M-triangle PQP models per worker) every configuration of the seeded batch goes through
`CudaCollisionChecker.CheckCollisionBatch` (C ABI, device kernels) and through `OracleChecker(kind="reference")` (unmodified
PQP) on all host cores, with the contact-band classifier on for EVERY configuration, and the record

  {compared, identical, near_contact, verdict_differs, first_pair_differs, mismatch_inside_band,
   mismatch_outside_band (must be 0), first_pair_mismatch (outside the band, must be 0)}

is written to <out-dir>/<tag>_parity_<workload>.json.  For C3 (BiRRT on wamtest1) the record counts plans that are
waypoint-for-waypoint identical between the batched GPU planner and the serial early-exit CPU loop.
Test infrastructure: this script lives under tests/ because it drives the oracle.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def sweep_collision(workload, n=None, workers=None, kind=None):
    import bench
    from oracle import parity
    from openrave_b200.checker import CudaCollisionChecker
    env, robot, nfull, seed = bench.load_world(workload)
    n = nfull if not n else min(n, nfull)
    q = bench.sample(robot, nfull, seed)[:n]           # a prefix of the very batch bench.py times on rank 0
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    cc.SetPairTables(False)                              # pass 1: every pair through the narrowphase
    t0 = time.perf_counter()
    got = cc.CheckCollisionBatch(robot, q)
    gpu_first_s = time.perf_counter() - t0
    cc.SetPairTables(True)                               # pass 2: free-space certificates in force (the bench configuration)
    cc.CheckCollisionBatch(robot, q[:min(n, 20000)])
    t0 = time.perf_counter()
    got2 = cc.CheckCollisionBatch(robot, q)
    gpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(got["verdict"], got2["verdict"]) and np.array_equal(got["first_env"], got2["first_env"])
                and np.array_equal(got["first_self"], got2["first_self"]))
    bq = got2["query"]
    ncert = bq.table_count()
    ref = parity.oracle_pass(workload, q, kind=kind, workers=workers, want_band=True)
    rec = parity.compare(got2, ref, lambda b: cc.body_from_id(b).name)
    rec.update(workload=workload, description=bench.WORKLOADS[workload][4], seed=seed, batch=nfull,
               gpu_wall_s=round(gpu_s, 3), gpu_wall_s_without_certificates=round(gpu_first_s, 3),
               gpu_pass_with_and_without_certificates_identical=same,
               pairs_env=int(bq.info.npairs_env), pairs_self=int(bq.info.npairs_self), pairs_hoisted=int(bq.info.npairs_hoisted),
               pairs_certified_free=int(bq.dropped_pairs), pairs_with_table=int(ncert - bq.dropped_pairs),
               ref_bv_tests_per_config=float(ref["counters_sum"][0]) / n, ref_tri_tests_per_config=float(ref["counters_sum"][1]) / n,
               oracle_configs_per_s_per_core_with_band_classifier=n / max(ref["cpu_s"], 1e-9))
    return rec


def sweep_birrt(workload, nproblems=None):
    import bench
    import oracle
    from openrave_b200 import planning as pl
    from openrave_b200.checker import CudaCollisionChecker
    scene, nfull, seed, desc = bench.PLAN_WORKLOADS[workload]
    m = nfull if not nproblems else min(nproblems, nfull)
    kind = "reference" if oracle.available("reference") else "port"
    env, robot, cc, q = bench.plan_problems(scene, nfull, seed, 0, lambda e: CudaCollisionChecker(e, 0))
    free = q[cc.CheckCollisionBatch(robot, q)["verdict"] == 0]
    starts, goals = free[0:2 * nfull:2][:m].copy(), free[1:2 * nfull:2][:m].copy()
    seeds = np.arange(m, dtype=np.uint32)
    t0 = time.perf_counter()
    gpu = pl.BirrtPlanner(robot, cc).PlanBatch(starts, goals, seeds)
    gpu_s = time.perf_counter() - t0
    env2, robot2, orc, _ = bench.plan_problems(scene, nfull, seed, 0, lambda e: oracle.OracleChecker(e, kind))
    orc.InitEnvironment()
    fn_blk = orc.group_checker(robot2)
    t0 = time.perf_counter()
    cpu = pl.BirrtPlanner(robot2, fn_blk).PlanBatch(starts, goals, seeds)
    cpu_s = time.perf_counter() - t0
    same = [bool(a["status"] == b["status"] and np.array_equal(a["path"], b["path"]) and a["configs_serial"] == b["configs_serial"])
            for a, b in zip(cpu, gpu)]
    invalid = 0
    for r in gpu:
        if r["status"] == pl.PLAN_SOLVED and orc.check_batch(robot2, r["path"])["verdict"].any():
            invalid += 1
    return {"workload": workload, "description": desc, "compared": m, "identical": int(sum(same)),
            "plans_differing": [int(i) for i, s in enumerate(same) if not s],
            "solved_gpu": int(sum(r["status"] == pl.PLAN_SOLVED for r in gpu)),
            "solved_cpu": int(sum(r["status"] == pl.PLAN_SOLVED for r in cpu)),
            "gpu_plans_with_a_colliding_waypoint_under_the_oracle": invalid, "against": "oracle (%s), serial early-exit loop" % kind,
            "gpu_wall_s": round(gpu_s, 3), "oracle_wall_s": round(cpu_s, 2),
            "note": "a plan can differ only where some sample of some edge sat inside the 1e-4 m contact band"}


DEFAULT_N = {"wam_clutter200_1250k": 20000}


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="all")
    ap.add_argument("--n", type=int, default=0, help="configurations per workload (0 = the whole BASELINE batch)")
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--kind", default=None, choices=[None, "port", "reference"])
    ap.add_argument("--out-dir", default=os.path.join(ROOT, "profiles"))
    ap.add_argument("--tag", default="r02")
    args = ap.parse_args()
    names = (sorted(bench.WORKLOADS) + sorted(bench.PLAN_WORKLOADS)) if args.workload == "all" else args.workload.split(",")
    os.makedirs(args.out_dir, exist_ok=True)
    rc = 0
    for w in names:
        t0 = time.perf_counter()
        if w in bench.PLAN_WORKLOADS:
            rec = sweep_birrt(w, args.n)
        else:
            rec = sweep_collision(w, args.n or DEFAULT_N.get(w, 0), args.workers or None, args.kind)
            if rec["mismatch_outside_band"] or not rec["gpu_pass_with_and_without_certificates_identical"]:
                rc = 1
        rec["sweep_wall_s"] = round(time.perf_counter() - t0, 1)
        rec["host_cores"] = os.cpu_count()
        path = os.path.join(args.out_dir, "%s_parity_%s.json" % (args.tag, w))
        with open(path, "w") as f:
            json.dump(rec, f, indent=1)
            f.write("\n")
        print(json.dumps(rec))
        sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(main())
