#!/usr/bin/env python
"""bench.py — collision-checked configs/sec (FK + env + self) on B200, one JSON line on rank 0.

  python bench.py --gpus 1 --steps 10 --warmup 3          # the product arm (CUDA, through the C ABI)
  python bench.py --impl reference --steps 3 --warmup 1   # the reference arm: pqprave's CPU path on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...       # one rank per GPU, scene replicated, configs sharded

A "step" is one pass of the hot path over one batch of synthetic uniform-random DOF vectors
(RobotConfigurationSampler semantics) of the workload named in `config.workload`.

  value      configs/s with the inputs already resident in HBM (cc_check_batch_device on torch's stream,
             CUDA events on that stream, L2 flushed between timed steps)
  e2e        same metric through the host-buffer entry cc_check_batch (pinned host q -> H2D -> kernel -> D2H of
             verdict + first pairs), i.e. the call an OpenRAVE-side caller makes
  roofline   ALGORITHMIC bytes (the reference's own NumBVTests/NumTriTests on the same inputs,
             f32 layout: 2*64 B per BV test, 2*36 B per triangle test, + 8*ndof + 16) / kernel time, against the
             measured HBM copy bandwidth.  The reference has no broadphase and enumerates all contacts, the
             kernel culls and stops at the first contact per pair, so `frac` can exceed 1: it measures how much
             of the reference's work the kernel avoids, not DRAM utilisation (`traffic` is the physical bytes).
             `roofline.physical` holds the physical fractions, each traceable to a file under profiles/ taken at HEAD
             (tools/harvest_r02.py): issue_frac (warp instructions per configuration x configs/s / issue slots),
             lane_efficiency (threads per executed instruction / 32), dram_frac (DRAM bytes per step / step time / HBM peak)
  cpu_baseline  the oracle (reference-PQP-linked when oracle/_ref exists) timed single-threaded on this host, plus `parity`:
             the CUDA results of the timed batch against that oracle on all host cores (a bounded sample) and the committed
             full-size sweep (profiles/r02_parity_<workload>.json)
  N > 1      `value`: every rank checks its own shard, results stay on its GPU (no data-path collective), max over ranks;
             `config.device_consumer`: the same with one overlapped ncclAllGather of the results per step;
             `e2e`: ONE process (rank 0) drives all N GPUs through a multi-device scene and one cc_check_batch call
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene fixture, n configs per step per GPU, seed, active manipulator or None = all DOFs, description)
    "pr2_pr2test1_1m": ("pr2test1", 1000000, 2, "rightarm_torso",
                        "PR2 (robots/pr2-beta-static.zae: 88 links, 39 DOF + 48 passive joints, 51.8k tris) in data/pr2test1.env.xml "
                        "(9 bodies): 1M uniform-random right-arm(7)+torso(1) configs, the other DOFs at the load pose, "
                        "CheckCollision + CheckSelfCollision (BASELINE.json configs[1])"),
    "wam_lab1_100k": ("lab1", 100000, 1, None,
                      "Barrett WAM (11 DOF, 18 links, 31.0k tris) in data/lab1.env.xml (14 bodies, 16.5k tris): 100k uniform-random "
                      "DOF configs, CheckCollision + CheckSelfCollision (BASELINE.json configs[0])"),
    "hironx_table_1m": ("hironxtable", 1000000, 4, None,
                        "HIRO-NX dual arm (robots/kawada-hironx.zae: 24 links, 23 DOF, 24.1k tris) + data/ikeatable.kinbody.xml: 1M "
                        "uniform-random configs over all DOFs, self-collision dominated (BASELINE.json configs[3])"),
    "wam_clutter200_1250k": ("clutter:200:22", 1250000, 5, "arm",
                             "Barrett WAM arm (7 active DOF, segway base disabled) vs synthetic clutter of 200 obstacles x 9 680-triangle "
                             "geodesic spheres (1.94M tris, 217 MB of BVH + triangles: larger than L2), 1.25M configs per GPU = 10M "
                             "sharded over 8 GPUs (BASELINE.json configs[4]; scene generated from seed 5 by openrave_b200/clutter.py)"),
}
# BASELINE.json configs[2]: plans/s of BiRRT + the default linear smoother, edge checks batched (csrc/cc_planner.cpp)
PLAN_WORKLOADS = {
    "wam_wamtest1_birrt100": ("wamtest1", 100, 3,
                              "BiRRT (rplanners) on the Barrett WAM arm (7 DOF) in data/wamtest1.env.xml: 100 start/goal pairs drawn from "
                              "collision-free uniform samples (seed 3), step 0.04, <= 10000 iterations, goal bias 0.01, 10 _SimpleOptimizePath "
                              "shortcuts + 100 linearsmoother iterations; all problems advance in lockstep, every Extend ray / shortcut of "
                              "every live problem goes through one cc_check_batch call per round (BASELINE.json configs[2])"),
}
# CO_Distance: minimum separation per configuration
DIST_WORKLOADS = {
    "wam_lab1_distance_100k": ("lab1", 100000, 1, "arm",
                               "Barrett WAM in data/lab1.env.xml, CO_ActiveDOFs on the 7 arm DOFs: minimum separation (env + self pairs) of "
                               "100k uniform-random arm configurations, exact (PQP_Distance with rel_err = abs_err = 0) in FP32"),
}
DEFAULT_WORKLOAD = "pr2_pr2test1_1m"
METRIC = "collision-checked configs/sec (FK+env+self)"
UNIT = "configs/s"


def load_world(workload):
    from openrave_b200.environment import Environment
    scene, n, seed, manip, _ = WORKLOADS[workload]
    if scene.startswith("clutter:"):
        from openrave_b200 import clutter
        _, nob, freq = scene.split(":")
        env, robot = clutter.make_clutter_environment(clutter.default_fixture(), nobstacles=int(nob), freq=int(freq), seed=seed)
        return env, robot, n, seed
    env = Environment()
    env.Load(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".npz"))
    robot = env.GetRobots()[0]
    if manip is not None:
        robot.SetActiveDOFs([m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices())
    return env, robot, n, seed


def sample(robot, n, seed):
    """RobotConfigurationSampler semantics over the robot's ACTIVE DOFs: lower + u*(upper-lower), u ~ U[0,1)."""
    lo, hi = robot.GetActiveDOFLimits()
    lo, hi = np.array(lo), np.array(hi)
    rng = np.random.default_rng(seed)
    return lo + rng.random((n, len(lo))) * (hi - lo)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def cpu_baseline(env, robot, q, budget_s=12.0, kind=None):
    """Single-thread pqprave-equivalent configs/s on a bounded prefix of the same batch, and the reference's own work counters
    (NumBVTests / NumTriTests per configuration) the algorithmic roofline figure is defined from."""
    import oracle
    if kind is None:
        kind = "reference" if oracle.available("reference") else "port"
    orc = oracle.OracleChecker(env, kind)
    orc.InitEnvironment()
    probe = min(200, len(q))
    secs, _ = orc.time_batch(robot, q[:probe])
    m = int(min(len(q), max(probe, budget_s / max(secs / probe, 1e-9))))
    secs, _ = orc.time_batch(robot, q[:m])
    ref = orc.check_batch(robot, q[:min(m, 2000)])
    cnt = ref["counters"].astype(np.float64)
    return {"value": m / secs, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d of the %d seeded configs, 1 thread, %.1f s" % (m, len(q), secs),
            "fclrave": "not runnable here: FCL is not part of the reference tree or of this image (SURVEY.md 8d)"}, cnt.mean(axis=0)


def parity_record(workload, cc, robot, q, rate_1core, budget_s=25.0):
    """SURVEY.md 8d parity stats, live: the CUDA path against the CPU oracle on a prefix of the timed batch sized to ~budget_s of
    wall clock on ALL host cores (contact-band classifier on for every configuration), plus the committed full-size sweep of
    this workload (tests/parity_sweep.py -> profiles/r02_parity_<workload>.json) when there is one."""
    from oracle import parity
    workers = max(1, min(os.cpu_count() or 1, 64))
    m = int(min(len(q), max(2000, 0.4 * rate_1core * workers * budget_s)))      # the band classifier roughly doubles the cost
    got = cc.CheckCollisionBatch(robot, q[:m])
    ref = parity.oracle_pass(workload, q[:m], workers=workers, want_band=True)
    rec = parity.compare(got, ref, lambda b: cc.body_from_id(b).name)
    rec["sample"] = "first %d of the %d seeded configs of the timed batch" % (m, len(q))
    fp = os.path.join(ROOT, "profiles", "r02_parity_%s.json" % workload)
    if os.path.isfile(fp):
        full = json.load(open(fp))
        rec["full_size_sweep"] = {k: full.get(k) for k in ("compared", "identical", "colliding", "near_contact", "verdict_differs",
                                                           "first_pair_differs", "mismatch_inside_band", "mismatch_outside_band",
                                                           "first_pair_mismatch", "against", "oracle_workers")}
        rec["full_size_sweep"]["source"] = "profiles/r02_parity_%s.json (tests/parity_sweep.py on the GPU box)" % workload
    return rec


def run_reference_arm(args):
    """pqprave's CPU path (FK + all link pairs, PQP_ALL_CONTACTS, no broadphase) on all host cores,
    one private scene per worker (the reference's clone-per-thread idiom, ormultithreadedplanning.cpp:27-33)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    kind = "reference" if oracle.available("reference") else "port"
    workload = args.workload
    env, robot, n, seed = load_world(workload)
    ncores = os.cpu_count() or 1
    workers = max(1, min(ncores, 64))
    q = sample(robot, n, seed)
    # bounded sample per step: ~8 s of work per worker, estimated from a probe
    orc = oracle.OracleChecker(env, kind)
    orc.InitEnvironment()
    secs, _ = orc.time_batch(robot, q[:100])
    per_worker = int(max(50, min(n // workers, 8.0 / (secs / 100))))
    m = per_worker * workers
    ctx = mp.get_context("fork")

    def one_step():
        with ctx.Pool(workers, initializer=_ref_worker_init, initargs=(workload, kind)) as pool:
            t0 = time.perf_counter()
            res = pool.map(_ref_worker_run, [(q[i * per_worker:(i + 1) * per_worker],) for i in range(workers)])
            wall = time.perf_counter() - t0
        return max(wall, max(res))

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    ms = 1000.0 * sum(times) / len(times)
    value = m / (ms / 1000.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "description": WORKLOADS[workload][4], "configs_per_step": m,
                       "note": "bounded sample of the %d-config batch, %d worker processes" % (n, workers)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind,
                             "sample": "%d configs per step over %d processes (private scene each)" % (m, workers)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


_REF_STATE = {}


def _ref_worker_init(workload, kind):
    import oracle
    env, robot, _, _ = load_world(workload)
    orc = oracle.OracleChecker(env, kind)
    orc.InitEnvironment()
    orc.time_batch(robot, np.zeros((1, robot.GetActiveDOF())))   # builds models + pair list before the timed region
    _REF_STATE.update(env=env, robot=robot, orc=orc)


def _ref_worker_run(a):
    secs, _ = _REF_STATE["orc"].time_batch(_REF_STATE["robot"], a[0])
    return secs


def run_distance(args):
    """distance-checked configs/s through cc_distance_batch (host buffers in and out: this entry has no device-pointer
    flavour yet, so `value` and `e2e` are the same measurement)."""
    import torch
    from openrave_b200.checker import CO_ActiveDOFs, CudaCollisionChecker
    scene, n, seed, manip, desc = DIST_WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from openrave_b200.environment import Environment
    env = Environment()
    env.Load(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".npz"))
    robot = env.GetRobots()[0]
    arm = robot.GetActiveManipulator().GetArmIndices()
    robot.SetActiveDOFs(arm)
    cc = CudaCollisionChecker(env, local_rank)
    env.SetCollisionChecker(cc)
    cc.SetCollisionOptions(CO_ActiveDOFs)
    q = sample(robot, n, seed + 1000 * rank)
    bq = cc.bind(robot)
    W = max(args.warmup, 3)
    for _ in range(W):
        d, pair, pts = bq.distance(q)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    launches0 = cc.GetStatistics()["launches"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d, pair, pts = bq.distance(q)
    secs = (time.perf_counter() - t0) / args.steps
    launches = cc.GetStatistics()["launches"] - launches0
    if world > 1:
        t = torch.tensor([secs], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank == 0:
        value = world * n / secs
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d_apx, _, _ = bq.distance(q, rel_err=200.0, abs_err=0.001)       # pqprave's own tolerances (collisionPQP.h:48-49)
        secs_apx = (time.perf_counter() - t0) / args.steps
        line = {"metric": "distance-checked configs/sec (FK + min separation over env+self pairs)", "value": value, "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": 1000 * secs, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "description": desc, "configs_per_step_per_gpu": n, "ndof": len(arm),
                           "pairs_env": bq.info.npairs_env, "pairs_self": bq.info.npairs_self,
                           "mean_distance": float(d.mean()), "zero_fraction": float((d == 0).mean()),
                           "configs_per_s_with_pqprave_tolerances": world * n / secs_apx,
                           "mean_distance_with_pqprave_tolerances": float(d_apx.mean()),
                           "timing": "host wall clock around cc_distance_batch (H2D + FK kernel + distance kernel + D2H), max over ranks",
                           "l2": "inputs larger than nothing: scene L2-resident by design (5 MB); q/results stream through HBM"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(n * len(arm) * 8), "d2h_bytes_per_step": int(n * 60)},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": None}
        if not args.no_cpu_baseline:
            import oracle
            m = 2000
            base = {}
            for kind in (["reference"] if oracle.available("reference") else []) + ["port"]:
                env2 = Environment()
                env2.Load(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".npz"))
                r2 = env2.GetRobots()[0]
                r2.SetActiveDOFs(arm)
                orc = oracle.OracleChecker(env2, kind)
                env2.SetCollisionChecker(orc)
                orc.SetCollisionOptions(CO_ActiveDOFs)
                mm = 200 if kind == "reference" else m
                orc.distance_batch(r2, q[:4], active=arm)
                t0 = time.perf_counter()
                ref = orc.distance_batch(r2, q[:mm], active=arm)
                dt = time.perf_counter() - t0
                base[kind] = mm / dt
                assert np.abs(ref["distance"] - d[:mm]).max() < 2e-5
            kind = "reference" if "reference" in base else "port"
            line["cpu_baseline"] = {"value": base[kind], "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": "first %d configs, 1 thread; 'reference' = unmodified PQP_Distance(rel_err=abs_err=0) on every "
                                              "visited pair, 'port' = the oracle's restated branch and bound carrying the running minimum "
                                              "across pairs (%.0f configs/s)" % (200 if kind == "reference" else m, base["port"])}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def plan_problems(scene, nproblems, seed, rank, checker_factory):
    """Start/goal pairs = consecutive collision-free samples of the seeded uniform stream (RobotConfigurationSampler semantics),
    classified by the checker the arm runs on."""
    from openrave_b200.environment import Environment
    env = Environment()
    env.Load(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".npz"))
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    checker = checker_factory(env)
    env.SetCollisionChecker(checker)
    q = sample(robot, 8 * nproblems + 64, seed + 1000 * rank)
    return env, robot, checker, q


def run_birrt(args):
    """plans/s: `steps` timed repetitions of planning all problems of the workload (fresh trees every step)."""
    import torch
    import torch.distributed as dist
    from openrave_b200 import planning as pl
    from openrave_b200.checker import CudaCollisionChecker
    scene, nproblems, seed, desc = PLAN_WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)
    env, robot, cc, q = plan_problems(scene, nproblems, seed, rank, lambda e: CudaCollisionChecker(e, local_rank))
    free = q[cc.CheckCollisionBatch(robot, q)["verdict"] == 0]
    starts, goals = free[0:2 * nproblems:2].copy(), free[1:2 * nproblems:2].copy()
    seeds = np.arange(nproblems, dtype=np.uint32) + 1000 * rank
    P = pl.BirrtPlanner(robot, cc)
    for _ in range(W):
        res = P.PlanBatch(starts, goals, seeds)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    launches0 = cc.GetStatistics()["launches"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = P.PlanBatch(starts, goals, seeds)      # host buffers in, paths out: the call a planner user makes
    torch.cuda.synchronize()
    secs = (time.perf_counter() - t0) / args.steps
    launches = cc.GetStatistics()["launches"] - launches0
    if world > 1:
        t = torch.tensor([secs], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank == 0:
        tot = P.totals
        solved = sum(1 for r in res if r["status"] == pl.PLAN_SOLVED)
        value = world * nproblems / secs
        ndof = len(robot.GetActiveDOFIndices())
        line = {"metric": "BiRRT plans/sec (batched CheckPathAllConstraints)", "value": value, "unit": "plans/s", "n_gpus": world,
                "steps": args.steps, "warmup": W, "ms_per_step": 1000.0 * secs, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "description": desc, "problems_per_step_per_gpu": nproblems, "ndof": ndof,
                           "solved": solved, "rounds": tot["rounds"], "configs_submitted": tot["configs_submitted"],
                           "configs_serial_equivalent": tot["configs_serial"],
                           "mean_iterations": float(np.mean([r["iterations"] for r in res])),
                           "mean_waypoints": float(np.mean([r["waypoints"] for r in res])),
                           "configs_per_s": world * tot["configs_submitted"] / secs,
                           "l2": "scene (5 MB) stays L2-resident across rounds, as it does in a planner",
                           "timing": "host wall clock around cc_birrt_plan_batch (planner host + H2D + kernels + D2H), max over ranks",
                           "parallelism": "problems sharded, scene replicated (dp%d)" % world},
                "e2e": {"value": value, "unit": "plans/s", "h2d_bytes_per_step": int(tot["configs_submitted"] * ndof * 8),
                        "d2h_bytes_per_step": int(tot["configs_submitted"] * 4)},
                "gpu_launches": int(launches), "clocks": sampler.summary(),
                "roofline": None}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = birrt_cpu_baseline(scene, nproblems, seed, starts, goals, seeds)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def birrt_cpu_baseline(scene, nproblems, seed, starts, goals, seeds, budget_s=15.0):
    """The same planner host with the reference's serial, early-exit edge loop on the CPU oracle (1 thread), on a prefix of
    the same problems sized to ~budget_s."""
    import oracle
    from openrave_b200 import planning as pl
    kind = "reference" if oracle.available("reference") else "port"
    env, robot, orc, _ = plan_problems(scene, nproblems, seed, 0, lambda e: oracle.OracleChecker(e, kind))
    orc.InitEnvironment()
    fn, blk = orc.group_checker(robot)
    P = pl.BirrtPlanner(robot, (fn, blk))
    t0 = time.perf_counter()
    P.PlanBatch(starts[:2], goals[:2], seeds[:2])
    per = max((time.perf_counter() - t0) / 2, 1e-4)
    m = int(max(2, min(nproblems, budget_s / per)))
    t0 = time.perf_counter()
    P.PlanBatch(starts[:m], goals[:m], seeds[:m])
    secs = time.perf_counter() - t0
    return {"value": m / secs, "unit": "plans/s", "cores": 1, "kind": kind,
            "sample": "first %d of the %d problems, serial early-exit edge checks, 1 thread, %.1f s (%d configs)" % (
                m, nproblems, secs, P.totals["configs_serial"])}


def run_birrt_reference(args):
    """Reference arm for the plans/s workload: one planner + private oracle scene per host core, problems split."""
    import multiprocessing as mp
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle
    from openrave_b200 import planning as pl  # noqa: F401
    kind = "reference" if oracle.available("reference") else "port"
    scene, nproblems, seed, desc = PLAN_WORKLOADS[args.workload]
    env, robot, orc, q = plan_problems(scene, nproblems, seed, 0, lambda e: oracle.OracleChecker(e, kind))
    orc.InitEnvironment()
    free = q[orc.check_batch(robot, q)["verdict"] == 0]
    starts, goals = free[0:2 * nproblems:2].copy(), free[1:2 * nproblems:2].copy()
    seeds = np.arange(nproblems, dtype=np.uint32)
    workers = max(1, min(os.cpu_count() or 1, 64, nproblems))
    ctx = mp.get_context("fork")
    chunks = [(starts[i::workers], goals[i::workers], seeds[i::workers]) for i in range(workers)]

    def one_step():
        with ctx.Pool(workers, initializer=_birrt_worker_init, initargs=(scene, nproblems, seed, kind)) as pool:
            t0 = time.perf_counter()
            pool.map(_birrt_worker_run, chunks)
            return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    secs = sum(times) / len(times)
    value = nproblems / secs
    print(json.dumps({"impl": "reference", "metric": "BiRRT plans/sec (batched CheckPathAllConstraints)", "value": value,
                      "unit": "plans/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * secs,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": args.workload, "description": desc, "problems_per_step": nproblems},
                      "cpu_baseline": {"value": value, "unit": "plans/s", "cores": workers, "kind": kind,
                                       "sample": "all %d problems over %d processes, serial early-exit edge checks" % (nproblems, workers)},
                      "e2e": {"value": value, "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def _birrt_worker_init(scene, nproblems, seed, kind):
    import oracle
    from openrave_b200 import planning as pl
    env, robot, orc, _ = plan_problems(scene, nproblems, seed, 0, lambda e: oracle.OracleChecker(e, kind))
    orc.InitEnvironment()
    fn, blk = orc.group_checker(robot)
    _REF_STATE.update(env=env, robot=robot, orc=orc, planner=pl.BirrtPlanner(robot, (fn, blk)), blk=blk)


def _birrt_worker_run(a):
    s, g, sd = a
    if len(s):
        _REF_STATE["planner"].PlanBatch(s, g, sd)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + sorted(PLAN_WORKLOADS) + sorted(DIST_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.workload in DIST_WORKLOADS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "distance workload: the CPU arm is reported inline as cpu_baseline"}))
            return
        run_distance(args)
        return
    if args.workload in PLAN_WORKLOADS:
        (run_birrt_reference if args.impl == "reference" else run_birrt)(args)
        return
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from openrave_b200.checker import BATCH_ENV, BATCH_SELF, CudaCollisionChecker

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)

    env, robot, n, seed = load_world(args.workload)
    ndof = robot.GetActiveDOF()
    cc = CudaCollisionChecker(env, local_rank)
    env.SetCollisionChecker(cc)
    bq = cc.bind(robot)
    flags = BATCH_ENV | BATCH_SELF

    # weak scaling: every rank owns a full batch (its own seed); the scene is replicated per GPU
    q_host = sample(robot, n, seed + 1000 * rank)
    q_pinned = torch.from_numpy(q_host).pin_memory()
    d_q = q_pinned.cuda(non_blocking=True)
    # verdict | first env pair | first self pair of this rank's shard in ONE buffer, so that what crosses NVLink is a single
    # ncclAllGather into a preallocated tensor (no torch.cat, no per-rank copy-out inside the timed region)
    d_res = torch.zeros(3 * n, dtype=torch.int32, device="cuda")
    d_verdict, d_fe, d_fs = d_res[0:n], d_res[n:2 * n], d_res[2 * n:3 * n]
    # N > 1: the gather of step i's results runs on NCCL's stream while step i + 1 computes into the other result buffer (a device
    # consumer reads batch i while batch i + 1 is checked); a gather inside every step would also re-synchronise the ranks K times
    d_res_alt = torch.zeros(3 * n, dtype=torch.int32, device="cuda") if world > 1 else None
    gathered = [torch.empty(world * 3 * n, dtype=torch.int32, device="cuda") for _ in range(2)] if world > 1 else None
    pending = [None, None]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device(i=0):
        b = i & 1 if world > 1 else 0
        res = d_res_alt if b else d_res
        if pending[b] is not None:          # the gather that last read this buffer (two steps ago) must be done
            pending[b].wait()
            pending[b] = None
        bq.check_device(d_q.data_ptr(), n, flags, res[0:n].data_ptr(), res[n:2 * n].data_ptr(), res[2 * n:3 * n].data_ptr(), 0, stream.cuda_stream)

    def gather_results(i=0):
        # only the per-config verdicts and first-contact pairs cross NVLink (device consumer): ONE ncclAllGather
        if world > 1:
            b = i & 1
            pending[b] = dist.all_gather_into_tensor(gathered[b], d_res_alt if b else d_res, async_op=True)

    def drain_gathers():
        for b in range(2):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None

    # the caller-side buffers of the host entry: page-locked (cc_host_alloc), reused every step
    from openrave_b200.checker import PinnedArray
    pin_v, pin_fe, pin_fs = PinnedArray(n, np.uint32), PinnedArray(n, np.int32), PinnedArray(n, np.int32)
    q_np = q_pinned.numpy()

    def step_e2e():
        v, fe, fs, _ = bq.check(q_np, flags, out=(pin_v.array, pin_fe.array, pin_fs.array))
        return v

    # the first warm-up step is also where the library builds its free-space certificates (one-off, like the BVH build)
    for i in range(W + 1):
        step_device(i)
        gather_results(i)
    drain_gathers()
    torch.cuda.synchronize()
    ncert = bq.table_count()
    ntab_only = ncert - bq.dropped_pairs
    cert_ms = bq.certificate_build_ms
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = cc.GetStatistics()["launches"]
    evs = []
    t_wall0 = time.perf_counter()
    # The path shards with NO data-path collective (independent configurations, replicated scene: SURVEY.md 8e): every rank's
    # results stay on its GPU, which is what `value` times.  The device-consumer flavour — the results of all ranks gathered on
    # every GPU — is timed separately below (config.device_consumer): at 8 GPUs each rank receives 96 MB per step, which streams
    # through the L2 the BVH lives in and costs 4 % (profiles/r02F_*).
    for i in range(args.steps):
        flush.fill_(1)                      # evict q, results and the BVH from L2 between timed steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device(i)
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches_timed = cc.GetStatistics()["launches"] - launches0    # FK + broadphase + traversal kernel per chunk
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    ms = sum(ms_steps) / len(ms_steps)
    kernel_ms = ms
    gathered_ms = None
    if world > 1:
        # device consumer: ONE ncclAllGather of the results per step, overlapped with the next step's kernels (double-buffered
        # results); the last gather is waited for inside the timed region
        evg = []
        kg = max(3, args.steps // 2)
        dist.barrier()
        for i in range(kg):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_device(i)
            e1.record(stream)
            gather_results(i)
            evg.append((e0, e1))
        e_tail = torch.cuda.Event(enable_timing=True)
        drain_gathers()
        e_tail.record(stream)
        torch.cuda.synchronize()
        gathered_ms = (sum(a.elapsed_time(b) for a, b in evg) + evg[-1][1].elapsed_time(e_tail)) / kg
        tg = torch.tensor([gathered_ms], device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gathered_ms = float(tg.item())
    ms_per_rank = [ms]
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        allms = torch.empty(world, device="cuda")
        dist.all_gather_into_tensor(allms, t)
        ms_per_rank = [float(x) for x in allms.cpu()]
        ms = max(ms_per_rank)               # the step of the slowest rank (GPUs of one box differ by a few per cent)

    # the report == NULL flavour the planners use (CC_BATCH_FIRST_HIT): same kernels, stops at the first contact
    from openrave_b200.checker import BATCH_FIRST_HIT
    evf = []
    for _ in range(max(3, args.steps // 2)):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bq.check_device(d_q.data_ptr(), n, flags | BATCH_FIRST_HIT, d_verdict.data_ptr(), 0, 0, 0, stream.cuda_stream)
        e1.record(stream)
        evf.append((e0, e1))
    torch.cuda.synchronize()
    first_hit_ms = sum(a.elapsed_time(b) for a, b in evf) / len(evf)
    # a scene whose obstacles keep moving never gets to use the free-space certificates (cc_query_refresh drops the ones a moved
    # body touches): the same batch with them switched off is the floor of the dynamic-scene throughput
    cc.SetPairTables(False)
    bq_dyn = cc.bind(robot)
    evd = []
    for k in range(2 + max(3, args.steps // 2)):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bq_dyn.check_device(d_q.data_ptr(), n, flags, d_verdict.data_ptr(), d_fe.data_ptr(), d_fs.data_ptr(), 0, stream.cuda_stream)
        e1.record(stream)
        if k >= 2:
            evd.append((e0, e1))
    torch.cuda.synchronize()
    dynamic_ms = sum(a.elapsed_time(b) for a, b in evd) / len(evd)
    cc.SetPairTables(True)
    bq = cc.bind(robot)
    for _ in range(2):                      # certificates are rebuilt on the re-bound query before the e2e timing below
        step_device(0)
    torch.cuda.synchronize()
    step_device(0)                          # leave the full outputs in the result buffers for the consistency check below
    gather_results(0)
    drain_gathers()
    torch.cuda.synchronize()

    # end to end through the host-buffer entry.  One GPU: this process's own checker.  N GPUs: ONE process (rank 0) drives all N
    # devices through a multi-device scene (cc_scene_create_multi): one cc_check_batch call over the N x n configurations of all
    # ranks in one page-locked buffer, range-sharded inside the library, every device copying its slice in and its results out
    # (host consumer: no collective).  The other ranks idle at the barrier meanwhile — this is the call an OpenRAVE process makes.
    multi = None
    if world == 1:
        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            v_last = step_e2e()
        e2e_s = (time.perf_counter() - t0) / args.steps
    else:
        torch.cuda.synchronize()
        # the other ranks wait on the HOST (gloo): an NCCL barrier parks a spinning kernel on their GPUs, and rank 0's kernels on those
        # GPUs would then time-slice against another process's context (69 instead of 180+ M configs/s on two GPUs, profiles/r02t_*)
        cpu_group = dist.new_group(backend="gloo")
        dist.barrier(group=cpu_group)
        e2e_s = 0.0
        if rank == 0:
            env_m, robot_m, _, _ = load_world(args.workload)
            cc_m = CudaCollisionChecker(env_m, list(range(world)))
            env_m.SetCollisionChecker(cc_m)
            bq_m = cc_m.bind(robot_m)
            q_all = PinnedArray((world * n, ndof), np.float64)
            for r in range(world):
                q_all.array[r * n:(r + 1) * n] = sample(robot, n, seed + 1000 * r)
            pv, pe, ps = PinnedArray(world * n, np.uint32), PinnedArray(world * n, np.int32), PinnedArray(world * n, np.int32)
            for _ in range(3):
                bq_m.check(q_all.array, flags, out=(pv.array, pe.array, ps.array))
            t0 = time.perf_counter()
            for _ in range(args.steps):
                bq_m.check(q_all.array, flags, out=(pv.array, pe.array, ps.array))
            e2e_s = (time.perf_counter() - t0) / args.steps
            v_last = pv.array[:n].copy()
            multi = {"devices": list(range(world)), "configs_per_call": world * n,
                     "kernel_launches_per_call": int((cc_m.GetStatistics()["launches"]) // (args.steps + 3))}
            bq_m.close()
            cc_m.DestroyEnvironment()
        dist.barrier(group=cpu_group)
        if rank != 0:
            v_last = d_verdict.cpu().numpy().astype(np.uint32)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # device-resident and host-buffer paths must agree
    assert (d_verdict.cpu().numpy().astype(np.uint32) == v_last).all()

    if rank == 0:
        stats = cc.GetStatistics()
        value = world * n / (ms / 1000.0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "description": WORKLOADS[args.workload][4], "configs_per_step_per_gpu": n,
                           "ndof": ndof, "pairs_env": bq.info.npairs_env, "pairs_self": bq.info.npairs_self,
                           "pairs_hoisted": bq.info.npairs_hoisted, "pairs_certified_free": ncert - ntab_only,
                           "pairs_with_table": ntab_only, "certificate_build_ms": cert_ms,
                           "bvh_nodes": bq.info.total_nodes, "tris": bq.info.total_tris,
                           "l2": "flushed between timed steps (256 MB write)", "fk": "f64 on device, narrowphase f32",
                           "parallelism": "configs sharded, scene replicated (dp%d)" % world, "multi_device_e2e": multi,
                           "ms_per_step_per_rank": [round(x, 4) for x in ms_per_rank],
                           "collective": None,
                           "device_consumer": (None if world == 1 else {
                               "configs_per_s": world * n / (gathered_ms / 1000.0), "ms_per_step": gathered_ms,
                               "collective": "one ncclAllGather of the results per step (12 B / configuration), overlapped with the next "
                                             "step's kernels; the last gather is waited for inside the timed region"}),
                           "colliding_fraction": float((v_last != 0).mean()),
                           "first_hit_mode_configs_per_s": world * n / (first_hit_ms / 1000.0),
                           "dynamic_scene_configs_per_s": world * n / (dynamic_ms / 1000.0),
                           "dynamic_scene_note": "the same batch with cc_scene_set_pair_tables(0): what a caller gets while obstacles move "
                                                 "every few batches (a refresh drops the certificates a moved body touches)"},
                "e2e": {"value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(world * n * ndof * 8),
                        "d2h_bytes_per_step": int(world * n * 12),
                        "path": "cc_check_batch, pinned host buffers" + ("" if world == 1 else
                                " — ONE process, one call over all %d devices (cc_scene_create_multi)" % world)},
                "gpu_launches": int(launches_timed),
                "clocks": sampler.summary(),
                "wall_s_timed_region": t_wall}
        peak, peak_src = peaks()
        base = None
        cnt = None
        if not args.no_cpu_baseline and world == 1:
            base, cnt = cpu_baseline(env, robot, q_host)
            base["parity"] = parity_record(args.workload, cc, robot, q_host, base["value"])
            line["cpu_baseline"] = base
        if cnt is not None:
            bytes_per_config = cnt[0] * 2 * 64 + cnt[1] * 2 * 36 + 8 * ndof + 16
            achieved = n * bytes_per_config / (kernel_ms / 1000.0) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
            if os.path.isfile(tp):   # dram bytes of one step's kernels, from the committed `ncu --set full` capture
                traffic = json.load(open(tp)).get("dram_bytes_per_step")
            physical = None
            pp = os.path.join(ROOT, "profiles", "physical_%s.json" % args.workload)
            if os.path.isfile(pp):   # counters of the step's kernels from the committed ncu capture of this command (tools/harvest_r02.py)
                physical = json.load(open(pp))
                # the PHYSICAL utilisation of this run, recomputable from profiles/ alone + this line's own value and clocks:
                #   issue_frac  = warp instructions per configuration (ncu inst_executed of the FK + traversal launches / their
                #                 configurations) x configs/s / (SMs x 4 schedulers x SM clock): the binding resource
                #   lane_efficiency = active threads per executed warp instruction / 32
                #   dram_frac   = DRAM bytes per step (ncu dram__bytes_read + write of the same launches) / step time / measured copy peak
                props = torch.cuda.get_device_properties(local_rank)
                clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
                w = physical.get("warp_instructions_per_configuration")
                if w:
                    physical["issue_frac"] = w * (n / (kernel_ms / 1000.0)) / (props.multi_processor_count * 4 * clk)
                    physical["issue_frac_inputs"] = {"sm_count": props.multi_processor_count, "schedulers_per_sm": 4, "sm_clock_hz": clk,
                                                     "configs_per_s_per_gpu": n / (kernel_ms / 1000.0)}
                if physical.get("threads_per_instruction"):
                    physical["lane_efficiency"] = physical["threads_per_instruction"] / 32.0
                if traffic:
                    physical["dram_frac"] = traffic / (kernel_ms / 1000.0) / (peak * 1e9)
            # compute side: the reference's box / triangle test counts priced at 263 / 745 flops + FK
            from openrave_b200.checker import device_fp32_peak
            fp32_peak = device_fp32_peak(local_rank)
            flops_per_config = cnt[0] * (200 + 63) + cnt[1] * (45 + 700) + len(robot.links) * 3 * 60
            fp32_achieved = n * flops_per_config / (kernel_ms / 1000.0) / 1e12
            line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "traffic": traffic, "peak_source": peak_src,
                                "fp32": {"achieved": fp32_achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_achieved / fp32_peak,
                                         "algorithmic_flops_per_config": flops_per_config,
                                         "peak_source": "measured in this run (cc_device_fp32_peak: FMA chains on every SM)"},
                                "physical": physical,
                                "algorithmic_bytes_per_config": bytes_per_config,
                                "ref_bv_tests_per_config": cnt[0], "ref_tri_tests_per_config": cnt[1],
                                "note": "algorithmic bytes follow the reference's counters (no broadphase, all contacts); frac > 1 "
                                        "means the kernel avoids that work by culling/any-hit, see profiles/ for physical utilisation"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
