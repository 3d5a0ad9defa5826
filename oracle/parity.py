"""Parity bookkeeping between the CUDA path and the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Imported by tests/parity_sweep.py (the full-size sweeps committed under profiles/), tests/ and bench.py's `cpu_baseline`
leg; never by the product.  It runs the oracle on every host core (one private scene per worker process: the reference's
clone-per-thread idiom, src/cppexamples/ormultithreadedplanning.cpp:27-33) over the SAME seeded configurations the CUDA
path checked, and reports what BASELINE.json's north_star asks for: verdicts and reported link pairs identical outside the
1e-4 m contact band, near-contact configurations counted and reported.

The contact band of a configuration: free and some visited pair closer than 1e-4 m (PQP_Tolerance on every
visited pair), or colliding with a penetration measure <= 1e-4 m (max over the intersecting triangle pairs of the first
colliding pair(s) of the minimum normalised overlap over TriContact's 17 axes, pqp/PQP.cpp:438-456).
"""
import multiprocessing as mp
import os
import time

import numpy as np

BAND = 1e-4
_STATE = {}


def _worker_init(workload, kind):
    import bench
    import oracle
    env, robot, _, _ = bench.load_world(workload)
    orc = oracle.OracleChecker(env, kind)
    orc.InitEnvironment()
    orc.time_batch(robot, np.zeros((1, robot.GetActiveDOF())))     # models + pair list built before any timing
    _STATE.update(env=env, robot=robot, orc=orc)


def _worker_run(args):
    q, want_band = args
    orc, robot = _STATE["orc"], _STATE["robot"]
    t0 = time.perf_counter()
    r = orc.check_batch(robot, q, want_band=want_band, band=BAND)
    secs = time.perf_counter() - t0
    # body ids -> names: the ids of two harness instances need not agree
    names = {bid: b.name for bid, b in orc._bodies_by_id().items()}
    return dict(verdict=r["verdict"], first_env=r["first_env"], first_self=r["first_self"], near=r["near"],
                penetration=r["penetration"], counters=r["counters"].sum(axis=0), secs=secs, names=names)


def oracle_pass(workload, q, kind=None, workers=None, want_band=True, slices_per_worker=4):
    """The oracle's answers for every row of q, computed on `workers` processes.  Returns a dict of concatenated arrays plus
    `wall_s`, `cpu_s` (sum over workers), `workers`, `kind`, `names` (oracle body id -> body name)."""
    import oracle
    if kind is None:
        kind = "reference" if oracle.available("reference") else "port"
    if workers is None:
        workers = max(1, min(os.cpu_count() or 1, 64))
    n = len(q)
    workers = max(1, min(workers, n))
    nsl = max(1, min(n, workers * slices_per_worker))       # several slices per worker: colliding configurations are cheaper than free ones
    bounds = [n * i // nsl for i in range(nsl + 1)]
    ctx = mp.get_context("spawn")                             # the parent may hold a CUDA context: never fork it
    t0 = time.perf_counter()
    with ctx.Pool(workers, initializer=_worker_init, initargs=(workload, kind)) as pool:
        t1 = time.perf_counter()
        parts = pool.map(_worker_run, [(q[bounds[i]:bounds[i + 1]], want_band) for i in range(nsl)], chunksize=1)
        wall = time.perf_counter() - t1
    out = {k: np.concatenate([p[k] for p in parts]) for k in ("verdict", "first_env", "first_self", "near", "penetration")}
    out["counters_sum"] = np.sum([p["counters"] for p in parts], axis=0)
    out.update(wall_s=wall, setup_s=t1 - t0, cpu_s=float(sum(p["secs"] for p in parts)), workers=workers, kind=kind,
               names=parts[0]["names"])
    return out


def band_mask(ref):
    return (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= BAND))


def compare(gpu, ref, gpu_body_name):
    """gpu: dict(verdict[n], first_env[n,3] {link, body id, link}, first_self[n,2]) from CudaCollisionChecker.CheckCollisionBatch;
    ref: oracle_pass() output; gpu_body_name(id) -> body name.  Returns the parity record."""
    n = len(ref["verdict"])
    gv = np.asarray(gpu["verdict"][:n], dtype=np.uint32)
    band = band_mask(ref)
    vbad = gv != ref["verdict"]
    # reported pairs: compare (link, other body NAME, other link)
    gfe, rfe = np.asarray(gpu["first_env"][:n]), ref["first_env"]
    fe_bad = (gfe[:, 0] != rfe[:, 0]) | (gfe[:, 2] != rfe[:, 2])
    both = (gfe[:, 0] >= 0) & (rfe[:, 0] >= 0) & ~fe_bad
    if both.any():
        gname = np.array([gpu_body_name(int(b)) for b in gfe[both, 1]])
        rname = np.array([ref["names"][int(b)] for b in rfe[both, 1]])
        sub = np.zeros(int(both.sum()), dtype=bool)
        sub[:] = gname != rname
        fe_bad[np.nonzero(both)[0][sub]] = True
    fs_bad = (np.asarray(gpu["first_self"][:n]) != ref["first_self"]).any(axis=1)
    pbad = ~vbad & (fe_bad | fs_bad)
    differ = vbad | pbad
    outside = differ & ~band
    rec = {
        "compared": int(n),
        "identical": int(n - differ.sum()),
        "colliding": int((ref["verdict"] != 0).sum()),
        "near_contact": int(band.sum()),
        "verdict_differs": int(vbad.sum()),
        "first_pair_differs": int(pbad.sum()),
        "mismatch_inside_band": int((differ & band).sum()),
        "mismatch_outside_band": int(outside.sum()),
        "first_pair_mismatch": int((pbad & ~band).sum()),
        "band_m": BAND,
        "against": "oracle (%s)" % ref["kind"],
        "oracle_workers": int(ref["workers"]),
        "oracle_wall_s": round(float(ref["wall_s"]), 2),
        "oracle_cpu_s": round(float(ref["cpu_s"]), 2),
    }
    if outside.any():
        rec["outside_band_indices"] = [int(i) for i in np.nonzero(outside)[0][:50]]
    if (differ & band).any():
        rec["inside_band_indices"] = [int(i) for i in np.nonzero(differ & band)[0][:50]]
    return rec
