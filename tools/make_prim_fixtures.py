#!/usr/bin/env python
"""Side fixtures for the analytic-geometry mode (fclrave semantics): for every link of the benchmark scenes
whose geometries are ALL primitives (box / cylinder / sphere), the primitive list {type, pose in the link frame, data} the
reference's fclrave would turn into fcl::Box / Cylinder / Sphere (plugins/fclrave/fclspace.cpp:475-602).  Parsed from
/root/reference by this repository's own XML loader (build container only); written to tests/golden/scenes/<scene>.prims.json.

  python tools/make_prim_fixtures.py
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from openrave_b200.environment import Environment          # noqa: E402
import make_fixtures                                        # noqa: E402

REF = "/root/reference"
OUT = make_fixtures.OUT


def main():
    tmp = os.environ.get("OPENRAVE_B200_MODELS") or tempfile.mkdtemp(prefix="orave_models_")
    if not os.path.isdir(os.path.join(tmp, "models")):
        make_fixtures.extract_models(tmp)
    for scene in ("lab1", "wamtest1", "pr2test1", "hironxtable"):
        env = Environment(data_dirs=[os.path.join(REF, "src"), tmp])
        env.Load("data/%s.env.xml" % scene)
        out = {}
        n = 0
        for b in env.GetBodies():
            for l in b.links:
                gs = l.geometries
                if not gs or any(g.type not in ("box", "cylinder", "sphere") for g in gs):
                    continue
                out.setdefault(b.name, {})["%d:%s" % (l.index, l.name)] = [dict(type=g.type, t=[list(g.t[0]), list(g.t[1])], data=[float(x) for x in g.data]) for g in gs]
                n += len(gs)
        path = os.path.join(OUT, scene + ".prims.json")
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)
        print("%s: %d primitives on %d links -> %s" % (scene, n, sum(len(v) for v in out.values()), path))


if __name__ == "__main__":
    main()
