#!/usr/bin/env python
"""Generate the committed scene fixtures from the reference tree (run in the build container only).

  python tools/make_fixtures.py            # writes tests/golden/scenes/*.npz

Reads /root/reference/src/{data,robots} and /root/reference/models.tgz (extracted to a temp dir);
nothing under /root/reference is copied verbatim — the outputs are the parsed, flattened scenes.
"""
import os
import sys
import tarfile
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from openrave_b200.environment import Environment          # noqa: E402
from openrave_b200.serialize import save_environment        # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "scenes")


def extract_models(dst):
    t = tarfile.open(os.path.join(REF, "models.tgz"), "r:gz")
    try:
        for m in t:
            if m.isfile():
                t.extract(m, dst)
    except Exception as e:          # the archive in the snapshot has trailing garbage after the last member
        print("models.tgz: stopped at", e)


def main():
    tmp = os.environ.get("OPENRAVE_B200_MODELS") or tempfile.mkdtemp(prefix="orave_models_")
    if not os.path.isdir(os.path.join(tmp, "models")):
        extract_models(tmp)
    os.makedirs(OUT, exist_ok=True)
    for scene in ("lab1", "wamtest1", "pr2test1", "hironxtable"):
        env = Environment(data_dirs=[os.path.join(REF, "src"), tmp])
        try:
            env.Load("data/%s.env.xml" % scene)
        except Exception as e:
            print("skip %s: %s" % (scene, e))
            continue
        path = os.path.join(OUT, scene + ".npz")
        save_environment(env, path)
        print("%s: %d bodies, %d tris -> %s (%.1f KB)" % (
            scene, len(env.GetBodies()), sum(l.idx.shape[0] for b in env.GetBodies() for l in b.links), path,
            os.path.getsize(path) / 1024.0))

    # the robot test/test_collision.py:94-181 builds from an XML string (segway box model joined to the WAM by a dummy joint):
    # the string is read from the reference's test file here, parsed by our loader, and only the flattened result is stored
    src = open(os.path.join(REF, "test", "test_collision.py")).read()
    a = src.index('testrobot_xml="""') + len('testrobot_xml="""')
    xml = src[a:src.index('"""', a)]
    f = tempfile.NamedTemporaryFile("w", suffix=".robot.xml", delete=False)
    f.write(xml)
    f.close()
    env = Environment(data_dirs=[os.path.join(REF, "src"), tmp])
    robot = env.ReadRobotURI(f.name)
    os.unlink(f.name)
    env.Add(robot)
    path = os.path.join(OUT, "joinxml_segwaywam.npz")
    save_environment(env, path)
    print("joinxml_segwaywam: %d links, %d tris -> %s (%.1f KB)" % (
        len(robot.links), sum(l.idx.shape[0] for l in robot.links), path, os.path.getsize(path) / 1024.0))

    # the scene of test/test_robot.py:625-660 (test_distancechecking): the WAM alone, one mug at the manipulator's transform,
    # a second mug named 'collision' at (0.5, 0, 0.5)
    env = Environment(data_dirs=[os.path.join(REF, "src"), tmp])
    robot = env.ReadRobotURI("robots/barrettwam.robot.xml")
    env.Add(robot)
    target = env.ReadKinBodyURI("data/mug1.kinbody.xml")
    env.Add(target)
    target.SetTransform(robot.GetActiveManipulator().GetTransform())
    coltarget = env.ReadKinBodyURI("data/mug1.kinbody.xml")
    coltarget.name = "collision"
    env.Add(coltarget)
    coltarget.SetTransform(((1.0, 0.0, 0.0, 0.0), (0.5, 0.0, 0.5)))
    path = os.path.join(OUT, "wam_two_mugs.npz")
    save_environment(env, path)
    print("wam_two_mugs: %d bodies -> %s (%.1f KB)" % (len(env.GetBodies()), path, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
