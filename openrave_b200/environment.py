"""Host-side mirror of the slice of OpenRAVE's Environment the collision path touches.

Reference: src/libopenrave-core/environment-core.h — body list in insertion order (`GetBodies`, the
order pqprave's env loop walks, collisionPQP.h:448), `SetCollisionChecker` -> `InitEnvironment`
(:1460-1485), `Add` -> `InitKinBody` (:1019,1060), `CheckCollision` overloads routed to the current
checker (:1491-1572).  This is the harness the tests and bench drive; inside a real OpenRAVE the
same calls arrive through plugins/cudacollision/.
"""
import os

from . import ravemath as rm
from .kinbody import KinBody, Link, Robot


class CollisionReport(object):
    """CollisionReport subset (include/openrave/collisionchecker.h:138-179): link pairs in the order the
    checker found them, `minDistance`, option bits."""

    def __init__(self):
        self.Reset()

    def Reset(self, options=0):
        self.options = options
        self.vLinkColliding = []      # [(link1, link2)]; link1 belongs to the queried body
        self.plink1 = None
        self.plink2 = None
        self.minDistance = 1e20
        self.numCols = 0
        self.nNumValidCollisions = 0
        self.contacts_by_pair = {}    # CO_Contacts: (link1, link2) -> [(pos, norm, depth)] (CollisionPairInfo::contacts)
        self.numWithinTol = 0         # CO_UseTolerance: bodies with a link pair within the tolerance
        self.contacts = []            # CO_Distance: [(pos, norm, depth)] of the closest pair (collisionPQP.h:633-647)

    def AddLinkCollision(self, link1, link2):
        """collisionchecker.cpp:577-596 — duplicates are not appended twice."""
        for i, (a, b) in enumerate(self.vLinkColliding):
            if a is link1 and b is link2:
                return i
        self.vLinkColliding.append((link1, link2))
        self.nNumValidCollisions = len(self.vLinkColliding)
        if self.plink1 is None:
            self.plink1, self.plink2 = link1, link2
        return len(self.vLinkColliding) - 1

    @property
    def collisionInfos(self):
        return [CollisionPairInfo(a, b) for a, b in self.vLinkColliding]


class CollisionPairInfo(object):
    def __init__(self, link1, link2):
        self.link1, self.link2 = link1, link2

    def ExtractFirstBodyLinkGeomNames(self):
        return (self.link1.body.name, self.link1.name, '')

    def ExtractSecondBodyLinkGeomNames(self):
        return (self.link2.body.name, self.link2.name, '')


class CollisionAction(object):
    """include/openrave/collisionchecker.h:40-44."""
    DefaultAction = 0
    Ignore = 1


class _CallbackHandle(object):
    """UserDataPtr returned by RegisterCollisionCallback: Close() (or losing the last reference) unregisters."""

    def __init__(self, env, fn):
        self._env, self._fn = env, fn

    def Close(self):
        if self._env is not None and self in self._env._collision_callbacks:
            self._env._collision_callbacks.remove(self)
        self._env = None


class Environment(object):
    def __init__(self, data_dirs=None):
        self._collision_callbacks = []
        self.bodies = []
        self.checker = None
        self._next_id = 1
        self.data_dirs = list(data_dirs) if data_dirs else default_data_dirs()

    # ---- loading
    def Load(self, filename):
        if filename.endswith(".npz"):
            from .serialize import load_environment_into
            load_environment_into(self, filename)
            # side fixture of the analytic-geometry mode: the primitive geometries of the links that consist of primitives only
            import json
            import os
            side = filename[:-4] + ".prims.json"
            if os.path.isfile(side):
                prims = json.load(open(side))
                for b in self.bodies:
                    for l in b.links:
                        l.prims = prims.get(b.name, {}).get("%d:%s" % (l.index, l.name), [])
            return True
        # EnvironmentBase::Load takes a scene, a robot or a kinbody file (environment-core.h: the root tag decides for XML; a COLLADA
        # file brings the bodies it holds)
        from .xmlscene import SceneLoader
        loader = SceneLoader(self.data_dirs)
        path = loader.find_file(filename, None)
        if path is None:
            raise IOError("cannot find %s" % filename)
        import os
        if os.path.splitext(path)[1].lower() in (".zae", ".dae"):
            # a document with an articulated system brings a robot, anything else a kinbody (colladareader.cpp Extract)
            from .colladascene import read_document
            is_robot = any(True for _ in read_document(path).iter("instance_articulated_system"))
            self.Add(loader.load_body_file(self, path, robot=is_robot))
            return True
        tag = loader._parse(path).tag
        if tag == "environment":
            loader.load_environment(self, path)
        elif tag in ("robot", "kinbody"):
            self.Add(loader.load_body_file(self, path, robot=(tag == "robot")))      # IAM_StrictNameChecking, xmlreaders-core.cpp:3624-3638
        else:
            raise ValueError("%s: root element is <%s>, expected <environment>, <robot> or <kinbody>" % (path, tag))
        return True

    def ReadKinBodyURI(self, filename):
        from .xmlscene import SceneLoader
        return SceneLoader(self.data_dirs).load_body_file(self, filename, robot=False)

    def ReadRobotURI(self, filename):
        from .xmlscene import SceneLoader
        return SceneLoader(self.data_dirs).load_body_file(self, filename, robot=True)

    # ---- bodies
    def Add(self, body, anonymous=False):
        names = {b.name for b in self.bodies}
        if body.name in names:
            if not anonymous:
                raise ValueError("body name %r already in the environment" % body.name)
            k = 1
            while "%s%d" % (body.name, k) in names:
                k += 1
            body.name = "%s%d" % (body.name, k)
        body.env = self
        body.env_id = self._next_id
        self._next_id += 1
        self.bodies.append(body)
        body.compute_internal_information()
        if self.checker is not None:
            self.checker.InitKinBody(body)

    def Remove(self, body):
        if body in self.bodies:
            self.bodies.remove(body)
            if self.checker is not None:
                self.checker.RemoveKinBody(body)

    def GetBodies(self):
        return list(self.bodies)

    def GetRobots(self):
        return [b for b in self.bodies if b.is_robot]

    def GetKinBody(self, name):
        for b in self.bodies:
            if b.name == name:
                return b
        return None

    GetRobot = GetKinBody

    # ---- checker
    def SetCollisionChecker(self, checker):
        if self.checker is not None:
            self.checker.DestroyEnvironment()
        self.checker = checker
        for b in self.bodies:
            b._nonadjacent_cache = {}      # _ResetInternalCollisionCache on checker change
        if checker is not None:
            checker.env = self
            checker.InitEnvironment()
        return True

    def GetCollisionChecker(self):
        return self.checker

    def CheckCollision(self, a, b=None, report=None, bodyexcluded=None, linkexcluded=None):
        """The CheckCollision overloads of EnvironmentBase (environment-core.h:1491-1572); `bodyexcluded` / `linkexcluded` select
        the (body | link, vbodyexcluded, vlinkexcluded) overloads (collisionchecker.h:249-263)."""
        cc = self.checker
        if cc is None:
            raise RuntimeError("no collision checker set")
        if bodyexcluded is not None or linkexcluded is not None:
            return cc.CheckCollisionExcluding(a, list(bodyexcluded or ()), list(linkexcluded or ()), report=report)
        if b is None:
            return cc.CheckCollision(a, report=report)
        return cc.CheckCollision(a, b, report=report)

    # ---- collision callbacks (environment-core.h RegisterCollisionCallback / GetRegisteredCollisionCallbacks)
    def RegisterCollisionCallback(self, fn):
        h = _CallbackHandle(self, fn)
        self._collision_callbacks.append(h)
        return h

    def HasRegisteredCollisionCallbacks(self):
        return len(self._collision_callbacks) > 0

    def GetRegisteredCollisionCallbacks(self):
        return [h._fn for h in self._collision_callbacks]

    def filter_collisions_through_callbacks(self, options, pairs):
        """What a checker does with a colliding link pair when callbacks are registered and CO_IgnoreCallbacks is off
        (plugins/pqprave/collisionPQP.h:569-617, plugins/fclrave/fclcollision.cpp:1002-1011): every callback sees a report
        holding that one pair; any answer other than DefaultAction drops the pair as if the links did not collide.
        `pairs` is an iterable of (link1, link2) in the checker's order; returns the pairs that stand."""
        CO_IgnoreCallbacks = 0x100                                  # collisionchecker.h CollisionOptions
        if (options & CO_IgnoreCallbacks) or not self._collision_callbacks:
            return list(pairs)
        kept = []
        for l1, l2 in pairs:
            rep = CollisionReport()
            rep.Reset(options)
            rep.AddLinkCollision(l1, l2)
            rep.minDistance = 0
            if all(fn(rep, False) == CollisionAction.DefaultAction for fn in self.GetRegisteredCollisionCallbacks()):
                kept.append((l1, l2))
        return kept

    def _body_changed(self, body, kind="structure"):
        """kind 'pose': only link transforms / DOF values moved; 'structure': enable flags, grabs, active DOFs, geometry."""
        if self.checker is not None:
            self.checker._on_body_changed(body, kind)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def excluding_pairs(env, a, bodyexcluded, linkexcluded):
    """Link pairs of CheckCollisionP with exclusion lists (plugins/pqprave/collisionPQP.h:432-498, 314-329) in the reference's order:
    for a body, each attached body (itself, what it grabs) against every other body that is not attached to it and not
    excluded; links named in `linkexcluded` are skipped on either side.  For a link, only that link (attached bodies are not
    walked, collisionchecker.h:250).  Returns [(query body, [(link1, link2), ...]), ...]."""
    from .kinbody import Link
    excl_b = set(id(b) for b in bodyexcluded)
    excl_l = set(id(l) for l in linkexcluded)
    if isinstance(a, Link):
        groups = [(a.body, [a])]
        attached = set([id(a.body)])
    else:
        bodies = [a] + [g.body for g in getattr(a, "grabbed", [])]
        attached = set(id(b) for b in bodies)
        groups = [(b, list(b.links)) for b in bodies]
    out = []
    for qb, links in groups:
        pairs = []
        for ob in env.GetBodies():
            if id(ob) in attached or id(ob) in excl_b:
                continue
            for l1 in links:
                if id(l1) in excl_l:
                    continue
                for l2 in ob.links:
                    if id(l2) in excl_l:
                        continue
                    pairs.append((l1, l2))
        out.append((qb, pairs))
    return out


def default_data_dirs():
    """$OPENRAVE_DATA-style search path (plugindatabase / RaveFindLocalFile).  In this image the
    reference tree provides src/ (data/, robots/) and the extracted models.tgz (models/)."""
    dirs = []
    env = os.environ.get("OPENRAVE_DATA", "")
    dirs += [d for d in env.split(os.pathsep) if d]
    for d in ("/root/reference/src", os.environ.get("OPENRAVE_B200_MODELS", "/tmp/openrave_models")):
        if d and os.path.isdir(d):
            dirs.append(d)
    return dirs
