"""OpenRAVE-XML (.env.xml / .robot.xml / .kinbody.xml) reader for the standalone harness.

Restates the subset of the reference's SAX readers that the benchmark scenes use
(src/libopenrave-core/xmlreaders-core.cpp): EnvironmentXMLReader :3319-3531, RobotXMLReader
:2878-3160, KinBodyXMLReader :2010-2445, LinkXMLReader :772-1353, JointXMLReader :1361-1753 and the
geometry reader src/libopenrave/xmlreaders.cpp:205-520.  Tag and attribute names are compared in
lower case like the reference's SAX front end.  Rules kept exactly:

  * a `file=` attribute loads that file INTO the same body before the element's own children;
  * link pose tags accumulate: <translation> adds, rotations pre-multiply the rotation only
    (:1265-1294); <offsetfrom> left-multiplies the named link's ORIGINAL transform at </body>
    (:1246-1260, `_vTransforms`);
  * body-level <translation>/<rotationaxis>/... are applied, at the end of the element, to the
    links created since the element started (`roottransoffset`, :2391-2396, :3119-3125);
  * joints: anchor/axis are mapped from the <offsetfrom> frame into the first body's frame and the
    joint frames are computed right away from the link poses of that moment (:1544-1548);
    limits beyond +-pi create the mid-range offset (:1525-1531); cylinders get the +90deg about-x
    axis fix (:1078-1086).

Load-time tooling, out of the per-config hot path.
"""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import ravemath as rm
from .ivloader import load_iv
from .kinbody import Geometry, Joint, KinBody, Link, Manipulator, MimicExpr, Robot, _pair


def _lower_tree(elem):
    elem.tag = elem.tag.lower()
    elem.attrib = {k.lower(): v for k, v in elem.attrib.items()}
    for c in elem:
        _lower_tree(c)


def _floats(text):
    out = []
    for tok in (text or "").split():
        try:
            out.append(float(tok))
        except ValueError:
            break   # stream extraction stops at the first bad token
    return out


def _is_false(s):
    return s.lower() == "false" or s == "0"


class SceneLoader(object):
    def __init__(self, data_dirs):
        self.data_dirs = [os.path.abspath(d) for d in data_dirs]
        self._mesh_cache = {}

    # ------------------------------------------------------------------ files
    def find_file(self, name, parse_dir):
        """RaveFindLocalFile: the current parse directory first, then the data directories."""
        cands = []
        if os.path.isabs(name):
            cands.append(name)
        if parse_dir:
            cands.append(os.path.join(parse_dir, name))
        for d in self.data_dirs:
            cands.append(os.path.join(d, name))
        for c in cands:
            if os.path.isfile(c):
                return os.path.abspath(c)
        return None

    def _parse(self, path):
        root = ET.parse(path).getroot()
        _lower_tree(root)
        return root

    def load_mesh(self, path):
        if path not in self._mesh_cache:
            ext = os.path.splitext(path)[1].lower()
            if ext not in (".iv", ".wrl", ".vrml"):      # (.wrl: the VRML 1.0 subset, i.e. Inventor's grammar; VRML 2.0 needs Coin3D in the reference too)
                raise NotImplementedError("mesh format %s (only Inventor .iv / VRML 1.0 through the ivcon path)" % ext)
            self._mesh_cache[path] = load_iv(path)
        return self._mesh_cache[path]

    # ------------------------------------------------------------------ environment
    def load_environment(self, env, filename):
        path = self.find_file(filename, None)
        if path is None:
            raise IOError("cannot find %s" % filename)
        root = self._parse(path)
        if root.tag != "environment":
            raise ValueError("%s: root element is <%s>, expected <environment>" % (path, root.tag))
        pdir = os.path.dirname(path)
        if "file" in root.attrib:
            # <Environment file="...">: that scene first, then this element's own bodies (InterfaceXMLReader's file attribute,
            # xmlreaders-core.cpp:1770-1812; hanoi_complex2.env.xml, pa10grasp2.env.xml)
            inner = self.find_file(root.attrib["file"], pdir)
            if inner is None:
                raise IOError("cannot find %s (from %s)" % (root.attrib["file"], pdir))
            self.load_environment(env, inner)
        for child in root:
            if child.tag == "robot":
                body = Robot(env)
                self.read_robot(child, body, pdir)
                env.Add(body)
                self._apply_jointvalues(body)
            elif child.tag == "kinbody":
                body = KinBody(env)
                self.read_kinbody(child, body, pdir)
                env.Add(body)
                self._apply_jointvalues(body)
        return env

    @staticmethod
    def _apply_jointvalues(body):
        """<jointvalues>: SetDOFValues once the body is in the environment, ignored when the count is not the DOF count
        (xmlreaders-core.cpp:3404-3411, 3418-3425)."""
        vals = getattr(body, "_xml_jointvalues", None)
        if vals is not None and len(vals) == body.GetDOF() and vals:
            body.SetDOFValues(list(vals))

    def load_body_file(self, env, filename, robot=False):
        path = self.find_file(filename, None)
        if path is None:
            raise IOError("cannot find %s" % filename)
        if os.path.splitext(path)[1].lower() in (".zae", ".dae"):
            from .colladascene import load_collada_into
            body = Robot(env) if robot else KinBody(env)
            body.uri = path
            return load_collada_into(body, path)
        root = self._parse(path)
        body = Robot(env) if (robot or root.tag == "robot") else KinBody(env)
        body.uri = path
        if root.tag == "robot":
            self.read_robot(root, body, os.path.dirname(path))
        else:
            self.read_kinbody(root, body, os.path.dirname(path))
        return body

    # ------------------------------------------------------------------ shared pieces
    @staticmethod
    def _apply_pose_tag(tag, text, t, scale=(1.0, 1.0, 1.0)):
        v = _floats(text)
        if tag == "translation":
            return (t[0], (t[1][0] + v[0] * scale[0], t[1][1] + v[1] * scale[1], t[1][2] + v[2] * scale[2]))
        if tag == "rotationaxis":
            fangle = v[3] if len(v) > 3 else 0.0
            tnew = (rm.quat_from_axis_angle((v[0], v[1], v[2]), fangle * rm.PI / 180.0), (0.0, 0.0, 0.0))
        elif tag == "rotationmat":
            tnew = (rm.quat_from_matrix(((v[0], v[1], v[2]), (v[3], v[4], v[5]), (v[6], v[7], v[8]))), (0.0, 0.0, 0.0))
        elif tag == "quat":
            tnew = (rm.normalize4((v[0], v[1], v[2], v[3])), (0.0, 0.0, 0.0))
        else:
            return t
        return (rm.t_mul(tnew, t)[0], t[1])

    def _load_into(self, body, fileattr, pdir, as_robot):
        path = self.find_file(fileattr, pdir)
        if path is None:
            raise IOError("cannot find %s (from %s)" % (fileattr, pdir))
        if os.path.splitext(path)[1].lower() in (".zae", ".dae"):
            from .colladascene import load_collada_into       # RaveParseColladaURI (xmlreaders-core.cpp:2076-2092)
            if not body.uri:
                body.uri = path
            load_collada_into(body, path)
            return
        root = self._parse(path)
        if not body.uri:
            body.uri = path
        sub = os.path.dirname(path)
        if root.tag == "robot":
            self.read_robot(root, body, sub)
        elif root.tag == "kinbody":
            self.read_kinbody(root, body, sub)
        else:
            raise ValueError("%s: unexpected root <%s>" % (path, root.tag))

    # ------------------------------------------------------------------ robot
    def read_robot(self, elem, robot, pdir):
        roottransoffset = len(robot.links)
        atts = elem.attrib
        prefix = atts.get("prefix", "")
        if "file" in atts:
            # the file is read with this element's other attributes (xmlreaders-core.cpp:1770-1812): its own reader applies the
            # prefix to what it loads, so the children of THIS element already see the prefixed names
            fileoffset = len(robot.links), len(robot.joints), len(robot.passive_joints)
            had = {id(m) for m in robot.manipulators}
            self._load_into(robot, atts["file"], pdir, True)
            if prefix:
                self._apply_prefix(robot, prefix, *fileoffset)
                self._prefix_manipulators(robot, prefix, had)
        rootoffset, rootj, rootjp = len(robot.links), len(robot.joints), len(robot.passive_joints)
        had_manips = {id(m) for m in robot.manipulators}
        trans = rm.IDENT_T
        for child in elem:
            tag = child.tag
            if tag == "robot":
                self.read_robot(child, robot, pdir)
            elif tag == "kinbody":
                self.read_kinbody(child, robot, pdir)
            elif tag == "manipulator":
                self._read_manipulator(child, robot)
            elif tag in ("translation", "rotationaxis", "rotationmat", "quat"):
                trans = self._apply_pose_tag(tag, child.text, trans)
            elif tag == "jointvalues":
                robot._xml_jointvalues = _floats(child.text)         # the robot's own tag wins over a nested kinbody's (:2985, 3019)
        if atts.get("name"):
            robot.name = atts["name"]
        if prefix:
            self._apply_prefix(robot, prefix, rootoffset, rootj, rootjp)
            self._prefix_manipulators(robot, prefix, had_manips)
        for l in robot.links[roottransoffset:]:
            l.t = rm.t_mul(trans, l.t)

    @staticmethod
    def _prefix_manipulators(robot, prefix, initial):
        """xmlreaders-core.cpp:3106-3116: manipulators that did not exist before this reader get the prefix on their name and on
        their gripper joint names (their base / effector are link objects here, renamed with the links)."""
        for m in robot.manipulators:
            if id(m) not in initial:
                m.name = prefix + m.name
                m.gripper_joint_names = [prefix + n for n in getattr(m, "gripper_joint_names", [])]

    def _read_manipulator(self, elem, robot):
        name = elem.attrib.get("name", "")
        m = None
        for mm in robot.manipulators:
            if mm.name == name and name:
                m = mm
        if m is None:
            m = Manipulator(robot, name)
            robot.manipulators.insert(0, m)   # newly declared manipulators go in front (:3036-3051)
        for c in elem:
            if c.tag == "base":
                m.base = robot.GetLink((c.text or "").strip())
            elif c.tag == "effector":
                m.effector = robot.GetLink((c.text or "").strip())
            elif c.tag == "joints":
                m.gripper_joint_names = (c.text or "").split()
            elif c.tag in ("translation", "rotationaxis", "rotationmat", "quat"):
                m.t_local = self._apply_pose_tag(c.tag, c.text, m.t_local)

    @staticmethod
    def _apply_prefix(body, prefix, rootoffset, rootj, rootjp):
        for l in body.links[rootoffset:]:
            l.name = prefix + l.name
        renamed = {}
        for j in body.joints[rootj:] + body.passive_joints[rootjp:]:
            renamed[j.name] = prefix + j.name
            j.name = prefix + j.name
        for j in body.joints[rootj:] + body.passive_joints[rootjp:]:
            if j.mimic is not None and renamed:
                # one pass, longest names first (utils::SearchAndReplace): "JF1mimic" must not be rewritten through "JF1"
                import re
                pat = re.compile("|".join(re.escape(k) for k in sorted(renamed, key=len, reverse=True)))
                j.mimic = MimicExpr(pat.sub(lambda m: renamed[m.group(0)], j.mimic.text))

    # ------------------------------------------------------------------ kinbody
    def read_kinbody(self, elem, body, pdir, scale=(1.0, 1.0, 1.0)):
        roottransoffset = len(body.links)
        atts = elem.attrib
        prefix = atts.get("prefix", "")
        if "file" in atts:
            fileoffset = len(body.links), len(body.joints), len(body.passive_joints)      # (see read_robot)
            self._load_into(body, atts["file"], pdir, False)
            if prefix:
                self._apply_prefix(body, prefix, *fileoffset)
        rootoffset, rootj, rootjp = len(body.links), len(body.joints), len(body.passive_joints)
        if "makejoinedlinksadjacent" in atts:
            body.make_joined_links_adjacent = int(atts["makejoinedlinksadjacent"]) != 0
        skip_before = getattr(self, "_skipgeometry", False)
        if "skipgeometry" in atts:                       # links without their <geom> children (:2037-2039; jack.kinbody.xml)
            self._skipgeometry = atts["skipgeometry"].strip().lower() in ("true", "1")
        vtransforms = [l.t for l in body.links]    # `_vTransforms`: original poses for <offsetfrom>
        trans = rm.IDENT_T
        modelsdir = ""
        for child in elem:
            tag = child.tag
            if tag == "kinbody":
                self.read_kinbody(child, body, pdir, scale)
                vtransforms = [l.t for l in body.links]
            elif tag == "body":
                link, torig = self._read_link(child, body, vtransforms, pdir, modelsdir, scale)
                if link.index < 0:
                    link.index = len(body.links)
                    body.links.append(link)
                    vtransforms.append(rm.IDENT_T)
                vtransforms[link.index] = torig
            elif tag == "joint":
                self._read_joint(child, body, vtransforms, scale)
            elif tag in ("translation", "rotationaxis", "rotationmat", "quat"):
                trans = self._apply_pose_tag(tag, child.text, trans, scale)
            elif tag == "adjacent":
                names = (child.text or "").split()
                l0, l1 = body.GetLink(names[0]), body.GetLink(names[1])
                if l0 is not None and l1 is not None:
                    body.forced_adjacent.add(_pair(l0.index, l1.index))
            elif tag == "modelsdir":
                modelsdir = (child.text or "").strip() + "/"
            elif tag == "jointvalues":
                body._xml_jointvalues = _floats(child.text)          # applied once the body is in the environment (:2332, 3404-3411)
        if atts.get("name"):
            body.name = atts["name"]
        if prefix:
            self._apply_prefix(body, prefix, rootoffset, rootj, rootjp)
        for l in body.links[roottransoffset:]:
            l.t = rm.t_mul(trans, l.t)
        self._skipgeometry = skip_before

    # ------------------------------------------------------------------ link
    def _read_link(self, elem, body, vtransforms, pdir, modelsdir, scale):
        atts = elem.attrib
        link = None
        name = atts.get("name", "")
        if name:
            for l in body.links:
                if l.name == name:
                    link = l
                    break
        if link is None:
            link = Link(body, name)
        if "type" in atts:
            link.static = atts["type"].lower() == "static"
        link.enabled = not _is_false(atts.get("enable", "true"))
        offsetfrom = None
        for child in elem:
            tag = child.tag
            if tag in ("translation", "rotationaxis", "rotationmat", "quat"):
                link.t = self._apply_pose_tag(tag, child.text, link.t, scale)
            elif tag == "offsetfrom":
                offsetfrom = body.GetLink((child.text or "").strip())
                if offsetfrom is None:
                    raise ValueError("offsetfrom link %r not found in %s" % (child.text, body.name))
            elif tag in ("geom", "geometry"):
                if getattr(self, "_skipgeometry", False):
                    continue
                g = self._read_geometry(child, pdir, modelsdir, scale)
                if g is not None:
                    link.geometries.append(g)
        if offsetfrom is not None:
            idx = offsetfrom.index
            root = vtransforms[idx] if 0 <= idx < len(vtransforms) else offsetfrom.t
            link.t = rm.t_mul(root, link.t)
        link.update_collision()
        return link, link.t

    # ------------------------------------------------------------------ geometry
    def _read_geometry(self, elem, pdir, modelsdir, scale):
        gtype = elem.attrib.get("type", "").lower()
        if gtype not in ("box", "cylinder", "sphere", "trimesh"):
            if gtype in ("", "none"):
                return None
            raise NotImplementedError("geometry type %r" % gtype)
        g = Geometry(gtype)
        g.visible = not _is_false(elem.attrib.get("render", "true"))
        filename, vscale = "", (1.0, 1.0, 1.0)
        renderfile, rscale = "", (1.0, 1.0, 1.0)
        for child in elem:
            tag = child.tag
            if tag in ("translation", "rotationaxis", "rotationmat", "quat"):
                g.t = self._apply_pose_tag(tag, child.text, g.t)
            elif tag == "name":
                g.name = (child.text or "").strip()
            elif gtype == "box" and tag in ("extents", "halfextents"):
                v = _floats(child.text)
                g.data = (v[0], v[1], v[2])
            elif gtype == "box" and tag == "fullextents":
                v = _floats(child.text)
                g.data = (v[0] * 0.5, v[1] * 0.5, v[2] * 0.5)
            elif gtype in ("cylinder", "sphere") and tag == "radius":
                g.data = (_floats(child.text)[0], g.data[1], g.data[2])
            elif gtype == "cylinder" and tag == "height":
                g.data = (g.data[0], _floats(child.text)[0], g.data[2])
            elif gtype == "trimesh" and tag in ("data", "collision"):
                if not filename:
                    filename, vscale = self._file_and_scale(child)
            elif gtype == "trimesh" and tag == "render":
                if not renderfile:
                    renderfile, rscale = self._file_and_scale(child)
        if gtype == "trimesh":
            use, usescale = (filename, vscale) if filename else (renderfile, rscale)
            if not use:
                return None
            path = None
            if modelsdir:
                path = self.find_file(use, os.path.join(pdir, modelsdir))
            if path is None:
                path = self.find_file(use, pdir)
            if path is None:
                raise IOError("mesh file %s not found" % use)
            v32, idx = self.load_mesh(path)
            v = v32.astype(np.float64)
            # CreateTriMeshFromFile: Vector(vscale.x*x, vscale.y*y, vscale.z*z) in double (:473-476)
            v = np.stack([usescale[0] * v[:, 0], usescale[1] * v[:, 1], usescale[2] * v[:, 2]], axis=1)
            if scale != (1.0, 1.0, 1.0):
                raise NotImplementedError("scalegeometry")
            g.verts, g.idx = v, idx.copy()
        else:
            if gtype == "cylinder":
                trot = (rm.quat_from_axis_angle((1.0, 0.0, 0.0), rm.PI / 2), (0.0, 0.0, 0.0))
                g.t = (rm.t_mul(g.t, trot)[0], g.t[1])
            g.init_collision_mesh()
        g.t = (g.t[0], (g.t[1][0] * scale[0], g.t[1][1] * scale[1], g.t[1][2] * scale[2]))
        return g

    @staticmethod
    def _file_and_scale(child):
        if "file" in child.attrib:
            toks = [child.attrib["file"]] + child.attrib.get("scale", "").split()
        else:
            toks = (child.text or "").split()
        if not toks:
            return "", (1.0, 1.0, 1.0)
        s = [1.0, 1.0, 1.0]
        nums = _floats(" ".join(toks[1:]))
        if len(nums) >= 1:
            s = [nums[0], nums[0], nums[0]]
        if len(nums) >= 3:
            s = [nums[0], nums[1], nums[2]]
        return toks[0], (s[0], s[1], s[2])

    # ------------------------------------------------------------------ joint
    def _read_joint(self, elem, body, vtransforms, scale):
        atts = elem.attrib
        jtype = atts.get("type", "hinge").lower()
        if jtype not in ("hinge", "slider"):
            raise NotImplementedError("joint type %r" % jtype)
        j = Joint(body, atts.get("name", ""), jtype)
        if "enable" in atts:
            j.is_active = not _is_false(atts["enable"])
        if "circular" in atts:
            j.circular = not _is_false(atts["circular"])
        for k, v in atts.items():
            if k.startswith("mimic_pos"):
                j.mimic = MimicExpr(v)
        fratio = 1.0 if jtype == "slider" else rm.PI / 180.0
        attached = []
        offsetfrom = None
        anchor = (0.0, 0.0, 0.0)
        axis = (0.0, 0.0, 0.0)
        initial = None
        for child in elem:
            tag = child.tag
            text = child.text or ""
            if tag == "body":
                lname = text.strip()
                found = None
                for l in body.links:
                    if l.name.lower() == lname.lower():
                        found = l
                        break
                if found is None:
                    raise ValueError("joint %s: link %r not found" % (j.name, lname))
                attached.append(found)
            elif tag == "offsetfrom":
                offsetfrom = body.GetLink(text.strip())
                if offsetfrom is None:
                    raise ValueError("joint %s: offsetfrom %r not found" % (j.name, text))
            elif tag in ("limits", "limitsrad", "limitsdeg"):
                fm = fratio if tag == "limitsdeg" else 1.0
                v = _floats(text)
                if len(v) == 2:
                    j.lower = fm * min(v[0], v[1])
                    j.upper = fm * max(v[0], v[1])
            elif tag == "resolution":
                v = _floats(text)
                j.resolution = (v[0] if v else 0.02) * fratio
            elif tag == "weight":
                v = _floats(text)
                if v:
                    j.weight = v[0]
            elif tag == "anchor":
                v = _floats(text)
                anchor = (v[0] * scale[0], v[1] * scale[1], v[2] * scale[2])
            elif tag == "axis":
                v = _floats(text)
                axis = rm.normalize3((v[0], v[1], v[2]))
            elif tag == "initial":
                v = _floats(text)
                initial = v[0] if v else None
        if len(attached) != 2:
            raise ValueError("joint %s needs two <body> tags" % j.name)
        if jtype == "hinge" and (j.lower < -rm.PI or j.upper > rm.PI):
            j.offset = 0.5 * (j.lower + j.upper)
        toff = rm.IDENT_T
        if offsetfrom is not None:
            idx = offsetfrom.index
            toff = vtransforms[idx] if 0 <= idx < len(vtransforms) else offsetfrom.t
        toff = rm.t_mul(rm.t_inv(attached[0].t), toff)
        axis = rm.quat_rotate(toff[0], axis)
        j.compute_internal(attached[0], attached[1], attached[0].t, attached[1].t, rm.t_apply(toff, anchor), axis, initial)
        if j.is_active:
            body.joints.append(j)
        else:
            body.passive_joints.append(j)
        return j
