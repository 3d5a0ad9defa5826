"""COLLADA 1.5 (.dae / .zae) robot ingestion — kinematics, collision meshes, adjacency extras.

Restates what the reference's reader does for documents written by the OpenRAVE COLLADA writer (the format of every
`src/robots/*.zae`), following src/libopenrave-core/colladaparser/colladareader.cpp:

  link / attachment_full recursion, link pose = parent * attachment * link transforms        :1872-1916, 2012-2157
  joints: type, axis (rotated into the parent link frame), anchor = attachment translation      :2093-2110, 2176-2187, 2381-2384
  limits: kinematics axis_info (soft) before the joint's own <limits>, degrees for revolute     :2285-2352
  active / locked / circular / discretization_resolution / planning_weight newparams           :2099-2106, 2267-2331
  <formula> -> mimic equations (OpenRAVE technique "position" equation, MathML)                 :1571-1660
  geometry: <triangles> soups per instance_geometry of the link's visual node, node-relative    :2391-2506, 2566-2646
  extras: ignore_link_pair, link_collision_state                                                :4975-5000, 5051-5070
  manipulators (frame_origin / frame_tip / gripper_joint)                                       :3388-3500
  DOF order: bound actuators first, in attach_actuator order                                    :3720-3792

Transform elements compose left to right (`_ExtractFullTransform`, :4543-4550); rotations are axis + degrees (:4463).
Documents that use other features (instance_joint across files, <polygons>, container / cage geometries, physics) raise
NotImplementedError instead of guessing.
"""
import math
import os
import re
import xml.etree.ElementTree as ET
import zipfile

import numpy as np

from . import ravemath as rm
from .kinbody import Geometry, Joint, Link, Manipulator, MimicExpr

_NS = re.compile(r"\{[^}]*\}")


def _strip_ns(root):
    for e in root.iter():
        if isinstance(e.tag, str):
            e.tag = _NS.sub("", e.tag)


def _floats(text):
    return [float(x) for x in (text or "").split()]


def read_document(path):
    """Return the parsed <COLLADA> root of a .dae, or of the root document of a .zae archive (manifest.xml names it)."""
    if zipfile.is_zipfile(path):
        z = zipfile.ZipFile(path)
        names = z.namelist()
        target = None
        if "manifest.xml" in names:
            m = ET.fromstring(z.read("manifest.xml"))
            _strip_ns(m)
            txt = (m.text or "").strip() if m.tag == "dae_root" else None
            if txt is None:
                e = m.find(".//dae_root")
                txt = (e.text or "").strip() if e is not None else None
            if txt:
                target = txt.lstrip("./")
        if target is None or target not in names:
            daes = [n for n in names if n.lower().endswith(".dae")]
            if not daes:
                raise IOError("%s: no .dae document inside the archive" % path)
            target = daes[0]
        root = ET.fromstring(z.read(target))
    else:
        root = ET.parse(path).getroot()
    _strip_ns(root)
    return root


def element_transform(elem, unit=1.0):
    """_ExtractFullTransform: product of the element's translate / rotate / matrix children, in order; lengths times the
    document's unit (_GetUnitScale, :4473-4486, 5805-5820)."""
    t = rm.IDENT_T
    for c in elem:
        if c.tag == "translate":
            v = _floats(c.text)
            t = rm.t_mul(t, (rm.IDENT_Q, (v[0] * unit, v[1] * unit, v[2] * unit)))
        elif c.tag == "rotate":
            v = _floats(c.text)
            t = rm.t_mul(t, (rm.quat_from_axis_angle((v[0], v[1], v[2]), v[3] * (rm.PI / 180.0)), (0.0, 0.0, 0.0)))
        elif c.tag == "matrix":
            v = _floats(c.text)
            q = rm.quat_from_matrix(((v[0], v[1], v[2]), (v[4], v[5], v[6]), (v[8], v[9], v[10])))
            t = rm.t_mul(t, (q, (v[3] * unit, v[7] * unit, v[11] * unit)))
        elif c.tag in ("scale", "skew", "lookat"):
            raise NotImplementedError("COLLADA <%s> transforms" % c.tag)
    return t


def _mathml_to_text(elem, jointname_of):
    """_ExtractMathML (colladareader.cpp:5275-5650): MathML content markup -> the fparser text OpenRAVE keeps as the equation."""
    tag = elem.tag
    if tag == "math":
        kids = list(elem)
        if len(kids) != 1:
            raise NotImplementedError("MathML <math> with %d children" % len(kids))
        return _mathml_to_text(kids[0], jointname_of)
    if tag == "cn":
        return repr(float((elem.text or "0").strip()))
    if tag == "pi":
        return "3.14159265358979323846"
    if tag == "csymbol":
        return jointname_of((elem.text or "").strip())
    if tag == "ci":
        return (elem.text or "").strip()
    if tag != "apply":
        raise NotImplementedError("MathML element <%s>" % tag)
    kids = list(elem)
    head = kids[0]
    op = head.tag
    arg = lambda k: _mathml_to_text(kids[k], jointname_of)
    args = lambda: [_mathml_to_text(k, jointname_of) for k in kids[1:]]
    nary = {"plus": "+", "times": "*", "and": "&", "or": "|"}
    binary = {"divide": "(%s/%s)", "quotient": "floor(%s/%s)", "rem": "(%s%%%s)", "power": "pow(%s,%s)", "eq": "(%s=%s)", "neq": "(%s!=%s)",
              "gt": "(%s>%s)", "lt": "(%s<%s)", "geq": "(%s>=%s)", "leq": "(%s<=%s)"}
    unary = {"not": "(!%s)", "floor": "floor(%s)", "ceiling": "ceil(%s)", "ln": "log(%s)", "arcsin": "asin(%s)", "arccos": "acos(%s)",
             "arctan": "atan(%s)", "arccosh": "acosh(%s)", "arcsinh": "asinh(%s)", "arctanh": "atanh(%s)"}
    if op in nary:
        return "(" + nary[op].join(args()) + ")"
    if op == "minus":
        return "(-%s)" % arg(1) if len(kids) == 2 else "(%s-%s)" % (arg(1), arg(2))
    if op in binary:
        return binary[op] % (arg(1), arg(2))
    if op in unary:
        return unary[op] % arg(1)
    if op in ("root", "log"):
        # <degree> / <logbase> qualifier first, if any
        qual = kids[1] if len(kids) > 2 and kids[1].tag in ("degree", "logbase") else None
        num = _mathml_to_text(kids[-1], jointname_of)
        q = _mathml_to_text(list(qual)[0], jointname_of) if qual is not None and len(qual) else None
        if op == "root":
            if q is None or float(q) == 2.0:
                return "sqrt(%s)" % num
            if float(q) == 1.0:
                return "(%s)" % num
            return "cbrt(%s)" % num if float(q) == 3.0 else "pow(%s,1.0/%s)" % (num, q)
        if q is None or float(q) == 10.0:
            return "log10(%s)" % num
        return "log2(%s)" % num if float(q) == 2.0 else "(log(%s)/log(%s))" % (num, q)
    if op == "csymbol":
        if head.attrib.get("encoding") == "text/xml":
            fn = (head.text or "").strip()
            if fn == "INRANGE":
                a, b, c = arg(1), arg(2), arg(3)
                return "((%s>=%s)&(%s<=%s))" % (a, b, a, c)
            if fn in ("SSSA", "SASA", "SASS"):
                return "%s(%s,%s,%s)" % (fn, arg(1), arg(2), arg(3))
            if fn == "atan2":
                return "atan2(%s,%s)" % (arg(1), arg(2))
            raise NotImplementedError("MathML csymbol function %r" % fn)
        return _mathml_to_text(head, jointname_of)
    if op in ("implies", "forall", "exists", "conjugate", "arg", "real", "imaginary", "lcm", "factorial", "xor", "arccot", "arccoth",
              "arccsc", "arccsch", "arcsec", "arcsech"):
        raise NotImplementedError("MathML operator <%s>" % op)       # (the reference emits acot() ...: names fparser does not know)
    return "%s(%s)" % (op, ",".join(args()))        # sin, cos, tan, abs, exp, min, max, ...: the tag is fparser's name


class ColladaRobotReader(object):
    def __init__(self, path):
        self.path = path
        self.root = read_document(path)
        self.by_id = {}
        self.parent = {}
        self._position_cache = {}
        # <asset><unit meter=...>: one unit per document (what the shipped files have); _GetUnitScale looks it up per element
        self.unit = 1.0
        units = set()
        for a in self.root.iter("asset"):
            u = a.find("unit")
            if u is not None and u.attrib.get("meter"):
                units.add(float(u.attrib["meter"]))
        if len(units) > 1:
            raise NotImplementedError("COLLADA document with several length units %s" % sorted(units))
        if units:
            self.unit = units.pop()
        for e in self.root.iter():
            i = e.attrib.get("id")
            if i:
                self.by_id[i] = e
            for c in e:
                self.parent[c] = e

    # ---------------------------------------------------------------- helpers
    def _url(self, url):
        if not url or not url.startswith("#"):
            raise NotImplementedError("external COLLADA reference %r" % url)
        e = self.by_id.get(url[1:])
        if e is None and "/" in url:
            # "#<id>/<id or sid of a descendant>/...": some exporters give every mesh's <vertices> the same id and address it through
            # its geometry (man1.zae, neuronics-katana.zae)
            parts = url[1:].split("/")
            e = self.by_id.get(parts[0])
            for name in parts[1:]:
                if e is None:
                    break
                e = next((d for d in e.iter() if d is not e and (d.attrib.get("id") == name or d.attrib.get("sid") == name)), None)
        if e is None:
            raise ValueError("%s: unresolved reference %s" % (self.path, url))
        return e

    @staticmethod
    def _openrave_technique(extra):
        for t in extra.findall("technique"):
            if t.attrib.get("profile") == "OpenRAVE":
                return t
        return None

    @staticmethod
    def _find_sid(scope, sid):
        for e in scope.iter():
            if e.attrib.get("sid") == sid:
                return e
        return None

    def _resolve_sidref(self, ref):
        """'kmodel1/joint3' or 'kmodel1/joint3/axis0' -> element (id first, then nested sids)."""
        parts = ref.split("/")
        e = self.by_id.get(parts[0])
        if e is None:
            return None
        for sid in parts[1:]:
            e = self._find_sid(e, sid)
            if e is None:
                return None
        return e

    # ---------------------------------------------------------------- main entry
    def load_into(self, body):
        """Append the document's first articulated system / kinematics model to `body` (a KinBody or Robot)."""
        root = self.root
        kmodels = root.findall("library_kinematics_models/kinematics_model")
        if not kmodels:
            raise ValueError("%s: no kinematics_model" % self.path)
        kmodel = kmodels[0]
        ktec = kmodel.find("technique_common")
        self.kmodel = kmodel
        self.kmodel_id = kmodel.attrib.get("id", "")
        # articulated systems: the kinematics one carries axis_info, the motion one wraps it
        askin, asmotion = None, None
        for a in root.findall("library_articulated_systems/articulated_system"):
            if a.find("kinematics") is not None and askin is None:
                askin = a
            if a.find("motion") is not None and asmotion is None:
                asmotion = a
        self.axis_info = {}
        if askin is not None:
            for ai in askin.findall("kinematics/technique_common/axis_info"):
                self.axis_info[ai.attrib.get("axis", "")] = ai
        # scene bindings: visual node of the model root + visual transform element of every joint axis
        self._read_bindings()
        if not body.name:
            nm = None
            for a in (asmotion, askin):
                if a is not None and a.attrib.get("name"):
                    nm = a.attrib["name"]
                    break
            ikm = root.find("library_kinematics_scenes/kinematics_scene/instance_articulated_system")
            if ikm is not None and ikm.attrib.get("name"):
                nm = ikm.attrib["name"]
            body.name = nm or kmodel.attrib.get("name") or self.kmodel_id
        self.body = body
        self.link_by_sid = {}
        self.joint_by_sid = {}
        self.joint_units = {}
        self.new_active, self.new_passive = [], []
        first_link = len(body.links)
        for ilink, le in enumerate(ktec.findall("link")):
            tnode = rm.IDENT_T
            node = self.root_node if ilink == 0 else None
            if node is not None:
                tnode = self._node_parent_transform(node)
            self._extract_link(le, node, tnode)
        self._extract_formulas(ktec)
        self._extract_collision_extras(kmodel)
        # DOF order: actuator-bound joints first, in attach_actuator order (:3720-3792)
        ordered = []
        for a in ([askin, asmotion] if askin is not None else [asmotion]):
            if a is None:
                continue
            for ex in a.findall("extra"):
                if ex.attrib.get("type") != "attach_actuator":
                    continue
                tec = self._openrave_technique(ex)
                if tec is None:
                    continue
                for b in tec.findall("bind_actuator"):
                    j = self._joint_from_ref(b.attrib.get("joint", ""))
                    if j is not None and j.is_active and j not in ordered:
                        ordered.append(j)
        rest = [j for j in self.new_active if j not in ordered]
        body.joints.extend(ordered + rest)
        body.passive_joints.extend(self.new_passive)
        if body.is_robot:
            for a in (askin, asmotion):
                if a is not None:
                    self._extract_manipulators(a)
        for l in body.links[first_link:]:
            l.update_collision()
        return body

    # ---------------------------------------------------------------- bindings
    def _read_bindings(self):
        root = self.root
        self.root_node = None
        self.axis_node = {}          # kinematic axis ref ('kmodel1/joint3/axis0') -> visual node carrying the bound transform
        scene = root.find("scene")
        iks = scene.find("instance_kinematics_scene") if scene is not None else None
        if iks is None:
            return
        vscene = None
        ivs = scene.find("instance_visual_scene")
        if ivs is not None:
            vscene = self._url(ivs.attrib.get("url"))
        # instantiate <instance_node> references in place so the visual hierarchy is one tree
        if vscene is not None:
            self._instantiate_nodes(vscene)
        bkm = iks.find("bind_kinematics_model")
        if bkm is not None and vscene is not None:
            self.root_node = self._resolve_visual(vscene, bkm.attrib.get("node", ""))
        for bja in iks.findall("bind_joint_axis"):
            target = bja.attrib.get("target", "")
            p = bja.find("axis/param")
            if p is None or vscene is None:
                continue
            # the param chain ends in '<...>_jointN.axisM'; the writer's newparam names keep that suffix
            m = re.search(r"(joint\d+)\.(axis\d+)$", (p.text or "").strip())
            if not m:
                continue
            vis = self._resolve_visual(vscene, target)
            if vis is None:
                continue
            node = self.parent.get(vis)
            self.axis_node["%s/%s/%s" % (self.kmodel_id_guess(), m.group(1), m.group(2))] = node

    def kmodel_id_guess(self):
        km = self.root.find("library_kinematics_models/kinematics_model")
        return km.attrib.get("id", "") if km is not None else ""

    def _instantiate_nodes(self, scope):
        for e in list(scope.iter("node")) + [scope]:
            for inst in list(e.findall("instance_node")):
                target = self._url(inst.attrib.get("url"))
                e.remove(inst)
                e.append(target)
                self.parent[target] = e
                self._instantiate_nodes(target)

    def _resolve_visual(self, vscene, ref):
        """'visual1/node0' : first part is a node id (or sid) in the visual scene, the rest nested sids."""
        parts = ref.split("/")
        e = self.by_id.get(parts[0])
        if e is None:
            e = self._find_sid(vscene, parts[0])
        for sid in parts[1:]:
            if e is None:
                return None
            e = self._find_sid(e, sid)
        return e

    def _node_parent_transform(self, node):
        t = rm.IDENT_T
        chain = []
        p = self.parent.get(node)
        while p is not None and p.tag == "node":
            chain.append(p)
            p = self.parent.get(p)
        for n in reversed(chain):
            t = rm.t_mul(t, element_transform(n, self.unit))
        return t

    def _node_world(self, node):
        return rm.t_mul(self._node_parent_transform(node), element_transform(node, self.unit))

    # ---------------------------------------------------------------- links and joints
    def _link_name(self, le):
        return le.attrib.get("name") or le.attrib.get("sid") or le.attrib.get("id") or ""

    def _extract_link(self, le, node, tparent):
        body = self.body
        name = self._link_name(le)
        link = body.GetLink(name)
        if link is None:
            link = Link(body, name)
            link.index = len(body.links)
            body.links.append(link)
        self.link_by_sid[le.attrib.get("sid", name)] = link
        link.t = rm.t_mul(tparent, element_transform(le, self.unit))
        if node is not None:
            self._extract_geometries(node, link)
        for att in le.findall("attachment_full"):
            jref = att.attrib.get("joint")
            if not jref:
                continue
            je = self._resolve_sidref(jref)
            if je is None:
                raise ValueError("%s: attachment joint %s not found" % (self.path, jref))
            if je.tag == "instance_joint":
                je = self._url(je.attrib.get("url"))
            tatt = element_transform(att, self.unit)
            axes = [c for c in je if c.tag in ("revolute", "prismatic")]
            if len(axes) != 1:
                raise NotImplementedError("joint %s has %d axes (only 1-DOF joints)" % (jref, len(axes)))
            ax = axes[0]
            axisref = "%s/%s" % (jref, ax.attrib.get("sid", "axis0"))
            info = self.axis_info.get(axisref)
            j = Joint(body, je.attrib.get("name") or "", "hinge" if ax.tag == "revolute" else "slider")
            j.sid = jref
            j.is_active = True
            if info is not None:
                act = self._resolve_bool(info, "active")
                if act is not None:
                    j.is_active = act
            if not j.name:
                j.name = "dummy%d" % (len(self.new_active) if j.is_active else -1)
            (self.new_active if j.is_active else self.new_passive).append(j)
            self.joint_by_sid[jref] = j
            self.joint_by_sid[jref.split("/")[-1]] = j
            child_le = att.find("link")
            if child_le is None:
                raise NotImplementedError("attachment_full without a child link")
            child_node = self.axis_node.get(axisref)
            child = self._extract_link(child_le, child_node, rm.t_mul(link.t, tatt))
            # axis / limits
            v = _floats(ax.find("axis").text)
            n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2]
            axis = rm.normalize3((v[0], v[1], v[2])) if n2 > 0 else (0.0, 0.0, 1.0)
            mult = rm.PI / 180.0 if j.type == "hinge" else self.unit          # (:2341)
            self.joint_units[j.name] = 1.0 if j.type == "hinge" else self.unit        # _mapJointUnits (:2091-2121)
            locked = self._resolve_bool(info, "locked") if info is not None else None
            has_soft = has_hard = False
            circular = None
            if info is not None:
                if locked:
                    j.lower = j.upper = 0.0
                else:
                    lim = info.find("limits")
                    if lim is not None:
                        has_soft = True
                        j.lower = mult * self._resolve_float(info, lim.find("min"))
                        j.upper = mult * self._resolve_float(info, lim.find("max"))
                        if j.type == "hinge" and (j.lower < -rm.PI or j.upper > rm.PI):
                            j.offset = 0.5 * (j.lower + j.upper)
                for np_ in info.findall("newparam"):
                    sid = np_.attrib.get("sid")
                    if sid == "circular":
                        b = np_.find("bool")
                        i = np_.find("int")
                        if b is not None:
                            circular = (b.text or "").strip().lower() in ("true", "1")
                        elif i is not None:
                            circular = int((i.text or "0").strip()) != 0
                    elif sid == "planning_weight":
                        f = np_.find("float")
                        if f is not None and float(f.text) > 0:
                            j.weight = float(f.text)
                    elif sid == "discretization_resolution":
                        f = np_.find("float")
                        if f is not None and float(f.text) > 0:
                            j.resolution = float(f.text)
            lim = ax.find("limits")
            if not locked and lim is not None:
                has_hard = True
                if not has_soft:
                    j.lower = float(lim.find("min").text) * mult
                    j.upper = float(lim.find("max").text) * mult
                    if j.type == "hinge" and (j.lower < -rm.PI or j.upper > rm.PI):
                        j.offset = 0.5 * (j.lower + j.upper)
            if circular is not None:
                j.circular = circular
            if not has_soft and not has_hard and not locked:
                if j.type == "hinge":
                    if circular is None:
                        j.circular = True
                    j.lower, j.upper = -rm.PI, rm.PI
                else:
                    j.lower, j.upper = -1000000.0, 1000000.0
            axis = rm.quat_rotate(tatt[0], axis)
            j.compute_internal(link, child, link.t, child.t, tatt[1], axis, None)
        return link

    @staticmethod
    def _resolve_bool(info, tag):
        e = info.find(tag)
        if e is None:
            return None
        b = e.find("bool")
        if b is not None:
            return (b.text or "").strip().lower() in ("true", "1")          # xsBoolean
        p = e.find("param")
        if p is not None:
            for np_ in info.findall("newparam"):
                if np_.attrib.get("sid") == (p.text or "").strip():
                    bb = np_.find("bool")
                    if bb is not None:
                        return (bb.text or "").strip().lower() in ("true", "1")
        return None

    @staticmethod
    def _resolve_float(info, e):
        f = e.find("float")
        if f is not None:
            return float(f.text)
        p = e.find("param")
        if p is not None:
            for np_ in info.findall("newparam"):
                if np_.attrib.get("sid") == (p.text or "").strip():
                    ff = np_.find("float")
                    if ff is not None:
                        return float(ff.text)
        if e.text and e.text.strip():
            return float(e.text)
        raise ValueError("unresolved float in axis_info")

    def _joint_from_ref(self, ref):
        j = self.joint_by_sid.get(ref)
        if j is None:
            j = self.joint_by_sid.get(ref.split("/")[-1])
        return j

    # ---------------------------------------------------------------- geometry
    def _extract_geometries(self, node, link):
        """ExtractGeometries (:2391-2506): instance_geometry of the node and of its non-joint child nodes."""
        bound = set(id(n) for n in self.axis_node.values() if n is not None)
        for child in node.findall("node"):
            if id(child) in bound:
                continue
            self._extract_geometries(child, link)
        igs = node.findall("instance_geometry")
        if not igs:
            return
        tnodegeom = rm.t_mul(rm.t_inv(link.t), self._node_world(node))
        for ig in igs:
            ge = self._url(ig.attrib.get("url"))
            tlocal = rm.IDENT_T
            primitive = None
            for ex in ge.findall("extra"):
                if ex.attrib.get("type") == "geometry_info":
                    tec = self._openrave_technique(ex)
                    if tec is not None:
                        for prim in ("container", "cage", "calibration_board"):
                            if tec.find(prim) is not None:
                                raise NotImplementedError("geometry_info primitive <%s>" % prim)
                        tlocal = element_transform(tec, self.unit)
                        primitive = self._geometry_info_primitive(tec, tlocal) or primitive
            if primitive is not None:
                # ExtractGeometry (:2956-3270): an analytic primitive in geometry_info replaces the mesh
                primitive.name = ge.attrib.get("name", "")
                primitive.t = rm.t_mul(tnodegeom, primitive.t)
                primitive.init_collision_mesh()
                link.geometries.append(primitive)
                continue
            mesh = ge.find("mesh")
            if mesh is None:
                continue
            if mesh.find("polygons") is not None and not any(mesh.find(k) is not None for k in ("triangles", "trifans", "tristrips", "polylist")):
                raise NotImplementedError("COLLADA <polygons> primitives")      # (the reference skips them too: :3291-3318)
            tlocalinv = rm.t_inv(tlocal)
            # one geometry per primitive element, in the reference's order of kinds (:3291-3318)
            for kind in ("triangles", "trifans", "tristrips", "polylist"):
                for tri in mesh.findall(kind):
                    soup = self._primitive_soup(tri, kind)
                    if soup is None:
                        continue
                    verts, idx = soup
                    g = Geometry("trimesh")
                    g.name = ge.attrib.get("name", "")
                    g.verts = rm_transform_points(tlocalinv, verts)
                    g.idx = idx
                    g.t = rm.t_mul(tnodegeom, tlocal)
                    link.geometries.append(g)

    @staticmethod
    def _geometry_info_primitive(tec, tlocal):
        """box / sphere / cylinder / cylinderz of a geometry_info technique (:2982-3050); None when it holds none."""
        def first_floats(elem, name, n):
            c = elem.find(name)
            v = _floats(c.text) if c is not None else []
            return v[:n] if len(v) >= n else None
        for child in tec:
            if child.tag == "box":
                v = first_floats(child, "half_extents", 3)
                if v:
                    g = Geometry("box"); g.data = (v[0], v[1], v[2]); g.t = tlocal
                    return g
            elif child.tag == "sphere":
                v = first_floats(child, "radius", 1)
                if v:
                    g = Geometry("sphere"); g.data = (v[0], 0.0, 0.0); g.t = tlocal
                    return g
            elif child.tag in ("cylinder", "cylinderz"):
                r, h = first_floats(child, "radius", 1), first_floats(child, "height", 1)
                if r and h:
                    g = Geometry("cylinder"); g.data = (r[0], h[0], 0.0)
                    # COLLADA cylinders run along y, OpenRAVE's along z (:3025-3027)
                    g.t = rm.t_mul(tlocal, (rm.quat_rotate_direction((0.0, 0.0, 1.0), (0.0, 1.0, 0.0)), (0.0, 0.0, 0.0))) if child.tag == "cylinder" else tlocal
                    return g
        return None

    def _primitive_soup(self, tri, kind):
        """Vertices and triangle indices of one <triangles> / <trifans> / <tristrips> / <polylist> element the way the reader
        emits them (:2566-2900): one vertex per index; fans and polygons as (first, i-1, i), strips with alternating winding."""
        stride, voff = 0, -1
        vsrc = None
        for inp in tri.findall("input"):
            off = int(inp.attrib.get("offset", "0"))
            if inp.attrib.get("semantic") == "VERTEX":
                voff = off
                vsrc = self._url(inp.attrib.get("source"))
            stride = max(stride, off)
        stride += 1
        if vsrc is None:
            return None
        pos = None
        for inp in vsrc.findall("input"):
            if inp.attrib.get("semantic") == "POSITION":
                pos = self._url(inp.attrib.get("source"))
        if pos is None:
            return None
        fl = self._position_cache.get(id(pos))                          # thousands of primitives may share one <source>
        if fl is None:
            fl = self._position_cache[id(pos)] = np.array((pos.find("float_array").text or "").split(), dtype=np.float64).reshape(-1, 3) * self.unit
        count = int(tri.attrib.get("count", "0"))
        plists = [np.array([int(x) for x in (p.text or "").split()], dtype=np.int64) for p in tri.findall("p")]
        if kind == "triangles":
            if not plists:
                return None
            vidx = plists[0][voff::stride][:3 * count]
            verts = fl[vidx]
            return verts, np.arange(verts.shape[0], dtype=np.int32).reshape(-1, 3)
        verts, tris = [], []
        if kind in ("trifans", "tristrips"):
            for p in plists[:count]:
                vidx = p[voff::stride]
                start = len(verts)
                verts.extend(fl[vidx])
                flip = False
                for iv in range(start + 2, len(verts)):
                    if kind == "trifans":
                        tris.append((start, iv - 1, iv))
                    else:
                        tris.append((iv - 2, iv, iv - 1) if flip else (iv - 2, iv - 1, iv))
                        flip = not flip
        else:   # polylist
            vc = tri.find("vcount")
            counts = [int(x) for x in (vc.text or "").split()] if vc is not None else []
            p = plists[0] if plists else np.zeros(0, dtype=np.int64)
            k = voff
            for nv in counts:
                if nv > 0 and k + (nv - 1) * stride < len(p):
                    start = len(verts)
                    for _ in range(nv):
                        verts.append(fl[p[k]])
                        k += stride
                    for iv in range(start + 2, len(verts)):
                        tris.append((start, iv - 1, iv))
        if not tris:
            return None
        return np.array(verts, dtype=np.float64).reshape(-1, 3), np.array(tris, dtype=np.int32).reshape(-1, 3)

    # ---------------------------------------------------------------- formulas, extras, manipulators
    def _extract_formulas(self, ktec):
        for f in ktec.findall("formula"):
            tp = f.find("target/param")
            if tp is None:
                continue
            j = self._joint_from_ref((tp.text or "").strip())
            if j is None:
                continue

            def jointname_of(ref):
                jj = self._joint_from_ref(ref)
                if jj is None:
                    raise ValueError("formula references unknown joint %s" % ref)
                u = self.joint_units.get(jj.name, 1.0)
                return jj.name if u == 1.0 else "(%r*%s)" % (1.0 / u, jj.name)        # (:5623-5628)
            tec = None
            for t in f.findall("technique"):
                if t.attrib.get("profile") == "OpenRAVE":
                    tec = t
            eq = None
            if tec is not None:
                for e in tec.findall("equation"):
                    if e.attrib.get("type") == "position":
                        eq = _mathml_to_text(list(e)[0], jointname_of)
            elif f.find("technique_common") is not None:
                kids = list(f.find("technique_common"))
                if kids:
                    eq = _mathml_to_text(kids[0], jointname_of)
            if eq:
                tu = self.joint_units.get(j.name, 1.0)
                if tu != 1.0:
                    eq = "%r*(%s)" % (tu, eq)                                        # (:1618-1620)
                j.mimic = MimicExpr(eq, names=[jj.name for jj in self.body.all_joints() + self.new_active + self.new_passive])

    def _link_from_ref(self, ref):
        e = self._resolve_sidref(ref)
        if e is not None and e.tag == "link":
            return self.body.GetLink(self._link_name(e))
        return self.link_by_sid.get(ref.split("/")[-1])

    def _extract_collision_extras(self, kmodel):
        for ex in kmodel.findall("extra"):
            if ex.attrib.get("type") != "collision":
                continue
            tec = self._openrave_technique(ex)
            if tec is None:
                continue
            for e in tec:
                if e.tag == "ignore_link_pair":
                    l0 = self._link_from_ref(e.attrib.get("link0", ""))
                    l1 = self._link_from_ref(e.attrib.get("link1", ""))
                    if l0 is not None and l1 is not None:
                        a, b = l0.index, l1.index
                        self.body.forced_adjacent.add((min(a, b), max(a, b)))
                elif e.tag == "link_collision_state":
                    l = self._link_from_ref(e.attrib.get("link", ""))
                    b = e.find("bool")
                    if l is not None and b is not None:
                        l.enabled = (b.text or "").strip().lower() in ("true", "1")
                elif e.tag == "bind_instance_geometry":
                    pass    # geometry groups are not on the collision path

    def _extract_manipulators(self, asys):
        for ex in asys.findall("extra"):
            if ex.attrib.get("type") != "manipulator":
                continue
            tec = self._openrave_technique(ex)
            if tec is None:
                continue
            m = Manipulator(self.body, ex.attrib.get("name", ""))
            fo, ft = tec.find("frame_origin"), tec.find("frame_tip")
            if fo is not None:
                m.base = self._link_from_ref(fo.attrib.get("link", ""))
            if ft is not None:
                m.effector = self._link_from_ref(ft.attrib.get("link", ""))
                m.t_local = element_transform(ft, self.unit)
            for gj in tec.findall("gripper_joint"):
                j = self._joint_from_ref(gj.attrib.get("joint", ""))
                if j is not None:
                    m.gripper_joint_names.append(j.name)
            if m.base is not None and m.effector is not None:
                self.body.manipulators.append(m)


def rm_transform_points(t, pts):
    from .kinbody import transform_points
    return transform_points(t, np.ascontiguousarray(pts, dtype=np.float64))


def load_collada_into(body, path):
    return ColladaRobotReader(path).load_into(body)
