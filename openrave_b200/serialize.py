"""Flat-scene fixtures: save / load a parsed Environment as one .npz.

The GPU box has no /root/reference, so the benchmark scenes are parsed once here (tools/make_fixtures.py)
and committed under tests/golden/scenes/.  The file stores what the XML reader produced BEFORE
`_ComputeInternalInformation` (link poses, link-frame meshes, raw joint frames), so loading re-runs the
same host logic as a fresh parse.  Identical meshes (the six mugs, the finger links) are stored once.
"""
import hashlib
import json

import numpy as np

from .kinbody import Geometry, Joint, KinBody, Link, Manipulator, MimicExpr, Robot


def _t_list(t):
    return [list(t[0]), list(t[1])]


def _t_tuple(v):
    return (tuple(float(x) for x in v[0]), tuple(float(x) for x in v[1]))


def body_to_dict(body, meshes):
    links = []
    for l in body.links:
        key = hashlib.sha1(l.verts.tobytes() + l.idx.tobytes()).hexdigest()[:16]
        if key not in meshes:
            meshes[key] = (l.verts, l.idx)
        links.append(dict(name=l.name, static=l.static, enabled=l.enabled, t=_t_list(l.raw_t), mesh=key,
                          ngeom=len(l.geometries)))
    joints = []
    for j in body.joints + body.passive_joints:
        r = j.raw
        joints.append(dict(name=j.name, type=j.type, active=j.is_active, links=[r["l0"], r["l1"]], lower=r["lower"],
                           upper=r["upper"], offset=j.offset, circular=j.circular, resolution=j.resolution, weight=j.weight,
                           mimic=j.mimic.text if j.mimic is not None else None, axis=list(r["axis"]), anchor=list(r["anchor"]),
                           tl=_t_list(r["tl"]), tr=_t_list(r["tr"]), tlo=_t_list(r["tlo"]), tro=_t_list(r["tro"])))
    d = dict(name=body.name, robot=body.is_robot, links=links, joints=joints,
             forced_adjacent=sorted(list(p) for p in body.forced_adjacent),
             make_joined_links_adjacent=body.make_joined_links_adjacent, uri=body.uri)
    if body.is_robot:
        d["manipulators"] = [dict(name=m.name, base=m.base.name if m.base else None, effector=m.effector.name if m.effector else None,
                                  gripper=m.gripper_joint_names, t_local=_t_list(m.t_local)) for m in body.manipulators]
    return d


def body_from_dict(d, meshes, env=None):
    body = Robot(env) if d["robot"] else KinBody(env)
    body.name = d["name"]
    body.uri = d.get("uri", "")
    body.make_joined_links_adjacent = d["make_joined_links_adjacent"]
    for i, ld in enumerate(d["links"]):
        l = Link(body, ld["name"])
        l.index = i
        l.static, l.enabled = ld["static"], ld["enabled"]
        l.t = _t_tuple(ld["t"])
        v, idx = meshes[ld["mesh"]]
        l.verts, l.idx = v.copy(), idx.copy()
        l.geometries = [Geometry("trimesh") for _ in range(ld["ngeom"])]   # count only (adjacency rule needs it)
        body.links.append(l)
    for jd in d["joints"]:
        j = Joint(body, jd["name"], jd["type"])
        j.is_active = jd["active"]
        j.links = [body.links[jd["links"][0]], body.links[jd["links"][1]]]
        j.lower, j.upper, j.offset = jd["lower"], jd["upper"], jd["offset"]
        j.circular, j.resolution, j.weight = jd["circular"], jd["resolution"], jd["weight"]
        j.mimic = MimicExpr(jd["mimic"], names=[x["name"] for x in d["joints"]]) if jd["mimic"] else None
        j.axis = tuple(jd["axis"])
        j.anchor = tuple(jd["anchor"])
        j.t_left_nooffset, j.t_right_nooffset = _t_tuple(jd["tl"]), _t_tuple(jd["tr"])
        j.t_left, j.t_right = _t_tuple(jd["tlo"]), _t_tuple(jd["tro"])
        (body.joints if j.is_active else body.passive_joints).append(j)
    body.forced_adjacent = {tuple(p) for p in d["forced_adjacent"]}
    if d["robot"]:
        for md in d.get("manipulators", []):
            m = Manipulator(body, md["name"])
            m.base = body.GetLink(md["base"]) if md["base"] else None
            m.effector = body.GetLink(md["effector"]) if md["effector"] else None
            m.gripper_joint_names = md["gripper"]
            m.t_local = _t_tuple(md["t_local"])
            body.manipulators.append(m)
    return body


def save_environment(env, path):
    meshes = {}
    bodies = [body_to_dict(b, meshes) for b in env.GetBodies()]
    arrays = {}
    for k, (v, i) in meshes.items():
        arrays["v_" + k] = v
        arrays["i_" + k] = i
    arrays["scene_json"] = np.frombuffer(json.dumps(dict(version=1, bodies=bodies)).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


def load_environment_into(env, path):
    z = np.load(path)
    scene = json.loads(bytes(z["scene_json"]).decode())
    meshes = {}
    for k in z.files:
        if k.startswith("v_"):
            meshes[k[2:]] = (z[k], z["i_" + k[2:]])
    for bd in scene["bodies"]:
        env.Add(body_from_dict(bd, meshes, env))
    return env
