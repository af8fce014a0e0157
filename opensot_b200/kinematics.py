"""Kinematic trees for the on-device Cartesian / CoM producers (include/opensot_b200.h: osb_kinematics_create,
osb_produce_kinematics_device): a URDF becomes the flat joint table the C ABI takes. Host-side, once per robot.

The reference reaches the same quantities through XBot::ModelInterface (getJacobian, getRelativeJacobian, getPose,
getCOM, getCOMJacobian in src/tasks/velocity/Cartesian.cpp:68-105 and CoM.cpp:59-74); joint order here is the URDF's
order of the movable joints."""
import ctypes as C
import xml.etree.ElementTree as ET

import numpy as np

FIXED, REVOLUTE, PRISMATIC = 0, 1, 2
CARTESIAN, COM = 0, 1


class KinTreeC(C.Structure):
    _fields_ = [("n_joints", C.c_int32), ("n_dof", C.c_int32), ("parent", C.c_void_p), ("type", C.c_void_p),
                ("dof", C.c_void_p), ("axis", C.c_void_p), ("origin", C.c_void_p), ("mass", C.c_void_p), ("com", C.c_void_p)]


class KinTaskC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("distal", C.c_int32), ("base", C.c_int32), ("level", C.c_int32), ("row0", C.c_int32),
                ("reserved", C.c_int32), ("lambda_", C.c_double), ("orientation_gain", C.c_double),
                ("desired", C.c_void_p), ("desired_stride", C.c_int64), ("desired_vel", C.c_void_p),
                ("desired_vel_stride", C.c_int64)]


class StatesConfigC(C.Structure):
    _fields_ = [("blob_template", C.c_void_p), ("q_ref", C.c_void_p), ("q_min", C.c_void_p), ("q_max", C.c_void_p),
                ("q_neutral", C.c_void_p), ("qdot_max", C.c_void_p), ("lambda_", C.c_double), ("bound_scaling", C.c_double),
                ("dT", C.c_double), ("postural_level", C.c_int32), ("joint_limit_set", C.c_int32),
                ("velocity_limit_set", C.c_int32), ("n_tasks", C.c_int32), ("tasks", C.c_void_p)]


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


class KinTree:
    """joints (fixed ones included) in parents-first order; joint j carries the frame of its child link.
    Row 0 is a fixed joint holding the root link at the world origin."""

    def __init__(self, names, links, parent, jtype, dof, axis, origin, mass, com, dof_names, lower, upper, velocity):
        self.names, self.links = names, links
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.type = np.ascontiguousarray(jtype, dtype=np.int32)
        self.dof = np.ascontiguousarray(dof, dtype=np.int32)
        self.axis = np.ascontiguousarray(axis, dtype=np.float64)
        self.origin = np.ascontiguousarray(origin, dtype=np.float64)
        self.mass = np.ascontiguousarray(mass, dtype=np.float64)
        self.com = np.ascontiguousarray(com, dtype=np.float64)
        self.dof_names = dof_names
        self.q_min, self.q_max, self.qdot_max = (np.asarray(v, dtype=np.float64) for v in (lower, upper, velocity))
        self.n_joints, self.n_dof = len(names), len(dof_names)

    def link(self, name):
        """index of the joint whose child link is `name` (what distal / base arguments take)"""
        return self.links.index(name)

    def c_struct(self):
        t = KinTreeC()
        t.n_joints, t.n_dof = self.n_joints, self.n_dof
        for f in ("parent", "type", "dof", "axis", "origin", "mass", "com"):
            setattr(t, f, getattr(self, f).ctypes.data)
        return t


def load_urdf(path):
    root = ET.parse(path).getroot()
    links = {}
    for l in root.findall("link"):
        m, c = 0.0, np.zeros(3)
        ine = l.find("inertial")
        if ine is not None:
            if ine.find("mass") is not None:
                m = float(ine.find("mass").get("value"))
            o = ine.find("origin")
            if o is not None and o.get("xyz"):
                c = np.array([float(v) for v in o.get("xyz").split()])
        links[l.get("name")] = (m, c)
    joints = []
    for j in root.findall("joint"):
        o = j.find("origin")
        xyz = np.array([float(v) for v in (o.get("xyz") or "0 0 0").split()]) if o is not None else np.zeros(3)
        rpy = [float(v) for v in (o.get("rpy") or "0 0 0").split()] if o is not None else [0.0, 0.0, 0.0]
        ax = j.find("axis")
        axis = np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None else np.array([1.0, 0.0, 0.0])
        lim = j.find("limit")
        lo = float(lim.get("lower", "-1e20")) if lim is not None else -1e20
        hi = float(lim.get("upper", "1e20")) if lim is not None else 1e20
        vel = float(lim.get("velocity", "1e20")) if lim is not None else 1e20
        t = j.get("type")
        if t not in ("fixed", "revolute", "continuous", "prismatic"):
            raise ValueError("joint %s: type %s is not supported (fixed-base trees of Euclidean joints)" % (j.get("name"), t))
        joints.append(dict(name=j.get("name"), type=t, parent=j.find("parent").get("link"), child=j.find("child").get("link"),
                           R=_rpy(*rpy), p=xyz, axis=axis, lo=lo, hi=hi, vel=vel))
    children = {j["child"] for j in joints}
    roots = [n for n in links if n not in children]
    if len(roots) != 1:
        raise ValueError("expected one root link, found %s" % roots)
    dof_names = [j["name"] for j in joints if j["type"] != "fixed"]           # URDF order of the movable joints
    names, lnk, parent, jtype, dof, axis, origin, mass, com = [], [], [], [], [], [], [], [], []

    def add(name, link, par, ty, d, ax, R, p):
        names.append(name); lnk.append(link); parent.append(par); jtype.append(ty); dof.append(d); axis.append(ax)
        origin.append(np.concatenate([R.reshape(-1), p])); mass.append(links[link][0]); com.append(links[link][1])
        return len(names) - 1
    idx = {roots[0]: add("__root__", roots[0], -1, FIXED, -1, np.array([1.0, 0, 0]), np.eye(3), np.zeros(3))}
    todo = list(joints)
    while todo:                                                               # parents first
        rest = []
        for j in todo:
            if j["parent"] in idx:
                ty = FIXED if j["type"] == "fixed" else (PRISMATIC if j["type"] == "prismatic" else REVOLUTE)
                idx[j["child"]] = add(j["name"], j["child"], idx[j["parent"]], ty,
                                      dof_names.index(j["name"]) if ty != FIXED else -1, j["axis"], j["R"], j["p"])
            else:
                rest.append(j)
        if len(rest) == len(todo):
            raise ValueError("disconnected joints: %s" % [j["name"] for j in rest])
        todo = rest
    byname = {j["name"]: j for j in joints}
    return KinTree(names, lnk, parent, jtype, dof, axis, origin, mass, com, dof_names,
                   [byname[n]["lo"] for n in dof_names], [byname[n]["hi"] for n in dof_names], [byname[n]["vel"] for n in dof_names])


def load_npz(path):
    """a tree saved with numpy.savez (tests/golden/robots/coman_tree.npz)"""
    z = np.load(path)
    return KinTree([str(v) for v in z["names"]], [str(v) for v in z["links"]], z["parent"], z["type"], z["dof"], z["axis"],
                   z["origin"], z["mass"], z["com"], [str(v) for v in z["dof_names"]], z["q_min"], z["q_max"], z["qdot_max"])
