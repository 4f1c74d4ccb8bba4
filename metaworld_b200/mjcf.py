"""MJCF -> mjModel-like constant tables (host side, run once per model).

This is the *model loader* for the hot path: it plays the role that MuJoCo's
XML compiler plays for the reference (``MujocoEnv.__init__`` ->
``MjModel.from_xml_path``; reference call site ``metaworld/sawyer_xyz_env.py:53-63``).
It understands exactly the MJCF feature set the 36 Meta-World task models use
(SURVEY.md Appendix B): includes, default classes / childclass, bodies with
pos + quat/euler/xyaxes, explicit ``<inertial>`` or geom-inferred inertia
(``inertiafromgeom=auto`` limited to ``inertiagrouprange``), hinge / slide /
free joints, primitive + mesh geoms (mesh -> convex hull, geom frame re-centred
on the mesh inertial frame), sites, one mocap body, weld equalities and
``position`` actuators.

The output (`Model`) uses MuJoCo's array names (body_pos, jnt_axis, geom_size,
...) so host glue can be written against the same vocabulary as the reference's
``self.model.body(name).pos`` / ``self.data.site(name).xpos`` accessors.

Nothing here touches the GPU.  The lowering of a `Model` to the flat device
blob lives in ``lower.py``.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

# geom type ids follow MuJoCo's mjtGeom enumeration
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
_GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4, "cylinder": 5, "box": 6, "mesh": 7}
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)
_JNT_TYPES = {"free": 0, "ball": 1, "slide": 2, "hinge": 3}

MINVAL = 1e-15


# --------------------------------------------------------------------------- small math
def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(v) < n and default is not None:
        d = np.array(default, dtype=np.float64)
        d[: len(v)] = v
        v = d
    return v


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_norm(q):
    n = np.linalg.norm(q)
    if n < MINVAL:
        return np.array([1.0, 0, 0, 0])
    return q / n


def quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat2quat(m):
    """Rotation matrix -> unit quaternion (w,x,y,z), w-largest branch ordering."""
    m = np.asarray(m)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        w = 0.5 * np.sqrt(1 + tr)
        q = np.array([w, (m[2, 1] - m[1, 2]) / (4 * w), (m[0, 2] - m[2, 0]) / (4 * w), (m[1, 0] - m[0, 1]) / (4 * w)])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        x = 0.5 * np.sqrt(1 + m[0, 0] - m[1, 1] - m[2, 2])
        q = np.array([(m[2, 1] - m[1, 2]) / (4 * x), x, (m[0, 1] + m[1, 0]) / (4 * x), (m[0, 2] + m[2, 0]) / (4 * x)])
    elif m[1, 1] > m[2, 2]:
        y = 0.5 * np.sqrt(1 - m[0, 0] + m[1, 1] - m[2, 2])
        q = np.array([(m[0, 2] - m[2, 0]) / (4 * y), (m[0, 1] + m[1, 0]) / (4 * y), y, (m[1, 2] + m[2, 1]) / (4 * y)])
    else:
        z = 0.5 * np.sqrt(1 - m[0, 0] - m[1, 1] + m[2, 2])
        q = np.array([(m[1, 0] - m[0, 1]) / (4 * z), (m[0, 2] + m[2, 0]) / (4 * z), (m[1, 2] + m[2, 1]) / (4 * z), z])
    return quat_norm(q)


def euler2quat(e):
    """Intrinsic x-y-z sequence (MuJoCo default eulerseq='xyz'), radians."""
    q = np.array([1.0, 0, 0, 0])
    for i, ang in enumerate(e):
        r = np.zeros(4)
        r[0] = np.cos(ang / 2)
        r[i + 1] = np.sin(ang / 2)
        q = quat_mul(q, r)
    return q


def eig3(mat):
    """Symmetric 3x3 eigen-decomposition by Jacobi rotations accumulated in a
    quaternion, eigenvalues sorted in decreasing order (the convention MuJoCo's
    compiler uses for inertial frames [3P]).  Returns (eigval, quat)."""
    eps = 1e-12
    quat = np.array([1.0, 0, 0, 0])
    mat = np.asarray(mat, dtype=np.float64)
    eigval = np.zeros(3)
    for _ in range(500):
        R = quat2mat(quat)
        D = R.T @ mat @ R
        eigval = np.array([D[0, 0], D[1, 1], D[2, 2]])
        a01, a02, a12 = abs(D[0, 1]), abs(D[0, 2]), abs(D[1, 2])
        if a01 > a02 and a01 > a12:
            rk, ck, rotk = 0, 1, 2
        elif a02 > a12:
            rk, ck, rotk = 0, 2, 1
        else:
            rk, ck, rotk = 1, 2, 0
        if abs(D[rk, ck]) < eps:
            break
        tau = (D[ck, ck] - D[rk, rk]) / (2 * D[rk, ck])
        if tau >= 0:
            t = 1.0 / (tau + np.sqrt(1 + tau * tau))
        else:
            t = -1.0 / (-tau + np.sqrt(1 + tau * tau))
        c = 1.0 / np.sqrt(1 + t * t)
        if c > 1.0 - eps:
            break
        tmp = np.zeros(4)
        tmp[rotk + 1] = -np.sqrt(0.5 - 0.5 * c) if tau >= 0 else np.sqrt(0.5 - 0.5 * c)
        if rotk == 1:
            tmp[rotk + 1] = -tmp[rotk + 1]
        tmp[0] = np.sqrt(1.0 - tmp[rotk + 1] ** 2)
        tmp = quat_norm(tmp)
        quat = quat_norm(quat_mul(quat, tmp))
    for j in range(3):
        j1 = j % 2
        if eigval[j1] + eps < eigval[j1 + 1]:
            eigval[j1], eigval[j1 + 1] = eigval[j1 + 1], eigval[j1]
            tmp = np.zeros(4)
            tmp[0] = 0.707106781186548
            tmp[(j1 + 2) % 3 + 1] = tmp[0]
            quat = quat_norm(quat_mul(quat, tmp))
    return eigval, quat


# --------------------------------------------------------------------------- meshes
def load_stl(path):
    """Binary STL -> (verts[nv,3], faces[nf,3]) with duplicate vertices merged."""
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack("<I", data[80:84])[0]
    if 84 + 50 * n != len(data):
        raise ValueError(f"{path}: not a binary STL")
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
    tri = rec["v"].astype(np.float64).reshape(-1, 3)
    verts, inv = np.unique(tri, axis=0, return_inverse=True)
    faces = inv.reshape(-1, 3)
    return verts, faces


def mesh_inertial(verts, faces):
    """Volume, centre of mass and inertia (unit density, about the COM) of a
    triangle mesh; tetrahedra are taken from the area-weighted face centroid and
    their volumes are summed in absolute value (MuJoCo 'legacy' convention [3P];
    identical to the exact signed sum for convex watertight meshes)."""
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    cen = (a + b + c) / 3
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    keep = area > 1e-14
    a, b, c, cen, area = a[keep], b[keep], c[keep], cen[keep], area[keep]
    facecen = (cen * area[:, None]).sum(0) / area.sum()
    a0, b0, c0 = a - facecen, b - facecen, c - facecen
    vol = np.abs(np.einsum("ij,ij->i", a0, np.cross(b0, c0))) / 6
    V = vol.sum()
    com = facecen + ((a0 + b0 + c0) / 4 * vol[:, None]).sum(0) / V
    # inertia about COM: integrate over tetrahedra (com, a, b, c)
    a1, b1, c1 = a - com, b - com, c - com
    # second moments P = int x x^T dV for tet with one vertex at origin:
    # vol/20 * (sum_i v_i v_i^T + (sum_i v_i)(sum_i v_i)^T)
    s = a1 + b1 + c1
    P = (np.einsum("n,ni,nj->ij", vol, a1, a1) + np.einsum("n,ni,nj->ij", vol, b1, b1)
         + np.einsum("n,ni,nj->ij", vol, c1, c1) + np.einsum("n,ni,nj->ij", vol, s, s)) / 20
    I = np.trace(P) * np.eye(3) - P
    return V, com, I


def convex_hull(verts):
    from scipy.spatial import ConvexHull

    h = ConvexHull(verts)
    idx = np.unique(h.simplices)
    remap = -np.ones(len(verts), dtype=np.int64)
    remap[idx] = np.arange(len(idx))
    hv = verts[idx]
    hf = remap[h.simplices]
    # orient faces outward
    c = hv.mean(0)
    a, b, cc = hv[hf[:, 0]], hv[hf[:, 1]], hv[hf[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, cc - a), a - c) < 0
    hf[flip] = hf[flip][:, ::-1]
    return hv, hf


# --------------------------------------------------------------------------- model container
@dataclass
class Model:
    """mjModel-like constant tables (float64 numpy) + name maps."""
    path: str = ""
    opt: dict = field(default_factory=dict)
    names: dict = field(default_factory=dict)  # kind -> list of names (index = id)
    arrays: dict = field(default_factory=dict)
    meshes: list = field(default_factory=list)  # per mesh id: dict(vert, face) hull in mesh frame

    def __getattr__(self, k):
        a = self.__dict__.get("arrays")
        if a is not None and k in a:
            return a[k]
        raise AttributeError(k)

    def name2id(self, kind, name):
        try:
            return self.names[kind].index(name)
        except ValueError:
            raise KeyError(f"no {kind} named {name!r} in {self.path}")

    @property
    def nq(self):
        return int(self.arrays["qpos0"].shape[0])

    @property
    def nv(self):
        return int(self.arrays["dof_jntid"].shape[0])

    @property
    def nbody(self):
        return len(self.names["body"])

    @property
    def ngeom(self):
        return len(self.names["geom"])


# --------------------------------------------------------------------------- parser
_GEOM_DEF = dict(type="sphere", size="0 0 0", pos="0 0 0", contype="1", conaffinity="1", condim="3", group="0",
                 priority="0", friction="1 0.005 0.0001", solmix="1", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2",
                 margin="0", gap="0", density="1000")
_JNT_DEF = dict(type="hinge", pos="0 0 0", axis="0 0 1", stiffness="0", springref="0", ref="0", damping="0",
                armature="0", margin="0", frictionloss="0", solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2")
_SITE_DEF = dict(pos="0 0 0")


class _Defaults:
    def __init__(self, parent=None):
        self.parent = parent
        self.d = {} if parent is None else {k: dict(v) for k, v in parent.d.items()}

    def update(self, tag, attrib):
        self.d.setdefault(tag, {}).update(attrib)

    def get(self, tag):
        return self.d.get(tag, {})


def _expand_includes(elem, base_dir):
    """Replace <include file=...> by the children of the included file's root
    (paths relative to the directory of the top-level model file)."""
    out = []
    for ch in list(elem):
        if ch.tag == "include":
            p = os.path.normpath(os.path.join(base_dir, ch.attrib["file"]))
            root = ET.parse(p).getroot()
            _expand_includes(root, base_dir)
            out.extend(list(root))
        else:
            _expand_includes(ch, base_dir)
            out.append(ch)
    for ch in list(elem):
        elem.remove(ch)
    for ch in out:
        elem.append(ch)


def _orientation(attrib):
    if "quat" in attrib:
        return quat_norm(_vec(attrib["quat"]))
    if "euler" in attrib:
        return euler2quat(_vec(attrib["euler"]))
    if "xyaxes" in attrib:
        v = _vec(attrib["xyaxes"])
        x = v[:3] / np.linalg.norm(v[:3])
        y = v[3:] - x * np.dot(x, v[3:])
        y /= np.linalg.norm(y)
        z = np.cross(x, y)
        return mat2quat(np.stack([x, y, z], axis=1))
    if "axisangle" in attrib or "zaxis" in attrib:
        raise NotImplementedError("axisangle/zaxis orientation")
    return np.array([1.0, 0, 0, 0])


def _geom_volume_inertia(gtype, size):
    """Volume and unit-density inertia diagonal in the geom frame."""
    if gtype == GEOM_SPHERE:
        r = size[0]
        v = 4.0 / 3.0 * np.pi * r ** 3
        return v, np.full(3, 0.4 * v * r * r)
    if gtype == GEOM_BOX:
        v = 8 * size[0] * size[1] * size[2]
        return v, v / 3.0 * np.array([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2, size[0] ** 2 + size[1] ** 2])
    if gtype == GEOM_CYLINDER:
        r, h = size[0], size[1]
        v = np.pi * r * r * 2 * h
        ixx = v * (3 * r * r + 4 * h * h) / 12
        return v, np.array([ixx, ixx, v * r * r / 2])
    if gtype == GEOM_CAPSULE:
        r, h = size[0], size[1]
        vc = np.pi * r * r * 2 * h
        vs = 4.0 / 3.0 * np.pi * r ** 3
        v = vc + vs
        # cylinder + two hemispheres (each hemisphere COM at 3r/8 from its flat face)
        izz = vc * r * r / 2 + vs * 0.4 * r * r
        ixx = vc * (3 * r * r + 4 * h * h) / 12 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
        return v, np.array([ixx, ixx, izz])
    if gtype == GEOM_ELLIPSOID:
        v = 4.0 / 3.0 * np.pi * size[0] * size[1] * size[2]
        return v, v / 5 * np.array([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2, size[0] ** 2 + size[1] ** 2])
    return 0.0, np.zeros(3)


def load(path) -> Model:
    path = os.path.abspath(path)
    base_dir = os.path.dirname(path)
    root = ET.parse(path).getroot()
    _expand_includes(root, base_dir)

    # ---- compiler / option
    comp = dict(angle="degree", inertiafromgeom="auto", inertiagrouprange="0 5", meshdir="")
    opt = dict(timestep=0.002, iterations=100, tolerance=1e-8, impratio=1.0, gravity=np.array([0, 0, -9.81]),
               cone="pyramidal", solver="Newton")
    for e in root.findall("compiler"):
        comp.update({k: v for k, v in e.attrib.items() if k in comp})
    for e in root.findall("option"):
        for k, v in e.attrib.items():
            if k in ("timestep", "tolerance", "impratio"):
                opt[k] = float(v)
            elif k == "iterations":
                opt[k] = int(v)
            elif k == "gravity":
                opt[k] = _vec(v)
            else:
                opt[k] = v
    if comp["angle"] != "radian":
        raise NotImplementedError("only angle=radian models are supported")
    grp_lo, grp_hi = (int(x) for x in comp["inertiagrouprange"].split())

    # ---- defaults
    classes = {"main": _Defaults()}

    def parse_default(elem, cur):
        for ch in elem:
            if ch.tag == "default":
                name = ch.attrib["class"]
                classes[name] = _Defaults(cur)
                parse_default(ch, classes[name])
            else:
                cur.update(ch.tag, ch.attrib)

    # two passes so nested classes see main-level settings made in any include
    for e in root.findall("default"):
        for ch in e:
            if ch.tag != "default":
                classes["main"].update(ch.tag, ch.attrib)
    for e in root.findall("default"):
        for ch in e:
            if ch.tag == "default":
                name = ch.attrib["class"]
                classes[name] = _Defaults(classes["main"])
                parse_default(ch, classes[name])

    def resolve(tag, elem, childclass, base):
        cls = elem.attrib.get("class", childclass or "main")
        a = dict(base)
        a.update(classes[cls].get(tag))
        a.update(elem.attrib)
        return a

    # ---- assets: meshes
    mesh_names, mesh_data = [], []
    for asset in root.findall("asset"):
        for e in asset.findall("mesh"):
            a = dict(classes["main"].get("mesh"))
            a.update(e.attrib)
            name = a.get("name") or os.path.splitext(os.path.basename(a["file"]))[0]
            if name in mesh_names:
                continue
            mesh_names.append(name)
            mesh_data.append(dict(file=os.path.normpath(os.path.join(base_dir, comp["meshdir"], a["file"])),
                                  scale=_vec(a.get("scale"), 3, [1, 1, 1]), loaded=False))

    def get_mesh(mid):
        md = mesh_data[mid]
        if not md["loaded"]:
            v, f = load_stl(md["file"])
            v = v * md["scale"]
            if np.prod(md["scale"]) < 0:
                f = f[:, ::-1]
            vol, com, I = mesh_inertial(v, f)
            ev, q = eig3(I)
            R = quat2mat(q)
            md.update(loaded=True, volume=vol, pos=com, quat=q, inertia=ev, vert=(v - com) @ R, face=f)
        return md

    # ---- bodies
    B = dict(name=["world"], parent=[0], pos=[np.zeros(3)], quat=[np.array([1.0, 0, 0, 0])], mocap=[False],
             inertial=[None])
    J = dict(name=[], body=[], type=[], pos=[], axis=[], range=[], limited=[], stiffness=[], springref=[], ref=[],
             damping=[], armature=[], margin=[], solref=[], solimp=[])
    G = dict(name=[], body=[], type=[], size=[], pos=[], quat=[], contype=[], conaffinity=[], condim=[], group=[],
             priority=[], friction=[], solmix=[], solref=[], solimp=[], margin=[], gap=[], mass=[], mesh=[])
    S = dict(name=[], body=[], pos=[], quat=[])

    def add_geom(e, bid, childclass):
        a = resolve("geom", e, childclass, _GEOM_DEF)
        gtype = _GEOM_TYPES[a["type"]]
        if "mesh" in a and "type" not in e.attrib and classes[e.attrib.get("class", childclass or "main")].get("geom").get("type") is None:
            gtype = GEOM_MESH
        size = _vec(a["size"], 3, [0, 0, 0])
        if "fromto" in a:
            raise NotImplementedError("geom fromto")
        pos = _vec(a["pos"])
        quat = _orientation(a)
        mid = -1
        if gtype == GEOM_MESH:
            mid = mesh_names.index(a["mesh"])
            md = get_mesh(mid)
            # geom frame is re-centred on the mesh inertial frame
            pos = pos + quat2mat(quat) @ md["pos"]
            quat = quat_norm(quat_mul(quat, md["quat"]))
            vol, inertia = md["volume"], md["inertia"]
        else:
            vol, inertia = _geom_volume_inertia(gtype, size)
        if "mass" in a:
            mass = float(a["mass"])
            inertia = inertia * (mass / vol) if vol > 0 else inertia * 0
        else:
            dens = float(a["density"])
            mass = dens * vol
            inertia = inertia * dens
        G["name"].append(a.get("name"))
        G["body"].append(bid)
        G["type"].append(gtype)
        G["size"].append(size)
        G["pos"].append(pos)
        G["quat"].append(quat)
        G["contype"].append(int(a["contype"]))
        G["conaffinity"].append(int(a["conaffinity"]))
        G["condim"].append(int(a["condim"]))
        G["group"].append(int(a["group"]))
        G["priority"].append(int(a["priority"]))
        G["friction"].append(_vec(a["friction"], 3, [1, 0.005, 0.0001]))
        G["solmix"].append(float(a["solmix"]))
        G["solref"].append(_vec(a["solref"], 2, [0.02, 1]))
        G["solimp"].append(_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2]))
        G["margin"].append(float(a["margin"]))
        G["gap"].append(float(a["gap"]))
        G["mass"].append((mass, inertia))
        G["mesh"].append(mid)

    def add_body(e, parent, childclass):
        bid = len(B["name"])
        childclass = e.attrib.get("childclass", childclass)
        B["name"].append(e.attrib.get("name", f"body{bid}"))
        B["parent"].append(parent)
        B["pos"].append(_vec(e.attrib.get("pos"), 3, [0, 0, 0]))
        B["quat"].append(_orientation(e.attrib))
        B["mocap"].append(e.attrib.get("mocap", "false") == "true")
        B["inertial"].append(None)
        for ch in e:
            if ch.tag == "inertial":
                if "fullinertia" in ch.attrib:
                    raise NotImplementedError("fullinertia")
                B["inertial"][bid] = dict(pos=_vec(ch.attrib.get("pos"), 3, [0, 0, 0]), quat=_orientation(ch.attrib),
                                          mass=float(ch.attrib["mass"]),
                                          inertia=_vec(ch.attrib.get("diaginertia"), 3, [0, 0, 0]))
            elif ch.tag in ("joint", "freejoint"):
                a = resolve("joint", ch, childclass, _JNT_DEF)
                if ch.tag == "freejoint":
                    a["type"] = "free"
                jt = _JNT_TYPES[a["type"]]
                rng = _vec(a.get("range"), 2, [0, 0])
                lim = a.get("limited", "auto")
                limited = (lim == "true") or (lim == "auto" and "range" in a and rng[0] < rng[1])
                if jt == JNT_FREE:
                    limited = False
                axis = _vec(a["axis"])
                axis = axis / max(np.linalg.norm(axis), MINVAL)
                J["name"].append(a.get("name"))
                J["body"].append(bid)
                J["type"].append(jt)
                J["pos"].append(_vec(a["pos"]))
                J["axis"].append(axis)
                J["range"].append(rng)
                J["limited"].append(limited)
                for k in ("stiffness", "springref", "ref", "damping", "armature", "margin"):
                    J[k].append(float(a[k]))
                J["solref"].append(_vec(a["solreflimit"], 2, [0.02, 1]))
                J["solimp"].append(_vec(a["solimplimit"], 5, [0.9, 0.95, 0.001, 0.5, 2]))
            elif ch.tag == "geom":
                add_geom(ch, bid, childclass)
            elif ch.tag == "site":
                a = resolve("site", ch, childclass, _SITE_DEF)
                S["name"].append(a.get("name"))
                S["body"].append(bid)
                S["pos"].append(_vec(a["pos"]))
                S["quat"].append(_orientation(a))
        for ch in e:
            if ch.tag == "body":
                add_body(ch, bid, childclass)

    for wb in root.findall("worldbody"):
        for ch in wb:
            if ch.tag == "geom":
                add_geom(ch, 0, None)
            elif ch.tag == "site":
                a = resolve("site", ch, None, _SITE_DEF)
                S["name"].append(a.get("name"))
                S["body"].append(0)
                S["pos"].append(_vec(a["pos"]))
                S["quat"].append(_orientation(a))
        # MuJoCo numbers bodies depth-first in document order
    for wb in root.findall("worldbody"):
        for ch in wb:
            if ch.tag == "body":
                add_body(ch, 0, None)

    # MuJoCo orders geoms/sites/joints by owning body id; our depth-first walk
    # appends a body's own elements before descending, but worldbody elements of
    # later <worldbody> blocks must still precede child-body elements -> stable sort.
    def reorder(T):
        order = np.argsort(np.array(T["body"]), kind="stable")
        for k in T:
            T[k] = [T[k][i] for i in order]

    reorder(G)
    reorder(S)
    reorder(J)

    nbody = len(B["name"])
    arr = {}
    arr["body_parentid"] = np.array(B["parent"], dtype=np.int32)
    arr["body_pos"] = np.array(B["pos"])
    arr["body_quat"] = np.array(B["quat"])
    mocapid = -np.ones(nbody, dtype=np.int32)
    nm = 0
    for i, m in enumerate(B["mocap"]):
        if m:
            mocapid[i] = nm
            nm += 1
    arr["body_mocapid"] = mocapid

    # ---- joints / dofs
    njnt = len(J["name"])
    jnt_qposadr, jnt_dofadr = [], []
    qpos0, dof_jnt, dof_body = [], [], []
    for j in range(njnt):
        jnt_qposadr.append(len(qpos0))
        jnt_dofadr.append(len(dof_jnt))
        b = J["body"][j]
        if J["type"][j] == JNT_FREE:
            qpos0.extend(list(B["pos"][b]) + list(B["quat"][b]))
            dof_jnt += [j] * 6
            dof_body += [b] * 6
        elif J["type"][j] == JNT_BALL:
            raise NotImplementedError("ball joint")
        else:
            qpos0.append(J["ref"][j])
            dof_jnt.append(j)
            dof_body.append(b)
    arr["jnt_type"] = np.array(J["type"], dtype=np.int32)
    arr["jnt_bodyid"] = np.array(J["body"], dtype=np.int32)
    arr["jnt_qposadr"] = np.array(jnt_qposadr, dtype=np.int32)
    arr["jnt_dofadr"] = np.array(jnt_dofadr, dtype=np.int32)
    arr["jnt_pos"] = np.array(J["pos"]).reshape(njnt, 3)
    arr["jnt_axis"] = np.array(J["axis"]).reshape(njnt, 3)
    arr["jnt_range"] = np.array(J["range"]).reshape(njnt, 2)
    arr["jnt_limited"] = np.array(J["limited"], dtype=np.int32)
    arr["jnt_stiffness"] = np.array(J["stiffness"])
    arr["jnt_margin"] = np.array(J["margin"])
    arr["jnt_solref"] = np.array(J["solref"]).reshape(njnt, 2)
    arr["jnt_solimp"] = np.array(J["solimp"]).reshape(njnt, 5)
    arr["qpos0"] = np.array(qpos0)
    arr["qpos_spring"] = arr["qpos0"].copy()
    for j in range(njnt):
        if J["type"][j] in (JNT_SLIDE, JNT_HINGE):
            arr["qpos_spring"][jnt_qposadr[j]] = J["springref"][j]
    arr["dof_jntid"] = np.array(dof_jnt, dtype=np.int32)
    arr["dof_bodyid"] = np.array(dof_body, dtype=np.int32)
    arr["dof_damping"] = np.array([J["damping"][j] for j in dof_jnt])
    arr["dof_armature"] = np.array([J["armature"][j] for j in dof_jnt])
    body_jntnum = np.zeros(nbody, dtype=np.int32)
    body_jntadr = -np.ones(nbody, dtype=np.int32)
    body_dofnum = np.zeros(nbody, dtype=np.int32)
    body_dofadr = -np.ones(nbody, dtype=np.int32)
    for j in range(njnt):
        b = J["body"][j]
        if body_jntnum[b] == 0:
            body_jntadr[b] = j
            body_dofadr[b] = jnt_dofadr[j]
        body_jntnum[b] += 1
        body_dofnum[b] += 6 if J["type"][j] == JNT_FREE else 1
    arr.update(body_jntnum=body_jntnum, body_jntadr=body_jntadr, body_dofnum=body_dofnum, body_dofadr=body_dofadr)
    # weld id: nearest ancestor-or-self that owns a joint (0 = static w.r.t. world)
    weld = np.zeros(nbody, dtype=np.int32)
    for b in range(1, nbody):
        weld[b] = b if body_jntnum[b] > 0 else weld[B["parent"][b]]
    arr["body_weldid"] = weld
    # dof parent chain
    dof_parent = -np.ones(len(dof_jnt), dtype=np.int32)
    last_dof_of_body = -np.ones(nbody, dtype=np.int32)
    for b in range(1, nbody):
        p = B["parent"][b]
        last = last_dof_of_body[p]
        for k in range(body_dofnum[b]):
            d = body_dofadr[b] + k
            dof_parent[d] = last
            last = d
        last_dof_of_body[b] = last
    arr["dof_parentid"] = dof_parent

    # ---- geoms
    ngeom = len(G["name"])
    arr["geom_bodyid"] = np.array(G["body"], dtype=np.int32)
    arr["geom_type"] = np.array(G["type"], dtype=np.int32)
    arr["geom_size"] = np.array(G["size"]).reshape(ngeom, 3)
    arr["geom_pos"] = np.array(G["pos"]).reshape(ngeom, 3)
    arr["geom_quat"] = np.array(G["quat"]).reshape(ngeom, 4)
    for k in ("contype", "conaffinity", "condim", "group", "priority"):
        arr["geom_" + k] = np.array(G[k], dtype=np.int32)
    arr["geom_friction"] = np.array(G["friction"]).reshape(ngeom, 3)
    arr["geom_solmix"] = np.array(G["solmix"])
    arr["geom_solref"] = np.array(G["solref"]).reshape(ngeom, 2)
    arr["geom_solimp"] = np.array(G["solimp"]).reshape(ngeom, 5)
    arr["geom_margin"] = np.array(G["margin"])
    arr["geom_gap"] = np.array(G["gap"])
    arr["geom_dataid"] = np.array(G["mesh"], dtype=np.int32)
    # bounding sphere radius (for pair culling)
    rb = np.zeros(ngeom)
    for g in range(ngeom):
        t, s = G["type"][g], G["size"][g]
        if t == GEOM_SPHERE:
            rb[g] = s[0]
        elif t == GEOM_CAPSULE:
            rb[g] = s[0] + s[1]
        elif t == GEOM_CYLINDER:
            rb[g] = np.hypot(s[0], s[1])
        elif t == GEOM_BOX:
            rb[g] = np.linalg.norm(s)
        elif t == GEOM_MESH:
            rb[g] = np.linalg.norm(get_mesh(G["mesh"][g])["vert"], axis=1).max()
    arr["geom_rbound"] = rb

    # ---- sites
    nsite = len(S["name"])
    arr["site_bodyid"] = np.array(S["body"], dtype=np.int32)
    arr["site_pos"] = np.array(S["pos"]).reshape(nsite, 3)
    arr["site_quat"] = np.array(S["quat"]).reshape(nsite, 4)

    # ---- body inertial properties
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    for b in range(1, nbody):
        ine = B["inertial"][b]
        if ine is not None:
            body_mass[b] = ine["mass"]
            body_ipos[b] = ine["pos"]
            body_iquat[b] = ine["quat"]
            body_inertia[b] = ine["inertia"]
            continue
        if comp["inertiafromgeom"] == "false":
            continue
        gs = [g for g in range(ngeom) if G["body"][g] == b and grp_lo <= G["group"][g] <= grp_hi]
        m = sum(G["mass"][g][0] for g in gs)
        if m <= MINVAL:
            continue
        com = sum(G["mass"][g][0] * G["pos"][g] for g in gs) / m
        I = np.zeros((3, 3))
        for g in gs:
            mg, ig = G["mass"][g]
            R = quat2mat(G["quat"][g])
            d = G["pos"][g] - com
            I += R @ np.diag(ig) @ R.T + mg * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        ev, q = eig3(I)
        body_mass[b] = m
        body_ipos[b] = com
        body_iquat[b] = q
        body_inertia[b] = ev
    arr.update(body_mass=body_mass, body_ipos=body_ipos, body_iquat=body_iquat, body_inertia=body_inertia)

    # ---- actuators (position servos on a joint: force = kp*(ctrl - q))
    A = dict(jnt=[], kp=[], range=[])
    for sec in root.findall("actuator"):
        for e in sec:
            if e.tag != "position":
                raise NotImplementedError(f"actuator {e.tag}")
            a = resolve("position", e, None, dict(kp="1", ctrllimited="false", ctrlrange="0 0"))
            # MuJoCo applies the *main* class only unless class= is given
            A["jnt"].append(J["name"].index(a["joint"]))
            A["kp"].append(float(a["kp"]))
            lim = a["ctrllimited"] == "true"
            A["range"].append(_vec(a["ctrlrange"]) if lim else np.array([-np.inf, np.inf]))
    arr["actuator_jntid"] = np.array(A["jnt"], dtype=np.int32)
    arr["actuator_kp"] = np.array(A["kp"])
    arr["actuator_ctrlrange"] = np.array(A["range"]).reshape(len(A["jnt"]), 2)

    # ---- equality (weld)
    E = dict(b1=[], b2=[], data=[], solref=[], solimp=[])
    for sec in root.findall("equality"):
        for e in sec:
            if e.tag != "weld":
                raise NotImplementedError(f"equality {e.tag}")
            a = dict(solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2")
            a.update(classes["main"].get("equality"))
            a.update(e.attrib)
            b1 = B["name"].index(a["body1"])
            b2 = B["name"].index(a["body2"]) if "body2" in a else 0
            E["b1"].append(b1)
            E["b2"].append(b2)
            # data = [anchor(3) in body2, relpose pos(3), relpose quat(4), torquescale]; the
            # env overwrites it immediately (reset_mocap_welds, sawyer_xyz_env.py:133-140)
            E["data"].append(np.array([0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1.0]))
            E["solref"].append(_vec(a["solref"], 2, [0.02, 1]))
            E["solimp"].append(_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2]))
    neq = len(E["b1"])
    arr["eq_obj1id"] = np.array(E["b1"], dtype=np.int32)
    arr["eq_obj2id"] = np.array(E["b2"], dtype=np.int32)
    arr["eq_data"] = np.array(E["data"]).reshape(neq, 11)
    arr["eq_solref"] = np.array(E["solref"]).reshape(neq, 2)
    arr["eq_solimp"] = np.array(E["solimp"]).reshape(neq, 5)

    m = Model(path=path, opt=opt, arrays=arr,
              names=dict(body=B["name"], joint=J["name"], geom=G["name"], site=S["name"], mesh=mesh_names))
    m.meshes = []
    used = set(int(i) for i in arr["geom_dataid"] if i >= 0)
    for mid in range(len(mesh_names)):
        if mid in used and _mesh_collides(arr, mid):
            md = get_mesh(mid)
            hv, hf = convex_hull(md["vert"])
            m.meshes.append(dict(vert=hv, face=hf))
        else:
            m.meshes.append(None)
    set_const(m)
    return m


def _mesh_collides(arr, mid):
    sel = arr["geom_dataid"] == mid
    return bool(np.any((arr["geom_contype"][sel] != 0) | (arr["geom_conaffinity"][sel] != 0)))


# --------------------------------------------------------------------------- qpos0-dependent constants
def kinematics(m: Model, qpos, mocap_pos=None, mocap_quat=None):
    """World poses of all bodies for a configuration (numpy; compile-time use)."""
    a = m.arrays
    nb = m.nbody
    xpos = np.zeros((nb, 3))
    xquat = np.tile(np.array([1.0, 0, 0, 0]), (nb, 1))
    for b in range(1, nb):
        p = a["body_parentid"][b]
        if a["body_mocapid"][b] >= 0 and mocap_pos is not None:
            xpos[b] = mocap_pos
            xquat[b] = quat_norm(mocap_quat)
            continue
        xpos[b] = xpos[p] + quat2mat(xquat[p]) @ a["body_pos"][b]
        xquat[b] = quat_mul(xquat[p], a["body_quat"][b])
        for k in range(a["body_jntnum"][b]):
            j = a["body_jntadr"][b] + k
            qa = a["jnt_qposadr"][j]
            t = a["jnt_type"][j]
            if t == JNT_FREE:
                xpos[b] = qpos[qa:qa + 3]
                xquat[b] = quat_norm(qpos[qa + 3:qa + 7])
            elif t == JNT_SLIDE:
                xpos[b] = xpos[b] + quat2mat(xquat[b]) @ a["jnt_axis"][j] * (qpos[qa] - a["qpos0"][qa])
            elif t == JNT_HINGE:
                R = quat2mat(xquat[b])
                anchor = xpos[b] + R @ a["jnt_pos"][j]
                ang = qpos[qa] - a["qpos0"][qa]
                ax = a["jnt_axis"][j]
                qr = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
                xquat[b] = quat_norm(quat_mul(xquat[b], qr))
                xpos[b] = anchor - quat2mat(xquat[b]) @ a["jnt_pos"][j]
    return xpos, xquat


def dense_jacobians(m: Model, xpos, xquat):
    """Per-body 6 x nv Jacobian at the body inertial-frame origin (rows: lin, ang)."""
    a = m.arrays
    nb, nv = m.nbody, m.nv
    xipos = np.array([xpos[b] + quat2mat(xquat[b]) @ a["body_ipos"][b] for b in range(nb)])
    Jp = np.zeros((nb, 3, nv))
    Jr = np.zeros((nb, 3, nv))
    for b in range(1, nb):
        c = b
        while c > 0:
            for k in range(a["body_jntnum"][c]):
                j = a["body_jntadr"][c] + k
                d = a["jnt_dofadr"][j]
                t = a["jnt_type"][j]
                R = quat2mat(xquat[c])
                if t == JNT_FREE:
                    for i in range(3):
                        Jp[b, i, d + i] = 1.0
                    for i in range(3):
                        ax = R[:, i]
                        Jr[b, :, d + 3 + i] = ax
                        Jp[b, :, d + 3 + i] = np.cross(ax, xipos[b] - xpos[c])
                elif t == JNT_SLIDE:
                    Jp[b, :, d] = R @ a["jnt_axis"][j]
                elif t == JNT_HINGE:
                    ax = R @ a["jnt_axis"][j]
                    anchor = xpos[c] + R @ a["jnt_pos"][j]
                    Jr[b, :, d] = ax
                    Jp[b, :, d] = np.cross(ax, xipos[b] - anchor)
            c = a["body_parentid"][c]
    return xipos, Jp, Jr


def mass_matrix(m: Model, xpos, xquat):
    a = m.arrays
    xipos, Jp, Jr = dense_jacobians(m, xpos, xquat)
    nv = m.nv
    M = np.zeros((nv, nv))
    for b in range(1, m.nbody):
        if a["body_mass"][b] <= 0 and not np.any(a["body_inertia"][b] > 0):
            continue
        R = quat2mat(quat_mul(xquat[b], a["body_iquat"][b]))
        I = R @ np.diag(a["body_inertia"][b]) @ R.T
        M += a["body_mass"][b] * Jp[b].T @ Jp[b] + Jr[b].T @ I @ Jr[b]
    M += np.diag(a["dof_armature"])
    return M, Jp, Jr


def set_const(m: Model):
    """body_invweight0 / dof_invweight0 / meaninertia at qpos0 (what MuJoCo's
    mj_setConst derives at compile time [3P]); they feed the constraint
    regulariser (efc_diagApprox) and the solver's tolerance scale."""
    a = m.arrays
    nb, nv = m.nbody, m.nv
    mp = mq = None
    for b in range(nb):
        if a["body_mocapid"][b] >= 0:
            mp, mq = a["body_pos"][b], a["body_quat"][b]
    xpos, xquat = kinematics(m, a["qpos0"], mp, mq)
    M, Jp, Jr = mass_matrix(m, xpos, xquat)
    Minv = np.linalg.inv(M) if nv else np.zeros((0, 0))
    inv0 = np.zeros((nb, 2))
    for b in range(1, nb):
        if a["body_weldid"][b] == 0:
            continue
        At = Jp[b] @ Minv @ Jp[b].T
        Ar = Jr[b] @ Minv @ Jr[b].T
        inv0[b, 0] = max(MINVAL, np.trace(At) / 3)
        inv0[b, 1] = max(MINVAL, np.trace(Ar) / 3)
    dinv = np.zeros(nv)
    for j in range(len(a["jnt_type"])):
        d = a["jnt_dofadr"][j]
        if a["jnt_type"][j] == JNT_FREE:
            dinv[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            dinv[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        else:
            dinv[d] = Minv[d, d]
    a["body_invweight0"] = inv0
    a["dof_invweight0"] = dinv
    a["dof_M0"] = np.diag(M).copy()
    m.opt["meaninertia"] = float(np.mean(np.diag(M))) if nv else 1.0


def asset_dir():
    """Directory holding Meta-World's ``assets`` tree (MJCF + STL).  The compiled
    tables are cached in ``metaworld_b200/models/*.npz`` so this is only needed
    when (re)building that cache."""
    for p in (os.environ.get("METAWORLD_ASSETS"), "/root/reference/metaworld/assets"):
        if p and os.path.isdir(p):
            return p
    try:
        import metaworld  # type: ignore

        return os.path.join(os.path.dirname(metaworld.__file__), "assets")
    except Exception:
        raise FileNotFoundError("Meta-World assets not found; set METAWORLD_ASSETS")
