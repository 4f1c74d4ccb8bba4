"""ctypes binding of the CPU oracle physics (oracle/libmjphys.so).

TEST INFRASTRUCTURE -- PARITY UNPINNED (see oracle/mjphys.h).  Gives the Python
restatement of ``SawyerXYZEnv`` (oracle/sawyer_env.py) the same verbs the
reference uses on ``mujoco.MjModel`` / ``MjData``: ``mj_step``, ``mj_forward``,
``mj_resetData`` and named accessors ``data.body(name).xpos`` etc.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _OContact(C.Structure):
    _fields_ = [("dist", C.c_double), ("pos", C.c_double * 3), ("frame", C.c_double * 9),
                ("includemargin", C.c_double), ("friction", C.c_double * 5), ("solref", C.c_double * 2),
                ("solimp", C.c_double * 5), ("mu", C.c_double), ("dim", C.c_int), ("geom1", C.c_int),
                ("geom2", C.c_int), ("efc_address", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "libmjphys.so")
    srcs = [os.path.join(_HERE, f) for f in ("mjphys.c", "mjcollide.c", "mjphys.h", "mjinternal.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.om_model_new.restype = C.c_void_p
        L.om_model_free.argtypes = [C.c_void_p]
        L.om_model_set_f64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.om_model_set_i32.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.om_model_add_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.om_model_finalize.argtypes = [C.c_void_p]
        L.om_model_f64.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.om_model_f64.restype = C.POINTER(C.c_double)
        L.om_data_new.argtypes = [C.c_void_p]
        L.om_data_new.restype = C.c_void_p
        L.om_data_free.argtypes = [C.c_void_p]
        L.om_reset_data.argtypes = [C.c_void_p, C.c_void_p]
        L.om_forward.argtypes = [C.c_void_p, C.c_void_p]
        L.om_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.om_data_f64.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.om_data_f64.restype = C.POINTER(C.c_double)
        L.om_data_ncon.argtypes = [C.c_void_p]
        L.om_data_nefc.argtypes = [C.c_void_p]
        L.om_data_solver_iter.argtypes = [C.c_void_p]
        L.om_data_flops.argtypes = [C.c_void_p]
        L.om_data_flops.restype = C.c_long
        L.om_data_contacts.argtypes = [C.c_void_p]
        L.om_data_contacts.restype = C.POINTER(_OContact)
        _LIB = L
    return _LIB


_F64_FIELDS = ["body_pos", "body_quat", "body_mass", "body_ipos", "body_iquat", "body_inertia", "body_invweight0",
               "jnt_pos", "jnt_axis", "jnt_range", "jnt_stiffness", "jnt_margin", "jnt_solref", "jnt_solimp", "qpos0",
               "qpos_spring", "dof_damping", "dof_armature", "dof_invweight0", "geom_size", "geom_pos", "geom_quat",
               "geom_friction", "geom_solmix", "geom_solref", "geom_solimp", "geom_margin", "geom_gap", "geom_rbound",
               "site_pos", "site_quat", "actuator_kp", "actuator_ctrlrange", "eq_data", "eq_solref", "eq_solimp"]
_I32_FIELDS = ["body_parentid", "body_mocapid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum",
               "body_dofadr", "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "dof_jntid",
               "dof_bodyid", "dof_parentid", "geom_bodyid", "geom_type", "geom_contype", "geom_conaffinity",
               "geom_condim", "geom_priority", "geom_dataid", "site_bodyid", "actuator_jntid", "eq_obj1id", "eq_obj2id"]


class _View:
    """``model.body("x").pos`` style accessor over live C memory."""

    def __init__(self, owner, kind, idx, fields):
        object.__setattr__(self, "_o", (owner, kind, idx, fields))

    @property
    def id(self):
        return self._o[2]

    def __getattr__(self, k):
        owner, kind, idx, fields = self._o
        arr, width = fields[k]
        a = owner._arr(arr)
        return a.reshape(-1, width)[idx] if width > 1 else a[idx:idx + 1]

    def __setattr__(self, k, v):
        owner, kind, idx, fields = self._o
        arr, width = fields[k]
        a = owner._arr(arr)
        if width > 1:
            a.reshape(-1, width)[idx] = np.asarray(v, dtype=np.float64)
        else:
            a[idx] = float(np.asarray(v).reshape(-1)[0])


class OModel:
    def __init__(self, m):
        """m: metaworld_b200.mjcf.Model"""
        L = lib()
        self.src = m
        self.names = m.names
        self.ptr = L.om_model_new()
        a = m.arrays
        opt = np.array([m.opt["timestep"], m.opt["tolerance"], m.opt["impratio"], m.opt["meaninertia"],
                        *m.opt["gravity"], m.opt["iterations"]], dtype=np.float64)
        L.om_model_set_f64(self.ptr, b"opt", opt.ctypes.data, len(opt))
        for k in _F64_FIELDS:
            v = np.ascontiguousarray(a[k], dtype=np.float64).reshape(-1)
            L.om_model_set_f64(self.ptr, k.encode(), v.ctypes.data, len(v))
        for k in _I32_FIELDS:
            v = np.ascontiguousarray(a[k], dtype=np.int32).reshape(-1)
            L.om_model_set_i32(self.ptr, k.encode(), v.ctypes.data, len(v))
        for me in m.meshes:
            if me is None:
                L.om_model_add_mesh(self.ptr, None, 0)
            else:
                v = np.ascontiguousarray(me["vert"], dtype=np.float64)
                L.om_model_add_mesh(self.ptr, v.ctypes.data, len(v))
        if L.om_model_finalize(self.ptr) != 0:
            raise RuntimeError("om_model_finalize failed")
        self.nq, self.nv, self.nbody = m.nq, m.nv, m.nbody
        self.nmocap = int((a["body_mocapid"] >= 0).sum())
        self.body_mocapid = a["body_mocapid"].copy()
        self.eq_type = np.ones(len(a["eq_obj1id"]), dtype=np.int32)  # all welds (mjEQ_WELD == 1)
        self.timestep = m.opt["timestep"]

    def _arr(self, name):
        n = C.c_int()
        p = lib().om_model_f64(self.ptr, name.encode(), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,))

    @property
    def eq_data(self):
        return self._arr("eq_data").reshape(-1, 11)

    @property
    def body_pos(self):
        return self._arr("body_pos").reshape(-1, 3)

    def body(self, name):
        i = name if isinstance(name, int) else self.names["body"].index(name)
        return _View(self, "body", i, dict(pos=("body_pos", 3), quat=("body_quat", 4)))

    def site(self, name):
        i = self.names["site"].index(name)
        return _View(self, "site", i, dict(pos=("site_pos", 3)))

    def geom(self, name):
        i = self.names["geom"].index(name)
        return _View(self, "geom", i, dict(pos=("geom_pos", 3), size=("geom_size", 3)))

    def joint(self, name):
        return self.names["joint"].index(name)

    def __del__(self):
        try:
            lib().om_model_free(self.ptr)
        except Exception:
            pass


class OData:
    def __init__(self, model: OModel):
        self.model = model
        self.ptr = lib().om_data_new(model.ptr)
        self._cache = {}

    def _arr(self, name):
        if name not in self._cache:
            n = C.c_int()
            p = lib().om_data_f64(self.ptr, name.encode(), C.byref(n))
            if not p:
                raise KeyError(name)
            self._cache[name] = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[: n.value]
        return self._cache[name]

    def __getattr__(self, k):
        if k.startswith("_") or k in ("model", "ptr"):
            raise AttributeError(k)
        a = self._arr(k)
        if k in ("mocap_pos",):
            return a.reshape(-1, 3)
        if k in ("mocap_quat",):
            return a.reshape(-1, 4)
        return a

    def __setattr__(self, k, v):
        if k in ("model", "ptr", "_cache"):
            object.__setattr__(self, k, v)
        else:
            a = self._arr(k)
            a[:] = np.asarray(v, dtype=np.float64).reshape(-1)

    def body(self, name):
        i = name if isinstance(name, int) else self.model.names["body"].index(name)
        return _View(self, "body", i, dict(xpos=("xpos", 3), xquat=("xquat", 4), xmat=("xmat", 9)))

    def geom(self, name):
        i = self.model.names["geom"].index(name)
        return _View(self, "geom", i, dict(xpos=("geom_xpos", 3), xmat=("geom_xmat", 9)))

    def site(self, name):
        i = self.model.names["site"].index(name)
        return _View(self, "site", i, dict(xpos=("site_xpos", 3), xmat=("site_xmat", 9)))

    def joint(self, name):
        j = self.model.names["joint"].index(name)
        qa = int(self.model.src.arrays["jnt_qposadr"][j])
        return _View(self, "joint", qa, dict(qpos=("qpos", 1)))

    @property
    def ncon(self):
        return lib().om_data_ncon(self.ptr)

    @property
    def nefc(self):
        return lib().om_data_nefc(self.ptr)

    @property
    def flops(self):
        """Floating-point operations of the physics passes run on this data so far (counted in mjphys.c)."""
        return int(lib().om_data_flops(self.ptr))

    @property
    def solver_iter(self):
        return lib().om_data_solver_iter(self.ptr)

    @property
    def contact(self):
        p = lib().om_data_contacts(self.ptr)
        return [p[i] for i in range(self.ncon)]

    def __del__(self):
        try:
            lib().om_data_free(self.ptr)
        except Exception:
            pass


def mj_step(model: OModel, data: OData, nstep=1):
    lib().om_step(model.ptr, data.ptr, int(nstep))


def mj_forward(model: OModel, data: OData):
    lib().om_forward(model.ptr, data.ptr)


def mj_resetData(model: OModel, data: OData):
    lib().om_reset_data(model.ptr, data.ptr)
