"""Stand-in for the `mujoco` Python bindings (TEST INFRASTRUCTURE; see oracle/refshim/README.md).

Only the surface the reference touches (SURVEY.md 8c "Reference call sites into MuJoCo"):
MjModel.from_xml_path, MjData, mj_step / mj_forward / mj_resetData / mj_name2id / mj_rnePostConstraint,
named accessors, data.contact, data.efc_force, mjtObj, mjtEq.  Backed by oracle/mjphys (float64 C restatement).
"""
from __future__ import annotations

import enum
import os
import types

import numpy as np

from oracle import mjphys as _P

__version__ = "3.3.0+refshim"


class mjtObj(enum.IntEnum):
    mjOBJ_UNKNOWN = 0
    mjOBJ_BODY = 1
    mjOBJ_XBODY = 2
    mjOBJ_JOINT = 3
    mjOBJ_DOF = 4
    mjOBJ_GEOM = 5
    mjOBJ_SITE = 6


class mjtEq(enum.IntEnum):
    mjEQ_CONNECT = 0
    mjEQ_WELD = 1
    mjEQ_JOINT = 2


_KIND = {mjtObj.mjOBJ_BODY: "body", mjtObj.mjOBJ_JOINT: "joint", mjtObj.mjOBJ_GEOM: "geom", mjtObj.mjOBJ_SITE: "site"}
_COMPILED: dict = {}


class _JointView:
    def __init__(self, model, j):
        a = model.src.arrays
        self.id = j
        self.qposadr = np.array([int(a["jnt_qposadr"][j])])
        self.dofadr = np.array([int(a["jnt_dofadr"][j])])


class MjModel(_P.OModel):
    @classmethod
    def from_xml_path(cls, path):
        from metaworld_b200 import mjcf

        path = os.path.abspath(str(path))
        if path not in _COMPILED:
            _COMPILED[path] = mjcf.load(path)          # compiled from the reference's own asset tree
        m = cls(_COMPILED[path])
        a = m.src.arrays
        m.nu = len(a["actuator_jntid"])
        m.na = 0
        m.njnt = len(a["jnt_type"])
        m.ngeom = len(a["geom_type"])
        m.nsite = len(a["site_bodyid"])
        m.opt = types.SimpleNamespace(timestep=m.timestep)
        m.actuator_ctrlrange = np.asarray(a["actuator_ctrlrange"], dtype=np.float64).reshape(-1, 2)
        return m

    def joint(self, name):
        return _JointView(self, self.names["joint"].index(name))


class MjData(_P.OData):
    def __init__(self, model):
        super().__init__(model)
        object.__setattr__(self, "time", 0.0)

    def __setattr__(self, k, v):
        if k == "time":
            object.__setattr__(self, k, v)
        else:
            super().__setattr__(k, v)


def mj_step(model, data, nstep=1):
    _P.mj_step(model, data, nstep)
    object.__setattr__(data, "time", data.time + nstep * model.timestep)


def mj_forward(model, data):
    _P.mj_forward(model, data)


def mj_resetData(model, data):
    _P.mj_resetData(model, data)
    object.__setattr__(data, "time", 0.0)


def mj_rnePostConstraint(model, data):   # cacc / cfrc_ext only; nothing on the path reads them
    return None


def mj_name2id(model, objtype, name):
    try:
        return model.names[_KIND[mjtObj(objtype)]].index(name)
    except ValueError:
        return -1
