"""Compiled model cache.

The reference loads MJCF at env construction (``MujocoEnv.__init__`` from
metaworld/sawyer_xyz_env.py:53-63).  Here the MJCF -> table compilation
(`mjcf.load`) is run once per model file and its result is cached as a small
``.npz`` under ``metaworld_b200/models/`` (derived constants: body tree, joint
tables, inertias, convex hull vertices ...), so the engine runs on machines that
do not have the Meta-World asset tree.  ``python -m metaworld_b200.modelzoo``
rebuilds the cache from an asset tree (METAWORLD_ASSETS or an installed
``metaworld`` package).
"""
from __future__ import annotations

import io
import json
import os

import numpy as np

from . import mjcf

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")
_CACHE: dict = {}

# the 36 model files the 50 V3 tasks point at (reference: `model_name` property of each env class)
USED_XML = [
    "sawyer_assembly_peg", "sawyer_basketball", "sawyer_bin_picking", "sawyer_box", "sawyer_button_press",
    "sawyer_button_press_topdown", "sawyer_button_press_topdown_wall", "sawyer_button_press_wall", "sawyer_coffee",
    "sawyer_dial", "sawyer_door_lock", "sawyer_door_pull", "sawyer_drawer", "sawyer_faucet", "sawyer_hammer",
    "sawyer_handle_press", "sawyer_handle_press_sideways", "sawyer_lever_pull", "sawyer_peg_insertion_side",
    "sawyer_peg_unplug_side", "sawyer_pick_out_of_hole", "sawyer_pick_place_v3", "sawyer_pick_place_wall_v3",
    "sawyer_plate_slide", "sawyer_plate_slide_sideway", "sawyer_push_back_v3", "sawyer_push_v3", "sawyer_push_wall_v3",
    "sawyer_reach_v3", "sawyer_reach_wall_v3", "sawyer_shelf_placing", "sawyer_soccer", "sawyer_stick_obj",
    "sawyer_sweep_v3", "sawyer_table_with_hole", "sawyer_window_horizontal",
]


def _save(m: mjcf.Model, path):
    out = {"arr_" + k: v for k, v in m.arrays.items()}
    meta = dict(names=m.names, opt={k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in m.opt.items()},
                nmesh=len(m.meshes), has_mesh=[me is not None for me in m.meshes])
    for i, me in enumerate(m.meshes):
        if me is not None:
            out[f"mesh{i}_vert"] = me["vert"].astype(np.float64)
            out[f"mesh{i}_face"] = me["face"].astype(np.int32)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)


def _load(path) -> mjcf.Model:
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    opt = meta["opt"]
    opt["gravity"] = np.array(opt["gravity"])
    m = mjcf.Model(path=path, opt=opt, names=meta["names"],
                   arrays={k[4:]: z[k] for k in z.files if k.startswith("arr_")})
    m.meshes = [dict(vert=z[f"mesh{i}_vert"], face=z[f"mesh{i}_face"]) if meta["has_mesh"][i] else None
                for i in range(meta["nmesh"])]
    return m


def full_model(xml_name: str) -> mjcf.Model:
    """xml_name like 'sawyer_reach_v3' (no directory, no extension)."""
    xml_name = os.path.splitext(os.path.basename(xml_name))[0]
    if xml_name not in _CACHE:
        path = os.path.join(_DIR, xml_name + ".npz")
        if os.path.exists(path):
            _CACHE[xml_name] = _load(path)
        else:
            _CACHE[xml_name] = mjcf.load(os.path.join(mjcf.asset_dir(), "sawyer_xyz", xml_name + ".xml"))
    return _CACHE[xml_name]


def build_cache(names=None):
    os.makedirs(_DIR, exist_ok=True)
    root = mjcf.asset_dir()
    for n in names or USED_XML:
        m = mjcf.load(os.path.join(root, "sawyer_xyz", n + ".xml"))
        _save(m, os.path.join(_DIR, n + ".npz"))
        print(f"{n}: nq={m.nq} nv={m.nv} nbody={m.nbody} ngeom={m.ngeom}")


if __name__ == "__main__":
    build_cache()
