"""Generates tests/golden/reward_utils.json from the REFERENCE's own metaworld/utils/reward_utils.py
(importable by file path in the build container; needs only numpy) and scipy's Rotation.  Run once here:
    python tests/golden/make_reward_goldens.py
The GPU box has no /root/reference; tests only read the committed JSON."""
import importlib.util, json, os
import numpy as np
from scipy.spatial.transform import Rotation

REF = "/root/reference/metaworld/utils/reward_utils.py"
spec = importlib.util.spec_from_file_location("ref_reward_utils", REF)
ru = importlib.util.module_from_spec(spec); spec.loader.exec_module(ru)
rng = np.random.default_rng(123)
out = {"tolerance": [], "hamacher": [], "rect_prism": [], "quat": []}
for sig in ("long_tail", "gaussian"):
    for _ in range(60):
        lo = float(rng.uniform(0, 0.05)); hi = lo + float(rng.uniform(0, 0.05)); m = float(rng.choice([0.0, rng.uniform(0.01, 0.5)]))
        x = float(rng.uniform(-0.1, 0.6))
        out["tolerance"].append([x, lo, hi, m, sig, float(ru.tolerance(x, bounds=(lo, hi), margin=m, sigmoid=sig))])
for _ in range(60):
    a, b = float(rng.uniform(0, 1)), float(rng.uniform(0, 1))
    out["hamacher"].append([a, b, float(ru.hamacher_product(a, b))])
out["hamacher"].append([0.0, 0.0, float(ru.hamacher_product(0.0, 0.0))])
for _ in range(40):
    z, o, c = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3), rng.uniform(-1.2, 1.2, 3)
    out["rect_prism"].append([c.tolist(), z.tolist(), o.tolist(), float(ru.rect_prism_tolerance(c, z, o))])
for _ in range(60):
    R = Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix()
    out["quat"].append([R.reshape(-1).tolist(), Rotation.from_matrix(R).as_quat().tolist()])
json.dump(out, open(os.path.join(os.path.dirname(__file__), "reward_utils.json"), "w"))
print({k: len(v) for k, v in out.items()})
