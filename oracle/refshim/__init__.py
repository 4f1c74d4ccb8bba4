"""See oracle/refshim/README.md.  `activate()` puts the shim packages and the reference tree on sys.path."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def activate(reference_root="/root/reference"):
    if not os.path.isdir(os.path.join(reference_root, "metaworld")):
        raise RuntimeError(f"reference tree not found at {reference_root} (the shim only works where it exists)")
    for p in (reference_root, _HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    stub = sys.modules.get("metaworld")
    if stub is not None and not str(getattr(stub, "__file__", "")).startswith(reference_root):
        for k in [k for k in sys.modules if k == "metaworld" or k.startswith("metaworld.")]:
            del sys.modules[k]          # e.g. the bare namespace stub tests/test_oracle.py uses to load the scripted policies
    for name in ("gymnasium", "mujoco"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(_HERE):
            raise RuntimeError(f"a real `{name}` is already imported; the shim is only for images without it")
    import metaworld  # noqa: F401  (the reference package, unmodified)

    return metaworld
