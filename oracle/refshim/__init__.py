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
    import metaworld  # noqa: F401  (the reference package, unmodified)

    return metaworld
