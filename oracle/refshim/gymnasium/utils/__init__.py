"""gymnasium.utils stand-in (TEST INFRASTRUCTURE; see oracle/refshim/README.md)."""
from . import ezpickle, seeding
from .ezpickle import EzPickle


class RecordConstructorArgs:
    def __init__(self, *, _disable_deepcopy=False, **kwargs):
        if not hasattr(self, "_saved_kwargs"):
            self._saved_kwargs = kwargs
