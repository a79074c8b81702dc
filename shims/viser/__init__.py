"""Import-only stub: the viser web viewer (reference visergui.py) is outside this framework's scope
(SURVEY.md §2 OUT OF SCOPE).  train.py imports visergui unconditionally (train.py:14) but only builds a
viewer under --gui 1."""
from . import transforms  # noqa: F401


class ViserServer:
    def __init__(self, *a, **k):
        raise NotImplementedError("viser is not available in this environment (GUI is out of scope)")
