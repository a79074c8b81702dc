"""Import-only stub of omegaconf (visergui.py:7)."""


class OmegaConf:
    @staticmethod
    def create(*a, **k):
        raise NotImplementedError("omegaconf is not available in this environment (GUI is out of scope)")
