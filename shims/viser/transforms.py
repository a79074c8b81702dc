"""Import-only stub of viser.transforms (visergui.py:6)."""
