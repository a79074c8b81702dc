"""Stand-in for pykdtree (reference splatter.py:18): see kdtree.py."""
