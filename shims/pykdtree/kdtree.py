"""`pykdtree.kdtree.KDTree` on top of scipy's cKDTree (reference splatter.py:18,382-385: mean distance
to the 3 nearest neighbours initialises the scales)."""
import numpy as np
from scipy.spatial import cKDTree


class KDTree:
    def __init__(self, data_pts, leafsize=16):
        self._tree = cKDTree(np.asarray(data_pts), leafsize=leafsize)

    def query(self, query_pts, k=1, eps=0.0, distance_upper_bound=None, sqr_dists=False, mask=None):
        kw = {} if distance_upper_bound is None else {"distance_upper_bound": distance_upper_bound}
        d, i = self._tree.query(np.asarray(query_pts), k=k, eps=eps, **kw)
        if sqr_dists:
            d = d * d
        return d.astype(np.asarray(query_pts).dtype, copy=False), i.astype(np.uint32, copy=False)
