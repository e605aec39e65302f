"""CPU restatement of the rolling-submap semantics -- TEST INFRASTRUCTURE ONLY.

What the reference does with its external mapper object (locus/src/Locus.cc:464-465,479-486,522-543;
locus/config/lo_settings.yaml:49-62).  The mapper package (point_cloud_mapper) is not vendored in the reference tree,
so this file pins the semantics the product documents in include/locus_b200.h, in plain sequential Python / numpy:

  insert(points)   for each point IN INPUT ORDER: skip it if it is not finite or if its voxel -- floor(p / resolution)
                   per axis, float32 division, world-anchored -- already holds a map point; otherwise append it.
  crop(c, half)    keep the points with  c - half <= p <= c + half  on every axis (pcl::CropBox, inclusive), order kept.
  neighbors(q)     exact nearest map point of every query (float32 squared distance ((dx*dx)+(dy*dy))+(dz*dz), ties ->
                   lowest map index), through oracle.KdTree.
  covariances      the k-NN covariance (gicp.hpp:85-154) of a map point is computed by the first registration after its
                   insertion, from the map as it is at that time, and cached until the point leaves the window.
"""
import numpy as np

from . import oracle as O


class SubmapOracle:
    def __init__(self, resolution):
        self.res = np.float32(resolution)
        self.pts = np.zeros((0, 3), dtype=np.float32)
        self.occ = set()
        self.cov = np.zeros((0, 3, 3))       # cached covariances of the first len(self.cov) points
        self.cov_key = None

    def _voxels(self, p):
        return np.floor(p.astype(np.float32) / self.res).astype(np.int64)

    def insert(self, points):
        p = np.ascontiguousarray(points, dtype=np.float32)[:, :3]
        vox = self._voxels(np.where(np.isfinite(p), p, 0))
        added = []
        for i in range(len(p)):
            if not np.isfinite(p[i]).all():
                continue
            key = (int(vox[i, 0]), int(vox[i, 1]), int(vox[i, 2]))
            if key in self.occ:
                continue
            self.occ.add(key)
            added.append(i)
        if added:
            self.pts = np.concatenate([self.pts, p[added]])
        return p[added]

    def crop(self, center, half):
        c = np.asarray(center, dtype=np.float32); h = np.float32(half)
        mn, mx = (c - h).astype(np.float32), (c + h).astype(np.float32)
        keep = ~(((self.pts < mn).any(axis=1)) | ((self.pts > mx).any(axis=1)))
        n_cov = len(self.cov)
        self.cov = self.cov[keep[:n_cov]]
        removed = int((~keep).sum())
        self.pts = np.ascontiguousarray(self.pts[keep])
        v = self._voxels(self.pts)
        self.occ = {(int(a), int(b), int(c_)) for a, b, c_ in v}
        return removed

    def neighbors(self, query, num_threads=8):
        idx, d2 = O.KdTree(self.pts).nn_batch(np.ascontiguousarray(query, dtype=np.float32)[:, :3], num_threads=num_threads)
        return self.pts[idx], idx, d2

    def covariances(self, k=20, eps=1e-3, num_threads=8):
        """cached covariances, computing those of the points inserted since the last call from the map as it is NOW"""
        if self.cov_key != (k, eps):
            self.cov = np.zeros((0, 3, 3)); self.cov_key = (k, eps)
        if len(self.cov) < len(self.pts):
            allc = O.covariances(self.pts, k, eps, num_threads)
            self.cov = np.concatenate([self.cov, allc[len(self.cov):]])
        return self.cov
