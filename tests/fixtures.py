"""Data-driven re-creations of the reference's test fixtures (no PCL needed).

  hollow_cube / cube : point_cloud_odometry/test/test_point_cloud_odometry.cpp:23-97
  plane              : point_cloud_localization/test/test_point_cloud_localization.cpp:26-43
  garage             : multithreaded_gicp/test/*_82_garage.pcd  (tests/golden/garage.npz)
"""
import os
import numpy as np

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cube(nx=10, ny=10, nz=10, sx=0.1, sy=0.1, sz=0.1, hollow=False):
    pts = []
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                if hollow and not (ix == 0 or iy == 0 or ix == nx - 1 or iy == ny - 1):
                    continue
                pts.append((np.float32(ix) * np.float32(sx), np.float32(iy) * np.float32(sy),
                            np.float32(iz) * np.float32(sz)))
    return np.array(pts, dtype=np.float32)


def hollow_cube(**kw):
    return cube(hollow=True, **kw)


def plane(nx=10, ny=10, sx=0.1, sy=0.1):
    """GeneratePlane: z=0, normal (0,0,1). returns xyz (n,3), normals (n,3)."""
    pts = [(np.float32(ix) * np.float32(sx), np.float32(iy) * np.float32(sy), np.float32(0))
           for ix in range(nx) for iy in range(ny)]
    xyz = np.array(pts, dtype=np.float32)
    nrm = np.zeros_like(xyz); nrm[:, 2] = 1.0
    return xyz, nrm


def garage():
    d = np.load(os.path.join(_GOLD, "garage.npz"))
    return d["query"], d["reference"]


def rot_zyx(roll, pitch, yaw):
    cr, sr = np.cos(roll), np.sin(roll); cp, sp = np.cos(pitch), np.sin(pitch); cy, sy = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]); Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def se3(t, rpy):
    T = np.eye(4); T[:3, :3] = rot_zyx(*rpy); T[:3, 3] = t
    return T


def pose_delta(Ta, Tb):
    """(|dt| in m, rotation angle of dR in rad) between two 4x4 poses."""
    Ta = np.asarray(Ta, dtype=np.float64); Tb = np.asarray(Tb, dtype=np.float64)
    dt = np.linalg.norm(Ta[:3, 3] - Tb[:3, 3])
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    c = np.clip((np.trace(dR) - 1) / 2, -1, 1)
    # small angles: use the skew part, acos is ill-conditioned near 1
    s = 0.5 * np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    return dt, float(np.arctan2(s, c))


def random_scene(n, seed, extent=(20.0, 15.0, 3.0)):
    """Points on the walls/floor of a box room plus a few boxes: surface-like cloud for quick tests."""
    rng = np.random.default_rng(seed)
    ex, ey, ez = extent
    pts = []
    per = n // 6
    u = rng.uniform(-1, 1, (per, 2)); pts.append(np.c_[u[:, 0] * ex, u[:, 1] * ey, np.full(per, -ez)])
    u = rng.uniform(-1, 1, (per, 2)); pts.append(np.c_[u[:, 0] * ex, u[:, 1] * ey, np.full(per, ez)])
    u = rng.uniform(-1, 1, (per, 2)); pts.append(np.c_[np.full(per, -ex), u[:, 0] * ey, u[:, 1] * ez])
    u = rng.uniform(-1, 1, (per, 2)); pts.append(np.c_[np.full(per, ex), u[:, 0] * ey, u[:, 1] * ez])
    u = rng.uniform(-1, 1, (per, 2)); pts.append(np.c_[u[:, 0] * ex, np.full(per, -ey), u[:, 1] * ez])
    rest = n - 5 * per
    u = rng.uniform(-1, 1, (rest, 2)); pts.append(np.c_[u[:, 0] * ex, np.full(rest, ey), u[:, 1] * ez])
    p = np.concatenate(pts) + rng.normal(0, 0.01, (n, 3))
    return p.astype(np.float32)


def c5_case():
    """BASELINE config 5 shape for the parity tests: 200 k-point scan vs 10 M-point map (dense stress).
    returns (src, tgt, T_true, cfg); the oracle's pose for exactly these inputs is tests/golden/c5_oracle_pose.npz"""
    tgt = random_scene(10_000_000, 5, extent=(100.0, 75.0, 15.0))
    rng = np.random.default_rng(11)
    sub = tgt[rng.choice(len(tgt), 200_000, replace=False)] + rng.normal(0, 0.005, (200_000, 3)).astype(np.float32)
    Tg = se3([0.05, 0.03, -0.01], [0.002, 0.003, -0.004])
    src = ((sub.astype(np.float64) - Tg[:3, 3]) @ Tg[:3, :3]).astype(np.float32)
    return src, tgt, Tg, dict(tf_eps=1e-5, corr_dist=0.3, max_iterations=50, max_inner=50)


def cloud_checksum(*clouds):
    """order-sensitive 64-bit checksum of the raw float32 bits (golden fixtures verify their regenerated inputs)"""
    h = np.uint64(1469598103934665603)
    for c in clouds:
        b = np.ascontiguousarray(c, dtype=np.float32).view(np.uint32).astype(np.uint64).ravel()
        w = np.arange(1, b.size + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            h = h * np.uint64(1099511628211) + (b * w).sum(dtype=np.uint64)
    return int(h)
