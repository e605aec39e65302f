"""Synthetic 64-beam lidar stream for the BASELINE.json configs.

Analytic scene: axis-aligned room 40 x 30 x 6 m + 12 boxes + 8 vertical
cylinders (positions from np.random.default_rng(seed)); HDL-64E-like sensor:
`beams` elevations linearly spaced in [-24.8 deg, +2.0 deg], `az` azimuth steps
-> beams*az rays per scan (64*2048 = 131072).  Range noise N(0, 0.02 m), 1 %
drop-outs -> NaN (exercises the finite check of the voxel filter).

Output layout = pcl::PointXYZI as it leaves the BodyFilter nodelet
(point_cloud_filter/src/body_filter.cc:36-39): point_step 32 bytes,
x@0 y@4 z@8 (float32), intensity@16 (float32), the rest padding.
"""
import numpy as np

POINT_STEP = 32
X_OFF, Y_OFF, Z_OFF, I_OFF = 0, 4, 8, 16
FLOAT_FIELDS = (0, 4, 8, 16)

ROOM_MIN = np.array([-20.0, -15.0, -1.5])
ROOM_MAX = np.array([20.0, 15.0, 4.5])


def make_scene(seed):
    rng = np.random.default_rng(seed)
    boxes = []
    for _ in range(12):
        c = np.array([rng.uniform(-17, 17), rng.uniform(-12, 12), 0.0])
        h = np.array([rng.uniform(0.4, 2.0), rng.uniform(0.4, 2.0), rng.uniform(0.5, 2.5)])
        lo = np.array([c[0] - h[0], c[1] - h[1], ROOM_MIN[2]])
        hi = np.array([c[0] + h[0], c[1] + h[1], ROOM_MIN[2] + 2 * h[2]])
        boxes.append((lo, hi))
    cyls = []
    for _ in range(8):
        cyls.append((rng.uniform(-17, 17), rng.uniform(-12, 12), rng.uniform(0.2, 0.8)))
    return {"boxes": boxes, "cyls": cyls, "seed": seed}


def _rot(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def pose_matrix(t, rpy):
    T = np.eye(4)
    T[:3, :3] = _rot(rpy)
    T[:3, 3] = t
    return T


def trajectory(seed, n, t_step=0.3, r_step_deg=2.0):
    """World poses of the sensor: random walk, ego-motion per frame t~U(-t_step,t_step) per
    axis, rpy~U(-r_step,r_step) (inside the 1.0 m / 1.0 rad gates of the callers)."""
    rng = np.random.default_rng(seed + 1000)
    T = np.eye(4)
    poses = [T.copy()]
    for _ in range(n - 1):
        for _try in range(100):
            dt = rng.uniform(-t_step, t_step, 3)
            dt[2] *= 0.2
            dr = np.deg2rad(rng.uniform(-r_step_deg, r_step_deg, 3))
            Tn = T @ pose_matrix(dt, dr)
            p = Tn[:3, 3]
            # stay well inside the room and near-level
            if (abs(p[0]) < 8 and abs(p[1]) < 6 and abs(p[2]) < 0.8 and abs(Tn[2, 2]) > 0.95):
                break
        T = Tn
        poses.append(T.copy())
    return poses


def _ray_dirs(beams, az):
    el = np.deg2rad(np.linspace(-24.8, 2.0, beams))
    a = np.linspace(0, 2 * np.pi, az, endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(a)[None, :], ce * np.sin(a)[None, :], se * np.ones_like(a)[None, :]], axis=-1)
    return d.reshape(-1, 3)


def _cast(scene, o, d):
    """o: (3,), d: (n,3) unit. returns range t (n,), inf if nothing hit."""
    n = d.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # room interior: exit distance
        t1 = (ROOM_MIN - o) * inv
        t2 = (ROOM_MAX - o) * inv
        texit = np.min(np.maximum(t1, t2), axis=1)
        t = texit.copy()
        for lo, hi in scene["boxes"]:
            a = (lo - o) * inv
            b = (hi - o) * inv
            tn = np.max(np.minimum(a, b), axis=1)
            tf = np.min(np.maximum(a, b), axis=1)
            hit = (tn <= tf) & (tn > 0.05)
            t = np.where(hit & (tn < t), tn, t)
        for cx, cy, r in scene["cyls"]:
            ox, oy = o[0] - cx, o[1] - cy
            A = d[:, 0] ** 2 + d[:, 1] ** 2
            B = 2 * (ox * d[:, 0] + oy * d[:, 1])
            Cc = ox * ox + oy * oy - r * r
            disc = B * B - 4 * A * Cc
            ok = (disc > 0) & (A > 1e-12)
            sq = np.sqrt(np.where(ok, disc, 0))
            tc = np.where(ok, (-B - sq) / (2 * A), np.inf)
            hit = ok & (tc > 0.05) & (tc < t)
            t = np.where(hit, tc, t)
    return t


def scan(scene, pose, seed, beams=64, az=2048, noise=0.02, dropout=0.01):
    """One scan in the SENSOR frame.  returns uint8 blob (n*32,), n = beams*az."""
    rng = np.random.default_rng(seed)
    d_s = _ray_dirs(beams, az)
    R = pose[:3, :3]
    o = pose[:3, 3]
    d_w = d_s @ R.T
    t = _cast(scene, o, d_w)
    t = t + rng.normal(0, noise, t.shape)
    drop = rng.random(t.shape) < dropout
    t = np.where(drop | ~np.isfinite(t) | (t > 120.0), np.nan, t)
    p = (d_s * t[:, None]).astype(np.float32)
    inten = (100.0 / (1.0 + np.nan_to_num(t))).astype(np.float32)
    n = p.shape[0]
    blob = np.zeros((n, POINT_STEP // 4), dtype=np.float32)
    blob[:, 0:3] = p
    blob[:, 3] = 1.0
    blob[:, 4] = inten
    return blob.view(np.uint8).reshape(-1)


def blob_xyz(blob, point_step=POINT_STEP):
    a = np.ascontiguousarray(blob).view(np.uint8).reshape(-1, point_step)
    return a[:, :12].copy().view(np.float32).reshape(-1, 3)


def stream(seed, n_scans, beams=64, az=2048):
    """(scene, poses, [blob...]) for a stream of n_scans."""
    scene = make_scene(seed)
    poses = trajectory(seed, n_scans)
    blobs = [scan(scene, poses[i], seed * 7919 + i, beams, az) for i in range(n_scans)]
    return scene, poses, blobs


def world_points(scene, poses, seed, beams=64, az=2048):
    """Union of posed scans in the world frame (finite points only), float32 (n,3)."""
    out = []
    for i, T in enumerate(poses):
        b = scan(scene, T, seed * 104729 + i, beams, az)
        p = blob_xyz(b)
        p = p[np.isfinite(p).all(axis=1)]
        out.append((p.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    return np.concatenate(out)


if __name__ == "__main__":
    import argparse
    import time
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--scans", type=int, default=2)
    a = ap.parse_args()
    t0 = time.time()
    scene, poses, blobs = stream(a.seed, a.scans)
    p = blob_xyz(blobs[0])
    fin = np.isfinite(p).all(axis=1)
    print("scans", len(blobs), "pts/scan", p.shape[0], "finite", int(fin.sum()),
          "bbox", p[fin].min(0), p[fin].max(0), "gen s/scan", (time.time() - t0) / a.scans)
