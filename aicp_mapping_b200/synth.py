"""Seeded synthetic lidar-shaped workloads for the parity tests and bench.py.

Nothing here is on the product path: these are the inputs that both the CUDA path and the CPU oracle are fed.
All generators are deterministic functions of (config, trial) through numpy.random.default_rng(1000*config + trial).

Configurations (BASELINE.json "configs"):
  C1  planar 2-D sample scans of aicp_core/data (scan_0{0,1,2}.csv) extruded to 3-D  (fixture tests/golden/c1_scans.npz)
  C2  Velodyne VLP-16 ANYmal-shaped clouds, 7 sweeps accumulated (aicp_ros/launch/aicp.launch:68), 0.08 m voxel
      filter (aicp_core/src/utils/filteringUtils.cpp:10-13), exactly 32 768 points
  C3  Velodyne HDL-64 KITTI-shaped clouds, ground removed, exactly 131 072 points  (the headline workload)
  C4  122 880-point reading against a 10 485 760-point fixed map
  C5  validation sweep pairs: 38 400-point cube cloud (aicp_core/src/tools/create_cube_cloud.cpp:16-80) perturbed as in
      aicp_lcm/examples/registration_main.cpp:331-343
"""
import numpy as np

__all__ = ["cube_cloud", "rigid", "apply_T", "lidar_scan", "street_scene", "room_scene", "make_pair",
           "perturbation_T", "voxel_downsample", "campus_map", "raw_sweep"]


# ----------------------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------------------
def rigid(tx=0.0, ty=0.0, tz=0.0, roll=0.0, pitch=0.0, yaw=0.0):
    """4x4 float64 rigid transform, R = Rz(yaw) Ry(pitch) Rx(roll), angles in radians."""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [tx, ty, tz]
    return T


def apply_T(T, xyz):
    """float64 application; result cast to float32 n x 3."""
    xyz = np.asarray(xyz, dtype=np.float64)[:, :3]
    return (xyz @ np.asarray(T, dtype=np.float64)[:3, :3].T + np.asarray(T, dtype=np.float64)[:3, 3]).astype(np.float32)


def perturbation_T(rng):
    """registration_main.cpp:331-343: x,y ~ N(0, 0.10 m), yaw ~ N(0, 0.10)*10 deg, through parseTransformationDeg
    (cloudIO.cpp:261-302)."""
    v = rng.normal(0.0, 0.10, 3)
    return rigid(tx=v[0], ty=v[1], yaw=np.deg2rad(v[2] * 10.0))


def cube_cloud():
    """create_cube_cloud.cpp:16-80, including its float32 loop counters. n x 3 float32."""
    lo, hi, step = np.float32(-2.0), np.float32(2.0), np.float32(0.05)
    ticks = []
    i = lo
    while i < hi:
        ticks.append(i)
        i = np.float32(i + step)
    t = np.array(ticks, dtype=np.float32)
    a, b = np.meshgrid(t, t, indexing="ij")
    a, b = a.ravel(), b.ravel()
    faces = [np.stack([a, b, np.full_like(a, lo)], 1), np.stack([a, b, np.full_like(a, hi)], 1),
             np.stack([np.full_like(a, lo), a, b], 1), np.stack([np.full_like(a, hi), a, b], 1),
             np.stack([a, np.full_like(a, lo), b], 1), np.stack([a, np.full_like(a, hi), b], 1)]
    return np.concatenate(faces, 0).astype(np.float32)


def voxel_downsample(xyz, leaf):
    """Centroid per occupied voxel (pcl::VoxelGrid semantics, filteringUtils.cpp:10-13); output ordered by voxel key."""
    xyz = np.asarray(xyz, dtype=np.float64)
    key = np.floor(xyz / leaf).astype(np.int64)
    key -= key.min(0)
    dims = key.max(0) + 1
    lin = (key[:, 0] * dims[1] + key[:, 1]) * dims[2] + key[:, 2]
    order = np.argsort(lin, kind="stable")
    lin_s = lin[order]
    starts = np.flatnonzero(np.r_[True, lin_s[1:] != lin_s[:-1]])
    counts = np.diff(np.r_[starts, lin_s.size])
    sums = np.add.reduceat(xyz[order], starts, axis=0)
    return (sums / counts[:, None]).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------
# scenes and ray casting
# ----------------------------------------------------------------------------------------------------------
def street_scene(rng, length=120.0, width=40.0, n_boxes=40):
    """Street canyon: two building fronts along x plus n_boxes boxes (cars, kiosks, poles) clear of the lane |y|<3."""
    boxes = [[-length / 2, -width / 2 - 5.0, 0.0, length / 2, -width / 2, 12.0],
             [-length / 2, width / 2, 0.0, length / 2, width / 2 + 5.0, 12.0]]
    for _ in range(n_boxes):
        cx = rng.uniform(-length / 2 + 2, length / 2 - 2)
        side = -1.0 if rng.random() < 0.5 else 1.0
        cy = side * rng.uniform(3.5, width / 2 - 2)
        sx, sy, sz = rng.uniform(0.4, 5.0), rng.uniform(0.4, 3.0), rng.uniform(0.5, 3.5)
        boxes.append([cx - sx / 2, cy - sy / 2, 0.0, cx + sx / 2, cy + sy / 2, sz])
    return np.array(boxes, dtype=np.float64)


def room_scene(rng, lx=30.0, ly=20.0, lz=4.0, n_boxes=12):
    """Closed room (the sensor is inside box 0, so its faces are hit from within) with n_boxes pieces of furniture."""
    boxes = [[-lx / 2, -ly / 2, 0.0, lx / 2, ly / 2, lz]]
    for _ in range(n_boxes):
        while True:
            cx, cy = rng.uniform(-lx / 2 + 1.5, lx / 2 - 1.5), rng.uniform(-ly / 2 + 1.5, ly / 2 - 1.5)
            if abs(cy) > 2.0 or abs(cx) > 6.0:      # keep the robot's path (along x near y=0) free
                break
        sx, sy, sz = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0), rng.uniform(0.4, 2.5)
        boxes.append([cx - sx / 2, cy - sy / 2, 0.0, cx + sx / 2, cy + sy / 2, sz])
    return np.array(boxes, dtype=np.float64)


def _raycast(origin, dirs, boxes, max_range, ground_z=0.0, chunk=65536):
    """Nearest hit distance per ray against the ground plane and axis-aligned boxes (entered from outside or inside)."""
    R = dirs.shape[0]
    t = np.full(R, np.inf)
    dz = dirs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(dz < -1e-12, (ground_z - origin[2]) / dz, np.inf)
    t = np.minimum(t, np.where(tg > 0, tg, np.inf))
    if len(boxes) > 200:
        return _raycast_culled(origin, dirs, boxes, max_range, t)
    lo, hi = boxes[:, :3][None], boxes[:, 3:][None]
    for s in range(0, R, chunk):
        d = dirs[s:s + chunk]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
        t1 = (lo - origin) * inv[:, None, :]
        t2 = (hi - origin) * inv[:, None, :]
        tn = np.nanmax(np.minimum(t1, t2), axis=2)
        tf = np.nanmin(np.maximum(t1, t2), axis=2)
        hit = tf >= np.maximum(tn, 0.0)
        tb = np.where(hit, np.where(tn > 1e-9, tn, tf), np.inf)
        t[s:s + chunk] = np.minimum(t[s:s + chunk], tb.min(axis=1))
    t[t > max_range] = np.inf
    return t


def _raycast_culled(origin, dirs, boxes, max_range, t):
    """The same slab test for scenes with many boxes: every box is tested only against the rays whose azimuth falls inside
    the azimuth interval its footprint subtends from the sensor (a superset of the rays that can hit it; per ray and box the
    arithmetic is that of _raycast, and a minimum does not depend on the order of its arguments)."""
    phi = np.arctan2(dirs[:, 1], dirs[:, 0])
    order = np.argsort(phi, kind="stable")
    phis = phi[order]
    ds = dirs[order]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / ds
    ts = t[order]
    for b in boxes:
        if b[0] <= origin[0] <= b[3] and b[1] <= origin[1] <= b[4]:
            sl = [slice(0, len(phis))]                                          # the sensor stands inside the footprint
        else:
            cx = np.array([b[0], b[0], b[3], b[3]]) - origin[0]
            cy = np.array([b[1], b[4], b[1], b[4]]) - origin[1]
            a = np.arctan2(cy, cx)
            mid = np.arctan2(0.5 * (b[1] + b[4]) - origin[1], 0.5 * (b[0] + b[3]) - origin[0])
            rel = (a - mid + np.pi) % (2 * np.pi) - np.pi                      # corners relative to the centre direction
            lo_a, hi_a = mid + rel.min() - 1e-6, mid + rel.max() + 1e-6
            if lo_a < -np.pi:
                sl = [slice(np.searchsorted(phis, lo_a + 2 * np.pi), len(phis)), slice(0, np.searchsorted(phis, hi_a, side="right"))]
            elif hi_a > np.pi:
                sl = [slice(np.searchsorted(phis, lo_a), len(phis)), slice(0, np.searchsorted(phis, hi_a - 2 * np.pi, side="right"))]
            else:
                sl = [slice(np.searchsorted(phis, lo_a), np.searchsorted(phis, hi_a, side="right"))]
        for q in sl:
            if q.stop <= q.start:
                continue
            t1 = (b[:3] - origin) * inv[q]
            t2 = (b[3:] - origin) * inv[q]
            tn = np.nanmax(np.minimum(t1, t2), axis=1)
            tf = np.nanmin(np.maximum(t1, t2), axis=1)
            hit = tf >= np.maximum(tn, 0.0)
            tb = np.where(hit, np.where(tn > 1e-9, tn, tf), np.inf)
            ts[q] = np.minimum(ts[q], tb)
    t[order] = ts
    t[t > max_range] = np.inf
    return t


def lidar_scan(pose, boxes, elevations_deg, n_azimuth, rng, max_range=120.0, noise=0.02, az_offset=0.0):
    """One sweep of a spinning lidar.  pose: 4x4 sensor pose in the world.  Returns world-frame hits, n x 3 float64."""
    el = np.deg2rad(np.asarray(elevations_deg, dtype=np.float64))
    az = az_offset + np.arange(n_azimuth) * (2 * np.pi / n_azimuth)
    A, E = np.meshgrid(az, el, indexing="ij")            # azimuth-major firing order, like a real sweep
    d_s = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    dirs = d_s @ pose[:3, :3].T
    origin = pose[:3, 3]
    t = _raycast(origin, dirs, boxes, max_range)
    ok = np.isfinite(t)
    tn = t[ok] + rng.normal(0.0, noise, ok.sum())
    return origin + dirs[ok] * tn[:, None]


HDL64_ELEV = np.linspace(2.0, -24.8, 64)
VLP16_ELEV = np.arange(-15.0, 15.1, 2.0)


def _exactly(xyz, n, what):
    """Deterministically thin (uniform stride) to exactly n points."""
    if xyz.shape[0] < n:
        raise ValueError("%s: only %d points generated, need %d" % (what, xyz.shape[0], n))
    sel = np.floor(np.arange(n) * (xyz.shape[0] / n)).astype(np.int64)
    return xyz[sel]


def _prior_error(rng):
    """Erroneous prior pose of section 8(d): t ~ U(-0.3,0.3) m per axis, yaw ~ U(-3,3) deg, roll/pitch ~ U(-1,1) deg."""
    t = rng.uniform(-0.3, 0.3, 3)
    yaw = np.deg2rad(rng.uniform(-3, 3))
    roll, pitch = np.deg2rad(rng.uniform(-1, 1, 2))
    return rigid(t[0], t[1], t[2], roll, pitch, yaw)


def campus_map(n_points, rng, lx=200.0, ly=200.0, lz=12.0, n_boxes=120, n_clutter=0):
    """C4 map: surfaces of a lx x ly ground plane and n_boxes buildings, area-uniformly sampled.  Returns (xyz, boxes).
    n_clutter: that many small boxes (cars, kiosks, bollards: 0.4 - 3 m) on top -- faces in every direction close to the lane,
    so that also the better half of the matches, which is all a trimmed ratio of 0.5 keeps, constrains the pose."""
    boxes = []
    for _ in range(n_boxes):
        while True:
            cx, cy = rng.uniform(-lx / 2 + 8, lx / 2 - 8), rng.uniform(-ly / 2 + 8, ly / 2 - 8)
            if abs(cy) > 6.0:                          # free boulevard along x
                break
        sx, sy, sz = rng.uniform(3, 16), rng.uniform(3, 16), rng.uniform(2.5, lz)
        boxes.append([cx - sx / 2, cy - sy / 2, 0.0, cx + sx / 2, cy + sy / 2, sz])
    for _ in range(n_clutter):
        while True:
            cx, cy = rng.uniform(-lx / 2 + 2, lx / 2 - 2), rng.uniform(-ly / 2 + 2, ly / 2 - 2)
            if abs(cy) > 2.5:                          # the lane itself stays free
                break
        sx, sy, sz = rng.uniform(0.4, 3.0), rng.uniform(0.4, 3.0), rng.uniform(0.8, 3.0)
        boxes.append([cx - sx / 2, cy - sy / 2, 0.0, cx + sx / 2, cy + sy / 2, sz])
    boxes = np.array(boxes, dtype=np.float64)
    # faces: ground + 5 visible faces per box
    areas = [lx * ly]
    for b in boxes:
        sx, sy, sz = b[3] - b[0], b[4] - b[1], b[5] - b[2]
        areas += [sx * sy, sx * sz, sx * sz, sy * sz, sy * sz]
    areas = np.array(areas)
    counts = np.floor(areas / areas.sum() * n_points).astype(np.int64)
    counts[0] += n_points - counts.sum()
    out = np.empty((n_points, 3), dtype=np.float32)
    pos = 0
    u = rng.random((counts[0], 2))
    out[pos:pos + counts[0]] = np.stack([(u[:, 0] - 0.5) * lx, (u[:, 1] - 0.5) * ly, np.zeros(counts[0])], 1)
    pos += counts[0]
    f = 1
    for b in boxes:
        for face in range(5):
            c = counts[f]
            f += 1
            u = rng.random((c, 2))
            if face == 0:
                p = np.stack([b[0] + u[:, 0] * (b[3] - b[0]), b[1] + u[:, 1] * (b[4] - b[1]), np.full(c, b[5])], 1)
            elif face in (1, 2):
                y = b[1] if face == 1 else b[4]
                p = np.stack([b[0] + u[:, 0] * (b[3] - b[0]), np.full(c, y), b[2] + u[:, 1] * (b[5] - b[2])], 1)
            else:
                x = b[0] if face == 3 else b[3]
                p = np.stack([np.full(c, x), b[1] + u[:, 0] * (b[4] - b[1]), b[2] + u[:, 1] * (b[5] - b[2])], 1)
            out[pos:pos + c] = p
            pos += c
    out += rng.normal(0.0, 0.01, out.shape).astype(np.float32)
    return out, boxes


# ----------------------------------------------------------------------------------------------------------
# configuration pairs
# ----------------------------------------------------------------------------------------------------------
def make_pair(config, trial=0, n_points=None, variants=0):
    """Returns dict(ref, read: n x 3 float32 world frame; ref_origin, read_origin: float64[3]; T_true: 4x4 float64, the
    correction that registerClouds should recover (maps the reading onto the reference); name).
    n_points overrides the nominal cloud size (used by small parity tests).  variants = V > 0 (configs 2 and 3): a list
    of V such dicts that share the two scans and differ in the erroneous prior pose of the reading."""
    rng = np.random.default_rng(1000 * config + trial)
    if config == 2:
        n = n_points or 32768
        boxes = room_scene(rng)
        x0 = rng.uniform(-4.0, -2.0)

        def accumulate(xstart):
            pts = []
            for s in range(7):                       # batch_size = 7 sweeps, 0.35 m apart (aicp.launch:68)
                pose = rigid(xstart + 0.35 * s, rng.uniform(-0.05, 0.05), 0.6, 0, 0, rng.uniform(-0.05, 0.05))
                # denser azimuth sampling than the nominal 1800 when a larger cloud is requested
                pts.append(lidar_scan(pose, boxes, VLP16_ELEV, 1800 if n <= 32768 else 3600, rng, max_range=100.0,
                                      az_offset=rng.uniform(0, 0.01)))
            return np.concatenate(pts, 0), np.array([xstart + 0.35 * 3, 0.0, 0.6])
        ref_w, ref_o = accumulate(x0)
        read_w, read_o = accumulate(x0 + 0.35 * 7)
        ref = _exactly(voxel_downsample(ref_w, 0.08), n, "C2 reference")
        read_true = _exactly(voxel_downsample(read_w, 0.08), n, "C2 reading")
        name = "C2 VLP-16 ANYmal-shaped, 7 sweeps, 0.08 m voxel, %d pts" % n
    elif config == 3:
        n = n_points or 131072
        boxes = street_scene(rng)
        n_az = 1200 if n <= 16384 else (4800 if n <= 65536 else 12000)
        pose_a = rigid(rng.uniform(-10, -5), rng.uniform(-1, 1), 1.73, 0, 0, rng.uniform(-0.1, 0.1))
        pose_b = pose_a @ rigid(rng.uniform(1.0, 2.0), rng.uniform(-0.2, 0.2), 0, 0, 0, rng.uniform(-0.05, 0.05))

        def scan(pose):
            p = lidar_scan(pose, boxes, HDL64_ELEV, n_az, rng, max_range=120.0, az_offset=rng.uniform(0, 0.001))
            return p[p[:, 2] >= 0.2]                 # ground removed (pcl_ground_removal in the KITTI tools)
        ref = _exactly(scan(pose_a), n, "C3 reference").astype(np.float32)
        read_true = _exactly(scan(pose_b), n, "C3 reading").astype(np.float32)
        ref_o, read_o = pose_a[:3, 3].copy(), pose_b[:3, 3].copy()
        name = "C3 HDL-64 KITTI-shaped, ground removed, %d pts" % n
    elif config == 5:
        cube = cube_cloud()
        n = n_points or cube.shape[0]
        base = cube if n == cube.shape[0] else _exactly(cube, n, "C5 cube")
        ref = base.copy()
        read_true = (base.astype(np.float64) + rng.normal(0.0, 0.005, base.shape)).astype(np.float32)
        ref_o = np.zeros(3)
        read_o = np.zeros(3)
        P = perturbation_T(rng)
        read = apply_T(P, read_true)
        return dict(ref=ref, read=read, ref_origin=ref_o, read_origin=apply_T(P, read_o[None])[0].astype(np.float64),
                    T_true=np.linalg.inv(P), name="C5 cube validation pair, %d pts" % n)
    else:
        raise ValueError("make_pair supports configs 2, 3, 5 (C1: tests/golden/c1_scans.npz, C4: make_map_case)")
    out = []
    for v in range(max(1, int(variants))):
        # variant 0 draws the prior error from the scene's own stream (the pair every test and golden uses); further variants
        # express the SAME two scans through other erroneous priors: distinct inputs with distinct ICP trajectories for
        # the price of one ray cast (bench.py)
        E = _prior_error(rng if v == 0 else np.random.default_rng(1000 * config + trial + 1000003 * v))
        out.append(dict(ref=np.ascontiguousarray(ref, dtype=np.float32), read=apply_T(E, read_true),
                        ref_origin=np.asarray(ref_o, dtype=np.float64), read_origin=(E[:3, :3] @ read_o + E[:3, 3]),
                        T_true=np.linalg.inv(E), name=name if v == 0 else name + ", prior variant %d" % v))
    return out if variants else out[0]


def raw_sweep(config, trial=0, n_sweeps=None):
    """The UNFILTERED cloud that App hands to the pre-filter (regionGrowingUniformPlaneSegmentationFilter, app.cpp:102-110):
    config 2 -- 7 accumulated VLP-16 sweeps in the room scene (aicp.launch:68), ~200 k points; config 3 -- one HDL-64 sweep
    of the street scene with the ground, ~230 k points.  Returns dict(cloud n x 3 float32 world frame, origin float64[3])."""
    rng = np.random.default_rng(1000 * config + trial + 500)
    if config == 2:
        boxes = room_scene(rng)
        x0 = rng.uniform(-4.0, -2.0)
        pts = []
        ns = n_sweeps or 7
        for s in range(ns):
            pose = rigid(x0 + 0.35 * s, rng.uniform(-0.05, 0.05), 0.6, 0, 0, rng.uniform(-0.05, 0.05))
            pts.append(lidar_scan(pose, boxes, VLP16_ELEV, 1800, rng, max_range=100.0, az_offset=rng.uniform(0, 0.01)))
        return dict(cloud=np.concatenate(pts, 0).astype(np.float32), origin=np.array([x0 + 0.35 * (ns // 2), 0.0, 0.6]),
                    name="raw VLP-16 accumulation, %d sweeps" % ns)
    if config == 3:
        boxes = street_scene(rng)
        pose = rigid(rng.uniform(-10, -5), rng.uniform(-1, 1), 1.73, 0, 0, rng.uniform(-0.1, 0.1))
        p = lidar_scan(pose, boxes, HDL64_ELEV, (n_sweeps or 1) * 4000, rng, max_range=120.0, az_offset=rng.uniform(0, 0.001))
        return dict(cloud=p.astype(np.float32), origin=pose[:3, 3].copy(), name="raw HDL-64 sweep with ground")
    raise ValueError("raw_sweep supports configs 2 and 3")


def make_map_case(n_map=10485760, n_read=122880, trial=0, n_poses=1, remove_ground=False, n_clutter=0):
    """C4: fixed map + reading(s) scanned inside it with the HDL-64 model, expressed through an erroneous prior.
    remove_ground: drop the ground hits of the reading (z < 0.2 m) as the KITTI tools do before registration; with the
    ground in, the trimmed half of the matches is almost all ground and the pose is not constrained along it."""
    rng = np.random.default_rng(1000 * 4 + trial)
    map_xyz, boxes = campus_map(n_map, rng, n_clutter=n_clutter)
    out = []
    for _ in range(n_poses):
        pose = rigid(rng.uniform(-60, 60), rng.uniform(-3, 3), 1.73, 0, 0, rng.uniform(-np.pi, np.pi))
        n_az = 1024 if n_read <= 16384 else (12000 if remove_ground else 6000)
        p = lidar_scan(pose, boxes, HDL64_ELEV, n_az, rng, max_range=100.0)
        p = p[(np.abs(p[:, 0]) < 100) & (np.abs(p[:, 1]) < 100)]
        if remove_ground:
            p = p[p[:, 2] >= 0.2]
        read_true = _exactly(p, n_read, "C4 reading").astype(np.float32)
        E = _prior_error(rng)
        out.append(dict(read=apply_T(E, read_true), T_true=np.linalg.inv(E), read_origin=E[:3, :3] @ pose[:3, 3] + E[:3, 3]))
    return dict(map=map_xyz, readings=out, name="C4 %d-pt reading vs %d-pt map" % (n_read, n_map))


def c1_pair(reading=1, fixture=None):
    """C1: the planar sample scans of aicp_core/data (scan_00 = reference, scan_01 / scan_02 = readings) extruded to 3-D:
    16 copies at z = 0.00 .. 0.75 m (34 592 wall points) plus a 0.1 m floor grid under the scan, because a flat z = 0
    lift makes every normal +-z (point-to-plane degenerate).  Scans come from tests/golden/c1_scans.npz."""
    import os
    fixture = fixture or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                      "c1_scans.npz")
    z = np.load(fixture)

    def extrude(xy):
        layers = [np.c_[xy, np.full(len(xy), np.float32(0.05 * k))] for k in range(16)]
        lo, hi = xy.min(0), xy.max(0)
        gx, gy = np.meshgrid(np.arange(lo[0], hi[0], 0.1), np.arange(lo[1], hi[1], 0.1), indexing="ij")
        floor = np.c_[gx.ravel(), gy.ravel(), np.zeros(gx.size)]
        return np.concatenate(layers + [floor], 0).astype(np.float32)
    ref = extrude(z["scan_00"])
    read = extrude(z["scan_0%d" % reading])
    return dict(ref=ref, read=read, ref_origin=np.zeros(3), read_origin=np.zeros(3), T_true=None,
                name="C1 aicp_core/data sample scans extruded to 3-D, %d / %d pts" % (len(ref), len(read)))
