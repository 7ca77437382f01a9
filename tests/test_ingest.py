"""Ingest either side of the path: the PCD / pose-file readers of the replay format
(App::processFromFile, app.cpp:250-279; poseFileReader.hpp:46-78) and the sweep accumulation
(VelodyneAccumulatorROS::processLidar, aicp_ros/src/velodyne_accumulator.cpp:31-73).
not gpu: the library's readers (host code, no CUDA) against independent numpy / scipy parsers and hand-written files.
gpu    : the accumulation on the device against the oracle, bit for bit, and chained into the pre-filter without a host copy."""
import os

import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth


def write_pcd_by_hand(path, cols, names, data_kind, crlf=False):
    """An independent PCD writer: arbitrary float32 fields."""
    n = cols.shape[0]
    nl = "\r\n" if crlf else "\n"
    hdr = ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS " + " ".join(names), "SIZE " + " ".join(["4"] * len(names)),
           "TYPE " + " ".join(["F"] * len(names)), "COUNT " + " ".join(["1"] * len(names)), "WIDTH %d" % n, "HEIGHT 1",
           "VIEWPOINT 0 0 0 1 0 0 0", "POINTS %d" % n, "DATA " + data_kind]
    with open(path, "wb") as f:
        f.write((nl.join(hdr) + nl).encode())
        if data_kind == "binary":
            f.write(np.ascontiguousarray(cols, dtype=np.float32).tobytes())
        else:
            for row in cols:
                f.write((" ".join("nan" if np.isnan(v) else repr(float(v)) for v in row) + nl).encode())


def test_pcd_reader_ascii_binary_extra_fields(tmp_path):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-50, 50, (1000, 3)).astype(np.float32)
    pts[7] = np.nan
    inten = rng.uniform(0, 255, (1000, 1)).astype(np.float32)
    for kind in ("ascii", "binary"):
        for crlf in (False, True):
            if kind == "binary" and crlf:
                continue
            p = tmp_path / ("a_%s_%d.pcd" % (kind, crlf))
            write_pcd_by_hand(p, pts, ["x", "y", "z"], kind, crlf)
            got = ab.readPCD(p)
            assert got.shape == (1000, 4) and np.all(got[:, 3] == 1.0)
            assert np.array_equal(got[:, :3].view(np.uint32), pts.view(np.uint32))      # repr(float32) round-trips exactly
            # x y z not first, an extra field in between
            p2 = tmp_path / ("b_%s_%d.pcd" % (kind, crlf))
            write_pcd_by_hand(p2, np.c_[inten, pts[:, 2], pts[:, 0], inten, pts[:, 1]], ["intensity", "z", "x", "ring", "y"], kind, crlf)
            assert np.array_equal(ab.readPCD(p2)[:, :3].view(np.uint32), pts.view(np.uint32))
    empty = tmp_path / "empty.pcd"
    write_pcd_by_hand(empty, np.zeros((0, 3), np.float32), ["x", "y", "z"], "binary")
    assert ab.readPCD(empty).shape == (0, 4)


def test_pcd_writer_round_trip_and_cube_cloud(tmp_path):
    cube = synth.cube_cloud()                                   # create_cube_cloud.cpp writes this cloud with PCDWriter (:84)
    p = tmp_path / "cube.pcd"
    ab.writePCD(p, cube)
    back = ab.readPCD(p)
    assert np.array_equal(back[:, :3], cube) and back.shape[0] == cube.shape[0]
    head = open(p, "rb").read(200).decode(errors="replace")
    assert "FIELDS x y z" in head and ("POINTS %d" % cube.shape[0]) in head and "DATA binary" in head


def test_pcd_reader_rejects_what_it_does_not_implement(tmp_path):
    pts = np.zeros((4, 3), np.float32)
    p = tmp_path / "c.pcd"
    write_pcd_by_hand(p, pts, ["x", "y", "z"], "binary_compressed")
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        ab.readPCD(p)
    write_pcd_by_hand(p, pts[:, :2], ["x", "y"], "ascii")
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        ab.readPCD(p)
    with open(p, "wb") as f:                                    # truncated binary payload
        f.write(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 10\nHEIGHT 1\nPOINTS 10\nDATA binary\n" + b"\0" * 50)
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        ab.readPCD(p)
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        ab.readPCD(tmp_path / "missing.pcd")


def test_ply_reader(tmp_path):
    """The prior map is a PLY (app_ros.cpp:301): ascii and binary_little_endian, extra properties, double coordinates, faces after."""
    rng = np.random.default_rng(6)
    pts = rng.uniform(-80, 80, (500, 3)).astype(np.float32)
    p = tmp_path / "a.ply"
    with open(p, "wb") as f:
        f.write(b"ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 500\nproperty float x\nproperty float y\nproperty float z\n"
                b"property uchar red\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n")
        for q in pts:
            f.write(("%r %r %r 255\n" % (float(q[0]), float(q[1]), float(q[2]))).encode())
        f.write(b"3 0 1 2\n")
    got = ab.readPLY(p)
    assert got.shape == (500, 4) and np.array_equal(got[:, :3].view(np.uint32), pts.view(np.uint32)) and np.all(got[:, 3] == 1)
    p = tmp_path / "b.ply"
    rec = np.zeros(500, dtype=[("i", "<f4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1")])
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    with open(p, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 500\nproperty float intensity\nproperty float x\nproperty float y\n"
                b"property float z\nproperty uchar ring\nend_header\n" + rec.tobytes())
    assert np.array_equal(ab.readPLY(p)[:, :3].view(np.uint32), pts.view(np.uint32))
    p = tmp_path / "c.ply"
    rec = np.zeros(500, dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8")])
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    with open(p, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 500\nproperty double x\nproperty double y\nproperty double z\nend_header\n" + rec.tobytes())
    assert np.array_equal(ab.readPLY(p)[:, :3], pts)
    with open(p, "wb") as f:
        f.write(b"ply\nformat binary_big_endian 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n" + b"\0" * 12)
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        ab.readPLY(p)
    with open(p, "wb") as f:
        f.write(b"not a ply\n")
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        ab.readPLY(p)


def test_pose_file_reader_matches_scipy(tmp_path):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(2)
    rows = []
    lines = ["# counter, sec, nsec, x, y, z, qx, qy, qz, qw"]
    for i in range(20):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.uniform(-100, 100, 3)
        rows.append((i, 1500000000 + i, 123456 * i, t, q))
        lines.append("%d, %d, %d, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g" % (i, 1500000000 + i, 123456 * i, *t, *q))
    p = tmp_path / "aicp_input_poses.csv"
    p.write_text("\n".join(lines) + "\n")
    got = ab.readPoseFile(p)
    assert len(got) == 20
    for (c, s, ns, pose), (i, sec, nsec, t, q) in zip(got, rows):
        assert (c, s, ns) == (i, sec, nsec)
        assert np.abs(pose[:3, :3] - Rotation.from_quat(q).as_matrix()).max() < 1e-14 and np.array_equal(pose[:3, 3], t)
        assert np.array_equal(pose[3], [0, 0, 0, 1])
    bad = tmp_path / "bad.csv"
    bad.write_text("0, 1, 2, 3\n")
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        ab.readPoseFile(bad)


def test_replay_loop_reads_the_reference_layout(tmp_path):
    """app.cpp:250-279: <dir>/aicp_input_poses.csv + <dir>/cloud_<counter>_<sec>_<nsec>.pcd."""
    rng = np.random.default_rng(3)
    clouds = [rng.uniform(-5, 5, (100 + i, 3)).astype(np.float32) for i in range(3)]
    lines = ["# header"]
    for i, c in enumerate(clouds):
        ab.writePCD(tmp_path / ("cloud_%d_%d_%d.pcd" % (i, 10 + i, 500 + i)), c)
        lines.append("%d,%d,%d,%g,0,0,0,0,0,1" % (i, 10 + i, 500 + i, float(i)))
    lines.append("3,13,503,3,0,0,0,0,0,1")                     # no cloud file for this row: the replay stops there
    (tmp_path / "aicp_input_poses.csv").write_text("\n".join(lines) + "\n")
    seen = list(ab.processFromFile(str(tmp_path)))
    assert len(seen) == 3
    for i, (utime, cloud, pose) in enumerate(seen):
        assert utime == int((10 + i) * 1e6 + 500 + i) and np.array_equal(cloud[:, :3], clouds[i]) and pose[0, 3] == float(i)


def test_oracle_pose_to_float_transform(orc):
    for seed in range(10):
        rng = np.random.default_rng(seed)
        P = synth.rigid(*rng.uniform(-50, 50, 3), *rng.uniform(-3.1, 3.1, 3))
        T = orc.pose_to_float_transform(P)
        assert np.abs(T - P).max() < 5e-6 and np.array_equal(T[3], [0, 0, 0, 1])


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_accumulation_parity_and_device_chain(orc):
    rng = np.random.default_rng(4)
    boxes = synth.room_scene(rng)
    sweeps, poses = [], []
    for s in range(7):
        pose = synth.rigid(-3.0 + 0.35 * s, rng.uniform(-0.05, 0.05), 0.6, rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), rng.uniform(-3, 3))
        world = synth.lidar_scan(pose, boxes, synth.VLP16_ELEV, 900, rng, max_range=100.0)
        local = (world - pose[:3, 3]) @ pose[:3, :3]                    # the sensor-frame sweep a driver delivers
        local[::53] = np.nan
        local[::101] *= 20.0                                            # far returns: beyond the +-30 m crop
        sweeps.append(local.astype(np.float32)); poses.append(pose)
    acc = ab.B200VelodyneAccumulator(batch_size=7, device=0)
    pf = ab.B200Prefilter(device=0)
    try:
        for sw, P in zip(sweeps, poses):
            acc.processLidar(sw, P)
        assert acc.getFinished() and acc.processLidar(sweeps[0], poses[0]) == 0      # finished: further sweeps are ignored
        want = orc.accumulate_sweeps(sweeps, poses)
        got = acc.download()
        assert got.shape[0] > 50000 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # device chain: accumulated cloud -> pre-filter, no host copy in between
        out = pf.filter(acc.getCloud())
        o = orc.prefilter(want, threads=8)
        assert np.array_equal(out.view(np.uint32), o.cloud.view(np.uint32))
        # clearCloud starts over
        acc.clearCloud()
        assert acc.getCloud().shape[0] == 0
        acc.processLidar(sweeps[2], poses[2])
        assert np.array_equal(acc.download().view(np.uint32), orc.accumulate_sweeps(sweeps[2:3], poses[2:3]).view(np.uint32))
        # an empty / fully rejected sweep adds nothing
        assert acc.processLidar(np.full((10, 3), 1000.0, np.float32), poses[0]) == 0
    finally:
        acc.close(); pf.close()


def test_readers_survive_mutated_files(tmp_path):
    """Host-side parsers (PCD, PLY, pose file, SVM model): truncations and byte flips of valid files either parse or fail with
    an error code -- never crash, never read past the output buffer (the process surviving 600 mutants is the assertion)."""
    from aicp_mapping_b200 import classification
    rng = np.random.default_rng(12)
    pts = rng.uniform(-5, 5, (64, 3)).astype(np.float32)
    seeds = {}
    p = tmp_path / "s.pcd"; write_pcd_by_hand(p, pts, ["x", "y", "z"], "ascii"); seeds["pcd_a"] = (p.read_bytes(), ab.readPCD)
    p = tmp_path / "t.pcd"; write_pcd_by_hand(p, pts, ["x", "y", "z"], "binary"); seeds["pcd_b"] = (p.read_bytes(), ab.readPCD)
    ply = b"ply\nformat ascii 1.0\nelement vertex 64\nproperty float x\nproperty float y\nproperty float z\nend_header\n" + \
        b"".join(("%r %r %r\n" % tuple(float(v) for v in q)).encode() for q in pts)
    seeds["ply"] = (ply, ab.readPLY)
    seeds["pose"] = (b"# c\n0,1,2,0.5,0.25,0.125,0,0,0,1\n1,2,3,1,2,3,0.1,0.2,0.3,0.9\n", ab.readPoseFile)
    model = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "svm_models", "svm_1000training_thresh60.xml")
    seeds["svm"] = (open(model, "rb").read(), classification.parse_model)
    n_ok = n_err = 0
    for name, (data, reader) in seeds.items():
        for k in range(120):
            b = bytearray(data)
            mode = k % 3
            if mode == 0:
                b = b[:rng.integers(0, len(b))]
            elif mode == 1:
                for _ in range(int(rng.integers(1, 8))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            else:
                i = int(rng.integers(0, len(b)))
                b[i:i] = bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
            f = tmp_path / ("m_%s_%d" % (name, k))
            f.write_bytes(bytes(b))
            try:
                reader(str(f))
                n_ok += 1
            except ab.capi.AicpError:
                n_err += 1
    assert n_ok + n_err == 600 and n_err > 100
