"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libaicp_b200.so), against the CPU oracle on
the same seeded inputs.  Bars (BASELINE.json north_star): correspondence indices bit-exact, final transforms within
1e-5 m and 1e-5 rad, overlap within 1e-4.  By construction (exact integer reductions, float64 solve from + - * / sqrt)
the whole trajectory is expected to be bit-identical, and the tests assert that too."""
import os

import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import capi, synth
from conftest import rot_angle

pytestmark = pytest.mark.gpu

TOL_M = 1e-5      # metres, BASELINE.json
TOL_RAD = 1e-5    # radians, BASELINE.json
NCPU = os.cpu_count() or 1
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def reg():
    r = ab.B200Registration()
    yield r
    r.close()


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_transform_close(T_gpu, T_orc):
    d = T_gpu.astype(np.float64) @ np.linalg.inv(T_orc.astype(np.float64))
    assert np.linalg.norm(d[:3, 3]) <= TOL_M, "translation differs by %g m" % np.linalg.norm(d[:3, 3])
    assert rot_angle(d[:3, :3]) <= TOL_RAD, "rotation differs by %g rad" % rot_angle(d[:3, :3])


# ---- exact NN ------------------------------------------------------------------------------------------------------
def test_match_bit_exact_random_lattice_duplicates(reg, orc):
    rng = np.random.default_rng(11)
    ref = rng.uniform(-5, 5, (5000, 3)).astype(np.float32)
    ref[100:200] = ref[0:100]                       # duplicates: ties must go to the lowest ORIGINAL index
    ref = np.round(ref * 4) / 4                     # lattice: many equidistant candidates
    qry = np.round(rng.uniform(-6, 6, (4000, 3)).astype(np.float32) * 8) / 8
    gi, gd = reg.match(ref, qry)
    oi, od = orc.match(ref, qry, use_kdtree=False)
    assert np.array_equal(gi, oi) and np.array_equal(u32(gd), u32(od))


@pytest.mark.parametrize("n_ref,n_qry", [(1, 5), (7, 3), (8, 8), (9, 100), (1000, 1), (4097, 513)])
def test_match_ragged_sizes(reg, orc, n_ref, n_qry):
    rng = np.random.default_rng(100 + n_ref)
    ref = rng.normal(0, 3, (n_ref, 3)).astype(np.float32)
    qry = rng.normal(0, 4, (n_qry, 3)).astype(np.float32)
    gi, gd = reg.match(ref, qry)
    oi, od = orc.match(ref, qry, use_kdtree=False)
    assert np.array_equal(gi, oi) and np.array_equal(u32(gd), u32(od))


def test_match_bit_exact_full_size_lidar(reg, orc, pair_cache):
    pair = pair_cache(3, 0)                         # 131 072 x 131 072, the headline workload
    gi, gd = reg.match(pair["ref"], pair["read"])
    oi, od = orc.match(pair["ref"], pair["read"], use_kdtree=True, threads=NCPU)
    assert np.array_equal(gi, oi) and np.array_equal(u32(gd), u32(od))
    # size-independent property: every reference point is its own nearest neighbour at distance 0
    si, sd = reg.match(pair["ref"], pair["ref"])
    dup = sd != 0
    assert not dup.any()
    same = si == np.arange(len(si))
    assert np.array_equal(pair["ref"][si[~same]], pair["ref"][~same])      # only exact duplicates map elsewhere (lower id)
    assert np.all(si[~same] < np.flatnonzero(~same))


# ---- normals ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("schedule", [1, 2])
def test_surface_normals_parity_small_and_degenerate(reg, orc, schedule):
    """Both k-NN kernel schedules (1 warp per query, 2 tile per warp) against the oracle."""
    reg.setKnnSchedule(schedule)
    rng = np.random.default_rng(12)
    pts = rng.uniform(-2, 2, (3000, 3)).astype(np.float32)
    pts[:, 2] = (0.3 * pts[:, 0] - 0.2 * pts[:, 1] + 0.01 * rng.normal(size=3000)).astype(np.float32)
    pts[500:520] = pts[0:20]                        # duplicates inside neighbourhoods
    gn, gk = reg.surfaceNormals(pts, 20)
    on, ok = orc.surface_normals(pts, 20, use_kdtree=False)
    assert np.array_equal(gk, ok)
    assert np.array_equal(u32(gn), u32(on))
    t = np.linspace(0, 1, 200, dtype=np.float32)
    line = np.c_[t, 2 * t, -t].astype(np.float32)   # collinear neighbourhoods -> (0,1,0), SURVEY.md A.2
    gn, _ = reg.surfaceNormals(line, 10)
    on, _ = orc.surface_normals(line, 10)
    assert np.array_equal(u32(gn), u32(on)) and np.array_equal(gn[:, :3], np.tile(np.float32([0, 1, 0]), (200, 1)))
    with pytest.raises(capi.AicpError, match="KNN_TOO_LARGE"):
        reg.surfaceNormals(pts[:20], 20)
    reg.setKnnSchedule(0)


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("n,knn", [(21, 20), (33, 20), (64, 10), (65, 32), (97, 5), (1000, 1), (4099, 30)])
def test_surface_normals_ragged_sizes_and_knn(reg, orc, schedule, n, knn):
    """Cloud sizes around the 32-point tile / chunk boundaries, knn from 1 to 32, lattice points (many exact ties)."""
    reg.setKnnSchedule(schedule)
    rng = np.random.default_rng(100 + n)
    pts = (np.round(rng.uniform(-3, 3, (n, 3)) * 4) / 4).astype(np.float32)
    gn, gk = reg.surfaceNormals(pts, knn)
    on, ok = orc.surface_normals(pts, knn, use_kdtree=False)
    assert np.array_equal(gk, ok)
    assert np.array_equal(u32(gn), u32(on))
    reg.setKnnSchedule(0)


@pytest.mark.parametrize("schedule", [1, 2])
def test_surface_normals_parity_full_size(reg, orc, pair_cache, schedule):
    reg.setKnnSchedule(schedule)
    pair = pair_cache(3, 0)
    gn, gk = reg.surfaceNormals(pair["ref"], 20)
    on, ok = orc.surface_normals(pair["ref"], 20, use_kdtree=True, threads=NCPU)
    assert np.array_equal(gk, ok)
    assert np.array_equal(u32(gn), u32(on))
    reg.setKnnSchedule(0)


# ---- trimmed quantile --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ratio", [0.25, 0.358818, 0.5, 0.7, 0.999, 1.0])
def test_trim_threshold_parity(reg, orc, ratio):
    rng = np.random.default_rng(13)
    d2 = rng.exponential(0.3, 200001).astype(np.float32) ** 2
    d2[:50] = 0.0
    d2[50:60] = np.inf
    d2[1000:3000] = d2[999]                         # a heavy tie inside the distribution
    gl, gn = reg.trimThreshold(d2, ratio)
    ol, on = orc.trim_threshold(d2, ratio)
    assert gn == on and u32([gl])[0] == u32([ol])[0]


def test_trim_threshold_edge_cases(reg, orc):
    assert reg.trimThreshold(np.float32([3.0]), 0.7) == orc.trim_threshold(np.float32([3.0]), 0.7)
    with pytest.raises(capi.AicpError, match="NO_VALID_MATCH"):
        reg.trimThreshold(np.zeros(100, dtype=np.float32), 0.5)
    tiny = np.float32([1e-45, 1e-38, 3e38, 1.0, 2.0])       # denormals and huge values order like their bit patterns
    for r in (0.3, 0.5, 0.9, 1.0):
        assert reg.trimThreshold(tiny, r) == orc.trim_threshold(tiny, r)


# ---- full chain ------------------------------------------------------------------------------------------------------
def run_both(reg, orc, ref, read, ratio=0.7, init=None, threads=NCPU, **kw):
    reg.setConfig(ratio=ratio, **kw)
    reg.enableMatchTrace(True)
    T = reg.registerClouds(ref, read) if init is None else reg.registerCloudsInit(ref, read, init)
    cfg = orc.default_config(ratio=ratio, threads=threads, **kw)
    o = orc.icp(ref, read, cfg, init_T=init, want_trace_idx=True, want_normals=True)
    assert o.rc == 0
    return T, o


def assert_full_parity(reg, T, o):
    st = reg.stats
    assert st.iterations == o.iterations and st.stop_reason == o.stop_reason
    assert_transform_close(T, o.T)
    # correspondences of EVERY iteration are bit-exact
    assert np.array_equal(reg.getTraceMatches(), o.trace_idx)
    tr = reg.trace()
    for g, c in zip(tr, o.trace):
        assert g["n_valid"] == c["n_valid"] and g["n_used"] == c["n_used"]
        assert u32([g["limit_d2"]])[0] == u32([c["limit_d2"]])[0]
        assert np.array_equal(u32(g["T_iter"]), u32(c["T_iter"]))
    assert np.array_equal(u32(T), u32(o.T))                               # by construction: bit-identical
    assert np.array_equal(u32(reg.getOutputReading()), u32(o.reading))    # pointmatcher_registration.cpp:128-131
    assert np.float32(reg.getWeightedPointUsedRatio()) == o.weighted_point_used_ratio
    assert np.array_equal(u32(reg.getReferenceNormals()), u32(o.normals))


@pytest.mark.parametrize("trial,ratio", [(0, 0.7), (1, 0.25), (2, 0.358818)])
def test_icp_parity_cube_pairs(reg, orc, trial, ratio):
    pair = synth.make_pair(5, trial)
    T, o = run_both(reg, orc, pair["ref"], pair["read"], ratio)
    assert_full_parity(reg, T, o)
    if ratio >= 0.7:          # with a low trimmed ratio the cube pair may settle in a local optimum -- in both paths alike
        err = T.astype(np.float64) @ np.linalg.inv(pair["T_true"])
        assert np.linalg.norm(err[:3, 3]) < 5e-3 and rot_angle(err[:3, :3]) < 2e-3      # and it actually registers


def test_icp_parity_c1_sample_scans(reg, orc):
    for reading in (1, 2):
        pair = synth.c1_pair(reading)
        T, o = run_both(reg, orc, pair["ref"], pair["read"], 0.7)
        assert_full_parity(reg, T, o)


@pytest.mark.parametrize("match_schedule,knn_schedule,loop_schedule", [(1, 1, 2), (2, 2, 2), (1, 1, 1), (2, 2, 1)])
def test_icp_parity_c2_vlp16(reg, orc, pair_cache, match_schedule, knn_schedule, loop_schedule):
    """Whole trajectory against the oracle with the per-thread kernels (1, 1) and with the tile kernels (2, 2), driven by the
    persistent loop kernel (2) and by three launches per iteration (1)."""
    reg.setMatchSchedule(match_schedule); reg.setKnnSchedule(knn_schedule); reg.setLoopSchedule(loop_schedule)
    pair = pair_cache(2, 0)
    T, o = run_both(reg, orc, pair["ref"], pair["read"], 0.7)
    assert_full_parity(reg, T, o)
    reg.setMatchSchedule(0); reg.setKnnSchedule(0); reg.setLoopSchedule(0)


def test_loop_schedules_agree_and_persistent_is_one_launch(reg, pair_cache):
    """The persistent loop kernel and the multi-launch loop give the same bits; the persistent one costs one launch for all
    iterations (the multi-launch loop: two per iteration), and fills the per-phase clocks of the stats."""
    pair = pair_cache(2, 0)
    reg.setConfig(ratio=0.7)
    out = {}
    for ls in (2, 1):
        reg.setLoopSchedule(ls)
        T = reg.registerClouds(pair["ref"], pair["read"])
        out[ls] = (u32(T).copy(), reg.stats.iterations, reg.stats.gpu_launches, reg.stats.ms_match, u32(reg.getOutputReading()).copy())
    reg.setLoopSchedule(0)
    assert np.array_equal(out[2][0], out[1][0]) and out[2][1] == out[1][1] and np.array_equal(out[2][4], out[1][4])
    assert out[1][2] - out[2][2] >= 2 * out[2][1] - 1        # multi-launch: search + (quantile, normal equations, solve) per iteration
    assert out[2][3] > 0.0


def test_wait_stream_orders_device_inputs_produced_on_another_stream(reg, pair_cache):
    """aicp_b200_wait_stream: a device cloud produced on a side stream (behind a long-running kernel) and handed over without
    synchronising that stream gives the same transform as the host arrays."""
    import torch
    pair = pair_cache(2, 0)
    reg.setConfig(ratio=0.7)
    T_host = reg.registerClouds(pair["ref"], pair["read"])
    side = torch.cuda.Stream()
    src_ref = torch.from_numpy(capi.to_xyzw(pair["ref"])).cuda()
    src_read = torch.from_numpy(capi.to_xyzw(pair["read"])).cuda()
    big = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(20):
            big.mul_(1.0001)                       # ~ms of queued work in front of the copies
        ref_dev = torch.zeros_like(src_ref); ref_dev.copy_(src_ref)
        read_dev = torch.zeros_like(src_read); read_dev.copy_(src_read)
    reg.waitStream(side)                           # the wrapper itself only synchronises the CURRENT (default) stream
    T_dev = reg.registerClouds(ref_dev, read_dev)
    assert np.array_equal(u32(T_dev), u32(T_host))
    side.synchronize()


@pytest.mark.parametrize("n_ref,n_qry", [(20, 7), (33, 40), (100, 31), (5000, 1000)])
def test_icp_tile_kernels_ragged_sizes(reg, orc, n_ref, n_qry):
    """Tile kernels on cloud sizes around the 32-point tile boundary (partial warps, single-chunk references)."""
    rng = np.random.default_rng(n_ref * 7 + n_qry)
    ref = rng.uniform(-1, 1, (n_ref, 3)).astype(np.float32)
    ref[:, 2] = (0.2 * ref[:, 0] + 0.02 * rng.normal(size=n_ref)).astype(np.float32)
    read = (ref[rng.integers(0, n_ref, n_qry)] + np.float32([0.03, -0.02, 0.01]) +
            0.002 * rng.normal(size=(n_qry, 3))).astype(np.float32)
    for sched in (1, 2):
        reg.setMatchSchedule(sched); reg.setKnnSchedule(sched)
        T, o = run_both(reg, orc, ref, read, 0.8, knn_normals=10, max_iterations=6)
        assert_full_parity(reg, T, o)
    reg.setMatchSchedule(0); reg.setKnnSchedule(0)
    reg.setConfig(knn_normals=20, max_iterations=20)          # the handle is shared by the module: back to the defaults


def test_icp_parity_c3_hdl64_full_size(reg, orc, pair_cache):
    pair = pair_cache(3, 0)
    T, o = run_both(reg, orc, pair["ref"], pair["read"], 0.7)
    assert_full_parity(reg, T, o)
    reg.setMatchSchedule(2); reg.setKnnSchedule(2)           # the schedules the batched path picks
    T, o2 = run_both(reg, orc, pair["ref"], pair["read"], 0.7)
    assert_full_parity(reg, T, o2)
    reg.setMatchSchedule(0); reg.setKnnSchedule(0)
    # second trial, auto-tuned ratio from the GPU overlap
    pair = pair_cache(3, 1, 65536)
    ov = ab.B200Overlap()
    ov.computeOverlap(pair["ref"], pair["read"], pair["ref_origin"], pair["read_origin"])
    ratio = ab.autotune_ratio(float(ov.getOverlap()))
    T, o = run_both(reg, orc, pair["ref"], pair["read"], ratio)
    assert_full_parity(reg, T, o)
    ov.close()


def test_icp_counter_stop_and_init_transform(reg, orc):
    pair = synth.make_pair(5, 4, 6000)
    T, o = run_both(reg, orc, pair["ref"], pair["read"], 0.6, max_iterations=3)
    assert reg.stats.iterations == 3 and reg.stats.stop_reason == capi.STOP_COUNTER
    assert_full_parity(reg, T, o)
    init = synth.rigid(0.02, -0.01, 0.0, 0, 0, 0.01).astype(np.float32)
    T, o = run_both(reg, orc, pair["ref"], pair["read"], 0.6, init=init, max_iterations=20)
    assert_full_parity(reg, T, o)
    assert np.array_equal(u32(reg.getInitializedReading()), u32(orc.transform_points(init, pair["read"])))


def test_setup_graph_replay_and_invalidation(orc):
    """The setup of a registration (index, normals, loop state, reading) is replayed as a CUDA graph from the third meeting
    of the same cloud sizes on (icp.cu, run_registration).  Same sizes with other DATA, other sizes in between (buffers grow,
    the graph must be dropped), an initial transform (another graph key), another knn, and profiling level 2 (plain
    launches): every result equals the oracle's, whichever way the setup ran."""
    reg = ab.B200Registration()
    try:
        a = [synth.make_pair(5, t, 5000) for t in range(3)]            # three different pairs of one size
        b = synth.make_pair(5, 7, 9000)                                # a larger pair: every buffer is reallocated
        want = {}

        def check(pair, tag, ratio=0.6, init=None, knn=20):
            reg.setConfig(ratio=ratio, max_iterations=20, knn_normals=knn)
            T = reg.registerClouds(pair["ref"], pair["read"]) if init is None else reg.registerCloudsInit(pair["ref"], pair["read"], init)
            if tag not in want:
                o = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=ratio, threads=NCPU, max_iterations=20, knn_normals=knn), init_T=init)
                assert o.rc == 0
                want[tag] = (o.T, o.iterations)
            assert np.array_equal(u32(T), u32(want[tag][0])) and reg.stats.iterations == want[tag][1], tag

        for rep in range(2):
            for t in range(3):
                check(a[t], "a%d" % t)                                 # rep 0: plain, plain (seen), captured; rep 1: replayed
        check(b, "b")                                                  # other sizes: reallocation, graph dropped
        for t in range(3):
            check(a[t], "a%d" % t)
        check(b, "b"); check(b, "b"); check(b, "b")
        init = synth.rigid(0.02, -0.01, 0.0, 0, 0, 0.01).astype(np.float32)
        for _ in range(3):
            check(a[1], "a1_init", init=init)
            assert np.array_equal(u32(reg.getInitializedReading()), u32(orc.transform_points(init, a[1]["read"])))
        for _ in range(3):
            check(a[2], "a2_knn10", knn=10)
        check(a[0], "a0")
        reg.setProfiling(2)
        check(a[0], "a0"); check(a[0], "a0")
        reg.setProfiling(0)
        check(a[0], "a0"); check(a[0], "a0"); check(a[0], "a0")
    finally:
        reg.close()


def test_icp_through_yaml_file_like_app(reg, orc, tmp_path):
    """App::computeRegistration path: clamp -> rewrite YAML -> updateConfigParams -> registerClouds (app.cpp:187-216)."""
    pair = synth.make_pair(5, 5, 8000)
    cfg_file = str(tmp_path / "icp_autotuned.yaml")
    T = ab.computeRegistration(reg, pair["ref"], pair["read"], 35.8818, os.path.join(GOLDEN, "icp_autotuned_default.yaml"), cfg_file)
    got = reg.getConfig()
    assert got.ratio == np.float32(0.358818) and got.knn_normals == 20 and got.max_iterations == 20
    o = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=float(np.float32(0.358818)), threads=NCPU))
    assert reg.stats.iterations == o.iterations and np.array_equal(u32(T), u32(o.T))
    reg.updateConfigParams(os.path.join(GOLDEN, "icp_3D_cfg_trimmed.yaml"))
    with pytest.raises(capi.AicpError, match="CONFIG"):
        reg.registerClouds(pair["ref"], pair["read"])
    reg.updateConfigParams("")


def test_icp_error_statuses_match_oracle(reg, orc):
    pair = synth.make_pair(5, 6, 3000)
    reg.setConfig(ratio=0.7, max_iterations=20)
    with pytest.raises(capi.AicpError, match="NO_VALID_MATCH"):             # identical clouds: every distance is zero
        reg.registerClouds(pair["ref"], pair["ref"])
    assert orc.icp(pair["ref"], pair["ref"], orc.default_config()).error == "NO_VALID_MATCH"
    bad = pair["read"].copy(); bad[17, 1] = np.nan
    with pytest.raises(capi.AicpError, match="NONFINITE_INPUT"):
        reg.registerClouds(pair["ref"], bad)
    far = pair["read"].copy(); far[:, 0] += 5000.0
    with pytest.raises(capi.AicpError, match="EXTENT"):
        reg.registerClouds(pair["ref"], far)
    assert orc.icp(pair["ref"], far, orc.default_config()).error == "EXTENT"
    with pytest.raises(capi.AicpError, match="KNN_TOO_LARGE"):
        reg.registerClouds(pair["ref"][:15], pair["read"])
    T = reg.registerClouds(pair["ref"], pair["read"])                        # the handle still works afterwards
    assert np.all(np.isfinite(T))


def test_fixed_reference_reuse(reg, orc):
    """aicp_b200_set_reference + register_to_reference give the same result as a full registerClouds."""
    pair = synth.make_pair(5, 7, 9000)
    reg.setConfig(ratio=0.6, max_iterations=20)
    T_full = reg.registerClouds(pair["ref"], pair["read"])
    reg.setReference(pair["ref"])
    T_a = reg.registerToReference(pair["read"])
    T_b = reg.registerToReference(pair["read"])
    assert np.array_equal(u32(T_full), u32(T_a)) and np.array_equal(u32(T_a), u32(T_b))


def test_device_pointers_are_used_in_place(reg, orc):
    import torch
    pair = synth.make_pair(5, 8, 5000)
    reg.setConfig(ratio=0.7, max_iterations=20)
    T_host = reg.registerClouds(pair["ref"], pair["read"])
    ref_d = torch.from_numpy(capi.to_xyzw(pair["ref"])).cuda()
    read_d = torch.from_numpy(capi.to_xyzw(pair["read"])).cuda()
    T_dev = reg.registerClouds(ref_d, read_d)
    assert np.array_equal(u32(T_host), u32(T_dev))


def test_register_batch_concurrent_streams_equals_sequential(reg, orc):
    """Independent pairs registered concurrently on 3 streams give bit-identical transforms to one-by-one calls, and a
    failing pair (identical clouds -> ConvergenceError) does not stop the others."""
    pairs = [synth.make_pair(5, t, 4000 + 500 * t) for t in range(5)]
    ratios = [0.7, 0.6, 0.5, 0.65, 0.7]
    reg.setConfig(max_iterations=20)
    seq = []
    for p, r in zip(pairs, ratios):
        reg.setConfig(ratio=r)
        seq.append(reg.registerClouds(p["ref"], p["read"]))
    reg.setConfig(ratio=0.7)
    T, stats, status, ms = reg.registerBatch([(p["ref"], p["read"]) for p in pairs], ratios=ratios, streams=3)
    assert ms > 0 and not status.any()
    for a, b in zip(seq, T):
        assert np.array_equal(u32(a), u32(b))
    o = orc.icp(pairs[2]["ref"], pairs[2]["read"], orc.default_config(ratio=0.5, threads=NCPU))
    assert stats[2].iterations == o.iterations and np.array_equal(u32(T[2]), u32(o.T))
    bad = [(pairs[0]["ref"], pairs[0]["read"]), (pairs[1]["ref"], pairs[1]["ref"]), (pairs[2]["ref"], pairs[2]["read"])]
    with pytest.raises(capi.AicpError, match="NO_VALID_MATCH"):
        reg.registerBatch(bad, ratios=ratios[:3], streams=2)


def test_register_batch_reuses_worker_buffers_across_sizes(reg):
    """Many small pairs of different sizes on 2 streams (every worker's staging and index buffers are reused by clouds of other
    sizes) give the transforms of one-by-one calls, twice in a row and again after a batch in which a pair in the middle failed."""
    pairs = [synth.make_pair(5, t, 3000 + 371 * (t % 5)) for t in range(11)]
    ratios = [0.7, 0.6, 0.5, 0.65, 0.7, 0.55, 0.7, 0.6, 0.5, 0.65, 0.7]
    reg.setConfig(max_iterations=20)
    seq = []
    for p, r in zip(pairs, ratios):
        reg.setConfig(ratio=r)
        seq.append(reg.registerClouds(p["ref"], p["read"]))
    for _ in range(2):
        T, stats, status, ms = reg.registerBatch([(p["ref"], p["read"]) for p in pairs], ratios=ratios, streams=2)
        assert not status.any()
        for a, b in zip(seq, T):
            assert np.array_equal(u32(a), u32(b))
    bad = [(p["ref"], p["read"]) for p in pairs]
    bad[3] = (pairs[3]["ref"], pairs[3]["ref"])
    with pytest.raises(capi.AicpError, match="NO_VALID_MATCH"):
        reg.registerBatch(bad, ratios=ratios, streams=2)
    T, stats, status, ms = reg.registerBatch([(p["ref"], p["read"]) for p in pairs], ratios=ratios, streams=2)
    assert not status.any() and all(np.array_equal(u32(a), u32(b)) for a, b in zip(seq, T))


def test_cpp_adapter_demo():
    """The C++ adapters (include/aicp_b200_adapter.hpp) compiled against the reference's own abstract plug-in headers:
    overlap -> updateConfigParams -> registerClouds -> getOutputReading through the factories, as App::runAicpPipeline does.
    The binary is built in the build container (tests/stubs/build_demo.sh needs the reference's headers) and travels."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs", "build", "adapter_demo")
    if not os.path.exists(exe):
        pytest.skip("tests/stubs/build/adapter_demo was not built (needs the reference headers at build time)")
    model = os.path.join(GOLDEN, "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")
    r = subprocess.run([exe, os.path.join(GOLDEN, "icp_autotuned_default.yaml"), "40", model], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    # B200SVM : AbstractClassification through the factory branch, against the numpy oracle of cv::ml::SVM
    from oracle import aicp_oracle_svm as svm_orc
    risk = float([l for l in r.stdout.splitlines() if l.startswith("risk ")][0].split()[1])
    assert abs(risk - svm_orc.test(svm_orc.load_model(model), np.array([[61.63, 50.02]]))[0]) <= 1e-6
    assert "prefilter 9600 -> " in r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("OK ")][0]
    T = np.array([float(x) for x in line.split("T=")[1].split()]).reshape(4, 4).T
    assert 4 <= int(line.split("iterations=")[1].split()[0]) <= 20
    # the demo's reading is the reference rotated by 0.02 rad about z and shifted by (0.05, -0.03): T undoes it
    assert abs(np.arctan2(T[1, 0], T[0, 0]) + 0.02) < 2e-3 and np.abs(T[:2, 3] - [-0.0506, 0.0290]).max() < 5e-3


# ---- overlap -------------------------------------------------------------------------------------------------------
def test_overlap_parity(orc, pair_cache):
    ov = ab.B200Overlap()
    cases = [pair_cache(2, 0, 8192), pair_cache(3, 1, 65536), synth.make_pair(5, 0)]
    for pair in cases:
        counts = ov.computeOverlap(pair["ref"], pair["read"], pair["ref_origin"], pair["read_origin"])
        o_ov, o_counts = orc.overlap(pair["ref"], pair["ref_origin"], pair["read"], pair["read_origin"])
        assert counts == o_counts, pair["name"]
        assert abs(float(ov.getOverlap()) - float(o_ov)) <= 1e-4 and u32([ov.getOverlap()])[0] == u32([o_ov])[0]
    # lattice points exactly on voxel boundaries, origin inside the cloud: the DDA overshoots its end voxels here
    t = (np.arange(40, dtype=np.float32) * np.float32(0.1) - np.float32(2.0))
    g = np.stack(np.meshgrid(t, t, indexing="ij"), -1).reshape(-1, 2)
    lattice = np.concatenate([np.c_[g, np.full(len(g), v)] for v in (-2.0, 2.0)] + [np.c_[np.full(len(g), v), g] for v in (-2.0, 2.0)] +
                             [np.c_[g[:, 0], np.full(len(g), v), g[:, 1]] for v in (-2.0, 2.0)]).astype(np.float32)
    shifted = lattice + np.float32([0.05, -0.03, 0.0007])
    counts = ov.computeOverlap(lattice, shifted, [0, 0, 0], [0, 0, 0])
    o_ov, o_counts = orc.overlap(lattice, [0, 0, 0], shifted, [0, 0, 0])
    assert counts == o_counts and u32([ov.getOverlap()])[0] == u32([o_ov])[0]
    # identical clouds -> 100 %, disjoint -> 0 %
    a = cases[0]["ref"]
    ov.computeOverlap(a, a, [0, 0, 0.6], [0, 0, 0.6])
    assert ov.getOverlap() == np.float32(100.0)
    ov.computeOverlap(a, a + np.float32(500.0), [0, 0, 0.6], [500, 500, 500.6])
    assert ov.getOverlap() == 0.0 and ov.counts[0] == 0
    ov.close()


# ---- randomised sweep -------------------------------------------------------------------------------------------------
def _random_scene(rng, n, kind):
    if kind == "plane":          # noisy tilted plane: well-conditioned normals
        p = rng.uniform(-2, 2, (n, 3)); p[:, 2] = 0.3 * p[:, 0] - 0.1 * p[:, 1] + 0.01 * rng.normal(size=n)
    elif kind == "corner":       # three orthogonal faces: constrains all six degrees of freedom
        f = rng.integers(0, 3, n); p = rng.uniform(0, 2, (n, 3)); p[np.arange(n), f] = 0.005 * rng.normal(size=n)
    elif kind == "lattice":      # quantised coordinates: many exactly equal distances (tie rule)
        f = rng.integers(0, 3, n); p = np.round(rng.uniform(0, 2, (n, 3)) * 16) / 16; p[np.arange(n), f] = 0.0
    elif kind == "clusters":     # dense blobs far apart: deep, unbalanced radix tree and large empty cells
        c = rng.uniform(-20, 20, (8, 3)); p = c[rng.integers(0, 8, n)] + 0.05 * rng.normal(size=(n, 3))
        p[:, 2] *= 0.05
    else:                        # duplicates: identical points inside every neighbourhood
        f = rng.integers(0, 3, n); p = rng.uniform(0, 2, (n, 3)); p[np.arange(n), f] = 0.004 * rng.normal(size=n)
        p[n // 2:] = p[:n - n // 2]
    return p.astype(np.float32)


@pytest.mark.parametrize("seed", range(24))
def test_icp_randomised_parity_sweep(reg, orc, seed):
    """Random scene type, sizes, knn, ratio, perturbation and kernel schedule; the whole trajectory must equal the oracle's.
    Error outcomes (ConvergenceError-type statuses) must match too."""
    rng = np.random.default_rng(9000 + seed)
    kind = ["plane", "corner", "lattice", "clusters", "duplicates"][seed % 5]
    n_ref, n_read = int(rng.integers(40, 6000)), int(rng.integers(1, 4000))
    knn = int(rng.integers(3, 33))
    if knn >= n_ref:
        knn = n_ref - 1
    ratio = float(np.float32(rng.choice([0.25, 0.4, 0.55, 0.7, 0.9, 1.0])))
    ref = _random_scene(rng, n_ref, kind)
    P = synth.rigid(*rng.uniform(-0.05, 0.05, 3), *np.deg2rad(rng.uniform(-1.5, 1.5, 3)))
    read = synth.apply_T(P, ref[rng.integers(0, n_ref, n_read)].astype(np.float64) + 0.003 * rng.normal(size=(n_read, 3)))
    sched = 1 + seed % 2
    reg.setMatchSchedule(sched); reg.setKnnSchedule(sched); reg.setLoopSchedule(1 + (seed // 2) % 2)
    reg.setConfig(ratio=ratio, knn_normals=knn, max_iterations=12)
    reg.enableMatchTrace(True)
    cfg = orc.default_config(ratio=ratio, threads=NCPU, knn_normals=knn, max_iterations=12)
    o = orc.icp(ref, read, cfg, want_trace_idx=True, want_normals=True)
    try:
        T = reg.registerClouds(ref, read)
        code = "OK"
    except capi.AicpError as e:
        T, code = None, e.code_name
    finally:
        reg.setMatchSchedule(0); reg.setKnnSchedule(0); reg.setLoopSchedule(0)
        reg.setConfig(knn_normals=20, max_iterations=20)
    assert code == o.error, (kind, n_ref, n_read, knn, ratio)
    if T is not None:
        assert_full_parity(reg, T, o)
