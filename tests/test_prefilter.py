"""regionGrowingUniformPlaneSegmentationFilter = pcl::VoxelGrid + NormalEstimation + RegionGrowing
(aicp_core/src/utils/filteringUtils.cpp:5-104), SURVEY.md 8(f) rank 1.
not gpu: the oracle (oracle/aicp_oracle_prefilter.c) against analytic cases and independent numpy statements; the
         equivalence between PCL's SEQUENTIAL region growing (as the oracle runs it) and the min-label fixed point that the
         CUDA path computes (csrc/prefilter.cu), on lidar neighbourhood graphs and on adversarial random directed graphs.
gpu    : the CUDA pre-filter through the C ABI against the oracle, bit for bit (voxel centroids, normals, curvature, labels,
         output cloud and its order)."""
import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth


# ---------------------------------------------------------------------------------------------------------------------
# independent statements
# ---------------------------------------------------------------------------------------------------------------------
def numpy_voxel_grid(pts, leaf=np.float32(0.08)):
    """pcl::VoxelGrid from its definition: float32 voxel coordinates floor(p * (1/leaf)), x fastest linear index, float64
    centroid per voxel.  Returns (centroids float64, counts)."""
    p = np.asarray(pts, dtype=np.float32)[:, :3]
    p = p[np.all(np.isfinite(p), axis=1)]
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(p * inv).astype(np.int64)
    ijk -= np.floor(p.min(0) * inv).astype(np.int64)
    div = np.floor(p.max(0) * inv).astype(np.int64) - np.floor(p.min(0) * inv).astype(np.int64) + 1
    lin = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(lin, kind="stable")
    ls = lin[order]
    starts = np.flatnonzero(np.r_[True, ls[1:] != ls[:-1]])
    counts = np.diff(np.r_[starts, ls.size])
    sums = np.add.reduceat(p[order].astype(np.float64), starts, axis=0)
    return sums / counts[:, None], counts


def min_label_regions(normals, knn, n_nb, cos_thr):
    """The fixed point the CUDA path computes: label(v) = lowest (curvature, index) rank among the points that reach v over
    the directed graph u -> w (w among u's first n_nb neighbours, |n_u . n_w| >= cos_thr).  Plain numpy sweeps."""
    m = knn.shape[0]
    order = np.lexsort((np.arange(m), normals[:, 3]))
    rank = np.empty(m, dtype=np.int64)
    rank[order] = np.arange(m)
    f = np.float32
    src = np.repeat(np.arange(m), n_nb)
    dst = knn[:, :n_nb].reshape(-1)
    a, b = normals[dst], normals[src]
    dot = np.abs((a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]).astype(f) + (a[:, 2] * b[:, 2]).astype(f))
    ok = ~(dot < f(cos_thr))
    src, dst = src[ok], dst[ok]
    label = rank.copy()
    while True:
        new = label.copy()
        np.minimum.at(new, dst, label[src])
        if np.array_equal(new, label):
            break
        label = new
    return label, rank


def condensed_min_label_regions(normals, knn, n_nb, cos_thr):
    """The way csrc/prefilter.cu reaches that fixed point: points joined by TWO-WAY edges are mutually reachable, so they share
    their ancestors and their label; components of the two-way graph first (scipy), then min-propagation between
    components along the remaining edges."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    m = knn.shape[0]
    order = np.lexsort((np.arange(m), normals[:, 3]))
    rank = np.empty(m, dtype=np.int64)
    rank[order] = np.arange(m)
    f = np.float32
    src = np.repeat(np.arange(m), n_nb)
    dst = knn[:, :n_nb].reshape(-1).astype(np.int64)
    a, b = normals[dst], normals[src]
    dot = np.abs((a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]).astype(f) + (a[:, 2] * b[:, 2]).astype(f))
    ok = ~(dot < f(cos_thr))
    src, dst = src[ok], dst[ok]
    edges = set(zip(src.tolist(), dst.tolist()))
    two_way = np.array([(u, w) for (u, w) in edges if u != w and (w, u) in edges], dtype=np.int64).reshape(-1, 2)
    g = coo_matrix((np.ones(len(two_way)), (two_way[:, 0], two_way[:, 1])), shape=(m, m))
    n_comp, comp = connected_components(g, directed=False)
    clabel = np.full(n_comp, m, dtype=np.int64)
    np.minimum.at(clabel, comp, rank)
    cs, cd = comp[src], comp[dst]
    keep = cs != cd
    cs, cd = cs[keep], cd[keep]
    passes = 0
    while True:
        new = clabel.copy()
        np.minimum.at(new, cd, clabel[cs])
        passes += 1
        if np.array_equal(new, clabel):
            break
        clabel = new
    return clabel[comp], n_comp, passes


def labels_from_min_label(label, min_size, max_size):
    """Cluster ordinals in seed order for regions whose size is within [min, max], else -1 (assembleRegions)."""
    seeds, counts = np.unique(label, return_counts=True)
    keep = (counts >= min_size) & (counts <= max_size)
    ordinal = np.full(label.max() + 1, -1, dtype=np.int64)
    ordinal[seeds[keep]] = np.arange(keep.sum())
    return ordinal[label].astype(np.int32), int(keep.sum())


def two_planes(rng, n=60000, noise=0.0):
    """Floor z = 0 (8 x 8 m) and wall x = 0 (8 x 3 m), plus a 30-point speck far away."""
    nf = n * 2 // 3
    floor = np.c_[rng.uniform(0.3, 8, nf), rng.uniform(-4, 4, nf), np.zeros(nf)]
    wall = np.c_[np.zeros(n - nf), rng.uniform(-4, 4, n - nf), rng.uniform(0.3, 3, n - nf)]
    speck = np.c_[rng.uniform(20, 20.3, 30), rng.uniform(0, 0.3, 30), np.full(30, 1.0)]
    pts = np.concatenate([floor, wall, speck], 0)
    pts += rng.normal(0, noise, pts.shape) if noise else 0.0
    return pts.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# oracle (CPU)
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_voxel_grid_analytic(orc):
    leaf = np.float32(0.5)
    pts = np.float32([[0.1, 0.1, 0.1], [0.3, 0.2, 0.4],          # voxel (0,0,0)
                      [0.6, 0.1, 0.1],                             # voxel (1,0,0)
                      [0.1, 0.7, 0.2], [0.2, 0.9, 0.3], [0.4, 0.6, 0.1],   # voxel (0,1,0)
                      [np.nan, 0.0, 0.0], [0.0, np.inf, 0.0],      # skipped
                      [0.2, 0.2, 0.8]])                            # voxel (0,0,1)
    out = orc.voxel_grid(pts, leaf)
    # ascending linear index, x fastest: (0,0,0), (1,0,0), (0,1,0), (0,0,1)
    want = np.float64([[0.2, 0.15, 0.25], [0.6, 0.1, 0.1], [0.7 / 3, 2.2 / 3, 0.2], [0.2, 0.2, 0.8]])
    assert out.shape == (4, 4) and np.all(out[:, 3] == 1.0)
    assert np.abs(out[:, :3] - want).max() < 1e-6
    assert orc.voxel_grid(np.zeros((0, 3), np.float32)).shape == (0, 4)
    assert orc.voxel_grid(np.full((5, 3), np.nan, np.float32)).shape == (0, 4)
    # negative coordinates: floor, not truncation
    out = orc.voxel_grid(np.float32([[-0.1, 0, 0], [0.1, 0, 0]]), leaf)
    assert out.shape[0] == 2 and out[0, 0] < 0 < out[1, 0]


def test_oracle_voxel_grid_matches_numpy_on_lidar(orc):
    cloud = synth.raw_sweep(2, 0, n_sweeps=2)["cloud"]
    cloud[::97] = np.nan
    out = orc.voxel_grid(cloud)
    want, counts = numpy_voxel_grid(cloud)
    assert out.shape[0] == want.shape[0] and counts.sum() == np.isfinite(cloud).all(1).sum()
    # fixed point 2^-20 m per term + the final float32 rounding
    assert np.abs(out[:, :3] - want).max() < 2e-6


def test_oracle_voxel_grid_leaf_too_small_and_extent(orc):
    far = np.float32([[-30000, -30000, -30000], [30000, 30000, 30000], [1, 2, 3]])
    out = orc.voxel_grid(far)                                      # 750001^3 voxels > INT32_MAX: PCL returns the input
    assert np.array_equal(out[:, :3], far)
    with pytest.raises(RuntimeError, match="EXTENT"):
        orc.voxel_grid(np.float32([[0, 0, 0], [40000, 0, 0]]))


def test_oracle_pcl_normals_on_a_tilted_plane(orc):
    rng = np.random.default_rng(5)
    xy = rng.uniform(-3, 3, (20000, 2))
    pts = np.c_[xy, 0.5 * xy[:, 0] + 2.0].astype(np.float32)
    o = orc.prefilter(pts, viewpoint=[0, 0, 10], threads=4)
    assert o.rc == 0 and o.n_clusters == 1
    n_true = np.float64([-0.5, 0, 1]) / np.sqrt(1.25)
    assert np.abs(o.normals[:, :3] @ n_true - 1.0).max() < 2e-3    # single-pass float32 covariance: ~1e-3 rad of noise
    assert o.normals[:, 3].max() < 1e-3                            # curvature ~ 0 on a plane
    o2 = orc.prefilter(pts, viewpoint=[0, 0, -10], threads=4)
    assert np.array_equal(o2.normals[:, :3], -o.normals[:, :3])    # flipNormalTowardsViewpoint
    assert np.array_equal(o2.cloud, o.cloud)                       # the segmentation ignores the sign


def test_oracle_normals_match_independent_numpy_statement(orc):
    """pcl::NormalEstimation from its definition, float64: 30 nearest neighbours (scipy cKDTree), covariance, smallest
    eigenvector (numpy.linalg.eigh), curvature = lambda_min / trace, flipped towards the view point."""
    from scipy.spatial import cKDTree
    r = synth.raw_sweep(2, 2, n_sweeps=2)
    vp = r["origin"].astype(np.float32)
    o = orc.prefilter(r["cloud"], viewpoint=vp, threads=4)
    pts = o.sampled[:, :3].astype(np.float64)
    _, nn = cKDTree(pts).query(pts, k=30)
    nb = pts[nn]                                                   # m x 30 x 3
    d = nb - nb.mean(1, keepdims=True)
    cov = np.einsum("mki,mkj->mij", d, d) / 30.0
    w, v = np.linalg.eigh(cov)
    n = v[:, :, 0]
    n = np.where((np.einsum("mi,mi->m", vp.astype(np.float64) - pts, n) < 0)[:, None], -n, n)
    curv = w[:, 0] / w.sum(1)
    well = (w[:, 1] - w[:, 0]) > 1e-3 * w[:, 2]                    # the smallest eigenvector is well defined
    cosang = np.abs(np.einsum("mi,mi->m", n, o.normals[:, :3].astype(np.float64)))
    assert well.mean() > 0.9 and cosang[well].min() > 1.0 - 2e-5   # float32 covariance of a shifted 0.5 m neighbourhood
    assert np.abs(curv - o.normals[:, 3])[well].max() < 2e-4
    same_side = np.einsum("mi,mi->m", n, o.normals[:, :3].astype(np.float64)) > 0
    assert same_side[well].mean() > 0.9999                         # flips differ only where (vp - p) . n is ~ 0


def test_oracle_prefilter_two_planes(orc):
    pts = two_planes(np.random.default_rng(1))
    o = orc.prefilter(pts, threads=4)
    assert o.rc == 0
    big = np.bincount(o.labels[o.labels >= 0])
    # the two planes dominate; the 30-point speck (< min cluster 50) is dropped
    assert np.sort(big)[-2:].sum() > 0.95 * o.sampled.shape[0]
    assert not np.any(o.cloud[:, 0] > 19.0)
    for c in np.argsort(big)[-2:]:
        nn = np.abs(o.normals[o.labels == c, :3]).mean(0)
        assert nn.max() > 0.999                                    # one axis-aligned normal per cluster
    # output = clusters in cluster order, ascending sampled index inside
    want = np.concatenate([o.sampled[o.labels == c] for c in range(o.n_clusters)], 0)
    assert np.array_equal(o.cloud, want)


def test_oracle_prefilter_small_clouds(orc):
    rng = np.random.default_rng(2)
    assert orc.prefilter(np.zeros((0, 3), np.float32)).cloud.shape == (0, 4)
    o = orc.prefilter(rng.uniform(0, 5, (25, 3)).astype(np.float32))      # fewer points than knn_normals
    assert o.rc == 0 and o.cloud.shape == (0, 4) and np.all(o.labels == -1)
    o = orc.prefilter(rng.uniform(0, 0.5, (2000, 3)).astype(np.float32))  # a solid blob: no planar region of 50 points
    assert o.rc == 0 and o.sampled.shape[0] > 31


@pytest.mark.parametrize("case", ["vlp16", "hdl64", "two_planes_noisy"])
def test_min_label_fixed_point_equals_sequential_region_growing_lidar(orc, case):
    """The claim the CUDA path rests on (csrc/prefilter.cu header), checked on real neighbourhood graphs."""
    if case == "vlp16":
        cloud = synth.raw_sweep(2, 1, n_sweeps=2)["cloud"]
    elif case == "hdl64":
        cloud = synth.raw_sweep(3, 1)["cloud"][::3]
    else:
        cloud = two_planes(np.random.default_rng(7), 40000, noise=0.01)
    o = orc.prefilter(cloud, threads=4)
    _, knn = orc.surface_normals(o.sampled, k=30, threads=4)
    cos_thr = np.cos(np.float32(3.0 / 180.0 * np.pi), dtype=np.float32)
    label, _ = min_label_regions(o.normals, knn, 15, cos_thr)
    labels, n_clusters = labels_from_min_label(label, 50, 1000000)
    assert n_clusters == o.n_clusters
    assert np.array_equal(labels, o.labels)
    label2, n_comp, passes = condensed_min_label_regions(o.normals, knn, 15, cos_thr)
    assert np.array_equal(label2, label)
    assert n_comp < o.sampled.shape[0] // 2 and passes < 40        # planes collapse into few components


@pytest.mark.parametrize("seed", range(8))
def test_min_label_fixed_point_equals_sequential_region_growing_random_digraphs(orc, seed):
    """Adversarial: sparse random DIRECTED graphs (nothing symmetric about them), few normal directions, many curvature ties."""
    rng = np.random.default_rng(100 + seed)
    m, k, n_nb = 3000, 6, 4
    knn = rng.integers(0, m, (m, k)).astype(np.int32)
    knn[:, 0] = np.arange(m)                                       # self first, like a k-NN list
    dirs = np.float32([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.999, 0.04, 0], [0.6, 0.8, 0]])
    normals = np.zeros((m, 4), dtype=np.float32)
    normals[:, :3] = dirs[rng.integers(0, len(dirs), m)]
    normals[:, 3] = rng.integers(0, 50, m).astype(np.float32) / 200.0       # ties broken by index
    cos_thr = np.float32(0.9986295)
    seq, n_seq = orc.region_growing(normals, knn, n_nb=n_nb, min_size=1, max_size=m, cos_thr=cos_thr)
    label, _ = min_label_regions(normals, knn, n_nb, cos_thr)
    par, n_par = labels_from_min_label(label, 1, m)
    assert n_seq == n_par and np.array_equal(seq, par)
    assert np.array_equal(condensed_min_label_regions(normals, knn, n_nb, cos_thr)[0], label)
    seq, n_seq = orc.region_growing(normals, knn, n_nb=n_nb, min_size=5, max_size=60, cos_thr=cos_thr)
    par, n_par = labels_from_min_label(label, 5, 60)
    assert n_seq == n_par and np.array_equal(seq, par)


@pytest.mark.parametrize("seed", range(6))
def test_min_label_fixed_point_equals_sequential_region_growing_geometric_digraphs(orc, seed):
    """Between the two extremes above: k-NN graphs of clustered 2-D points (a mix of two-way and one-way edges, density jumps
    make many edges one-way), noisy normals around a few directions so that the smoothness test cuts the graph irregularly."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(500 + seed)
    m, k, n_nb = 4000, 12, 7
    centres = rng.uniform(0, 10, (12, 2))
    pts = np.concatenate([c + rng.normal(0, rng.uniform(0.05, 0.8), (m // 12 + 1, 2)) for c in centres])[:m]
    _, knn = cKDTree(pts).query(pts, k=k)
    knn = knn.astype(np.int32)
    base = np.float32([[1, 0, 0], [0.9995, 0.0316, 0], [0.998, 0.0632, 0], [0, 1, 0]])
    normals = np.zeros((m, 4), dtype=np.float32)
    nn = base[rng.integers(0, len(base), m)] + rng.normal(0, 0.004, (m, 3)).astype(np.float32)
    normals[:, :3] = nn / np.linalg.norm(nn, axis=1, keepdims=True)
    normals[:, 3] = (rng.integers(0, 30, m) / 100.0).astype(np.float32)
    cos_thr = np.float32(0.9986295)
    for min_size, max_size in ((1, m), (20, 500)):
        seq, n_seq = orc.region_growing(normals, knn, n_nb=n_nb, min_size=min_size, max_size=max_size, cos_thr=cos_thr)
        label, _ = min_label_regions(normals, knn, n_nb, cos_thr)
        par, n_par = labels_from_min_label(label, min_size, max_size)
        assert n_seq == n_par and np.array_equal(seq, par)
    label2, n_comp, passes = condensed_min_label_regions(normals, knn, n_nb, cos_thr)
    assert np.array_equal(label2, label) and 1 < n_comp < m


def test_prefilter_config_defaults_match_reference():
    """filteringUtils.cpp:12,22,27-34."""
    cfg = ab.default_prefilter_config()
    assert cfg.leaf_size == np.float32(0.08) and cfg.knn_normals == 30 and cfg.n_neighbours == 15
    assert cfg.min_cluster_size == 50 and cfg.max_cluster_size == 1000000
    assert cfg.smoothness_threshold == np.float32(3.0 / 180.0 * np.pi) and cfg.curvature_threshold == 1.0


# ---------------------------------------------------------------------------------------------------------------------
# CUDA path (GPU)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def pf():
    p = ab.B200Prefilter(device=0)
    yield p
    p.close()


def _assert_prefilter_parity(orc, pf, cloud, view_point=None, cfg=None):
    if cfg is not None:
        pf.cfg = cfg
    ocfg = None
    if cfg is not None:
        ocfg = orc.prefilter_default_config(**{k: getattr(cfg, k) for k, _ in cfg._fields_})
    try:
        out = pf.filter(cloud, view_point)
        sampled, normals, labels, clusters = pf.segments()
    finally:
        if cfg is not None:
            pf.cfg = ab.default_prefilter_config()
    o = orc.prefilter(cloud, cfg=ocfg, viewpoint=view_point, threads=8)
    assert o.rc == 0
    assert np.array_equal(sampled.view(np.uint32), o.sampled.view(np.uint32)), "voxel grid differs"
    if sampled.shape[0] > 30:
        assert np.array_equal(normals.view(np.uint32), o.normals.view(np.uint32)), "normals / curvature differ"
    assert np.array_equal(labels, o.labels), "segmentation differs"
    assert len(clusters) == o.n_clusters == pf.info.n_clusters
    assert np.array_equal(out.view(np.uint32), o.cloud.view(np.uint32)), "output cloud differs"
    assert pf.info.n_sampled == o.sampled.shape[0] and pf.info.n_out == o.cloud.shape[0]
    return o


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1023, 1025, 4099, 70001])
def test_voxel_grid_parity_ragged_sizes(orc, pf, n):
    rng = np.random.default_rng(n)
    pts = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    pts[rng.integers(0, n, max(1, n // 40))] = np.nan
    for leaf in (np.float32(0.08), np.float32(0.5)):
        got = pf.voxelGrid(pts, leaf)
        want = orc.voxel_grid(pts, leaf)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_voxel_grid_parity_lidar_and_edge_cases(orc, pf):
    for c in (2, 3):
        cloud = synth.raw_sweep(c, 0)["cloud"]
        got, want = pf.voxelGrid(cloud), orc.voxel_grid(cloud)
        assert got.shape[0] > 50000 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # every point lands in exactly one voxel: the centroids' count-weighted mean is the cloud's mean
        cen, counts = numpy_voxel_grid(cloud)
        assert got.shape[0] == counts.size and np.abs(got[:, :3] - cen).max() < 2e-6
    assert pf.voxelGrid(np.zeros((0, 3), np.float32)).shape == (0, 4)
    assert pf.voxelGrid(np.full((300, 3), np.nan, np.float32)).shape == (0, 4)
    far = np.float32([[-30000, -30000, -30000], [30000, 30000, 30000], [1, 2, 3]])
    assert np.array_equal(pf.voxelGrid(far)[:, :3], far)           # PCL's "leaf size too small": input returned unchanged
    with pytest.raises(ab.capi.AicpError, match="EXTENT"):
        pf.voxelGrid(np.float32([[0, 0, 0], [40000, 0, 0]]))
    # duplicates: many points in one voxel, sums exact in any order
    dup = np.repeat(np.float32([[1.01, 2.02, 3.03], [1.02, 2.03, 3.01]]), 5000, axis=0)
    assert np.array_equal(pf.voxelGrid(dup).view(np.uint32), orc.voxel_grid(dup).view(np.uint32))
    # voxels holding 1 .. 5000 distinct points, around the length where the warp takes over a voxel's run from its head thread
    # (48) and across tile boundaries (2048): the whole run is still summed exactly
    rng = np.random.default_rng(77)
    parts = []
    for v, cnt in enumerate([1, 47, 48, 49, 50, 79, 80, 81, 96, 97, 2047, 2048, 2049, 5000, 3, 640]):
        corner = np.float32([0.08 * (3 * v + 1), 0.08 * (2 * v + 1), 0.08 * (v % 5)])
        parts.append((corner + rng.uniform(0.005, 0.075, (cnt, 3))).astype(np.float32))
    blob = rng.permutation(np.concatenate(parts, 0))
    got, want = pf.voxelGrid(blob), orc.voxel_grid(blob)
    assert got.shape[0] == 16 and np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["vlp16_raw", "hdl64_raw", "c3_131072", "cube", "two_planes"])
def test_prefilter_parity(orc, pf, pair_cache, case):
    if case == "vlp16_raw":
        cloud = synth.raw_sweep(2, 0)["cloud"]
    elif case == "hdl64_raw":
        cloud = synth.raw_sweep(3, 0)["cloud"]
    elif case == "c3_131072":
        cloud = pair_cache(3)["ref"]
    elif case == "cube":
        cloud = synth.cube_cloud()
    else:
        cloud = two_planes(np.random.default_rng(1))
    o = _assert_prefilter_parity(orc, pf, cloud)
    assert o.n_clusters >= 1 and pf.info.passes >= 1 and pf.info.gpu_launches > 20


@pytest.mark.gpu
def test_prefilter_parity_viewpoint_and_configs(orc, pf):
    r = synth.raw_sweep(2, 3, n_sweeps=3)
    _assert_prefilter_parity(orc, pf, r["cloud"], view_point=r["origin"].astype(np.float32))
    _assert_prefilter_parity(orc, pf, r["cloud"], cfg=ab.default_prefilter_config(leaf_size=0.12, knn_normals=20, n_neighbours=10,
                                                                                 min_cluster_size=30, max_cluster_size=5000,
                                                                                 smoothness_threshold=np.float32(0.1)))
    # second-overload mirror: sampled cloud with normals + clusters
    cloud8, clusters = ab.regionGrowingUniformPlaneSegmentationFilter(r["cloud"], view_point=r["origin"], prefilter=pf)
    o = orc.prefilter(r["cloud"], viewpoint=r["origin"].astype(np.float32), threads=8)
    assert cloud8.shape == (o.sampled.shape[0], 8) and len(clusters) == o.n_clusters
    for c, idx in enumerate(clusters):
        assert np.array_equal(idx, np.flatnonzero(o.labels == c))


@pytest.mark.gpu
def test_prefilter_small_and_degenerate_clouds(orc, pf):
    rng = np.random.default_rng(4)
    assert pf.filter(np.zeros((0, 3), np.float32)).shape == (0, 4)
    for n in (1, 25, 31, 60, 500):
        _assert_prefilter_parity(orc, pf, rng.uniform(0, 4, (n, 3)).astype(np.float32))
    _assert_prefilter_parity(orc, pf, rng.uniform(0, 0.5, (2000, 3)).astype(np.float32))     # solid blob
    nan_cloud = two_planes(rng, 20000)
    nan_cloud[::11] = np.nan
    _assert_prefilter_parity(orc, pf, nan_cloud)
    with pytest.raises(ab.capi.AicpError, match="BAD_ARG"):
        pf.cfg = ab.default_prefilter_config(knn_normals=40)
        try:
            pf.filter(nan_cloud)
        finally:
            pf.cfg = ab.default_prefilter_config()


@pytest.mark.gpu
def test_prefilter_is_deterministic_and_order_invariant_where_pcl_is(orc, pf):
    """Repeated runs give the same bits; shuffling the INPUT leaves the voxel grid (and everything after it) unchanged,
    because the centroid sums are exact."""
    cloud = synth.raw_sweep(2, 5, n_sweeps=3)["cloud"]
    a = pf.filter(cloud)
    b = pf.filter(cloud)
    c = pf.filter(cloud[np.random.default_rng(0).permutation(cloud.shape[0])])
    assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.gpu
def test_prefiltered_cloud_feeds_registration_on_device(orc, pf, pair_cache):
    """App::setAndFilterReading -> registerClouds: the device-resident filter output goes straight into the registration."""
    pair = pair_cache(2)
    view = pf.filter(pair["read"], keep_on_device=True)
    o = orc.prefilter(pair["read"], threads=8)
    assert view.shape[0] == o.cloud.shape[0]
    reg = ab.B200Registration(device=0)
    try:
        T_dev = reg.registerClouds(pair["ref"], view)
        T_host = reg.registerClouds(pair["ref"], o.cloud)
        assert np.array_equal(T_dev, T_host)
    finally:
        reg.close()


@pytest.mark.gpu
def test_map_prefilter_on_device(orc):
    """app.cpp:476-493: merge an aligned cloud into the map, re-filter the map in place."""
    m = ab.B200Map(device=0)
    try:
        a = synth.raw_sweep(2, 6, n_sweeps=2)["cloud"]
        b = synth.raw_sweep(2, 7, n_sweeps=2)["cloud"]
        m.updateCloud(a)
        m.append(b)
        info = m.prefilter()
        o = orc.prefilter(np.concatenate([a, b], 0), threads=8)
        assert m.size() == o.cloud.shape[0] == info.n_out and info.n_clusters == o.n_clusters
        assert np.array_equal(m.download().view(np.uint32), o.cloud.view(np.uint32))
    finally:
        m.close()


@pytest.mark.gpu
def test_voxel_grid_and_prefilter_at_map_scale(orc, pf):
    """2 M-point campus map (the C4 map generator at reduced size): parity with the oracle at a size where tiles, scans and the
    sort run many blocks deep, plus properties that hold at any size."""
    case = synth.make_map_case(n_map=2 * 1024 * 1024, n_read=1024, trial=1, n_poses=1)
    cloud = case["map"]
    got = pf.voxelGrid(cloud)
    want = orc.voxel_grid(cloud)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    cen, counts = numpy_voxel_grid(cloud)
    assert counts.sum() == cloud.shape[0] and got.shape[0] == counts.size
    out = pf.filter(cloud)
    o = orc.prefilter(cloud, threads=8)
    assert np.array_equal(out.view(np.uint32), o.cloud.view(np.uint32))
    sampled, normals, labels, clusters = pf.segments()
    assert all(len(c) >= 50 for c in clusters) and sum(len(c) for c in clusters) == out.shape[0]
    assert np.all(np.abs(np.linalg.norm(normals[:, :3], axis=1) - 1.0) < 1e-5)
