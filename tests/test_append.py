"""Incremental insert into the device-resident reference (aicp_b200_reference_append, csrc/append.cu):
after any sequence of appends the handle must be in the state a fresh set_reference(all points) + rebuild gives -- same
normals, same correspondences, same transform, bit for bit -- whether the append took the incremental path (new points
inside the old bounding box) or the fallback (outside: full rebuild).  The registration against the oracle on the union is
the parity anchor."""
import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def split_with_extremes_first(cloud, n_base, rng):
    """A random base subset that contains the extreme point of every axis (so the rest lies inside its bounding box)."""
    ext = set(int(i) for i in np.r_[cloud.argmin(0), cloud.argmax(0)])
    rest = np.array([i for i in rng.permutation(len(cloud)) if i not in ext])
    order = np.r_[np.array(sorted(ext)), rest]
    return cloud[order[:n_base]], cloud[order[n_base:]]


@pytest.mark.parametrize("config,n,knn", [(3, 60000, 20), (2, 30000, 10)])
def test_appends_equal_a_full_rebuild(pair_cache, orc, config, n, knn):
    pair = pair_cache(config, 2, n)
    rng = np.random.default_rng(5)
    base, extra = split_with_extremes_first(pair["ref"], int(0.7 * n), rng)
    batches = np.array_split(extra, 3)
    inc = ab.B200Registration()
    inc.setConfig(ratio=0.6, knn_normals=knn)
    inc.setReference(base)
    inc.registerToReference(pair["read"])                      # builds the index and the normals of the base
    for b in batches[:2]:
        info = inc.appendToReference(b)
        assert info.incremental == 1 and 0 < info.n_recomputed < info.n_total and info.ms > 0
        assert info.n_recomputed >= len(b)
    # a registration in between must not disturb later appends
    inc.registerToReference(pair["read"])
    info = inc.appendToReference(batches[2])
    assert info.incremental == 1 and info.n_total == n
    inc.enableMatchTrace(True)
    T_inc = inc.registerToReference(pair["read"])
    union = np.concatenate([base] + batches, 0)
    full = ab.B200Registration()
    full.setConfig(ratio=0.6, knn_normals=knn)
    full.setReference(union)
    full.enableMatchTrace(True)
    T_full = full.registerToReference(pair["read"])
    assert np.array_equal(u32(T_inc), u32(T_full)) and inc.stats.iterations == full.stats.iterations
    assert np.array_equal(inc.getTraceMatches(), full.getTraceMatches())
    assert np.array_equal(u32(inc.getReferenceNormals()), u32(full.getReferenceNormals()))
    assert np.array_equal(u32(inc.getOutputReading()), u32(full.getOutputReading()))
    o = orc.icp(union, pair["read"], orc.default_config(ratio=0.6, knn_normals=knn, threads=8), want_normals=True)
    assert o.rc == 0 and np.array_equal(u32(o.T), u32(T_inc)) and np.array_equal(u32(o.normals), u32(inc.getReferenceNormals()))
    # the fallback: a point outside the bounding box changes the Morton quantisation of every point
    far = (union.max(0) + np.float32([3.0, 1.0, 0.5]))[None, :].astype(np.float32)
    info = inc.appendToReference(far)
    assert info.incremental == 0 and info.n_total == n + 1
    full.setReference(np.concatenate([union, far], 0))
    assert np.array_equal(u32(inc.registerToReference(pair["read"])), u32(full.registerToReference(pair["read"])))
    assert np.array_equal(u32(inc.getReferenceNormals()), u32(full.getReferenceNormals()))
    # ... and an append before any registration only grows the stored cloud
    fresh = ab.B200Registration()
    fresh.setConfig(ratio=0.6, knn_normals=knn)
    fresh.setReference(base)
    assert fresh.appendToReference(extra).incremental == 0
    assert np.array_equal(u32(fresh.registerToReference(pair["read"])), u32(T_full))
    for r in (inc, full, fresh):
        r.close()


def test_append_duplicates_and_tiny_batches(orc):
    """Appending points that already exist (equal Morton keys, exact distance ties) and single points."""
    rng = np.random.default_rng(11)
    base = np.round(rng.uniform(-4, 4, (5000, 3)) * 8).astype(np.float32) / 8        # lattice: many ties
    base[:, 2] = np.float32(0.25) * base[:, 0] + np.float32(0.02) * rng.normal(size=5000).astype(np.float32)
    base[0] = base.min(0) - 1; base[1] = base.max(0) + 1                              # roomy bounding box
    read = (base[rng.integers(0, 5000, 3000)] + np.float32([0.05, -0.03, 0.02])).astype(np.float32)
    inc = ab.B200Registration()
    inc.setConfig(ratio=0.8, knn_normals=12)
    inc.setReference(base)
    inc.registerToReference(read)
    pieces = [base[100:140].copy(), base[7:8].copy(), (base[200:260] + np.float32(1e-3)).astype(np.float32)]
    for p in pieces:
        assert inc.appendToReference(p).incremental == 1
    union = np.concatenate([base] + pieces, 0)
    T_inc = inc.registerToReference(read)
    o = orc.icp(union, read, orc.default_config(ratio=0.8, knn_normals=12, threads=8), want_normals=True)
    assert o.rc == 0 and np.array_equal(u32(o.T), u32(T_inc))
    assert np.array_equal(u32(o.normals), u32(inc.getReferenceNormals()))
    inc.close()
