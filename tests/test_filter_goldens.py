"""Golden vectors of the widened rows (tests/golden/filter_goldens.json, frozen by tests/golden/make_filter_goldens.py).
  not gpu : the oracle still reproduces every vector and the seeded inputs are unchanged
  gpu     : the CUDA path, through the C ABI, reproduces every vector WITHOUT the oracle running."""
import json
import os
import sys

import numpy as np
import pytest

import aicp_mapping_b200 as ab

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_filter_goldens as mk  # noqa: E402

with open(os.path.join(HERE, "golden", "filter_goldens.json")) as f:
    GOLD = json.load(f)


@pytest.mark.parametrize("name", sorted(GOLD["prefilter"]))
def test_oracle_reproduces_prefilter_golden(orc, name):
    g = GOLD["prefilter"][name]
    cloud, vp = mk.prefilter_inputs(name)
    assert mk.crc(cloud) == g["input_crc"], "the synthetic generators changed: regenerate the goldens"
    o = orc.prefilter(cloud, viewpoint=vp, threads=os.cpu_count() or 1)
    assert (o.sampled.shape[0], o.n_clusters, o.cloud.shape[0]) == (g["n_sampled"], g["n_clusters"], g["n_out"])
    assert (mk.crc(o.sampled), mk.crc(o.normals), mk.crc(o.labels), mk.crc(o.cloud)) == (g["sampled_crc"], g["normals_crc"], g["labels_crc"], g["cloud_crc"])


@pytest.mark.parametrize("name", sorted(GOLD["pairs"]))
def test_oracle_reproduces_alignability_golden(orc, name):
    g = GOLD["pairs"][name]
    a, b, PA, PB, rng_m, view = mk.pair_inputs(name)
    assert [mk.crc(a), mk.crc(b)] == g["input_crc"], "the synthetic generators changed: regenerate the goldens"
    ov, fa, fb = orc.fov_overlap(a, b, PA, PB, rng_m, view)
    assert mk.fbits(ov) == g["fov_overlap_bits"] and [mk.crc(fa), mk.crc(fb)] == g["accepted_crc"]
    al, matching, info = orc.alignability(fa, fb, PA, PB, threads=os.cpu_count() or 1)
    assert mk.fbits(al) == g["alignability_bits"] and [int(m) for m in matching] == g["matching"] and list(info) == g["info"]


def test_oracle_reproduces_accumulation_golden(orc):
    g = GOLD["accumulate"]["vlp16_4sweeps"]
    sweeps, poses = mk.sweep_inputs()
    assert [mk.crc(s) for s in sweeps] == g["input_crc"]
    acc = orc.accumulate_sweeps(sweeps, poses)
    assert acc.shape[0] == g["n"] and mk.crc(acc) == g["cloud_crc"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["prefilter"]))
def test_gpu_reproduces_prefilter_golden(name):
    g = GOLD["prefilter"][name]
    cloud, vp = mk.prefilter_inputs(name)
    pf = ab.B200Prefilter(device=0)
    try:
        out = pf.filter(cloud, vp)
        sampled, normals, labels, clusters = pf.segments()
        assert (sampled.shape[0], len(clusters), out.shape[0]) == (g["n_sampled"], g["n_clusters"], g["n_out"])
        assert (mk.crc(sampled), mk.crc(normals), mk.crc(labels), mk.crc(out)) == (g["sampled_crc"], g["normals_crc"], g["labels_crc"], g["cloud_crc"])
    finally:
        pf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["pairs"]))
def test_gpu_reproduces_alignability_golden(name):
    g = GOLD["pairs"][name]
    a, b, PA, PB, rng_m, view = mk.pair_inputs(name)
    al = ab.B200Alignability(device=0)
    try:
        ov, fa, fb = al.overlapFilter(a, b, PA, PB, rng_m, view)
        assert mk.fbits(ov) == g["fov_overlap_bits"] and [mk.crc(fa), mk.crc(fb)] == g["accepted_crc"]
        ali, matching, info = al.alignabilityFilter(fa, fb, PA, PB)
        assert mk.fbits(ali) == g["alignability_bits"] and [int(m) for m in matching] == g["matching"] and list(info) == g["info"]
    finally:
        al.close()


@pytest.mark.gpu
def test_gpu_reproduces_accumulation_golden():
    g = GOLD["accumulate"]["vlp16_4sweeps"]
    sweeps, poses = mk.sweep_inputs()
    acc = ab.B200VelodyneAccumulator(batch_size=4, device=0)
    try:
        for s, P in zip(sweeps, poses):
            acc.processLidar(s, P)
        got = acc.download()
        assert got.shape[0] == g["n"] and mk.crc(got) == g["cloud_crc"]
    finally:
        acc.close()
