"""Golden vectors (tests/golden/icp_goldens.json, frozen by tests/golden/make_goldens.py).
  not gpu : the oracle still reproduces every vector (the checker does not drift) and the seeded inputs are unchanged
  gpu     : the CUDA path, through the C ABI, reproduces every vector WITHOUT the oracle running."""
import json
import os
import zlib

import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "icp_goldens.json")) as f:
    GOLD = json.load(f)
_INPUTS = {}


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def bits(a):
    return [int(x) for x in np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel()]


def inputs(name):
    if name not in _INPUTS:
        spec = GOLD[name]["spec"]
        _INPUTS[name] = synth.c1_pair(int(spec.split(":")[1])) if isinstance(spec, str) else synth.make_pair(*spec)
    p = _INPUTS[name]
    assert [crc(p["ref"]), crc(p["read"])] == GOLD[name]["input_crc"], "the synthetic generators changed: regenerate the goldens"
    return p


def check(g, T, iterations, stop_reason, trace, trace_idx, normals, reading, wpur):
    assert iterations == g["iterations"] and stop_reason == g["stop_reason"]
    assert bits(T) == g["T_bits"]
    assert [bits([t["limit_d2"]])[0] for t in trace] == g["limit_bits"]
    assert [int(t["n_used"]) for t in trace] == g["n_used"]
    assert [crc(row) for row in trace_idx] == g["match_crc"]
    assert crc(normals) == g["normals_crc"] and crc(reading) == g["reading_crc"]
    assert bits([wpur])[0] == g["weighted_point_used_ratio_bits"]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden(orc, name):
    g, p = GOLD[name], inputs(name)
    ov, counts = orc.overlap(p["ref"], p["ref_origin"], p["read"], p["read_origin"])
    assert bits([ov])[0] == g["overlap_bits"] and list(counts) == g["overlap_counts"]
    assert bits([orc.autotune_ratio(float(ov))[0]])[0] == g["autotuned_ratio_bits"]
    o = orc.icp(p["ref"], p["read"], orc.default_config(ratio=g["ratio"], threads=os.cpu_count() or 1, **g["config"]),
                want_trace_idx=True, want_normals=True)
    assert o.rc == 0
    check(g, o.T, o.iterations, o.stop_reason, o.trace, o.trace_idx, o.normals, o.reading, o.weighted_point_used_ratio)


@pytest.mark.gpu
@pytest.mark.parametrize("schedules", [(1, 1, 2), (2, 2, 2), (2, 2, 1)])      # (match, k-NN, loop): loop 2 persistent kernel, 1 multi-launch
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_reproduces_golden(name, schedules):
    g, p = GOLD[name], inputs(name)
    ovl = ab.B200Overlap()
    counts = ovl.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
    assert bits([ovl.getOverlap()])[0] == g["overlap_bits"] and list(counts) == g["overlap_counts"]
    assert bits([ab.autotune_ratio(float(ovl.getOverlap()))])[0] == g["autotuned_ratio_bits"]
    ovl.close()
    reg = ab.B200Registration()
    reg.setConfig(ratio=g["ratio"], **g["config"])
    reg.setMatchSchedule(schedules[0]); reg.setKnnSchedule(schedules[1]); reg.setLoopSchedule(schedules[2])
    reg.enableMatchTrace(True)
    T = reg.registerClouds(p["ref"], p["read"])
    check(g, T, reg.stats.iterations, reg.stats.stop_reason, reg.trace(), reg.getTraceMatches(), reg.getReferenceNormals(),
          reg.getOutputReading(), reg.getWeightedPointUsedRatio())
    reg.close()
