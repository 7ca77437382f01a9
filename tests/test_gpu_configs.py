"""GPU tests for the BASELINE.json configurations beyond the headline pair: C4 (localisation against a fixed map) and C5
(batched validation sweep).  Reduced sizes are compared bit for bit with the oracle; BASELINE's full sizes are checked
through size-independent properties (exact NN against a brute-force sample, decrease of the trimmed objective, determinism,
idempotence of the output cloud, batch == sequential)."""
import os

import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import capi, synth
from conftest import rot_angle

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def reg():
    r = ab.B200Registration()
    yield r
    r.close()


def test_c4_reduced_map_parity_with_oracle(reg, orc):
    """Fixed map (400k points) + two readings (30k): set_reference once, register twice; every reading's trajectory is
    bit-identical to the oracle's full ICP on (map, reading)."""
    case = synth.make_map_case(n_map=400_000, n_read=30_000, trial=1, n_poses=2)
    reg.setConfig(ratio=0.5, max_iterations=20)          # app.cpp:123-127: overlap forced to 50 % against a prior map
    reg.setReference(case["map"])
    reg.enableMatchTrace(True)
    for rd in case["readings"]:
        T = reg.registerToReference(rd["read"])
        o = orc.icp(case["map"], rd["read"], orc.default_config(ratio=0.5, threads=8), want_trace_idx=True)
        assert o.rc == 0 and reg.stats.iterations == o.iterations
        assert np.array_equal(reg.getTraceMatches(), o.trace_idx)
        assert np.array_equal(u32(T), u32(o.T))
        # and the registration undoes the injected prior error (the map is noisy: centimetres, not micrometres)
        d = T.astype(np.float64) @ np.linalg.inv(rd["T_true"])
        assert np.linalg.norm(d[:3, 3]) < 0.05 and rot_angle(d[:3, :3]) < 0.01
    reg.enableMatchTrace(False)


def test_c4_full_size_map_properties(reg, orc):
    """BASELINE config 4 at full size: 122 880-point reading against a 10 485 760-point map."""
    # the campus WITH street clutter (1500 small boxes): with buildings and ground alone the better half of the matches -- all a
    # trimmed ratio of 0.5 keeps -- lies on surfaces parallel to the boulevard and the registration slides along it by
    # decimetres, in the oracle exactly as on the GPU (round 1 accepted < 1 m here)
    # Trimmed ICP at ratio 0.5 from a 0.2 - 0.4 m prior error is bimodal on this scene (tools/c4_scene_probe.py: sub-millimetre or
    # stuck decimetres away, for the oracle as for the GPU); trial 1 is a scene whose three poses all converge.
    case = synth.make_map_case(n_map=10_485_760, n_read=122_880, trial=1, n_poses=1, n_clutter=1500)
    mp, rd = case["map"], case["readings"][0]
    # (1) exact NN on the full map, ALL 122 880 queries: indices and squared distances bit for bit against the oracle's
    #     kd-tree search (5 s on 8 cores), which is itself pinned by a float32 brute force with the oracle's operation order
    #     on a sample of the queries
    rng = np.random.default_rng(7)
    idx, d2 = reg.match(mp, rd["read"])
    o_idx, o_d2 = orc.match(mp, rd["read"], use_kdtree=True, threads=NCPU)
    assert np.array_equal(idx, o_idx) and np.array_equal(u32(d2), u32(o_d2))
    for k in rng.choice(len(rd["read"]), 64, replace=False):
        df = mp - rd["read"][k][None, :]
        dd = (df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1]) + df[:, 2] * df[:, 2]
        j = int(np.argmin(dd))                       # first minimum == lowest index on ties
        assert o_idx[k] == j and np.float32(o_d2[k]) == dd[j]
    # (2) localisation lowers the trimmed point-to-map objective and RECOVERS THE TRUE POSE: the injected prior error
    #     (decimetres, degrees) is undone to within 5 cm / 0.01 rad of the noisy map
    reg.setConfig(ratio=0.5, max_iterations=20)
    reg.setReference(mp)
    T = reg.registerToReference(rd["read"])
    it = reg.stats.iterations
    aligned = reg.getOutputReading()[:, :3]
    probe = rng.choice(len(aligned), 4096, replace=False)

    def trimmed(cloud):
        _, dd = reg.match(mp, cloud[probe])
        return float(np.sort(dd)[:len(dd) // 2].mean())
    before, after = trimmed(rd["read"]), trimmed(aligned)
    assert after < 0.5 * before, (before, after)
    d = T.astype(np.float64) @ np.linalg.inv(rd["T_true"])
    assert np.linalg.norm(d[:3, 3]) < 0.05 and rot_angle(d[:3, :3]) < 0.01, (np.linalg.norm(d[:3, 3]), rot_angle(d[:3, :3]))
    reg.setReference(mp)                                   # reg.match used the scratch index; the map index is kept
    T = reg.registerToReference(rd["read"])
    # (3) deterministic and reference reuse is stateless: the same call again gives the same bits
    T2 = reg.registerToReference(rd["read"])
    assert np.array_equal(u32(T), u32(T2)) and reg.stats.iterations == it
    # (4) idempotence: registering the aligned output is (nearly) a no-op
    out = reg.getOutputReading()[:, :3]
    T3 = reg.registerToReference(out)
    assert np.linalg.norm(T3[:3, 3]) < 0.02 and rot_angle(T3[:3, :3]) < 0.005


def test_c5_batch_of_perturbed_pairs(reg, orc):
    """BASELINE config 5 (scaled to 48 pairs of 38 400 points): the concurrent batch equals one-by-one calls bit for bit,
    three of the pairs are checked against the oracle, and every pair recovers its perturbation."""
    ov = ab.B200Overlap()
    pairs = [synth.make_pair(5, t) for t in range(48)]
    ratios = []
    for p in pairs:
        ov.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
        ratios.append(ab.autotune_ratio(float(ov.getOverlap())))
    ov.close()
    reg.setConfig(max_iterations=20)
    T, stats, status, ms = reg.registerBatch([(p["ref"], p["read"]) for p in pairs], ratios=ratios, streams=8)
    assert all(s == 0 for s in status) and ms > 0
    for t in (0, 17, 47):
        reg.setConfig(ratio=ratios[t], max_iterations=20)
        Ts = reg.registerClouds(pairs[t]["ref"], pairs[t]["read"])
        assert np.array_equal(u32(Ts), u32(T[t]))
        o = orc.icp(pairs[t]["ref"], pairs[t]["read"], orc.default_config(ratio=ratios[t], threads=8))
        assert o.rc == 0 and np.array_equal(u32(o.T), u32(T[t])) and o.iterations == stats[t].iterations
    for p, Tp in zip(pairs, T):
        d = Tp.astype(np.float64) @ np.linalg.inv(p["T_true"])
        assert np.linalg.norm(d[:3, 3]) < 0.02 and rot_angle(d[:3, :3]) < 0.01


def test_aicp_batch_whole_step_equals_separate_calls(reg, orc):
    """aicp_b200_aicp_batch (overlap -> auto-tuned ratio -> registration per pair, concurrently) against the same steps made
    one by one, and one pair against the oracle's overlap + registration."""
    ov = ab.B200Overlap()
    pairs = [synth.make_pair(5, t, 6000 + 700 * t) for t in range(6)] + [synth.make_pair(2, 0, 8192)]
    reg.setConfig(max_iterations=20)
    T, overlap, stats, status, ms = reg.aicpBatch([(p["ref"], p["read"]) for p in pairs],
                                                 [(p["ref_origin"], p["read_origin"]) for p in pairs], streams=4)
    assert all(s == 0 for s in status) and ms > 0
    for i, p in enumerate(pairs):
        ov.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
        assert np.float32(ov.getOverlap()) == overlap[i]
        reg.setConfig(ratio=ab.autotune_ratio(float(ov.getOverlap())), max_iterations=20)
        Ts = reg.registerClouds(p["ref"], p["read"])
        assert np.array_equal(u32(Ts), u32(T[i])) and reg.stats.iterations == stats[i].iterations
    p = pairs[-1]
    o_ov, _ = orc.overlap(p["ref"], p["ref_origin"], p["read"], p["read_origin"])
    o = orc.icp(p["ref"], p["read"], orc.default_config(ratio=float(orc.autotune_ratio(float(o_ov))[0]), threads=8))
    assert np.float32(o_ov) == overlap[-1] and np.array_equal(u32(o.T), u32(T[-1]))
    ov.close()


def test_register_batch_over_all_devices(reg):
    """aicp_b200_register_batch_devices (pair i on GPU devices[i % G], one worker pool per GPU, one process) against the
    single-device batch: same transforms, iteration counts and per-pair status bit for bit, for host and for device inputs,
    with a failing pair in the middle.  Runs on every GPU the box has (one on the driver's test box, eight on the scaling
    box)."""
    import torch
    G = min(torch.cuda.device_count(), 8)
    pairs = [synth.make_pair(5, t, 5000 + 300 * t) for t in range(13)]
    clouds = [(p["ref"], p["read"]) for p in pairs]
    bad = clouds[5][1].copy(); bad[3, 1] = np.inf
    clouds[5] = (clouds[5][0], bad)                                      # NONFINITE_INPUT in pair 5 only
    ratios = [0.55 + 0.01 * t for t in range(13)]
    reg.setConfig(max_iterations=20)

    def run(devices, inputs):
        try:
            return reg.registerBatch(inputs, ratios=ratios, streams=3, devices=devices)
        except ab.capi.AicpError as e:                                   # the call reports the first failing pair ...
            assert e.code_name == "NONFINITE_INPUT"
            return None
    assert run(None, clouds) is None and run(list(range(G)), clouds) is None
    clouds[5] = (pairs[5]["ref"], pairs[5]["read"])                      # ... and with it repaired everything agrees
    T1, s1, st1, ms1 = reg.registerBatch(clouds, ratios=ratios, streams=3)
    TG, sG, stG, msG = reg.registerBatch(clouds, ratios=ratios, streams=3, devices=list(range(G)))
    assert np.array_equal(u32(T1), u32(TG)) and list(st1) == list(stG) == [0] * 13 and msG > 0
    assert [s.iterations for s in s1] == [s.iterations for s in sG]
    dev_in = [(torch.from_numpy(ab.capi.to_xyzw(r)).cuda(), torch.from_numpy(ab.capi.to_xyzw(q)).cuda()) for r, q in clouds]
    TD, _, stD, _ = reg.registerBatch(dev_in, ratios=ratios, streams=3, devices=list(range(G))[::-1])
    assert np.array_equal(u32(T1), u32(TD)) and list(stD) == [0] * 13
