"""CPU tests of the host-side logic and of the C-ABI boundary (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import capi
from aicp_mapping_b200.build import build

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build()
    return capi.lib()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "aicp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(aicp_b200_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.aicp_b200_version()


def test_struct_layouts_match_header():
    # sizes implied by include/aicp_b200.h on LP64
    assert C.sizeof(capi.IcpConfig) == 32
    assert C.sizeof(capi.IterTrace) == 104
    assert C.sizeof(capi.Stats) == 96 + 104 * capi.MAX_ITERS   # 92 bytes of scalars, padded to the 8-byte alignment of the trace records


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """Every struct of include/aicp_b200.h that crosses the ABI: sizeof and the offset of every field, as gcc lays them out,
    against the ctypes mirrors in capi.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    structs = {"aicp_b200_icp_config": capi.IcpConfig, "aicp_b200_iter_trace": capi.IterTrace, "aicp_b200_stats": capi.Stats,
               "aicp_b200_prefilter_config": capi.PrefilterConfig, "aicp_b200_prefilter_info": capi.PrefilterInfo,
               "aicp_b200_svm_summary": capi.SvmSummary, "aicp_b200_append_info": capi.AppendInfo}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "aicp_b200.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for ln in out.splitlines():
        cname, what, val = ln.split()
        ct = structs[cname]
        if what == "sizeof":
            assert C.sizeof(ct) == int(val), cname
        else:
            assert getattr(ct, what).offset == int(val), (cname, what)
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.AicpError, match="no CPU fallback"):
        ab.B200Registration()


def _parse(lib, path):
    cfg = capi.IcpConfig()
    err = C.create_string_buffer(512)
    rc = lib.aicp_b200_parse_icp_yaml(path.encode() if path else None, C.byref(cfg), err, 512)
    return rc, cfg, err.value.decode()


def test_parse_shipped_chain_files_verbatim(lib):
    rc, cfg, err = _parse(lib, os.path.join(GOLDEN, "icp_autotuned.yaml"))
    assert rc == 0, err
    assert cfg.knn_normals == 20 and cfg.max_iterations == 20 and cfg.smooth_length == 4
    assert cfg.ratio == np.float32(0.358818)                          # icp_autotuned.yaml:35
    assert cfg.min_diff_rot == np.float32(0.001) and cfg.min_diff_trans == np.float32(0.01)
    assert cfg.matcher_epsilon == np.float32(3.16)                    # recorded; the GPU search is always exact
    rc, cfg, err = _parse(lib, os.path.join(GOLDEN, "icp_autotuned_default.yaml"))
    assert rc == 0 and cfg.ratio == np.float32(0.70)
    rc, cfg, err = _parse(lib, None)                                  # empty path -> chain defaults
    assert rc == 0 and cfg.knn_normals == 20 and cfg.ratio == np.float32(0.70)


def test_parse_rejects_unsupported_modules_loudly(lib, tmp_path):
    rc, cfg, err = _parse(lib, os.path.join(GOLDEN, "icp_3D_cfg_trimmed.yaml"))
    assert rc == 7 and "MaxDensityDataPointsFilter" in err
    rc, cfg, err = _parse(lib, str(tmp_path / "missing.yaml"))
    assert rc == 7 and "Cannot open config file" in err              # pointmatcher_registration.cpp:60-64 (without exit(1))
    base = open(os.path.join(GOLDEN, "icp_autotuned_default.yaml")).read()
    cases = {"PointToPlaneErrorMinimizer:": ("PointToPointErrorMinimizer:", "PointToPlaneErrorMinimizer"),
             "    knn: 1": ("    knn: 3", "knn: 1"),
             "- TrimmedDistOutlierFilter:": ("- MaxDistOutlierFilter:", "MaxDistOutlierFilter"),
             "    #maxDist: 0.25": ("    maxDist: 0.25", "maxDist")}
    for old, (new, needle) in cases.items():
        assert old in base
        p = tmp_path / "bad.yaml"
        p.write_text(base.replace(old, new, 1) if old != "PointToPlaneErrorMinimizer:" else
                     base.replace("  PointToPlaneErrorMinimizer:", "  PointToPointErrorMinimizer:"))
        rc, cfg, err = _parse(lib, str(p))
        assert rc == 7 and needle in err, (old, err)


def test_replace_ratio_config_file_is_the_reference_rewrite(lib, tmp_path):
    src = os.path.join(GOLDEN, "icp_autotuned_default.yaml")
    dst = str(tmp_path / "icp_autotuned.yaml")
    ab.replaceRatioConfigFile(src, dst, np.float32(35.8818 / 100.0))
    out = open(dst).read()
    # the committed icp_autotuned.yaml is literally a leftover of this rewrite (plus the extra newline per rewrite)
    assert "      ratio: 0.358818\n" in out
    assert out.rstrip("\n") == open(os.path.join(GOLDEN, "icp_autotuned.yaml")).read().rstrip("\n")
    rc, cfg, err = _parse(lib, dst)
    assert rc == 0 and cfg.ratio == np.float32(0.358818)
    # 11 characters are replaced: a shorter print leaves no residue of "0.70"
    ab.replaceRatioConfigFile(src, dst, np.float32(0.25))
    assert "      ratio: 0.25\n" in open(dst).read()


def test_autotune_ratio_clamp_and_text_roundtrip(lib):
    assert ab.autotune_ratio(35.8818) == float(np.float32(0.358818))
    assert ab.autotune_ratio(5.0) == 0.25 and ab.autotune_ratio(99.0) == float(np.float32(0.7))
    assert ab.autotune_ratio(50.0) == 0.5


def test_parse_transformation_deg():
    T = ab.parseTransformationDeg("[0.5, -0.25; 90]")
    assert np.allclose(T[:3, :3], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-7) and T[0, 3] == 0.5 and T[1, 3] == -0.25
    assert np.array_equal(ab.parseTransformationDeg("garbage"), np.eye(4, dtype=np.float32))


def test_factories_reject_unknown_types(capsys):
    assert ab.create_registrator(ab.RegistrationParams(type="GICP")) is None
    assert "Invalid registration type GICP." in capsys.readouterr().err
    assert ab.create_overlapper(ab.OverlapParams(type="Nope")) is None


def test_synthetic_workloads_are_seeded_and_sized(pair_cache):
    from aicp_mapping_b200 import synth
    a, b = synth.make_pair(5, 1, 2000), synth.make_pair(5, 1, 2000)
    assert np.array_equal(a["read"], b["read"]) and a["ref"].shape == (2000, 3)
    c2 = pair_cache(2, 0, 8192)
    assert c2["ref"].shape == (8192, 3) and c2["read"].shape == (8192, 3) and c2["ref"].dtype == np.float32
    c1 = synth.c1_pair(1)
    assert c1["ref"].shape[0] >= 34592 and c1["read"].shape[1] == 3
    assert synth.cube_cloud().shape == (6 * 81 * 81, 3)      # create_cube_cloud.cpp's float loop yields 81 ticks per axis
