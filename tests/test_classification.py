"""The alignment-risk classifier: aicp::SVM::load / test (aicp_core/src/classification/svm.cpp:53-107), SURVEY.md 8(f) rank 2.
PARITY PINNED: the reference ships its own inputs and outputs for this function (data/labels/testing_labelled_27Aug.txt ->
data/classification/probs_opencv3.txt), frozen with OpenCV's own cv2.ml.SVM outputs in tests/golden/svm_goldens.npz by
tests/golden/make_svm_goldens.py.
not gpu: the numpy oracle against those goldens; the library's model reader against the oracle's on every shipped model file.
gpu    : the CUDA path through the C ABI against the reference's golden file, cv2's raw decision values and the oracle."""
import glob
import os

import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import classification

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = sorted(glob.glob(os.path.join(GOLDEN, "svm_models", "*.xml")))
DEFAULT_MODEL = os.path.join(GOLDEN, "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")   # aicp.launch:19
# the golden file holds 6 significant digits of a probability near 0.5: half a unit of the last digit, plus the float32
# pow difference between OpenCV's cv::pow and a float64 pow rounded to float32 (1e-7 on the decision value)
TOL_PROB = 1e-6


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "svm_goldens.npz"))


@pytest.fixture(scope="module")
def svm_orc():
    from oracle import aicp_oracle_svm
    return aicp_oracle_svm


def test_oracle_reproduces_the_references_own_probabilities(svm_orc, gold):
    m = svm_orc.load_model(DEFAULT_MODEL)
    p = svm_orc.test(m, gold["testing_features"])
    assert p.shape == (269,)
    assert np.abs(p - gold["reference_probs_opencv3"]).max() <= TOL_PROB
    # the reference's six digits, reproduced as text for nearly every sample (a last-digit flip needs |diff| ~ 5e-7)
    same_text = sum(("%g" % a) == ("%g" % b) for a, b in zip(p.astype(np.float32), gold["reference_probs_opencv3"]))
    assert same_text >= 260


@pytest.mark.parametrize("model", MODELS, ids=[os.path.basename(m)[:-4] for m in MODELS])
def test_oracle_matches_opencv_raw_decision_values(svm_orc, gold, model):
    m = svm_orc.load_model(model)
    x = np.concatenate([gold["testing_features"], gold["grid_features"]], 0)
    raw, scale = svm_orc.predict_raw(m, x, return_scale=True)
    want = gold["cv2_raw_" + os.path.basename(model)[:-4]]
    # K is rounded to float32 before the float64 sum: judge the difference against sum |alpha_i K_i| (the 13-vector models
    # have K ~ 1e12 and alpha ~ 1e-10, cancelling to O(1)); 3 float32 ulps of that, plus the float32 rounding of the result
    assert (np.abs(raw.astype(np.float64) - want) / (scale + np.abs(want))).max() <= 4e-7


@pytest.mark.parametrize("model", MODELS, ids=[os.path.basename(m)[:-4] for m in MODELS])
def test_library_model_reader_matches_oracle_reader(svm_orc, model):
    s = classification.parse_model(model)
    m = svm_orc.load_model(model)
    assert (s.kernel, s.dim, s.sv_total, s.sv_count) == (m["kernel"], m["dim"], m["sv"].shape[0], len(m["alpha"]))
    assert (s.degree, s.gamma, s.coef0, s.rho) == (m["degree"], m["gamma"], m["coef0"], m["rho"])
    assert s.alpha_sum == float(np.cumsum(m["alpha"])[-1]) and s.sv_sum == float(np.cumsum(m["sv"].ravel().astype(np.float64))[-1])
    assert (s.index_first, s.index_last) == (int(m["index"][0]), int(m["index"][-1]))


def test_library_model_reader_rejects_what_it_does_not_implement(tmp_path):
    text = open(DEFAULT_MODEL).read()
    bad = tmp_path / "rbf.xml"
    bad.write_text(text.replace("<type>POLY</type>", "<type>RBF</type>"))
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        classification.parse_model(str(bad))
    trunc = tmp_path / "trunc.xml"
    trunc.write_text(text[:4000])
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        classification.parse_model(str(trunc))
    with pytest.raises(ab.capi.AicpError, match="CONFIG"):
        classification.parse_model(str(tmp_path / "missing.xml"))


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_reproduces_the_references_own_probabilities(gold):
    svm = ab.create_classifier(ab.ClassificationParams(type="SVM", svm=ab.SVMParams(threshold=0.5, saveFile=DEFAULT_MODEL)), device=0)
    try:
        p = svm.test(gold["testing_features"], labels=gold["testing_labels"])
        assert np.abs(p - gold["reference_probs_opencv3"]).max() <= TOL_PROB
        tp, tn, fp, fn = svm.confusion
        assert tp + tn + fp + fn == 269
        # svm.cpp:84-97 on the reference's own probabilities gives the same confusion matrix
        g, lab = gold["reference_probs_opencv3"], gold["testing_labels"]
        want = (int(np.sum((g >= 0.5) & (lab == 1))), int(np.sum((g < 0.5) & (lab == 0))), int(np.sum((g >= 0.5) & (lab != 1))),
                int(np.sum((g < 0.5) & (lab != 0))))
        assert abs(tp - want[0]) + abs(tn - want[1]) + abs(fp - want[2]) + abs(fn - want[3]) <= 2     # samples within 1e-6 of 0.5
    finally:
        svm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model", MODELS, ids=[os.path.basename(m)[:-4] for m in MODELS])
def test_gpu_matches_opencv_and_oracle_on_every_shipped_model(svm_orc, gold, model):
    svm = ab.B200SVM(device=0)
    try:
        svm.load(model)
        x = np.concatenate([gold["testing_features"], gold["grid_features"]], 0)
        p, raw = svm.test(x, want_raw=True)
        want = gold["cv2_raw_" + os.path.basename(model)[:-4]]
        o_raw, scale = svm_orc.predict_raw(svm_orc.load_model(model), x, return_scale=True)
        assert (np.abs(raw.astype(np.float64) - want) / (scale + np.abs(want))).max() <= 4e-7      # OpenCV itself
        # the oracle: same arithmetic; only a fractional-degree pow comes from two different libms
        assert (np.abs(raw.astype(np.float64) - o_raw) / (scale + np.abs(o_raw))).max() <= 1.3e-7
        assert np.abs(p - svm_orc.probability(raw)).max() <= 1e-15
        # batch sizes around the block size, and a single sample like App::computeAlignmentRisk (app.cpp:175-181)
        for n in (1, 127, 128, 129):
            assert np.array_equal(svm.test(x[:n]), p[:n])
    finally:
        svm.close()


@pytest.mark.gpu
def test_gpu_classifier_error_paths():
    svm = ab.B200SVM(device=0)
    try:
        with pytest.raises(ab.capi.AicpError, match="CONFIG"):
            svm.test(np.zeros((1, 2)))                                  # no model loaded
        svm.load(DEFAULT_MODEL)
        with pytest.raises(ab.capi.AicpError, match="BAD_ARG"):
            svm.test(np.zeros((3, 3)))                                  # wrong number of features
        assert svm.test(np.zeros((0, 2))).shape == (0,)
        assert ab.create_classifier(ab.ClassificationParams(type="Forest")) is None
    finally:
        svm.close()
