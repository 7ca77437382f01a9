"""CPU restatement of the alignment-risk classifier (aicp_core/src/classification/svm.cpp:53-107).  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.

aicp::SVM::test (svm.cpp:53-101) runs, per sample, cv::ml::SVM::predict(sample, output, 1) -- the raw decision value -- and
turns it into probability = 1.0 - 1.0 / (1.0 + exp(-output)) (svm.cpp:82); aicp::SVM::load (svm.cpp:103-107) is
cv::ml::SVM::load.  The arithmetic is OpenCV's (find_package(OpenCV), aicp_core/CMakeLists.txt:13, >= 3.0, unpinned; not
vendored under /root/reference), restated here from its published algorithm [UPSTREAM modules/ml/src/svm.cpp]:
  SVMKernelImpl::calc_non_rbf_base   s = sum_k sample[k] * sv[k]  (float32 products, float64 accumulator),
                                     K = (float32)(s * gamma + coef0)
  SVMKernelImpl::calc_poly           K = cv::pow(K, degree): an integer degree is binary exponentiation in float32 (iPow: a = 1,
                                     b = K; while p > 1: if p & 1: a *= b; b *= b; p >>= 1; a *= b), restated exactly; a
                                     fractional degree uses OpenCV's own float32 exp/log, here float64 pow -> float32
  SVMImpl::PredictBody               sum = -rho + sum_i alpha[i] * K[index[i]]  (float64); two classes + RAW_OUTPUT: (float32)sum

PARITY PINNED for this function: the reference ships its own inputs and outputs for it --
aicp_core/data/labels/testing_labelled_27Aug.txt (269 samples) -> aicp_core/data/classification/probs_opencv3.txt, produced by
aicp_core/src/classification/main.cpp:133-154 with the model svm_1000training_thresh50_cross_validation_opencv3.xml
(aicp_ros/launch/aicp.launch:19-20) -- and OpenCV's own implementation is importable in the build container (cv2 4.13):
tests/golden/make_svm_goldens.py freezes both into tests/golden/svm_goldens.npz.
"""
import xml.etree.ElementTree as ET

import numpy as np

LINEAR, POLY = 0, 1


def load_model(path):
    """cv::ml::SVM::load for the OpenCV 3 layout (<opencv_ml_svm>, svmType) and the legacy 2.4 layout (<my_svm>, svm_type)."""
    root = ET.parse(path).getroot()[0]
    svm_type = (root.findtext("svmType") or root.findtext("svm_type") or "").strip()
    if svm_type != "C_SVC":
        raise ValueError("only C_SVC models are restated (svm.cpp:9)")
    k = root.find("kernel")
    ktype = k.findtext("type").strip()
    m = dict(kernel={"LINEAR": LINEAR, "POLY": POLY}[ktype], degree=0.0, gamma=1.0, coef0=0.0)
    if ktype == "POLY":
        m.update(degree=float(k.findtext("degree")), gamma=float(k.findtext("gamma")), coef0=float(k.findtext("coef0")))
    m["dim"] = int(root.findtext("var_count"))
    if int(root.findtext("class_count")) != 2:
        raise ValueError("two classes expected")
    m["sv"] = np.array([[float(x) for x in e.text.split()] for e in root.find("support_vectors")], dtype=np.float32)
    assert m["sv"].shape == (int(root.findtext("sv_total")), m["dim"])
    df = root.find("decision_functions")[0]
    m["rho"] = float(df.findtext("rho"))
    m["alpha"] = np.array([float(x) for x in df.findtext("alpha").split()], dtype=np.float64)
    idx = df.findtext("index")
    m["index"] = np.array([int(x) for x in idx.split()], dtype=np.int64) if idx is not None else np.arange(len(m["alpha"]))
    assert len(m["alpha"]) == len(m["index"]) == int(df.findtext("sv_count"))
    return m


def ipow_f32(k, p):
    """cv::pow with an integer power on float32 data (iPow)."""
    a = np.ones_like(k, dtype=np.float32)
    b = k.astype(np.float32).copy()
    while p > 1:
        if p & 1:
            a = (a * b).astype(np.float32)
        b = (b * b).astype(np.float32)
        p >>= 1
    return (a * b).astype(np.float32)


def predict_raw(m, features, return_scale=False):
    """cv::ml::SVM::predict(samples, out, RAW_OUTPUT): float32 decision values, one per row of `features`.
    return_scale: also sum_i |alpha_i K_i| per sample, the magnitude against which float32 rounding of K is to be judged."""
    x = np.asarray(features, dtype=np.float64).reshape(-1, m["dim"]).astype(np.float32)      # svm.cpp:72-74
    out = np.zeros(x.shape[0], dtype=np.float32)
    scale = np.zeros(x.shape[0], dtype=np.float64)
    sv = m["sv"]
    for t in range(x.shape[0]):
        s = np.zeros(sv.shape[0], dtype=np.float64)
        for d in range(m["dim"]):
            s = s + (sv[:, d] * x[t, d]).astype(np.float64)          # float32 product, float64 accumulation, k ascending
        kv = (s * m["gamma"] + m["coef0"]).astype(np.float32)
        if m["kernel"] == POLY:
            deg = m["degree"]
            if abs(round(deg) - deg) < 2.220446049250313e-16 and deg >= 1:
                kv = ipow_f32(kv, int(round(deg)))
            else:
                kv = np.power(kv.astype(np.float64), deg).astype(np.float32)
        acc = -m["rho"]
        for a, i in zip(m["alpha"], m["index"]):
            acc = acc + a * float(kv[i])
        out[t] = np.float32(acc)
        scale[t] = abs(m["rho"]) + float(np.sum(np.abs(m["alpha"]) * np.abs(kv[m["index"]].astype(np.float64))))
    return (out, scale) if return_scale else out


def probability(raw):
    """svm.cpp:82."""
    return 1.0 - 1.0 / (1.0 + np.exp(-np.asarray(raw, dtype=np.float32).astype(np.float64)))


def test(m, features):
    """aicp::SVM::test(testing_data, &probabilities) (svm.cpp:46-51)."""
    return probability(predict_raw(m, features))
