#!/usr/bin/env python
"""Freezes golden vectors for the alignment-risk classifier (aicp_core/src/classification/svm.cpp) into
tests/golden/svm_goldens.npz and copies the reference's model files into tests/golden/svm_models/.
Run in the BUILD container only (it reads /root/reference and imports cv2, neither of which exists on the GPU box):

    python tests/golden/make_svm_goldens.py

Contents:
  testing_features        269 x 2: columns 1 and 100 * column 2 of aicp_core/data/labels/testing_labelled_27Aug.txt, exactly as
                          aicp_core/src/classification/main.cpp:136-139 builds them
  reference_probs_opencv3 the reference's OWN output for those samples, aicp_core/data/classification/probs_opencv3.txt
                          (written by main.cpp:149-153 with the model of aicp_ros/launch/aicp.launch:19), 6 significant digits
  grid_features           441 x 2 grid over [0, 100]^2 (overlap %, alignability %)
  cv2_raw_<model>         cv2.ml.SVM.predict(x, flags=RAW_OUTPUT) of OpenCV itself (cv2 4.13 in the build container) on the
                          testing samples followed by the grid, for every model file the reference ships
"""
import glob
import os
import shutil

import cv2
import numpy as np

REF = "/root/reference/aicp_core/data"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    t = np.loadtxt(os.path.join(REF, "labels", "testing_labelled_27Aug.txt"))
    feats = np.c_[t[:, 1], 100.0 * t[:, 2]]
    golden = np.loadtxt(os.path.join(REF, "classification", "probs_opencv3.txt"))[:, 1]
    g = np.arange(0.0, 100.1, 5.0)
    grid = np.array([[a, b] for a in g for b in g])
    allx = np.concatenate([feats, grid], 0).astype(np.float32)
    out = dict(testing_features=feats, testing_labels=t[:, 3], reference_probs_opencv3=golden, grid_features=grid,
               cv2_version=np.array(cv2.__version__))
    os.makedirs(os.path.join(HERE, "svm_models"), exist_ok=True)
    for f in sorted(glob.glob(os.path.join(REF, "classification", "*.xml"))):
        name = os.path.basename(f)
        dst = os.path.join(HERE, "svm_models", name)
        shutil.copyfile(f, dst)
        os.chmod(dst, 0o644)
        svm = cv2.ml.SVM_load(f)
        _, raw = svm.predict(allx, flags=cv2.ml.STAT_MODEL_RAW_OUTPUT)
        out["cv2_raw_" + name[:-4]] = raw.ravel().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "svm_goldens.npz"), **out)
    print("wrote", os.path.join(HERE, "svm_goldens.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
