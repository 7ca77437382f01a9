"""Host-side mirror of the reference's classifier plug-in (aicp_core/include/aicp_classification/abstract_classification.hpp:8-18,
classification.hpp:8-19, common.hpp:34-44): AbstractClassification::load / test, selected by ClassificationParams.type == "SVM".
The decision values are computed on the GPU (csrc/svm.cu) through the C ABI; this module only mirrors the interface for the
test and benchmark harness.  train() is the reference's offline cv::ml::SVM::trainAuto tool and is not provided."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class SVMParams:
    threshold: float = 0.5          # on the prediction probability (aicp_ros/launch/aicp.launch:16)
    trainingFile: str = ""
    testingFile: str = ""
    saveFile: str = ""              # svm.cpp:61-63: test() loads this file first when it is set
    saveProbs: str = ""
    modelLocation: str = ""


@dataclass
class ClassificationParams:
    type: str = "SVM"
    svm: SVMParams = field(default_factory=SVMParams)


class B200SVM:
    """aicp::SVM (svm.hpp) on the GPU."""

    def __init__(self, params=None, device=-1):
        self.params = params or ClassificationParams()
        self._lib = capi.lib()
        h = C.c_void_p()
        rc = self._lib.aicp_b200_create(None, int(device), C.byref(h))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(None).decode())
        self._h = h
        self.confusion = None

    def close(self):
        if self._h:
            self._lib.aicp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(self._h).decode())

    def train(self, training_data, labels):
        raise NotImplementedError("SVM::train (cv::ml::SVM::trainAuto, svm.cpp:18-51) is the reference's offline tool; "
                                  "load() a model it produced")

    def load(self, filename):
        """svm.cpp:103-107."""
        self._check(self._lib.aicp_b200_svm_load(self._h, str(filename).encode()))

    def test(self, testing_data, labels=None, want_raw=False):
        """svm.cpp:46-101: returns the probabilities (n float64).  With non-zero labels also fills self.confusion =
        (tp, tn, fp, fn) against params.svm.threshold like svm.cpp:84-97."""
        if self.params.svm.saveFile:
            self.load(self.params.svm.saveFile)
        x = np.ascontiguousarray(np.asarray(testing_data, dtype=np.float64))
        if x.ndim == 1:
            x = x.reshape(1, -1)
        n, dim = x.shape
        probs = np.zeros(n, dtype=np.float64)
        raw = np.zeros(n, dtype=np.float32)
        self._check(self._lib.aicp_b200_svm_predict(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), n, dim,
                                                    probs.ctypes.data_as(C.POINTER(C.c_double)), raw.ctypes.data_as(C.POINTER(C.c_float))))
        self.confusion = None
        if labels is not None and np.any(np.asarray(labels) != 0):
            lab = np.asarray(labels, dtype=np.float64).reshape(-1)
            high = probs >= self.params.svm.threshold
            self.confusion = (int(np.sum(high & (lab == 1.0))), int(np.sum(~high & (lab == 0.0))),
                              int(np.sum(high & (lab != 1.0))), int(np.sum(~high & (lab != 0.0))))
        return (probs, raw) if want_raw else probs


def create_classifier(params, device=-1):
    """classification.hpp:8-19: "SVM" is the reference's only type; anything else prints and returns None."""
    if params.type == "SVM":
        return B200SVM(params, device)
    print("Invalid classification type %s." % params.type)
    return None


def parse_model(filename):
    """The library's reading of a model file (no GPU needed): capi.SvmSummary."""
    s = capi.SvmSummary()
    err = C.create_string_buffer(512)
    rc = capi.lib().aicp_b200_svm_parse(str(filename).encode(), C.byref(s), err, 512)
    if rc:
        raise capi.AicpError(rc, err.value.decode(errors="replace"))
    return s
