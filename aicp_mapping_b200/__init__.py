"""aicp_mapping_b200 -- B200-native (sm_100a) drop-in for AICP's registration hot path.

The product is libaicp_b200.so (csrc/, built by build.py) behind the C ABI of include/aicp_b200.h; registration.py and
overlap.py mirror the reference's AbstractRegistrator / AbstractOverlapper plug-in interfaces on top of it for the
Python harness.  There is no CPU fallback and nothing here imports the oracle.
"""
from .classification import B200SVM, ClassificationParams, SVMParams, create_classifier  # noqa: F401
from .filtering import (B200Alignability, B200CropBox, B200Map, B200Prefilter, DeviceCloudView, default_prefilter_config,  # noqa: F401
                        getPointsInOrientedBox, regionGrowingUniformPlaneSegmentationFilter)
from .ingest import B200VelodyneAccumulator, processFromFile, readPCD, readPLY, readPoseFile, writePCD  # noqa: F401
from .overlap import B200Overlap, OverlapParams, create_overlapper  # noqa: F401
from .registration import (B200Registration, RegistrationParams, autotune_ratio, computeRegistration,  # noqa: F401
                           create_registrator, parseTransformationDeg, replaceRatioConfigFile)
