"""sfa_b200 — B200-native (sm_100a) implementation of SFA3D's point-cloud-side hot path, behind the
reference's own Python call signatures (SAGARCHRY0777/lidar-image_object-detection_-fpn_resnet-yolov8):

  stage A  get_filtered_lidar + makeBEVMap   (data_process/kitti_data_utils.py:228-241,
                                              data_process/kitti_bev_utils.py:22-55)
  stage B  _nms, _topk, decode, post_processing (utils/evaluation_utils.py:21-163)

Sub-modules mirror the reference's layout (config/, data_process/, utils/) so call sites switch by
changing an import root; `fast` holds the batched device-resident API.  All compute happens in
libsfa_b200.so (csrc/, C ABI in include/sfa_b200.h).  No CPU fallback, no Triton, no torch.compile.

The directory name contains '-', so import it with importlib (or through the `sfa_b200` alias
module at the repository root):

    import importlib; sfa = importlib.import_module("lidar-image_object-detection_-fpn_resnet-yolov8_b200")
"""
from . import _lib, geometry  # noqa: F401
from .geometry import BevGeometry, from_config  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
