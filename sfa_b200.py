"""Alias: `import sfa_b200` == the package directory `lidar-image_object-detection_-fpn_resnet-yolov8_b200/`
(whose name is not a valid identifier for the `import` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("lidar-image_object-detection_-fpn_resnet-yolov8_b200")
sys.modules[__name__] = _pkg
