import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "lidar-image_object-detection_-fpn_resnet-yolov8_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pkg(sub=None):
    return importlib.import_module(PKG if sub is None else PKG + "." + sub)


@pytest.fixture(scope="session")
def sfa():
    return pkg()


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads libsfa_b200.so."""
    build = importlib.import_module(PKG + ".build")
    build.build()
    return pkg("_lib").load()


@pytest.fixture(scope="session")
def cuda_device(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a test marked `gpu` ran without a CUDA device")
    return torch.device("cuda", 0)
