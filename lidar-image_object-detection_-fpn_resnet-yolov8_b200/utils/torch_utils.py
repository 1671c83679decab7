"""`_sigmoid` of the reference's utils/torch_utils.py:44-45.  It sits on the backbone side of the
hot-path boundary and stays stock PyTorch: every reference caller applies it to
hm_cen / cen_offset before `decode` (test.py:150,167)."""
import torch


def _sigmoid(x):
    return torch.clamp(x.sigmoid_(), min=1e-4, max=1 - 1e-4)
