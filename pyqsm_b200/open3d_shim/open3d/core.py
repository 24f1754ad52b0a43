"""``open3d.core`` names used by ``pyQSM/viz/ray_casting.py`` (:43, :62): ``Tensor`` and dtypes."""
import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
uint32 = torch.uint32
uint8 = torch.uint8
bool = torch.bool  # noqa: A001


def Tensor(data, dtype=None, device=None):
    """``o3c.Tensor(array, o3c.float32)`` -> a CPU torch tensor (has ``.numpy()``, ``.isfinite()``, ...)."""
    if isinstance(data, torch.Tensor):
        t = data
    else:
        t = torch.from_numpy(np.ascontiguousarray(data))
    return t.to(dtype) if dtype is not None else t
