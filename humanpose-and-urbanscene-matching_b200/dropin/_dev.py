"""Host <-> device plumbing shared by the drop-in modules (torch tensors as buffer carriers only)."""
import numpy as np
import torch

from .. import engine


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("the drop-in modules run on a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def up(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(device())


def down(t):
    return t.cpu().numpy()


__all__ = ["engine", "device", "up", "down"]
