"""Host<->device plumbing shared by the layer mirrors (torch = memory + streams only)."""
import numpy as np
import torch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("efficientdet_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_device(x, dtype=torch.float32):
    """-> (contiguous CUDA tensor of `dtype`, came_from_host)."""
    if isinstance(x, torch.Tensor):
        host = not x.is_cuda
        t = x.to(device=device(), dtype=dtype).contiguous()
        return t, host
    a = np.ascontiguousarray(x)
    t = torch.from_numpy(a).to(device=device(), dtype=dtype, non_blocking=False)
    return t.contiguous(), True


def give_back(t, host):
    return t.cpu().numpy() if host else t
