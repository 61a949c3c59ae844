"""Device-memory plumbing (PyTorch is used for allocation, streams and H2D/D2H copies only)."""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("vip_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def free_memory_bytes():
    """Device memory this process can still use, in bytes: free at the driver plus the blocks PyTorch's caching
    allocator holds but has not handed out (sizing of mini-batches, ``check_memory``)."""
    require_cuda()
    free = int(torch.cuda.mem_get_info()[0])
    return free + int(torch.cuda.memory_reserved()) - int(torch.cuda.memory_allocated())


def stream_ptr():
    return int(torch.cuda.current_stream().cuda_stream)


def to_device_f32(array, device=None):
    """numpy array (any float dtype) or torch tensor -> contiguous fp32 CUDA tensor."""
    device = device or require_cuda()
    if isinstance(array, torch.Tensor):
        return array.to(device=device, dtype=torch.float32).contiguous()
    a = np.ascontiguousarray(array)
    if a.dtype != np.float32:
        a = a.astype(np.float32)
    return torch.from_numpy(a).to(device, non_blocking=False)


def to_device(array, dtype, device=None):
    device = device or require_cuda()
    return torch.as_tensor(np.ascontiguousarray(array), dtype=dtype).to(device)


def empty(shape, dtype=torch.float32, device=None):
    return torch.empty(shape, dtype=dtype, device=device or require_cuda())


def ptr(t):
    return 0 if t is None else int(t.data_ptr())


def to_host(t, dtype=None):
    a = t.detach().cpu().numpy()
    if dtype is not None and a.dtype != dtype:
        a = a.astype(dtype)
    return a
