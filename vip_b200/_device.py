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
    if a.nbytes >= _STAGED_MIN_BYTES and torch.device(device).type == "cuda":
        return upload_staged(a, device)
    return torch.from_numpy(a).to(device, non_blocking=False)


_STAGED_MIN_BYTES = 16 << 20


def upload_staged(a, device):
    """C-contiguous numpy array -> CUDA tensor through ``vb_memcpy_h2d_staged`` (multi-threaded pinned staging of
    pageable sources; large uploads run at PCIe speed instead of the driver's ~10 GB/s single-threaded staging)."""
    from . import _cabi
    out = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].reshape(-1)).dtype, device=device)
    with torch.cuda.device(out.device):
        _cabi.check(_cabi.lib().vb_memcpy_h2d_staged(int(out.data_ptr()), int(a.ctypes.data), int(a.nbytes),
                                                      stream_ptr()), "vb_memcpy_h2d_staged")
    return out


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


class gpu_local_cpus:
    """Context manager: run the enclosed host code on the CPU cores local to CUDA device ``index`` (NVML's CPU affinity
    of the GPU = the NUMA node of its PCIe root port), then restore the previous affinity.

    Pinned host buffers are placed by first touch: allocating them inside this context puts the pages next to the
    GPU that will DMA from them.  With one process per GPU and eight GPUs uploading at once, buffers scattered over
    both sockets halve the aggregate PCIe rate (measured at config 5: 24 GB/s per GPU at N = 8 against 56 GB/s at
    N = 1, profiles/r02m_c5_full_scaling_8gpu_box.jsonl).  Silently does nothing when NVML or the affinity call is
    unavailable."""

    def __init__(self, index):
        self.index = int(index)
        self.saved = None

    def __enter__(self):
        try:
            import os
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(self.index)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus:
                self.saved = allowed
                os.sched_setaffinity(0, cpus)
        except Exception:                                       # noqa: BLE001 - best effort
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                import os
                os.sched_setaffinity(0, self.saved)
            except Exception:                                   # noqa: BLE001
                pass
        return False
