"""Device plumbing shared by the host-side wrappers: torch is used only for
device memory, streams and workspace caching."""
from __future__ import annotations

import torch

from . import _lib

_workspaces = {}
lib = _lib.load
check = _lib.check


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("posfeat_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def to_device(t: torch.Tensor, dtype=torch.float32):
    """Return (cuda tensor, original device)."""
    require_cuda()
    dev = t.device
    if dev.type != "cuda":
        t = t.to("cuda", non_blocking=True)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t, dev


def workspace(key: str, nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer (uint8), one per (name, device, CURRENT STREAM): two host threads (or one
    thread alternating between streams) that call the same op on different streams never share scratch
    memory, so the calls are re-entrant ("safe from multiple host threads on different streams").
    Calls that hand data to each other through a workspace (sampler -> matcher operands) run on one stream
    and therefore see the same buffer."""
    d = torch.device(device)
    idx = d.index if d.index is not None else torch.cuda.current_device()
    k = (key, idx, torch.cuda.current_stream(idx).cuda_stream)
    buf = _workspaces.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=d)
        _workspaces[k] = buf
    return buf


def release_workspaces():
    """Drop every cached scratch buffer (e.g. after a one-off very large call)."""
    _workspaces.clear()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def map_ptr(t: torch.Tensor) -> int:
    """Address a kernel can dereference: the tensor's own pointer for device memory, the device-side
    alias (posfeat_host_device_pointer) for a pinned host tensor; pageable host memory raises."""
    if t.device.type == "cuda":
        return t.data_ptr()
    import ctypes as C
    out = C.c_void_p(0)
    check(lib().posfeat_host_device_pointer(t.data_ptr(), C.byref(out)))
    return int(out.value)


