"""Device plumbing shared by the host-side wrappers: torch is used only for
device memory, streams and workspace caching."""
from __future__ import annotations

import threading

import torch

from . import _lib

_workspaces = {}
_scope = threading.local()
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
    k = (key, idx, torch.cuda.current_stream(idx).cuda_stream, getattr(_scope, "token", None))
    buf = _workspaces.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=d)
        _workspaces[k] = buf
    return buf


class workspace_scope:
    """Scratch buffers private to one owner: inside the ``with`` block ``workspace()`` hands out buffers no other
    caller sees.  A captured CUDA graph bakes the addresses of its scratch memory in, so every graph object captures
    inside its own scope and keeps the scope alive (``close()`` releases the buffers) -- the shared grow-only cache
    may replace a buffer at any later call."""

    def __init__(self):
        self.token = object()
        self._prev = None

    def __enter__(self):
        self._prev = getattr(_scope, "token", None)
        _scope.token = self.token
        return self

    def __exit__(self, *exc):
        _scope.token = self._prev
        return False

    def close(self):
        for k in [k for k in _workspaces if k[3] is self.token]:
            del _workspaces[k]

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def release_workspaces():
    """Drop every cached scratch buffer of the shared cache (e.g. after a one-off very large call); buffers that
    belong to a live ``workspace_scope`` (a captured graph) stay."""
    for k in [k for k in _workspaces if k[3] is None]:
        del _workspaces[k]


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


