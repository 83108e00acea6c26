"""Handles shared by the checkers and the node store: one CUDA context object per
device, reference-counted occupancy maps, and the host/device array plumbing.

PyTorch appears here only as the owner of device buffers and streams
(``Tensor.data_ptr()``, ``torch.cuda.current_stream()``); every computation is a
kernel of ``libsbpgpu.so`` reached through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import SbpGpuError, check

_contexts = {}
_lock = threading.Lock()


def _is_torch_tensor(a) -> bool:
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


class Context:
    """Wraps ``sbp_ctx``: device ordinal + the stream the kernels are launched on."""

    def __init__(self, device: int = 0):
        lib = _lib.load()
        self.device = int(device)
        h = C.c_void_p()
        check(lib.sbp_ctx_create(self.device, None, C.byref(h)), "sbp_ctx_create")
        self._h = h
        self._stream = None  # private stream created by the library

    @property
    def handle(self):
        return self._h

    def use_torch_stream(self):
        """Launch on torch's current stream so device tensors are ordered correctly."""
        import torch

        s = torch.cuda.current_stream(self.device).cuda_stream
        if self._stream != s:
            check(_lib.load().sbp_ctx_set_stream(self._h, C.c_void_p(s)), "sbp_ctx_set_stream")
            self._stream = s

    def sync(self):
        check(_lib.load().sbp_ctx_sync(self._h), "sbp_ctx_sync")

    def info(self):
        sm, launches = C.c_int32(), C.c_int64()
        check(_lib.load().sbp_ctx_info(self._h, C.byref(sm), C.byref(launches)), "sbp_ctx_info")
        return {"sm_count": sm.value, "launches": launches.value}

    @property
    def launches(self) -> int:
        return self.info()["launches"]

    def tune(self, key, value: int) -> None:
        """Set a tuning / instrumentation knob of THIS context (csrc/common.cuh: sbp_tuning);
        none of them changes a result."""
        if isinstance(key, str):
            key = key.encode()
        check(_lib.load().sbp_ctx_tune(self._h, key, int(value)), "sbp_ctx_tune")

    def kernel_time(self):
        """``(total_ms, launches)`` of the kernels bracketed since the last call
        (``tune("time_kernels", which)``); synchronises the stream."""
        ms, n = C.c_double(), C.c_int64()
        check(_lib.load().sbp_ctx_kernel_time(self._h, C.byref(ms), C.byref(n)), "sbp_ctx_kernel_time")
        return float(ms.value), int(n.value)

    @property
    def scratch_generation(self) -> int:
        g = C.c_int64()
        check(_lib.load().sbp_ctx_scratch_generation(self._h, C.byref(g)), "sbp_ctx_scratch_generation")
        return int(g.value)


def tune(key, value: int, device: int = 0) -> None:
    """``get_context(device).tune(key, value)``."""
    get_context(device).tune(key, value)


def get_context(device: int = 0) -> Context:
    """Process-wide context of a device (created on first use; raises without a B200)."""
    with _lock:
        ctx = _contexts.get(int(device))
        if ctx is None:
            ctx = Context(device)
            _contexts[int(device)] = ctx
        return ctx


class DeviceMap:
    """Reference-counted ``sbp_map``: the bit-packed occupancy grid on the device.

    ``free_xy`` is the checker's ``_img`` ([W][H], truthy where ``_img == 1``).
    Immutable after creation, so deep copies of a checker share it.
    """

    def __init__(self, free_xy, device: int = 0):
        self.ctx = get_context(device)
        lib = _lib.load()
        h = C.c_void_p()
        if _is_torch_tensor(free_xy):
            import torch

            t = free_xy
            if t.dtype != torch.uint8:
                t = (t == 1).to(torch.uint8)
            t = t.contiguous()
            assert t.is_cuda and t.dim() == 2
            self.ctx.use_torch_stream()
            W, H = int(t.shape[0]), int(t.shape[1])
            check(lib.sbp_map_create_device(self.ctx.handle, C.c_void_p(t.data_ptr()), W, H,
                                            C.byref(h)), "sbp_map_create_device")
        else:
            a = np.ascontiguousarray((np.asarray(free_xy) == 1).astype(np.uint8))
            if a.ndim != 2:
                raise ValueError("occupancy grid must be 2-D [W][H]")
            W, H = a.shape
            check(lib.sbp_map_create(self.ctx.handle, a.ctypes.data_as(C.c_void_p), W, H,
                                     C.byref(h)), "sbp_map_create")
        self._h = h
        self.W, self.H = int(W), int(H)

    @property
    def handle(self):
        if self._h is None:
            raise SbpGpuError("occupancy map handle already destroyed")
        return self._h

    def info(self):
        W, H, sm = C.c_int32(), C.c_int32(), C.c_int32()
        nb = C.c_int64()
        check(_lib.load().sbp_map_info(self.handle, C.byref(W), C.byref(H), C.byref(nb), C.byref(sm)),
              "sbp_map_info")
        return {"W": W.value, "H": H.value, "packed_bytes": nb.value, "smem_resident": bool(sm.value)}

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.load().sbp_map_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass


def as_host_f64(a, cols: int) -> np.ndarray:
    """C-contiguous float64 [n][cols] view/copy of a host array-like."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError(f"expected an array of shape [n][{cols}], got {a.shape}")
    return a


def as_device_f64(t, cols: int, device: int = None, allow_pinned: bool = False):
    """Contiguous float64 CUDA tensor [n][cols] (no host round trip).  A host tensor is refused
    (its address means nothing to a kernel) unless ``allow_pinned`` and it is pinned, float64 and
    contiguous already -- any conversion would produce a pageable temporary."""
    import torch

    if not t.is_cuda:
        if not (allow_pinned and t.is_pinned() and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError("expected a CUDA tensor (host data goes through the numpy entry points)")
    elif device is not None and t.device.index != int(device):
        raise ValueError(f"tensor lives on cuda:{t.device.index}, the handle on cuda:{int(device)}")
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    if t.dim() == 1:
        t = t.reshape(1, -1)
    if t.dim() != 2 or t.shape[1] != cols:
        raise ValueError(f"expected a tensor of shape [n][{cols}], got {tuple(t.shape)}")
    return t.contiguous()


def ptr(a):
    """Raw address of a numpy array or a torch tensor (``None`` -> NULL)."""
    if a is None:
        return None
    if _is_torch_tensor(a):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def raise_for_flags(flags: int, what: str, nan_is_error: bool = True) -> None:
    """Surface what CPython would have raised for the batch."""
    if flags & _lib.FLAG_PEER_TIMEOUT:
        raise RuntimeError(f"{what}: a peer rank never signalled the in-kernel rendezvous")
    if flags & _lib.FLAG_OVERFLOW:
        raise OverflowError(f"{what}: cannot convert float infinity to integer")
    if nan_is_error and flags & _lib.FLAG_NAN:
        raise ValueError(f"{what}: cannot convert float NaN to integer")
