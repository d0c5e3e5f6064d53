"""Shared host logic of the three drop-in modules: parameter container -> native handle.

The modules keep real `nn.Parameter`s / buffers with the reference's exact `state_dict`
keys so `load_state_dict(ckpt['model_state_dict'])` works unchanged
(inference.py:52-53).  `forward` folds/packs them once per (device, parameter version) via
`ar_model_create` and then runs `ar_model_forward` on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import threading

import torch
import torch.nn as nn

from .. import _lib


class _Workspace:
    """Grow-only device scratch, one buffer per (device, CUDA stream).

    Work on one stream is ordered, so forwards issued on the same stream can share scratch; forwards on different streams
    (or threads using their own streams) get different buffers and never overwrite each other's activations -- the
    reference `nn.Module`s are stream-safe and so are these.  A buffer that is replaced by a larger one is handed back
    to the caching allocator with `record_stream`, so its memory is not reused before the kernels still reading it end."""
    _bufs: dict = {}
    _lock = threading.Lock()

    @classmethod
    def get(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        stream = torch.cuda.current_stream(device)
        key = (device.index, stream.cuda_stream)
        with cls._lock:
            buf = cls._bufs.get(key)
            if buf is None or buf.numel() < nbytes:
                if buf is not None:
                    buf.record_stream(stream)
                cls._bufs[key] = None
                del buf
                buf = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=device)
                cls._bufs[key] = buf
            return buf

    @classmethod
    def clear(cls):
        with cls._lock:
            cls._bufs.clear()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class GraphedForward:
    """A module forward captured into a CUDA graph for one fixed input shape (`NativeModule.capture`).

    The three models are 15-45 kernel launches each; at the small batches of interactive use (BASELINE configs 1-3:
    2-16 chunks) the kernels take a few microseconds and the launch train dominates.  Replaying the captured graph issues
    the whole forward with one driver call.  `__call__(x)` copies `x` into the graph's static input and returns the
    static output tensor (valid until the next call)."""

    def __init__(self, module, example: torch.Tensor):
        x = module._check_input(example)
        self.module = module
        self.x = x.clone()
        with torch.no_grad():
            side = torch.cuda.Stream(x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):               # warm up off the capture: packs weights, sets kernel attributes
                module(self.x)
            torch.cuda.current_stream(x.device).wait_stream(side)
            torch.cuda.synchronize(x.device)
            self.graph = torch.cuda.CUDAGraph()
            before = set(_Workspace._bufs)
            with torch.cuda.graph(self.graph):
                self.y = module(self.x)
            # the scratch allocated during capture lives in the graph's private pool: it belongs to this graph, not to
            # whatever later runs on a stream that happens to reuse the capture stream's handle
            with _Workspace._lock:
                self._scratch = [_Workspace._bufs.pop(k) for k in set(_Workspace._bufs) - before]
        self._key = module._param_key(x.device)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.x.shape:
            raise RuntimeError(f"graph was captured for input {tuple(self.x.shape)}, got {tuple(x.shape)}")
        if self.module._param_key(self.x.device) != self._key:
            raise RuntimeError("module parameters changed since the graph was captured; capture again")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y


class NativeModule(nn.Module):
    KIND = -1

    def __init__(self):
        super().__init__()
        self._handle = None
        self._handle_key = None

    # ------------------------------------------------------------------ handle management
    def _param_key(self, device):
        eps = tuple(m.eps for m in self.modules() if isinstance(m, nn.BatchNorm1d))
        return (str(device), eps) + tuple((k, v._version, v.data_ptr()) for k, v in self.state_dict(keep_vars=True).items())

    def native_handle(self, device: torch.device):
        key = self._param_key(device)
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self._release()
        L = _lib.lib()
        sd = {k: v.detach().to("cpu", torch.float32).contiguous()
              for k, v in self.state_dict().items() if v.dtype.is_floating_point}
        for name, mod in self.named_modules():       # BatchNorm eps is a module attribute, not a state_dict entry
            if isinstance(mod, nn.BatchNorm1d) and mod.eps != 1e-5:
                sd[name + ".eps"] = torch.tensor([mod.eps], dtype=torch.float32)
        arr = (_lib.ArTensor * len(sd))()
        keep = []
        for i, (k, v) in enumerate(sd.items()):
            name = k.encode()
            keep.append((name, v))
            arr[i].name = name
            arr[i].data = v.data_ptr()
            arr[i].ndim = v.dim()
            for d in range(v.dim()):
                arr[i].shape[d] = v.shape[d]
        h = C.c_void_p()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(dev_index):
            _lib.check(L.ar_model_create(self.KIND, arr, len(sd), dev_index, C.byref(h)))
        self._handle, self._handle_key = h, key
        return h

    def _release(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            try:
                _lib.lib().ar_model_destroy(h)
            except Exception:
                pass
        # plain attribute writes: nn.Module.__setattr__ is not usable while the interpreter shuts down
        self.__dict__["_handle"] = None
        self.__dict__["_handle_key"] = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # ------------------------------------------------------------------ forward plumbing
    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError(
                f"{type(self).__name__}: the B200 path implements eval-mode inference only "
                "(call .eval(); training is out of scope, SURVEY.md section 2 #9)")
        if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[1] != 1:
            raise RuntimeError(f"expected input [batch, 1, samples], got {tuple(getattr(x, 'shape', ()))}")
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: input must be a CUDA tensor -- this build has no CPU fallback")
        if x.dtype != torch.float32:
            raise RuntimeError(f"expected float32 input, got {x.dtype}")
        return x.contiguous()

    def _out_shape(self, B: int, T: int):
        raise NotImplementedError

    def capture(self, example: torch.Tensor) -> GraphedForward:
        """Capture `forward` for inputs shaped like `example` into a CUDA graph (see `GraphedForward`)."""
        return GraphedForward(self, example)

    def audit(self, x) -> dict:
        """Dynamic-range audit (`ar_model_audit_*`): run `forward(x)` layer by layer and return `{layer: max |activation|}`
        for every fp16 tensor the kernels store.  Activations saturate at +-65504 in fp16 storage; a value of 65504 here
        means that layer clipped on this input ("dynamic range").  Synchronises the device."""
        x = self._check_input(x)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = self.native_handle(x.device)
            _lib.check(L.ar_model_audit_enable(h, 1))
            try:
                self.forward(x)
                vals = (C.c_float * _lib.AUDIT_MAX_LAYERS)()
                n = C.c_int()
                _lib.check(L.ar_model_audit_read(h, vals, _lib.AUDIT_MAX_LAYERS, C.byref(n)))
                return {L.ar_model_audit_name(h, i).decode(): float(vals[i]) for i in range(n.value)}
            finally:
                L.ar_model_audit_enable(h, 0)

    def forward(self, x):
        x = self._check_input(x)
        B, _, T = x.shape
        L = _lib.lib()
        if B == 0:      # the reference modules map an empty batch to an empty output (ATen convs accept N = 0): nothing to launch
            return torch.empty(self._out_shape(0, T), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            h = self.native_handle(x.device)
            need = C.c_size_t()
            _lib.check(L.ar_model_workspace_bytes(h, B, T, C.byref(need)))
            ws = _Workspace.get(x.device, need.value)
            y = torch.empty(self._out_shape(B, T), dtype=torch.float32, device=x.device)
            _lib.check(L.ar_model_forward(h, x.data_ptr(), y.data_ptr(), B, T, ws.data_ptr(), ws.numel(),
                                          _stream_ptr(x.device)))
        return y
