"""Device-resident fields with a numpy-facing façade.

The reference keeps every field in a numpy array that examples read and write
in place (``flow_sim.velocity_field[...] = ...``, ``np.amax(vorticity_field[idx])``;
reference ``examples/3d_examples/PointSourceAdvectAndDiffuseCase/point_source_advection_diffusion.py:65,121``).
Here the source of truth is a torch CUDA tensor; :class:`DeviceField` gives it the
small part of the ndarray interface those call sites use.  Host reads copy D2H,
host writes copy H2D; kernels use the device pointer directly.
"""
import ctypes
import types

import numpy as np
import torch

_NP_TO_TORCH = {
    np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64,
    np.dtype(np.int64): torch.int64,
    np.dtype(np.int32): torch.int32,
}


def torch_dtype(real_t):
    return _NP_TO_TORCH[np.dtype(real_t)]


def default_device():
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_HAS_CUDA = None


def current_stream_ptr(device=None):
    """The current CUDA stream of `device` as a ``void*``.  Called once per kernel launch, so it takes
    the raw-handle fast path (one C call) instead of building a ``torch.cuda.Stream`` object."""
    global _HAS_CUDA
    if _HAS_CUDA is None:
        _HAS_CUDA = torch.cuda.is_available()
    if not _HAS_CUDA:
        return ctypes.c_void_p(0)
    if _RAW_STREAM is not None:
        index = getattr(device, "index", None) if device is not None else None
        if index is None:
            index = torch.cuda.current_device()
        return ctypes.c_void_p(_RAW_STREAM(index))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _unwrap(value):
    if isinstance(value, DeviceField):
        return value.tensor
    return value


class DeviceField:
    """A view of a torch tensor that behaves like the reference's numpy field."""

    __array_priority__ = 1000

    def __init__(self, tensor, state=None):
        self.tensor = tensor
        self.flags = types.SimpleNamespace(writeable=True)
        # shared between all views of one allocation: host-side writes bump "version"
        self._state = {"version": 0} if state is None else state

    def _touch(self):
        self._state["version"] += 1

    @property
    def version(self):
        return self._state["version"]

    # ---- ndarray-like metadata
    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def ndim(self):
        return self.tensor.ndim

    @property
    def size(self):
        return self.tensor.numel()

    @property
    def dtype(self):
        return np.dtype(str(self.tensor.dtype).replace("torch.", ""))

    def __len__(self):
        return self.tensor.shape[0]

    # ---- host access
    def __array__(self, dtype=None, copy=None):
        a = self.tensor.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def numpy(self):
        return self.__array__()

    def __getitem__(self, idx):
        r = self.tensor[_unwrap(idx)]
        return DeviceField(r, self._state)

    def __setitem__(self, idx, value):
        value = _unwrap(value)
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value)).to(
                device=self.tensor.device, dtype=self.tensor.dtype)
        self.tensor[_unwrap(idx)] = value
        self._touch()

    def view(self):
        return DeviceField(self.tensor, self._state)

    def copy(self):
        return DeviceField(self.tensor.clone())

    def fill(self, value):
        self.tensor.fill_(value)
        self._touch()

    def item(self):
        return self.tensor.item()

    def __float__(self):
        return float(self.tensor.item())

    # ---- in-place arithmetic used by examples (field += ..., field *= ...)
    def _other(self, other):
        other = _unwrap(other)
        if isinstance(other, np.ndarray):
            other = torch.from_numpy(np.ascontiguousarray(other)).to(
                device=self.tensor.device, dtype=self.tensor.dtype)
        return other

    def __iadd__(self, other):
        self.tensor += self._other(other)
        self._touch()
        return self

    def __isub__(self, other):
        self.tensor -= self._other(other)
        self._touch()
        return self

    def __imul__(self, other):
        self.tensor *= self._other(other)
        self._touch()
        return self

    # ---- out-of-place arithmetic returns host arrays (diagnostics cadence)
    def __add__(self, other):
        return np.asarray(self) + np.asarray(other)

    def __sub__(self, other):
        return np.asarray(self) - np.asarray(other)

    def __mul__(self, other):
        return np.asarray(self) * np.asarray(other)

    def __neg__(self):
        return -np.asarray(self)

    __radd__ = __add__
    __rmul__ = __mul__

    def __rsub__(self, other):
        return np.asarray(other) - np.asarray(self)

    def __repr__(self):
        return f"DeviceField(shape={self.shape}, dtype={self.dtype}, device={self.tensor.device})"


def zeros(shape, real_t, device=None):
    device = default_device() if device is None else device
    return DeviceField(torch.zeros(tuple(int(s) for s in shape), dtype=torch_dtype(real_t), device=device))


def zeros_like(field):
    return DeviceField(torch.zeros_like(_unwrap(field)))


class Staged:
    """Give kernels a contiguous device tensor for ``x``; numpy inputs are uploaded
    and (for outputs) copied back, so the operator API also accepts host arrays as
    the reference's does."""

    def __init__(self, device):
        self.device = device
        self._writebacks = []

    def __call__(self, x, out=False):
        if x is None:
            return None
        if isinstance(x, DeviceField):
            if out:
                # an operator is about to rewrite this field on the device: whatever was derived from its
                # old contents (the cached max |u| of the stable timestep, "ghost planes are current")
                # is stale, exactly as after a host-side write
                x._touch()
            x = x.tensor
        if isinstance(x, torch.Tensor):
            if not x.is_contiguous():
                raise ValueError("fields handed to kernels must be contiguous")
            return x
        host = np.asarray(x)
        t = torch.from_numpy(np.ascontiguousarray(host)).to(self.device)
        if out:
            self._writebacks.append((host, t))
        return t

    def finish(self):
        for host, t in self._writebacks:
            host[...] = t.cpu().numpy()
        self._writebacks.clear()


def dptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())
