"""Minimal DLPack consumer/producer in ctypes, so tensors can cross into the C ABI without PyTorch.

Consumer: ``view(obj)`` borrows the DLTensor behind ``obj.__dlpack__()`` (torch CUDA tensor, CuPy array,
our own DeviceTensor …) and returns pointer/shape/dtype/device. The capsule is kept alive by the returned
object and is never renamed to ``used_dltensor``, so its own destructor releases it: a pure borrow.

Producer: ``DeviceTensor`` owns device memory obtained through ``ishara_device_malloc`` and exports it
through ``__dlpack__`` / ``__dlpack_device__`` (``torch.from_dlpack(t)`` works when torch is present).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib

kDLCPU, kDLCUDA, kDLCUDAHost = 1, 2, 3
kDLInt, kDLUInt, kDLFloat, kDLBfloat = 0, 1, 2, 4


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("device", DLDevice),
        ("ndim", C.c_int32),
        ("dtype", DLDataType),
        ("shape", C.POINTER(C.c_int64)),
        ("strides", C.POINTER(C.c_int64)),
        ("byte_offset", C.c_uint64),
    ]


class DLManagedTensor(C.Structure):
    pass


_DELETER = C.CFUNCTYPE(None, C.POINTER(DLManagedTensor))
DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p), ("deleter", _DELETER)]

_api = C.pythonapi
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_New.restype = C.py_object
_api.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_api.PyCapsule_GetName.restype = C.c_char_p
_api.PyCapsule_GetName.argtypes = [C.py_object]

_DTYPES = {
    (kDLFloat, 32): "float32",
    (kDLFloat, 64): "float64",
    (kDLFloat, 16): "float16",
    (kDLBfloat, 16): "bfloat16",
    (kDLInt, 32): "int32",
    (kDLInt, 64): "int64",
    (kDLUInt, 8): "uint8",
    (kDLUInt, 16): "uint16",
}
_DTYPE_CODES = {v: k for k, v in _DTYPES.items()}
_ITEMSIZE = {"float32": 4, "float64": 8, "float16": 2, "bfloat16": 2, "int32": 4, "int64": 8, "uint8": 1, "uint16": 2}


class TensorView:
    """Borrowed view of a DLPack tensor: .ptr, .shape, .dtype (str), .device_type, .device_id."""

    def __init__(self, capsule, keep):
        self._capsule = capsule
        self._keep = keep
        p = _api.PyCapsule_GetPointer(capsule, b"dltensor")
        mt = C.cast(p, C.POINTER(DLManagedTensor)).contents
        t = mt.dl_tensor
        self.shape: Tuple[int, ...] = tuple(int(t.shape[i]) for i in range(t.ndim))
        key = (int(t.dtype.code), int(t.dtype.bits))
        if key not in _DTYPES or t.dtype.lanes != 1:
            raise TypeError(f"unsupported DLPack dtype code={t.dtype.code} bits={t.dtype.bits} lanes={t.dtype.lanes}")
        self.dtype = _DTYPES[key]
        self.device_type = int(t.device.device_type)
        self.device_id = int(t.device.device_id)
        self.ptr = int(t.data or 0) + int(t.byte_offset)
        if bool(t.strides):
            exp = 1
            for i in range(t.ndim - 1, -1, -1):
                if self.shape[i] != 1 and int(t.strides[i]) != exp:
                    raise ValueError("ishara_b200 needs dense row-major (C-contiguous) tensors")
                exp *= self.shape[i]

    @property
    def on_cuda(self) -> bool:
        return self.device_type == kDLCUDA

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) * _ITEMSIZE[self.dtype]


def view(obj, stream: Optional[int] = None) -> TensorView:
    """Borrow `obj` (anything with __dlpack__). `stream`: consumer stream handle per the DLPack protocol
    (None = no ordering requested beyond the producer's default)."""
    if not hasattr(obj, "__dlpack__"):
        raise TypeError(f"{type(obj).__name__} does not implement __dlpack__")
    try:
        cap = obj.__dlpack__(stream=stream) if stream is not None else obj.__dlpack__()
    except TypeError:
        cap = obj.__dlpack__()
    if not _api.PyCapsule_IsValid(cap, b"dltensor"):
        raise ValueError("__dlpack__ did not return a fresh 'dltensor' capsule")
    return TensorView(cap, obj)


# ---------------------------------------------------------------------------------------------
# producer
# ---------------------------------------------------------------------------------------------
_live_exports = {}  # id -> (managed struct, shape array, owner) kept alive until the consumer's deleter runs


@_DELETER
def _export_deleter(mt_ptr):
    key = C.cast(mt_ptr, C.c_void_p).value
    _live_exports.pop(key, None)


@C.CFUNCTYPE(None, C.c_void_p)
def _capsule_destructor(cap_ptr):
    # called when an UNCONSUMED capsule dies: name is still "dltensor" -> run the deleter ourselves
    cap = C.cast(cap_ptr, C.py_object)
    if _api.PyCapsule_IsValid(cap, b"dltensor"):
        p = _api.PyCapsule_GetPointer(cap, b"dltensor")
        _live_exports.pop(p, None)


class DeviceTensor:
    """Dense row-major tensor in device memory owned by libishara_b200 (no PyTorch needed)."""

    def __init__(self, shape: Sequence[int], dtype: str = "float32", device: int = 0):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.device = int(device)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * _ITEMSIZE[dtype]
        out = C.c_void_p()
        _lib.check(_lib.load().ishara_device_malloc(self.device, self.nbytes, C.byref(out)))
        self.ptr = int(out.value or 0)

    def __del__(self):
        try:
            if getattr(self, "ptr", 0):
                _lib.load().ishara_device_free(self.device, C.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass

    # -- host transfer -------------------------------------------------------------------------
    def copy_from_host(self, a: np.ndarray, stream: int = 0) -> "DeviceTensor":
        np_dt = "uint16" if self.dtype == "bfloat16" else self.dtype
        a = np.ascontiguousarray(a, dtype=np_dt)
        if a.nbytes != self.nbytes:
            raise ValueError(f"copy_from_host: {a.nbytes} bytes into a {self.nbytes}-byte tensor")
        lib = _lib.load()
        _lib.check(lib.ishara_memcpy_async(C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p), self.nbytes, 1, C.c_void_p(stream)))
        _lib.check(lib.ishara_stream_synchronize(self.device, C.c_void_p(stream)))
        return self

    def numpy(self, stream: int = 0) -> np.ndarray:
        np_dt = "uint16" if self.dtype == "bfloat16" else self.dtype
        out = np.empty(self.shape, dtype=np_dt)
        lib = _lib.load()
        _lib.check(lib.ishara_memcpy_async(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), self.nbytes, 2, C.c_void_p(stream)))
        _lib.check(lib.ishara_stream_synchronize(self.device, C.c_void_p(stream)))
        return out

    # -- DLPack producer -----------------------------------------------------------------------
    def __dlpack_device__(self):
        return (kDLCUDA, self.device)

    def __dlpack__(self, stream=None, **_):
        mt = DLManagedTensor()
        shape = (C.c_int64 * len(self.shape))(*self.shape)
        code, bits = _DTYPE_CODES[self.dtype]
        mt.dl_tensor.data = C.c_void_p(self.ptr)
        mt.dl_tensor.device = DLDevice(kDLCUDA, self.device)
        mt.dl_tensor.ndim = len(self.shape)
        mt.dl_tensor.dtype = DLDataType(code, bits, 1)
        mt.dl_tensor.shape = C.cast(shape, C.POINTER(C.c_int64))
        mt.dl_tensor.strides = None
        mt.dl_tensor.byte_offset = 0
        mt.manager_ctx = None
        mt.deleter = _export_deleter
        addr = C.addressof(mt)
        _live_exports[addr] = (mt, shape, self)  # `self` keeps the device memory alive for the consumer
        return _api.PyCapsule_New(addr, b"dltensor", C.cast(_capsule_destructor, C.c_void_p))


class BorrowedTensor(DeviceTensor):
    """DLPack producer over device memory owned by someone else (e.g. the model handle's flat gradient buffer).
    `owner` is kept alive as long as any consumer holds the capsule; nothing is freed here."""

    def __init__(self, ptr: int, shape: Sequence[int], dtype: str, device: int, owner=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.device = int(device)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * _ITEMSIZE[dtype]
        self.ptr = int(ptr)
        self.owner = owner

    def __del__(self):
        self.ptr = 0


def from_host(a: np.ndarray, device: int = 0, dtype: Optional[str] = None) -> DeviceTensor:
    a = np.ascontiguousarray(a)
    t = DeviceTensor(a.shape, dtype or str(a.dtype), device)
    return t.copy_from_host(a)
