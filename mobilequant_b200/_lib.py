"""ctypes binding of libmqb200.so (include/mqb200.h).  No compute happens in Python: every op below forwards raw
device pointers + the current torch CUDA stream to a hand-written sm_100a kernel.  There is NO CPU fallback: a
missing library or a CPU tensor raises."""
import ctypes, os, re, threading
from ctypes import c_int, c_int32, c_int64, c_float, c_void_p, c_char_p, POINTER, Structure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmqb200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "mqb200.h")


class MQError(RuntimeError):
    pass


class mq_qcfg(Structure):
    _fields_ = [("bitwidth", c_int32), ("is_symmetric", c_int32)]


_lib = None
_ctx = {}
_lock = threading.Lock()


def declared_symbols(header_path=HEADER_PATH):
    """Every function name declared in include/mqb200.h (used by the CPU tests to check the exports)."""
    txt = open(header_path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MQError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback for the MobileQuant hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mq_version.restype = c_int
    lib.mq_get_error_description.restype = c_char_p
    lib.mq_get_error_description.argtypes = [c_int]
    lib.mq_get_last_error_extra_info.restype = c_char_p
    lib.mq_get_last_error_extra_info.argtypes = [c_int, c_void_p]
    lib.mq_setup.argtypes = [POINTER(c_void_p), c_int]
    lib.mq_release.argtypes = [c_void_p]
    lib.mq_ref_context.argtypes = [c_void_p]
    _lib = lib
    return lib


def ctx(device_index=None):
    """One context per CUDA device, created lazily."""
    import torch
    if not torch.cuda.is_available():
        raise MQError("mobilequant_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device_index is None:
        device_index = torch.cuda.current_device()
    with _lock:
        if device_index not in _ctx:
            lib = load()
            h = c_void_p()
            rc = lib.mq_setup(ctypes.byref(h), int(device_index))
            if rc != 0:
                extra = lib.mq_get_last_error_extra_info(rc, h)
                raise MQError(f"mq_setup failed: {lib.mq_get_error_description(rc).decode()} "
                              f"({extra.decode() if extra else ''})")
            _ctx[device_index] = h
        return _ctx[device_index]


def check(rc, h):
    if rc != 0:
        lib = load()
        d = lib.mq_get_error_description(rc)
        e = lib.mq_get_last_error_extra_info(rc, h)
        raise MQError(f"libmqb200: {d.decode() if d else rc}: {e.decode() if e else ''}")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=None):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise MQError("mobilequant_b200 kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise MQError("mobilequant_b200 kernels take contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise MQError(f"expected {dtype}, got {t.dtype}")
    return c_void_p(t.data_ptr())
