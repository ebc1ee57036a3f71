"""ctypes binding of libpdr.so (the C ABI declared in include/pdr.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the
caller gets an exception (SURVEY.md §8b "Error conventions").
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# PDR_LIB_PATH: load another build of the same ABI (A/B timing of kernel variants in one run)
LIB_PATH = os.environ.get("PDR_LIB_PATH") or os.path.join(_HERE, "libpdr.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pdr.h")

_lib = None


class PdrError(RuntimeError):
    pass


def load():
    """Load libpdr.so once. Raises PdrError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PdrError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for this path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pdr_last_error.restype = ctypes.c_char_p
        _lib.pdr_launch_count.restype = ctypes.c_ulonglong
    return _lib


def declared_symbols():
    """Names of all functions declared in include/pdr.h (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdr_[a-z0-9_]+)\s*\(", text)))


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) pointer of a contiguous tensor, or NULL for None."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_contiguous():
        raise PdrError("tensor passed to libpdr must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def check(status, what):
    if status != 0:
        msg = load().pdr_last_error().decode(errors="replace")
        raise PdrError(f"{what} failed (status {status}): {msg}")


def launch_count():
    return int(load().pdr_launch_count())


def call(name, *args):
    """Call `name(*args, stream)` on the current torch CUDA stream and raise on failure.

    Tensors may be passed directly (converted to their data pointer here, so temporaries stay
    alive until the launch is enqueued); None becomes a NULL pointer; Python ints/bools map to
    C int, Python floats to C float — pass ctypes.c_double / c_size_t / ... explicitly otherwise.
    """
    import torch
    lib = load()
    fn = getattr(lib, name)
    conv = []
    for a in args:
        if a is None:
            conv.append(ctypes.c_void_p(0))
        elif isinstance(a, torch.Tensor):
            conv.append(ptr(a))
        elif isinstance(a, float):
            conv.append(ctypes.c_float(a))
        elif isinstance(a, bool):
            conv.append(ctypes.c_int(int(a)))
        elif isinstance(a, int):
            conv.append(ctypes.c_int(a))
        else:
            conv.append(a)
    check(fn(*conv, _stream()), name)
