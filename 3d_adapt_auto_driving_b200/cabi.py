"""ctypes door onto libpn2_b200.so (the C-ABI declared in include/pn2_b200.h).

There is deliberately NO fallback: if the CUDA library has not been built the import of any
op raises, and every op raises if handed a CPU tensor.  torch is used only for device memory
and the current stream.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpn2_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pn2_b200.h")

_lib = None


class Pn2Error(RuntimeError):
    pass


def declared_symbols(header_path=HEADER_PATH):
    """Function names declared in include/pn2_b200.h (used by the symbol-export test)."""
    with open(header_path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Pn2Error(
                "libpn2_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "there is no CPU or PyTorch fallback for these ops")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pn2_last_error.restype = ctypes.c_char_p
    return _lib


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise Pn2Error("expected a CUDA tensor: these ops have no CPU path")
    return ctypes.c_void_p(t.data_ptr())


def check(status, what):
    if status != 0:
        raise Pn2Error("%s failed (status %d): %s" % (what, status, lib().pn2_last_error().decode()))


# ---- launch accounting (bench.py: gpu_launches, per-kernel CUDA-event timing) ----
launch_count = 0
_profile = None   # None, or {"names": set() | None (= all), "records": [(name, work, start_evt, end_evt)]}


def profile_start(names=None):
    """Bracket every C-ABI launch (or only `names`) with CUDA events on the launching stream."""
    global _profile
    _profile = {"names": set(names) if names else None, "records": []}


def profiling():
    return _profile is not None


def profile_stop():
    """-> {name: {"launches", "ms", "work"}} ; synchronises."""
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for name, work, s, e in prof["records"]:
        d = out.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
        d["launches"] += 1
        d["ms"] += s.elapsed_time(e)
        d["work"] += float(work or 0.0)
    return out


def call(name, *args, work=None):
    """Call lib.<name>(*args, current_stream) and raise on a non-zero status.  `work` is the
    algorithmic FLOPs or bytes of this launch (only used by the profiler)."""
    global launch_count
    fn = getattr(lib(), name)
    launch_count += 1
    if _profile is not None and (_profile["names"] is None or name in _profile["names"]):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        check(fn(*args, stream_ptr()), name)
        e.record()
        _profile["records"].append((name, work, s, e))
        return
    check(fn(*args, stream_ptr()), name)


def f32(x):
    return ctypes.c_float(float(x))


def i32(x):
    return ctypes.c_int(int(x))
