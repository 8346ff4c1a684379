"""ctypes loader for the C-ABI library libmtdgan_sm100a.so.

The prototypes are parsed from include/mtdgan_b200.h, so the header is the single source of truth for
the boundary.  There is no fallback of any kind: if the library is missing, or a tensor is not a
contiguous fp32 CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_PKG_DIR)
HEADER = os.path.join(_REPO, "include", "mtdgan_b200.h")
LIB_NAME = "libmtdgan_sm100a.so"

_CTYPES = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float}


def lib_path() -> str:
    return os.path.join(_PKG_DIR, LIB_NAME)


def is_built() -> bool:
    return os.path.isfile(lib_path())


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """{symbol: (return type, [argument types])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"\b(int|long long)\s+(mtd_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append("ptr")
                else:
                    types.append(re.sub(r"\s+\w+$", "", a).replace("const ", "").strip())
        protos[name] = (ret, types)
    return protos


_lib = None
_protos = None


def load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not is_built():
        raise RuntimeError(
            f"{LIB_NAME} is not built (expected at {lib_path()}). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU or PyTorch fallback for the MTD-GAN hot path.")
    lib = ctypes.CDLL(lib_path())
    _protos = parse_header()
    for name, (ret, types) in _protos.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = _CTYPES[ret]
        fn.argtypes = [ctypes.c_void_p if t == "ptr" else _CTYPES[t] for t in types]
    _lib = lib
    if os.environ.get("MTD_PDL", "1") == "0":       # A/B switch for the programmatic-dependent-launch path
        lib.mtd_set_pdl(0)
    return lib


def require_cuda_extension():
    """Raise unless the CUDA library is loadable and a B200-class device is present."""
    lib = load()
    if not torch.cuda.is_available():
        raise RuntimeError("mtdgan_b200 needs a CUDA device (sm_100a); no CPU path exists.")
    ok = lib.mtd_device_ok()
    if ok != 1:
        raise RuntimeError(f"mtdgan_b200 kernels are built for sm_100a only (mtd_device_ok() = {ok}).")
    return lib


class MtdError(RuntimeError):
    pass


_profile = None           # when a list: (entry point, args, start event, end event) per call (bench.py)


def kernel_launch_count() -> int:
    """CUDA kernels launched by the library so far in this process."""
    return int((_lib if _lib is not None else load()).mtd_kernel_launch_count())


_trace = None             # when a list: (entry point, args) per call, no events (legal during CUDA-graph capture)


def start_trace():
    """Record (entry point, args) of every C-ABI call.  Taken while a step is being captured into a CUDA graph, the
    pointers belong to the graph's private pool and stay valid for the life of the graph, so bench.py can re-issue
    one kernel family of the step as its own graph and time it in isolation."""
    global _trace
    _trace = []


def stop_trace():
    global _trace
    rec, _trace = _trace, None
    return rec


def start_profile():
    """Bracket every C-ABI call with CUDA events on the launching stream (for bench.py's per-kernel times;
    adds two event records per call, so never enabled inside a timed region)."""
    global _profile
    _profile = []


def stop_profile():
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    return [(n, a, s.elapsed_time(e)) for n, a, s, e in rec]


def call(name: str, *args):
    """Invoke an int-returning entry point; raise on a non-zero status."""
    lib = _lib if _lib is not None else load()
    if _profile is not None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = getattr(lib, name)(*args)
        e.record()
        _profile.append((name, args, s, e))
    else:
        if _trace is not None:
            _trace.append((name, args))
        rc = getattr(lib, name)(*args)
    if rc != 0:
        if rc < 0:
            raise MtdError(f"{name}: invalid argument / unsupported shape (status {rc})")
        raise MtdError(f"{name}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")
    return rc


def ptr(t):
    """Device pointer of a tensor that must be a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MtdError("mtdgan_b200 ops take CUDA tensors only (no CPU fallback exists)")
    if not t.is_contiguous():
        raise MtdError("mtdgan_b200 ops take contiguous tensors")
    return t.data_ptr()


def fptr(t):
    if t is not None and t.dtype != torch.float32:
        raise MtdError(f"expected float32, got {t.dtype}")
    return ptr(t)


def stream():
    return torch.cuda.current_stream().cuda_stream


_capture_keepalive = []       # pinned staging buffers referenced by memcpy nodes of captured CUDA graphs


def device_table(rows, dtype, device):
    """Small host-built table -> device through PINNED staging (a pageable H2D copy is illegal during CUDA-graph
    capture).  While capturing, the staging buffer is kept alive for the life of the process: the graph's memcpy
    node re-reads it on every replay."""
    host = torch.tensor(rows, dtype=dtype).pin_memory()
    dev = torch.empty(host.shape, dtype=dtype, device=device)
    dev.copy_(host, non_blocking=True)
    if torch.cuda.is_current_stream_capturing():
        _capture_keepalive.append(host)
    return dev
