"""ctypes binding of ``libgrapes_b200.so``.  The prototypes are parsed from
``include/grapes_b200.h`` so the header stays the single source of truth.

There is NO CPU fallback (BASELINE.json north_star): if the library is missing this module
raises, and ``grapes_ctx_create`` fails loudly on a box without a Blackwell GPU.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "grapes_b200.h")
LIB_PATH = os.path.join(HERE, "libgrapes_b200.so")

_SCALARS = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
            "double": ctypes.c_double, "uint32_t": ctypes.c_uint32}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """name -> (return type, [(param type, param name)])."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int)\s+(grapes_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        plist = []
        params = " ".join(params.split())
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                mm = re.match(r"(.*?)(\w+)$", p)
                plist.append((mm.group(1).strip(), mm.group(2)))
        out[name] = ("char*" if "char" in ret else ("int64" if "64" in ret else "int"), plist)
    return out


def _ctype(t: str):
    t = t.replace("const", "").strip()
    t = " ".join(t.split())
    if t.endswith("*"):
        return ctypes.c_void_p
    return _SCALARS[t]


class GrapesError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        # no-op when the binary carries the id of this source tree; rebuilds (one process at a time) when it does not, and
        # raises without nvcc: prototypes parsed from a header the binary was not built from would corrupt arguments
        from .build import build_library, built_id, tree_id
        build_library()
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.cdll.grapes_build_id.restype = ctypes.c_char_p
        if self.cdll.grapes_build_id().decode() != tree_id():
            raise GrapesError("libgrapes_b200.so loaded in this process does not match the source tree "
                              f"(library {built_id()[:12]}, tree {tree_id()[:12]}): restart after the rebuild")
        self.protos = parse_header()
        for name, (ret, params) in self.protos.items():
            fn = getattr(self.cdll, name)
            fn.restype = {"char*": ctypes.c_char_p, "int64": ctypes.c_int64}.get(ret, ctypes.c_int)
            fn.argtypes = [_ctype(t) for t, _ in params]
        if os.environ.get("GRAPES_PDL", "") != "":
            self.cdll.grapes_set_pdl(int(os.environ["GRAPES_PDL"], 0))   # bit mask per source file; 0 = plain stream order
        self.launches = 0          # number of C-ABI compute calls issued
        self.profiling = False     # when True every call is bracketed by CUDA events (bench.py breakdown)
        self._events = []
        self.tag = ""              # appended to the profiled name (the engine tags its classifier-sized launches)

    def last_error(self) -> str:
        return self.cdll.grapes_last_error().decode()

    def __getattr__(self, name):
        if not name.startswith("grapes_"):
            raise AttributeError(name)
        fn = getattr(self.cdll, name)
        if self.protos[name][0] != "int":
            return fn

        def call(*args):
            if self.profiling:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args)
                e1.record()
                self._events.append((name + self.tag, e0, e1))
            else:
                rc = fn(*args)
            if rc != 0:
                raise GrapesError(f"{name} failed ({rc}): {self.last_error()}")
            self.launches += 1
        call.__name__ = name
        setattr(self, name, call)
        return call


    def profile_summary(self, reset: bool = True):
        """name -> (total ms, calls) of the event-bracketed calls since the last reset (syncs)."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self._events:
            ms, n = out.get(name, (0.0, 0))
            out[name] = (ms + e0.elapsed_time(e1), n + 1)
        if reset:
            self._events = []
        return out


_LIB = None


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def ptr(t) -> int:
    """device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()
