"""ctypes binding of libbbbp_b200.so.

Prototypes are derived from include/bbbp_b200.h at import time, so the header is the single
source of truth for the ABI.  There is no fallback: if the library is missing or a symbol
is absent the import fails, and every call that returns a negative status raises
``RuntimeError(bbbp_last_error())``.
"""
from __future__ import annotations

import ctypes
import os
import re

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libbbbp_b200.so")
HEADER_PATH = os.path.join(PKG_DIR, "..", "include", "bbbp_b200.h")

_SCALARS = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t, "uint64_t": ctypes.c_uint64,
    "int64_t": ctypes.c_int64, "long": ctypes.c_longlong, "int32_t": ctypes.c_int32, "bbbp_stream_t": ctypes.c_void_p, "bbbp_comm_t": ctypes.c_void_p,
}
_PROTO = re.compile(r"\n((?:int|size_t|uint64_t|const char\s*\*)\s*)(bbbp_\w+)\s*\(([^;]*?)\)\s*;", re.S)


def parse_header(path: str = HEADER_PATH) -> dict[str, tuple]:
    """{name: (restype, [argtypes], [argnames])} for every prototype in the header."""
    text = re.sub(r"/\*.*?\*/", " ", open(path, "r", encoding="utf-8").read(), flags=re.S)
    out = {}
    for ret, name, args in _PROTO.findall(text):
        ret = ret.strip()
        restype = ctypes.c_char_p if "char" in ret else _SCALARS[ret]
        argtypes, argnames = [], []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                argnames.append(re.findall(r"\w+", a)[-1])
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = [t for t in a.replace("const", " ").split()[:-1]]
                    argtypes.append(_SCALARS[base[-1]])
        out[name] = (restype, argtypes, argnames)
    return out


class LibraryMissing(ImportError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this package)")
    lib = ctypes.CDLL(LIB_PATH)
    protos = parse_header()
    for name, (restype, argtypes, _) in protos.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise LibraryMissing(f"{LIB_PATH} does not export {name} declared in include/bbbp_b200.h") from e
        fn.restype = restype
        fn.argtypes = argtypes
    return lib, protos


lib, PROTOTYPES = _load()
ABI_VERSION = lib.bbbp_abi_version()


def last_error() -> str:
    return (lib.bbbp_last_error() or b"").decode("utf-8", "replace")


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise RuntimeError(f"bbbp_b200 {what}: status {status}: {last_error()}")
