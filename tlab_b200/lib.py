"""ctypes binding of libtlab_gpu.so.

The prototypes are read from include/tlab_gpu.h, so the Python mirror can never drift from the
C ABI the Fortran host binds (fortran/tlab_gpu_mod.f90).  Loading fails loudly when the library
has not been built; nothing in this package computes on the CPU instead.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "tlab_gpu.h")
LIBPATH = os.path.join(HERE, "libtlab_gpu.so")


class TlabError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tlab_gpu error %d: %s" % (code, msg))
        self.code = code


def _ctype(decl):
    d = decl.strip()
    if "[" in d or "*" in d:
        if re.match(r"^const\s+char\s*\*", d):
            return ctypes.c_char_p
        return ctypes.c_void_p
    t = d.rsplit(" ", 1)[0].replace("const", "").strip() if " " in d else d
    return {"int": ctypes.c_int, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
            "tlab_plan_t": ctypes.c_void_p, "tlab_dns_t": ctypes.c_void_p, "tlab_trp_t": ctypes.c_void_p,
            "long long": ctypes.c_longlong, "int64_t": ctypes.c_int64}[t]


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    for m in re.finditer(r"\b(int|const char\s*\*)\s+(tlab_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes = [] if args in ("void", "") else [_ctype(a) for a in args.split(",")]
        protos[name] = (ctypes.c_char_p if "char" in ret else ctypes.c_int, argtypes)
    return protos


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise ImportError("libtlab_gpu.so is not built (run `python -m tlab_b200.build`); "
                          "tlab_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIBPATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in parse_header().items():
        fn = getattr(lib, name)          # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise TlabError(rc, load().tlab_gpu_last_error().decode())
    return rc
