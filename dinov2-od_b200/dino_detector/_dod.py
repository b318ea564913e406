"""ctypes binding of libdod.so (the C ABI declared in include/dod.h).

The argument structs are generated from the header itself, so the Python side
cannot drift from the C ABI.  There is NO fallback: if the library is missing
or an op fails, a DodError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)                      # dinov2-od_b200/
_REPO_ROOT = os.path.dirname(_PKG_ROOT)
HEADER = os.path.join(_REPO_ROOT, "include", "dod.h")
LIB_PATH = os.environ.get("DOD_LIB", os.path.join(_PKG_ROOT, "lib", "libdod.so"))

DOD_BF16, DOD_F32 = 0, 1
ACT_NONE, ACT_GELU_ERF, ACT_RELU, ACT_SWIGLU = 0, 1, 2, 3


class DodError(RuntimeError):
    pass


_SCALARS = {"int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "float": ctypes.c_float}


def _parse_header(path):
    """-> ({struct_name: [(field, ctype)]}, [function names])"""
    with open(path) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    structs = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"^(const\s+)?(\w+)\s*(\*?)\s*(.*)$", decl)
            base, star, names = m.group(2), m.group(3), m.group(4)
            for nm in names.split(","):
                nm = nm.strip()
                ptr = bool(star)
                while nm.startswith("*"):
                    ptr, nm = True, nm[1:].strip()
                fields.append((nm, ctypes.c_void_p if ptr else _SCALARS[base]))
        structs[name] = fields
    funcs = re.findall(r"DOD_API\s+[\w\s\*]+?\b(dod_\w+)\s*\(", src)
    return structs, funcs


STRUCT_FIELDS, FUNCTIONS = _parse_header(HEADER)


def _make_struct(name, fields):
    return type(name, (ctypes.Structure,), {"_fields_": fields})


STRUCTS = {n: _make_struct(n, f) for n, f in STRUCT_FIELDS.items()}

_lib = None
_lock = threading.Lock()


def lib():
    """Load libdod.so once.  Raises DodError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DodError(
                f"libdod.so not found at {LIB_PATH}: build it with "
                f"`python {os.path.join(_PKG_ROOT, 'build.py')}` (there is no CPU/PyTorch fallback)")
        l = ctypes.CDLL(LIB_PATH)
        l.dod_last_error.restype = ctypes.c_char_p
        l.dod_version.restype = ctypes.c_int32
        l.dod_launch_count.restype = ctypes.c_int64
        l.dod_launch_count_reset.restype = None
        l.dod_device_check.argtypes = [ctypes.c_int32]
        l.dod_device_check.restype = ctypes.c_int32
        for fn in FUNCTIONS:
            sname = fn + "_args"
            alt = {"dod_gemm_bf16": "dod_gemm_args", "dod_patchify14": "dod_patchify_args",
                   "dod_pos_resize_bicubic": "dod_pos_resize_args", "dod_fmha_fwd": "dod_fmha_args",
                   "dod_cast_pad_bf16": "dod_cast_pad_args", "dod_split3_bf16": "dod_split3_args",
                   "dod_lsap_jv": "dod_lsap_args", "dod_transpose_bf16": "dod_transpose_args", "dod_adam_step": "dod_adam_args"}.get(fn, sname)
            if alt in STRUCTS:
                f = getattr(l, fn)
                f.argtypes = [ctypes.POINTER(STRUCTS[alt]), ctypes.c_void_p]
                f.restype = ctypes.c_int32
                _OP_STRUCT[fn] = STRUCTS[alt]
        _lib = l
    return _lib


_OP_STRUCT = {}
_checked_devices = set()


def check_device(index: int):
    if index in _checked_devices:
        return
    rc = lib().dod_device_check(index)
    if rc != 0:
        raise DodError(lib().dod_last_error().decode())
    _checked_devices.add(index)


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    return x.data_ptr()  # torch.Tensor


_OP_CACHE = {}   # fn -> (struct type, {field: is_pointer}, bound C function)


def call(fn: str, stream: int, **fields):
    """Fill dod_<op>_args from keyword arguments and launch on `stream`.  Only the fields given are
    touched (ctypes zero-initialises the rest): this runs once per kernel launch, so it is kept short."""
    ent = _OP_CACHE.get(fn)
    if ent is None:
        l = lib()
        stype = _OP_STRUCT[fn]
        ent = _OP_CACHE[fn] = (stype, {n: t is ctypes.c_void_p for n, t in stype._fields_}, getattr(l, fn))
    stype, is_ptr, cfn = ent
    st = stype()
    try:
        for name, v in fields.items():
            if is_ptr[name]:
                if v is None:
                    continue
                if not isinstance(v, int):
                    v = v.data_ptr()  # torch.Tensor
            setattr(st, name, v)
    except KeyError:
        raise TypeError(f"{fn}: unknown fields {sorted(set(fields) - set(is_ptr))}") from None
    rc = cfn(ctypes.byref(st), stream)
    if rc != 0:
        raise DodError(f"{fn} failed ({rc}): {lib().dod_last_error().decode()}")


def launch_count() -> int:
    return int(lib().dod_launch_count())


def launch_count_reset():
    lib().dod_launch_count_reset()
