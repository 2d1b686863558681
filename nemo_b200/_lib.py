"""ctypes binding of libb200conformer.so (C ABI declared in include/b200conformer.h).

The product path has NO fallback: if the shared library is missing, or a call returns an error,
an exception is raised.  Tensors are passed as raw device pointers; the CUDA stream is torch's current one.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200conformer.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200conformer.h")


class B2CError(RuntimeError):
    pass


class Epilogue(C.Structure):
    _fields_ = [
        ("out0", C.c_void_p), ("out1", C.c_void_p), ("aux", C.c_void_p), ("bias", C.c_void_p),
        ("lengths", C.c_void_p), ("ldo", C.c_longlong), ("rows_per_batch", C.c_int), ("rows_inner", C.c_int),
        ("alpha", C.c_float), ("seed", C.c_ulonglong), ("drop_site", C.c_uint), ("drop_keep", C.c_uint),
        ("drop_scale", C.c_float),
    ]


class TensorEntry(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("offset", C.c_longlong), ("numel", C.c_longlong), ("flags", C.c_int), ("reserved", C.c_int)]


class PrepEntry(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("A", C.c_int), ("B", C.c_int), ("C", C.c_int), ("flags", C.c_int),
                ("sa", C.c_longlong), ("sb", C.c_longlong), ("sc", C.c_longlong)]


class AdamWHyper(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("noam_d_model", C.c_int), ("noam_warmup", C.c_int), ("noam_min_lr", C.c_float)]


CHUNK = 2048  # elements per block of the multi-tensor kernels (csrc/optim.cu)


def upload_table(entries, numels, device):
    """(table, blocks, n_blocks): a ctypes struct array and its block -> (tensor, chunk) map as device tensors."""
    import torch

    arr = (type(entries[0]) * len(entries))(*entries)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    blocks = []
    for i, n in enumerate(numels):
        blocks += [(i, c) for c in range((int(n) + CHUNK - 1) // CHUNK)]
    bl = torch.tensor(blocks, dtype=torch.int32).to(device)
    return raw, bl, len(blocks)


EPI_BF16, EPI_SILU, EPI_RESID_F32, EPI_GLU_MASK, EPI_DSILU, EPI_RELU_MASK, EPI_ACCUM_F32, EPI_RELU_BWD = range(8)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -m nemo_b200.build` (nvcc, sm_100a). "
        "nemo_b200 has no CPU / PyTorch fallback."
    )
lib = C.CDLL(LIB_PATH)

_CTYPE = {
    "int": C.c_int, "float": C.c_float, "long long": C.c_longlong, "unsigned long long": C.c_ulonglong,
    "unsigned int": C.c_uint, "size_t": C.c_size_t, "int64_t": C.c_int64, "int32_t": C.c_int32, "double": C.c_double,
}


def declared_symbols():
    """[(name, restype, [argtypes])] parsed from the public header (pointers → c_void_p)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = []
    for m in re.finditer(r"^(const char\*|int|long long|size_t)\s+(b2c_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.M | re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(C.c_void_p)
                else:
                    ty = re.sub(r"\b\w+$", "", a).strip()
                    ty = ty.replace("const ", "")
                    argtypes.append(_CTYPE[ty])
        restype = {"const char*": C.c_char_p, "int": C.c_int, "long long": C.c_longlong, "size_t": C.c_size_t}[ret]
        out.append((name, restype, argtypes))
    return out


for _name, _res, _args in declared_symbols():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch (tested in tests/test_abi.py)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.b2c_last_error().decode()


def check(rc: int, what: str = ""):
    if rc != 0:
        raise B2CError(f"{what}: rc={rc}: {last_error()}")


def ptr(t):
    """device pointer of a torch tensor (None → NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Call `b2c_<name>(*args, stream)`; tensors are converted to device pointers."""
    import torch

    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    rc = getattr(lib, "b2c_" + name)(*conv, stream())
    if rc != 0:
        raise B2CError(f"b2c_{name}: rc={rc}: {last_error()}")


def drop_keep(p: float) -> int:
    return 0 if p <= 0.0 else int((1.0 - p) * 65536.0 + 0.5)


def gemm(A, B, M, N, K, lda, ldb, a_mn, b_mn, mode, *, out0, out1=None, aux=None, bias=None, lengths=None, ldo=None,
         rows_per_batch=1, rows_inner=1, alpha=1.0, seed=0, drop_site=0, drop_p=0.0):
    e = Epilogue()
    e.out0, e.out1, e.aux, e.bias, e.lengths = ptr(out0), ptr(out1), ptr(aux), ptr(bias), ptr(lengths)
    e.ldo = N if ldo is None else ldo
    e.rows_per_batch, e.rows_inner, e.alpha = rows_per_batch, rows_inner, alpha
    e.seed, e.drop_site, e.drop_keep = seed, drop_site, drop_keep(drop_p)
    e.drop_scale = 1.0 / (1.0 - drop_p) if drop_p > 0 else 1.0
    rc = lib.b2c_gemm(ptr(A), ptr(B), M, N, K, lda, ldb, a_mn, b_mn, mode, C.byref(e), stream())
    check(rc, "b2c_gemm")
