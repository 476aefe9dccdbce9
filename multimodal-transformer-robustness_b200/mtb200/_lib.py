"""ctypes binding of libmultb200.so (include/multb200.h).  No torch types cross this
boundary: device pointers travel as integers, sizes as ints, the stream as a void*.

The product path has NO CPU fallback: if the shared library is missing the import of this
module raises, and every op raises when handed a non-CUDA tensor."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MTB_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libmultb200.so")   # MTB_LIB: instrumented debug builds

MAX_GROUP = 24
ABI_VERSION = 15


class MtbError(RuntimeError):
    pass


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint64), ("dev", C.c_void_p)]


MAX_SEGS = 16


class Segs(C.Structure):
    _fields_ = [("len", C.c_int32), ("n", C.c_int32), ("seg", C.c_int32 * MAX_SEGS)]


class EmbedDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("sl", C.c_int64), ("sb", C.c_int64), ("se", C.c_int64),
                ("y", C.c_void_p), ("L", C.c_int), ("B", C.c_int), ("E", C.c_int),
                ("scale", C.c_float), ("p", C.c_float), ("rng", Rng)]


class AddNDesc(C.Structure):
    _fields_ = [("src", C.c_void_p * 3), ("ld_src", C.c_int64 * 3), ("n_src", C.c_int),
                ("dst", C.c_void_p), ("ld_dst", C.c_int64), ("T", C.c_int), ("E", C.c_int), ("accumulate", C.c_int),
                ("src_bf16", C.c_int * 3), ("dst_bf16", C.c_int)]


class ResLnDesc(C.Structure):
    _fields_ = [("res", C.c_void_p), ("ld_res", C.c_int64), ("a", C.c_void_p), ("ld_a", C.c_int64),
                ("x_new", C.c_void_p), ("ld_x", C.c_int64), ("y", C.c_void_p), ("ld_y", C.c_int64),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("idx", C.c_void_p),
                ("mean", C.c_void_p), ("rstd", C.c_void_p), ("T", C.c_int), ("E", C.c_int),
                ("eps", C.c_float), ("p", C.c_float), ("rng", Rng), ("a_bf16", C.c_int), ("y_bf16", C.c_int)]


class ResLnBwdDesc(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("ld_dy", C.c_int64), ("d_xnew", C.c_void_p), ("ld_dx", C.c_int64),
                ("x_new", C.c_void_p), ("ld_x", C.c_int64), ("mean", C.c_void_p), ("rstd", C.c_void_p),
                ("gamma", C.c_void_p), ("idx", C.c_void_p),
                ("d_res", C.c_void_p), ("ld_dres", C.c_int64), ("d_a", C.c_void_p), ("ld_da", C.c_int64),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("T", C.c_int), ("E", C.c_int),
                ("p", C.c_float), ("rng", Rng), ("dbias", C.c_void_p), ("dy_bf16", C.c_int), ("da_bf16", C.c_int)]


class LinearDesc(C.Structure):
    _fields_ = [("X", C.c_void_p), ("ldx", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64),
                ("bias", C.c_void_p), ("row_idx", C.c_void_p), ("col_idx", C.c_void_p),
                ("Y", C.c_void_p), ("ldy", C.c_int64), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("act", C.c_int), ("p", C.c_float), ("rng", Rng), ("row_segs", Segs), ("col_segs", Segs),
                ("in_bf16", C.c_int), ("out_bf16", C.c_int)]


class LinearBwdDesc(C.Structure):
    _fields_ = [("dY", C.c_void_p), ("ldy", C.c_int64), ("Yact", C.c_void_p), ("ldyact", C.c_int64),
                ("X", C.c_void_p), ("ldx", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64),
                ("row_idx", C.c_void_p), ("col_idx", C.c_void_p),
                ("dX", C.c_void_p), ("lddx", C.c_int64), ("accumulate_dx", C.c_int),
                ("dW", C.c_void_p), ("db", C.c_void_p), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("act", C.c_int), ("p", C.c_float), ("scratch", C.c_void_p), ("row_segs", Segs), ("col_segs", Segs),
                ("in_bf16", C.c_int), ("dx_bf16", C.c_int)]


class AttnDesc(C.Structure):
    _fields_ = [("q", C.c_void_p), ("ldq", C.c_int64), ("k", C.c_void_p), ("ldk", C.c_int64),
                ("v", C.c_void_p), ("ldv", C.c_int64), ("o", C.c_void_p), ("ldo", C.c_int64),
                ("lse", C.c_void_p), ("Lq", C.c_int), ("Lk", C.c_int), ("B", C.c_int), ("H", C.c_int),
                ("hd", C.c_int), ("scale", C.c_float), ("p", C.c_float), ("rng", Rng), ("keep_bits", C.c_void_p),
                ("bf16", C.c_int)]


class AttnBwdDesc(C.Structure):
    _fields_ = [("q", C.c_void_p), ("ldq", C.c_int64), ("k", C.c_void_p), ("ldk", C.c_int64),
                ("v", C.c_void_p), ("ldv", C.c_int64), ("o", C.c_void_p), ("ldo", C.c_int64),
                ("d_o", C.c_void_p), ("lddo", C.c_int64), ("lse", C.c_void_p), ("delta", C.c_void_p),
                ("dq", C.c_void_p), ("lddq", C.c_int64), ("dk", C.c_void_p), ("lddk", C.c_int64),
                ("dv", C.c_void_p), ("lddv", C.c_int64), ("Lq", C.c_int), ("Lk", C.c_int), ("B", C.c_int),
                ("H", C.c_int), ("hd", C.c_int), ("scale", C.c_float), ("p", C.c_float), ("rng", Rng), ("keep_bits", C.c_void_p),
                ("bf16", C.c_int)]


class AdamDesc(C.Structure):
    _fields_ = [("chunk_param", C.c_void_p), ("chunk_off", C.c_void_p), ("chunk_n", C.c_void_p), ("chunk_pid", C.c_void_p),
                ("active", C.c_void_p), ("steps", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("partial", C.c_void_p), ("scalars", C.c_void_p),
                ("n_chunks", C.c_int), ("n_params", C.c_int), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("weight_decay", C.c_float), ("max_norm", C.c_float), ("shadow", C.c_void_p)]


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("descs", C.c_void_p), ("side", C.c_int32), ("reserved", C.c_int32)]


OP_KIND = {"mtb_embed_fwd": 0, "mtb_embed_bwd": 1, "mtb_addn": 2, "mtb_resln_fwd": 3, "mtb_resln_bwd": 4,
           "mtb_linear_fwd": 5, "mtb_linear_bwd": 6, "mtb_attn_fwd": 7, "mtb_attn_bwd": 8}


# name -> (argtypes, restype); every symbol include/multb200.h declares
SYMBOLS = {
    "mtb_abi_version": ([], C.c_int),
    "mtb_last_error": ([], C.c_char_p),
    "mtb_sm_count": ([], C.c_int),
    "mtb_set_gemm_mode": ([C.c_int], C.c_int),
    "mtb_get_gemm_mode": ([], C.c_int),
    "mtb_set_attn_mode": ([C.c_int], C.c_int),
    "mtb_get_attn_mode": ([], C.c_int),
    "mtb_launch_count": ([], C.c_uint64),
    "mtb_preload": ([], C.c_int),
    "mtb_dropout_mask": ([Rng, C.c_float, C.c_int64, C.c_void_p, C.c_void_p], C.c_int),
    "mtb_rng_advance": ([C.c_void_p, C.c_uint64, C.c_void_p], C.c_int),
    "mtb_embed_fwd": ([C.POINTER(EmbedDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_embed_bwd": ([C.POINTER(EmbedDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_addn": ([C.POINTER(AddNDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_resln_fwd": ([C.POINTER(ResLnDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_resln_bwd": ([C.POINTER(ResLnBwdDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_linear_fwd": ([C.POINTER(LinearDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_linear_bwd": ([C.POINTER(LinearBwdDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_attn_fwd": ([C.POINTER(AttnDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_attn_bwd": ([C.POINTER(AttnBwdDesc), C.c_int, C.c_void_p], C.c_int),
    "mtb_adam_step": ([C.POINTER(AdamDesc), C.c_void_p], C.c_int),
    "mtb_run_ops": ([C.POINTER(OpDesc), C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "mtb_graph_capture": ([C.POINTER(OpDesc), C.c_int, C.c_int, C.POINTER(C.c_void_p)], C.c_int),
    "mtb_graph_launch": ([C.c_void_p, C.c_void_p], C.c_int),
    "mtb_graph_destroy": ([C.c_void_p], C.c_int),
}


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"libmultb200.so not found at {LIB_PATH}; build it with "
            f"`bash {os.path.join(os.path.dirname(_HERE), 'build.sh')}` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = ABI drift between header and library
        fn.argtypes = argtypes
        fn.restype = restype
        fn.mtb_name = name
    v = lib.mtb_abi_version()
    if v != ABI_VERSION:
        raise ImportError(f"libmultb200.so ABI version {v} != binding version {ABI_VERSION}; rebuild")
    return lib


lib = _load()


def check(rc: int, what: str):
    if rc != 0:
        raise MtbError(f"{what} failed ({rc}): {lib.mtb_last_error().decode()}")


def call_group(fn, desc_type, descs, stream: int, what: str):
    """Run one grouped launch per <= MAX_GROUP problems."""
    for i in range(0, len(descs), MAX_GROUP):
        chunk = descs[i:i + MAX_GROUP]
        arr = (desc_type * len(chunk))(*chunk)
        check(fn(arr, len(chunk), C.c_void_p(stream)), what)
