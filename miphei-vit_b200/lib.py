"""ctypes binding of libmiphei_b200.so (the C ABI declared in include/miphei_b200.h).

There is no fallback: if the shared library is missing, or no sm_100 device is visible, the first compute call raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# MIPHEI_B200_LIB selects another build of the SAME library (the diagnostic twin libmiphei_b200_prof.so) — never a fallback
LIB_PATH = os.environ.get("MIPHEI_B200_LIB") or os.path.join(_HERE, "libmiphei_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_float = ctypes.c_float


class GemmArgs(ctypes.Structure):
    """mv_gemm_args (include/miphei_b200.h)."""

    _fields_ = [
        ("a", c_void_p), ("lda", c_i64),
        ("b", c_void_p), ("ldb", c_i64),
        ("m", c_i32), ("n", c_i32), ("k", c_i32),
        ("mode", c_i32), ("act", c_i32), ("out_f32", c_i32),
        ("out", c_void_p), ("ldo", c_i64),
        ("aux", c_void_p), ("ldaux", c_i64),
        ("scale", c_void_p), ("shift", c_void_p),
        ("resid", c_void_p), ("ldr", c_i64),
        ("in2", c_void_p), ("ldin2", c_i64),
        ("rows_per_group", c_i32), ("group_stride", c_i32), ("row_offset", c_i32), ("resid_row_mod", c_i32),
        ("block_n", c_i32), ("conv", c_i32),
        ("a2", c_void_p),
        ("conv_batch", c_i32), ("conv_h", c_i32), ("conv_w", c_i32), ("conv_stride", c_i32),
        ("conv_c0", c_i32), ("conv_c1", c_i32),
        ("reserved_splits", c_i32), ("reserved2", c_i32),
        ("colstats", c_void_p),
        ("kskip_begin", c_i32), ("kskip_end", c_i32),
        ("ab_f16", c_i32), ("reserved3", c_i32),
        ("workspace", c_void_p), ("workspace_bytes", c_i64),
    ]


# name -> (restype, argtypes); every symbol include/miphei_b200.h declares must appear here (tests check both ways)
_SIGNATURES = {
    "mv_init": (c_int, [c_int]),
    "mv_last_error": (ctypes.c_char_p, []),
    "mv_version": (c_int, []),
    "mv_num_sms": (c_int, []),
    "mv_launch_count": (c_i64, []),
    "mv_reset_launch_count": (None, []),
    "mv_gemm_bf16": (c_int, [ctypes.POINTER(GemmArgs), c_void_p]),
    "mv_gemm_set_profile_buffer": (None, [c_void_p]),
    "mv_attn_set_profile_buffer": (None, [c_void_p]),
    "mv_layernorm_fwd": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_int,
                                 c_int, c_float, c_void_p]),
    "mv_layernorm_bwd": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_void_p, c_i64, c_void_p, c_i64,
                                 c_void_p, c_i64, c_int, c_int, c_float, c_void_p]),
    "mv_attn_fwd": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "mv_prep_input": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mv_prep_input_u8": (c_int, [c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_void_p, c_int, c_void_p, c_int,
                                 c_int, c_int, c_void_p]),
    "mv_fill_prefix": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mv_tokens_to_map": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mv_upsample2x": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mv_tokens_to_map_bwd": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mv_attn_bwd": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64,
                            c_int, c_int, c_int, c_float, c_void_p]),
    "mv_lora_grads_workspace_bytes": (c_i64, [c_int, c_int]),
    "mv_lora_grads": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_i64, c_void_p]),
    "mv_bn_finalize": (c_int, [c_void_p, ctypes.c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                               c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mv_gram32": (c_int, [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p]),
    "mv_heads_bn_from_gram": (c_int, [c_void_p, ctypes.c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mv_heads_bwd_algebra": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double,
                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                     c_void_p, c_void_p]),
    "mv_adam_schedule": (c_int, [c_void_p, c_float, c_i64, c_i64, c_float, c_float, c_void_p, c_void_p]),
    "mv_adam_clip_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_float, c_void_p, c_float,
                                      c_float, c_float, c_void_p]),
    "mv_lora_refresh": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_i64, c_i64, c_void_p]),
    "mv_gather_cast": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "mv_add_i64": (c_int, [c_void_p, c_int, c_i64, c_void_p]),
    "mv_memset_async": (c_int, [c_void_p, c_int, c_i64, c_void_p]),
    "mv_bn_relu_apply": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_void_p]),
    "mv_bn_relu_bwd": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_i64, c_int, c_void_p]),
    "mv_transpose_bf16": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_int, c_int, c_int, c_void_p]),
    "mv_f16_to_bf16": (c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "mv_upsample2x_bwd": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mv_zero_insert2x": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mv_add_bf16": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_void_p]),
    "mv_heads_ds": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mv_heads_bwd_stencil": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_void_p]),
    "mv_cell_means": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "mv_cell_means_workspace_bytes": (c_i64, [c_int, c_int, c_int]),
    "mv_cell_means_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p]),
    "mv_cell_means_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "mv_thumb_std_hist": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "mv_otsu_threshold": (c_int, [c_void_p, c_i64, c_void_p, c_void_p]),
    "mv_tile_tissue": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "mv_stitch_tiles": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_i64, c_i64, c_int, c_void_p]),
    "mv_host_alloc_mapped": (c_int, [c_i64, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "mv_host_free": (c_int, [c_void_p]),
    "mv_host_register": (c_int, [c_void_p, c_i64]),
    "mv_host_unregister": (c_int, [c_void_p]),
    "mv_loss_workspace_floats": (c_i64, [c_int, c_int, c_int]),
    "mv_loss_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                c_void_p, c_void_p, c_i64, c_void_p]),
    "mv_grad_norm": (c_int, [c_void_p, c_i64, c_float, c_void_p, c_void_p, c_void_p]),
    "mv_adam_clip_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_float, c_float, c_float,
                                  c_float, c_float, c_int, c_void_p]),
}

_lib = None
_lock = threading.Lock()
_inited = set()


class MipheiB200Error(RuntimeError):
    pass


def load():
    """dlopen the library and bind signatures (no device needed)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MipheiB200Error(
                "%s not found: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
                "miphei_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def last_error():
    return load().mv_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise MipheiB200Error("%s failed (code %d): %s" % (what, rc, last_error()))


def init(device=0):
    """Bind the library to a CUDA device. Raises if no sm_100 GPU is visible."""
    lib = load()
    if device not in _inited:
        check(lib.mv_init(int(device)), "mv_init(%d)" % device)
        _inited.add(device)
    return lib


def launch_count():
    return int(load().mv_launch_count())


def reset_launch_count():
    load().mv_reset_launch_count()
