"""ctypes binding of lib/libadvshadow_b200.so (the C ABI declared in include/advshadow_b200.h).

Only plain pointers and sizes cross this boundary: callers pass `tensor.data_ptr()` and the raw
`cudaStream_t` of torch's current stream.  There is NO fallback: if the shared library is missing
or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libadvshadow_b200.so")

F32, BF16, F16 = 0, 1, 2
CONV_PLAN_BYTES = 2048
ATTN_PLAN_BYTES = 1024


class AdvsError(RuntimeError):
    pass


class ConvSeg(C.Structure):
    _fields_ = [("x", C.c_void_p), ("w", C.c_void_p), ("C", C.c_int32), ("taps", C.c_int32)]


class ConvParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("Cout", C.c_int32), ("stride", C.c_int32), ("nseg", C.c_int32),
        ("seg", ConvSeg * 3),
        ("bias", C.c_void_p), ("temb", C.c_void_p),
        ("temb_stride", C.c_int32), ("out_mode", C.c_int32),
        ("residual", C.c_void_p), ("y", C.c_void_p),
        ("q", C.c_void_p), ("k", C.c_void_p), ("vt", C.c_void_p),
        ("heads", C.c_int32), ("qk_scale", C.c_float),
        ("dtype", C.c_int32), ("cout_valid", C.c_int32),
        ("stats_partial", C.c_void_p),
        ("up_phase", C.c_int32), ("stats_gran", C.c_int32),
        ("y_lo", C.c_void_p),
        ("operand_f16", C.c_int32), ("qkv_dh_pad", C.c_int32),
    ]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/advshadow_b200.h declares
SIGNATURES = {
    "advs_version": (C.c_int, []),
    "advs_last_error": (C.c_char_p, []),
    "advs_device_is_sm100": (C.c_int, []),
    "advs_set_pdl": (C.c_int, [_i]),
    "advs_timestep_embedding": (C.c_int, [_vp, _i, _vp, _i, _vp, _vp]),
    "advs_linear_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "advs_pack_conv_weight": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "advs_pack_upconv_weight": (C.c_int, [_vp, _vp, _i, _i, _i, _vp]),
    "advs_conv3x3_stem": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "advs_conv3x3_head": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "advs_groupnorm_workspace_bytes": (_sz, [_i, _i, _i]),
    "advs_groupnorm_stats": (C.c_int, [_vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "advs_groupnorm_partial_parts": (C.c_int, [_i, _i]),
    "advs_groupnorm_partial": (C.c_int, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "advs_groupnorm_finalize": (C.c_int, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "advs_groupnorm_finalize_ex": (C.c_int, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "advs_stem_im2col": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "advs_stem_im2col_ex": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "advs_pack_stem_weight": (C.c_int, [_vp, _vp, _i, _i, _vp]),
    "advs_pack_stem_weight_ex": (C.c_int, [_vp, _vp, _i, _i, _i, _vp]),
    "advs_conv_sm100_stats_parts": (C.c_int, [_i, _i, _i]),
    "advs_groupnorm_apply": (C.c_int, [_vp, _i, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp]),
    "advs_groupnorm_apply_wide": (C.c_int, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp]),
    "advs_conv_simt": (C.c_int, [C.POINTER(ConvParams), _vp]),
    "advs_conv_sm100_plan": (C.c_int, [C.POINTER(ConvParams), _vp]),
    "advs_conv_sm100_launch": (C.c_int, [_vp, _vp]),
    "advs_selftest_umma_row_shift": (C.c_int, [_i, _i, _vp, _vp]),
    "advs_upsample_nearest2x": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "advs_attention_simt_workspace_bytes": (_sz, [_i, _i, _i]),
    "advs_attention_simt": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _i, _vp]),
    "advs_attention_sm100_plan": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "advs_attention_sm100_plan_ex": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "advs_attention_sm100_launch": (C.c_int, [_vp, _vp]),
    "advs_ddim_step": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _i, _i, _vp]),
    "advs_ddpm_step": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _i, _i, _vp]),
    "advs_select_row": (C.c_int, [_vp, _i, _vp, _vp, _i, _vp]),
    "advs_shadow_disk_mask": (C.c_int, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "advs_gaussian_blur5": (C.c_int, [_vp, _vp, _i, _i, _i, _vp]),
    "advs_shadow_composite": (C.c_int, [_vp, _vp, _vp, _i, _vp, _f, _vp, _vp, _i, _i, _i, _i, _vp]),
    "advs_shadow_composite_generated": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "advs_shadow_composite_generated_ex": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "advs_ddim_step_composite": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i,
                                           _vp]),
    "advs_maxpool2x2": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "advs_upsample_bilinear2x": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "advs_copy_channels": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "advs_layernorm": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _i, _f, _i, _vp]),
    "advs_groupnorm_apply_ex": (C.c_int, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _vp]),
    "advs_groupnorm_apply_ex16": (C.c_int, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _vp]),
    "advs_layernorm_f16out": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _i, _f, _vp]),
    "advs_pos_encoding_ex": (C.c_int, [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "advs_activation": (C.c_int, [_vp, _vp, _sz, _i, _i, _vp]),
    "advs_pos_encoding": (C.c_int, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "advs_cfg_lerp": (C.c_int, [_vp, _vp, _f, _vp, _sz, _vp]),
    "advs_to_uint8": (C.c_int, [_vp, _vp, _sz, _vp]),
    "advs_success_flags": (C.c_int, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises AdvsError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise AdvsError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C <package>/csrc`). There is no fallback path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().advs_last_error().decode("utf-8", "replace")
        raise AdvsError(f"{what or 'advs call'} failed (rc={rc}): {msg}")


def call(name, *args):
    """Invoke an int-returning entry point and raise on a non-zero return code."""
    check(getattr(lib(), name)(*args), name)


class PlanBuffer:
    """64-byte aligned host memory for an opaque launch plan (TMA descriptors + arguments)."""

    def __init__(self, nbytes):
        self._raw = C.create_string_buffer(nbytes + 64)
        addr = C.addressof(self._raw)
        self.ptr = (addr + 63) & ~63
