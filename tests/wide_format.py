"""TEST INFRASTRUCTURE: host restatement (torch integer ops) of the "wide" pre-norm storage format of
include/advshadow_b200.h (advs_conv_params.y_lo): value = as_float((bits(bf16) << 16) + (int8 << 8))."""
import torch


def wide_decode(hi_bf16, lo_i8):
    """the decode rule stated in the header; what k_gn_apply<..., WIDE> computes per element"""
    bits = (hi_bf16.view(torch.int16).to(torch.int32) << 16) + (lo_i8.to(torch.int32) << 8)
    return bits.view(torch.float32)


def wide_encode(x_f32):
    """the conv epilogue's encoder (csrc/common.cuh): u = bits(x) + 0x8000; hi = u >> 16 (bf16, nearest, ties away from
    zero); lo = int8(byte 1 of u) - 128"""
    u = x_f32.contiguous().view(torch.int32) + 0x8000
    hi = (u >> 16).to(torch.int16).view(torch.bfloat16)
    lo = (((u >> 8) & 0xFF) - 128).to(torch.int8)
    return hi, lo
