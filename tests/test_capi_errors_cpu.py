"""C-ABI error behaviour and host-side queries, without a GPU: every launch function validates its arguments BEFORE
it touches CUDA, so bad calls must come back as ADVS_ERR_ARG with a message naming the entry point, and the pure
host queries (workspace sizes, partial-row counts, the PDL switch) can be exercised anywhere."""
import ctypes as C

import pytest


@pytest.fixture(scope="module")
def L():
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import _capi as capi
    return capi.lib()


def _err(L):
    return L.advs_last_error().decode()


def test_bad_arguments_are_rejected_before_any_launch(L):
    from advshadow_b200 import _capi as capi
    ERR_ARG = -1
    assert L.advs_ddim_step(None, None, None, None, 0, None, None, 0, 1, None) == ERR_ARG and "ddim_step" in _err(L)
    # float4 path: x / eps / out (and noise) must be 16-byte aligned
    assert L.advs_ddim_step(0x1004, 0x2000, None, 0x1004, 16, 0x3000, 0x4000, 0, 1, None) == ERR_ARG and "aligned" in _err(L)
    assert L.advs_ddpm_step(None, None, None, None, 4, None, None, 0, 1, None) == ERR_ARG and "ddpm_step" in _err(L)
    assert L.advs_gaussian_blur5(0x1000, 0x1000, 1, 8, 8, None) == ERR_ARG and "in-place" in _err(L)
    assert L.advs_shadow_composite(0x10, 0x10, 0x10, 2, None, 0.33, 0x10, 0x10, 1, 3, 8, 8, None) == ERR_ARG and "Cm" in _err(L)
    assert L.advs_success_flags(None, None, 0, 0, None, None, None) == ERR_ARG and "success_flags" in _err(L)
    assert L.advs_groupnorm_stats(0x1000, 12, None, 0, 2, 64, 4, 1e-5, 0x1, 0x1, 0x1, 0x1, 1 << 20, capi.BF16, None) == ERR_ARG
    assert "multiples of 8" in _err(L)
    assert L.advs_groupnorm_stats(0x1000, 64, None, 0, 2, 64, 32, 1e-5, 0x1, 0x1, 0x1, 0x1, 16, capi.BF16, None) == ERR_ARG
    assert "workspace too small" in _err(L)
    assert L.advs_stem_im2col_ex(0x1000, 0x1000, 1, 8, 8, 8, capi.BF16, None) == ERR_ARG and "9*Cin" in _err(L)
    assert L.advs_stem_im2col_ex(0x1000, 0x1000, 1, 8, 8, 3, capi.F32, None) == ERR_ARG and "dtype" in _err(L)
    # attention: the v^T rows are TMA rows, T must be a multiple of 8
    assert L.advs_attention_sm100_plan_ex(0x1000, 0x1000, 0x1000, 0x1000, 1, 4, 100, 64, 64, 0x1000) == ERR_ARG
    assert "multiple of 8" in _err(L)
    p = capi.ConvParams()
    assert L.advs_conv_sm100_plan(C.byref(p), None) == ERR_ARG and "conv_sm100_plan" in _err(L)
    assert L.advs_conv_simt(C.byref(p), None) == ERR_ARG and "conv_simt" in _err(L)
    # an error text is the calling thread's and survives a later successful host query
    L.advs_version()
    assert "conv_simt" in _err(L)


def test_host_queries(L):
    assert L.advs_version() >= 100
    assert L.advs_device_is_sm100() in (0, 1)                         # 0 here: no device
    # GroupNorm workspace grows linearly with the batch; the statistics chunking itself does not depend on it
    w1, w2 = L.advs_groupnorm_workspace_bytes(1, 4096, 128), L.advs_groupnorm_workspace_bytes(2, 4096, 128)
    assert w1 > 0 and w2 == 2 * w1
    for hw in (64, 1024, 4096, 65536):
        assert len({L.advs_groupnorm_partial_parts(b, hw) for b in (1, 2, 7, 64)}) == 1
        assert 1 <= L.advs_groupnorm_partial_parts(1, hw) <= 64
    # conv-epilogue statistics rows: one per 128-pixel tile, independent of the batch (bit-reproducible trajectories
    # across batch sizes); images smaller than one tile have none
    assert [L.advs_conv_sm100_stats_parts(b, 256, 256) for b in (1, 2, 64)] == [512] * 3
    assert [L.advs_conv_sm100_stats_parts(b, 8, 8) for b in (1, 64)] == [0, 0]
    assert L.advs_attention_simt_workspace_bytes(1, 4, 64) == 4 * 64 * 64 * 4


def test_pdl_switch_roundtrip(L):
    first = L.advs_set_pdl(1)
    assert L.advs_set_pdl(0) == 1 and L.advs_set_pdl(-1) == 0 and L.advs_set_pdl(first) == -1
