/*
 * advshadow_b200 -- C ABI of the B200-native DDIM shadow sampler hot path.
 *
 * The reference (Raineasy/AdvShadow, pure Python/PyTorch) has no FFI of its own;
 * its boundary is the Python surface of diff_model.py / ddim2/diff_model2.py.
 * Each entry point below replaces one group of PyTorch calls on that surface and
 * cites it (file:line relative to the reference root; dm1 = diff_model.py,
 * dm2 = ddim2/diff_model2.py, ts = tools/train_shadow.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - no allocation, no synchronisation, no host callbacks inside any launch
 *     function: all of them are CUDA-graph capturable; scratch memory is
 *     supplied by the caller;
 *   - return value 0 = ok, <0 = error (advs_last_error() gives the text; the message buffer is
 *     thread-local, so it is the last error of the CALLING thread and stays valid until that thread's
 *     next failing call);
 *   - host threads: plans are immutable after *_plan returns and launches keep no host state, so
 *     several threads may launch concurrently -- with one exception: the FIRST launch of each kernel
 *     family on a device sets that kernel's shared-memory attribute; make it from one thread (a
 *     warm-up call, which graph capture needs anyway) before launching the family from others.
 *     Two launches on the same scratch / state buffers are the caller's to order, as with any
 *     stream work;
 *   - activations are NHWC ("pixel-major"): element (b,h,w,c) of a [B,H,W,C]
 *     tensor lives at ((b*H+h)*W+w)*C+c.  Sampler state x_t / eps stay in the
 *     reference's NCHW fp32 layout;
 *   - `dtype`: ADVS_F32 (fp32 storage, SIMT fp32 math: the <=1e-4 mode) or
 *     ADVS_BF16 (bf16 storage, fp32 accumulate: the throughput mode).
 */
#ifndef ADVSHADOW_B200_H
#define ADVSHADOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVS_F32 0
#define ADVS_BF16 1
/* fp16 is an OPERAND format only (packed weights, GroupNorm outputs): values that are bounded by construction keep
 * three more mantissa bits than in bf16 at the same tensor-core rate (tcgen05 kind::f16 takes either format per
 * operand); activations whose range is not bounded (conv outputs, the residual stream, q/k/v) stay bf16. */
#define ADVS_F16 2

#define ADVS_OK 0
#define ADVS_ERR_ARG (-1)
#define ADVS_ERR_CUDA (-2)
#define ADVS_ERR_UNSUPPORTED (-3)

/* ---- library ---------------------------------------------------------------------------- */
int advs_version(void);
const char* advs_last_error(void);
/* 1 if the device behind the current context is compute capability 10.x */
int advs_device_is_sm100(void);
/* Programmatic dependent launch between the library's kernels (each kernel's prologue overlaps the previous
 * kernel's tail; a captured CUDA graph keeps the setting it was captured with).  on: 1 / 0, or -1 = follow the
 * environment (ADVS_PDL=1, default off).  Returns the previous setting.  Pays at batch 1-2 only. */
int advs_set_pdl(int on);

/* ---- K6: timestep embedding + small linears (dm1:16-33, dm1:184-188, dm1:77-80) ----------- */
/* out[i, 0:half] = cos(t_i * f_j), out[i, half:2*half] = sin(t_i * f_j)  (dm1:25-30, "cos first").
 * freqs[half] is the reference's table exp(-ln(max_period) * arange(half) / half); the reference
 * always evaluates it on the host (dm1:25-28: torch.arange has no device argument) and uploads
 * it, so the host side does the same and passes the device copy here. */
int advs_timestep_embedding(const int64_t* t, int nt, const float* freqs, int half, float* out,
                            void* stream);
/* y[r, o] = act_out( sum_i act_in(x[r, i]) * w[o, i] + b[o] ), act = SiLU when the flag is set.
 * Replaces nn.Linear / nn.SiLU in UNetModel.time_embed (dm1:184-188) and
 * ResidualBlock.time_emb (dm1:77-80).  fp32 throughout. */
int advs_linear_f32(const float* x, const float* w, const float* b, float* y, int rows, int in_f,
                    int out_f, int silu_in, int silu_out, void* stream);

/* ---- weight packing --------------------------------------------------------------------- */
/* nn.Conv2d weight OIHW fp32 -> [O][kh*kw][I] in `dtype` (K-major rows for the implicit GEMM); dtype may be
 * ADVS_F16 for the tcgen05 path (advs_conv_params.operand_f16 bit 2). */
int advs_pack_conv_weight(const float* w_oihw, void* dst, int O, int I, int kh, int kw, int dtype,
                          void* stream);

/* 3x3 weight OIHW fp32 -> four phase kernels [phase = 2a+b][O][2*2][I]: tap (p,q) of phase (a,b) is the sum
 * of the 3x3 taps that read the same low-res pixel after nearest-2x upsampling
 * (rows: a=0: {0},{1,2}; a=1: {0,1},{2}; same for columns). */
int advs_pack_upconv_weight(const float* w_oihw, void* dst, int O, int I, int dtype, void* stream);

/* ---- K1 edge layers: Cin=3 stem and Cout=3 head (dm1:192, dm1:240-242) -------------------- */
/* x NCHW fp32 [B,Cin,H,W], w fp32 [Cout][9][Cin], y NHWC `dtype` [B,H,W,Cout]. 3x3, pad 1. */
int advs_conv3x3_stem(const float* x_nchw, const float* w, const float* bias, void* y, int B,
                      int H, int W, int Cin, int Cout, int dtype, void* stream);
/* The same stem on the tensor cores (9*Cin <= 64): advs_stem_im2col turns the fp32 NCHW input into bf16 rows
 * col[B,H,W,64] (entry = part*9*Cin + tap*Cin + ci; part 0 = bf16(x), part 1 = bf16(x - bf16(x)) when 18*Cin <= 64,
 * rest 0), advs_pack_stem_weight writes the matching bf16 rows [Cout][64]; the stem is then a 1x1 convolution
 * (advs_conv_sm100_plan with one segment, C = 64, taps = 1) with the usual fused epilogue. */
int advs_stem_im2col(const float* x_nchw, void* col, int B, int H, int W, int Cin, void* stream);
/* dtype = ADVS_F16: fp16 rows (hi = fp16(x), lo = fp16(x - hi)) for fp16 stem weights; the sampler state is bounded */
int advs_stem_im2col_ex(const float* x_nchw, void* col, int B, int H, int W, int Cin, int dtype, void* stream);
int advs_pack_stem_weight(const float* w_oihw, void* dst, int O, int I, void* stream);
int advs_pack_stem_weight_ex(const float* w_oihw, void* dst, int O, int I, int dtype, void* stream); /* ADVS_BF16 | ADVS_F16 */
/* x NHWC `dtype` [B,H,W,Cin], w fp32 [Cout][9][Cin], y NCHW fp32 [B,Cout,H,W]. 3x3, pad 1. */
int advs_conv3x3_head(const void* x, const float* w, const float* bias, float* y_nchw, int B,
                      int H, int W, int Cin, int Cout, int dtype, void* stream);

/* ---- K5: GroupNorm(32,C) [+ SiLU] over NHWC, optionally over a virtual concat of two
 *      sources (dm1:62-63, 71-72, 83-84, 113, 240-241; torch.cat at dm1:265) ----------------- */
/* scratch floats needed by advs_groupnorm_stats */
size_t advs_groupnorm_workspace_bytes(int B, int HW, int C);
/* per-(b,channel) affine  y = x*scale + shift  with scale = rstd*gamma, shift = beta-mean*rstd*gamma
 * written to scale_shift[B][C][2].  x1 may be NULL (c1 = 0). */
int advs_groupnorm_stats(const void* x0, int c0, const void* x1, int c1, int B, int HW, int groups,
                         float eps, const float* gamma, const float* beta, float* scale_shift,
                         void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* The two halves of advs_groupnorm_stats, usable separately:
 * advs_groupnorm_partial: per-(image, pixel-chunk, channel) {sum, sumsq} of ONE source,
 *   part[B][advs_groupnorm_partial_parts(B,HW)][C][2];
 * advs_groupnorm_finalize: combines the partial rows of up to two concatenated sources (from
 *   advs_groupnorm_partial or from a conv epilogue's stats_partial) into scale_shift[B][c0+c1][2]. */
int advs_groupnorm_partial_parts(int B, int HW);
int advs_groupnorm_partial(const void* x, int C, int B, int HW, float* part, int dtype, void* stream);
int advs_groupnorm_finalize(const float* part0, int c0, int parts0, const float* part1, int c1, int parts1,
                            int B, int HW, int groups, float eps, const float* gamma, const float* beta,
                            float* scale_shift, void* stream);
/* same, for partial rows that hold one pair per gran_s consecutive channels (gran_s = 1 or 4; see
 * advs_conv_params.stats_gran): part_s is [B][parts_s][c_s / gran_s][2]. */
int advs_groupnorm_finalize_ex(const float* part0, int c0, int parts0, int gran0, const float* part1, int c1, int parts1,
                               int gran1, int B, int HW, int groups, float eps, const float* gamma, const float* beta,
                               float* scale_shift, void* stream);
int advs_groupnorm_apply(const void* x0, int c0, const void* x1, int c1, int B, int HW,
                         const float* scale_shift, int silu, void* y, int dtype, void* stream);
/* bf16 only: same, where lo0 / lo1 (each may be NULL) are the int8 mantissa extensions written by a conv epilogue
 * through advs_conv_params.y_lo for x0 / x1. */
int advs_groupnorm_apply_wide(const void* x0, const void* lo0, int c0, const void* x1, const void* lo1, int c1, int B,
                              int HW, const float* scale_shift, int silu, void* y, int y_dtype, void* stream);
/* y_dtype: ADVS_BF16, or ADVS_F16 when the consumer is a tcgen05 conv with operand_f16 bit 0 set (a normalised
 * tensor is bounded by sqrt(group size) * |gamma| + |beta|, far inside the fp16 range). */

/* ---- K1/K2/K3: convolution as implicit GEMM (dm1:73, 86, 90, 114-115, 134, 148) ----------- */
/* D[pixel, cout] = sum over K-segments s, taps, channels of  X_s[pixel + tap, c] * W_s[cout, tap, c]
 * followed by the fused epilogue  + bias[cout] + temb[b, cout] + residual[pixel, cout].
 *   segment 0      : the convolution proper (3x3 pad 1, stride 1 or 2; or 1x1)
 *   segments 1..2  : optional 1x1 "shortcut" convolutions of the block input accumulated into
 *                    the same output (ResidualBlock.shortcut, dm1:89-92,103; two segments when
 *                    the block input is the virtual concat [h, skip], dm1:265).
 * out_mode 0: y NHWC [B,H,W,Cout].
 * out_mode 1: "qkv split" for AttentionBlock (dm1:114,120-121): cout = head*3*dh + which*dh + d;
 *             q,k -> [B,heads,T,dh] multiplied by qk_scale (= dh^-1/4), v -> vt [B,heads,dh,T].
 * out_mode 2: y is fp32 NCHW [B,cout_valid,H,W] (the UNet head, dm1:240-243): only the first
 *             cout_valid output channels are stored; Cout may be zero-padded up to the GEMM tile. */
typedef struct advs_conv_seg {
  const void* x; /* NHWC [B, Hin, Win, C]; Hin = H*stride for segment 0, H otherwise */
  const void* w; /* [Cout][taps][C] in dtype */
  int32_t C;
  int32_t taps; /* 9 (3x3 pad 1), 1 (1x1) or 4 (2x2 phase of an upsample-conv, see up_phase) */
} advs_conv_seg;

typedef struct advs_conv_params {
  int32_t B, H, W; /* output spatial size */
  int32_t Cout;
  int32_t stride; /* 1 or 2, segment 0 only */
  int32_t nseg;
  advs_conv_seg seg[3];
  const float* bias;      /* [Cout] or NULL */
  const float* temb;      /* [B or 1][temb_stride] or NULL: per-image per-channel bias (dm1:101) */
  int32_t temb_stride;    /* floats between images; 0 = broadcast one row */
  int32_t out_mode;       /* 0 NHWC, 1 qkv split, 2 fp32 NCHW */
  const void* residual;   /* NHWC [B,H,W,Cout] or NULL */
  void* y;                /* out_mode 0 */
  void* q;                /* out_mode 1 */
  void* k;
  void* vt;
  int32_t heads;
  float qk_scale;
  int32_t dtype;
  int32_t cout_valid;     /* out_mode 2: number of real output channels (<= Cout) */
  /* optional (sm100 path, out_mode 0): per-tile, per-channel {sum, sum of squares} of the stored output,
   * layout [B][advs_conv_sm100_stats_parts(B,H,W)][Cout][2] fp32 -- the GroupNorm statistics of the
   * tensor come for free from the conv epilogue (feed to advs_groupnorm_finalize) */
  float* stats_partial;
  /* Upsample(nearest 2x) + conv3x3 (dm1:137-139) as four 2x2 "phase" convolutions on the LOW-RES
   * input: up_phase = 1 + 2*a + b selects output pixels (2h+a, 2w+b); 0 = ordinary convolution.
   * With up_phase != 0: B,H,W are the INPUT size, y is [B,2H,2W,Cout], segment 0 has taps = 4 and
   * weights from advs_pack_upconv_weight; stats_partial rows are [B][4 * parts][Cout][2]. */
  int32_t up_phase;
  /* granularity of stats_partial: 0 or 1 = one {sum, sumsq} pair per output channel (layout above);
   * 4 = one pair per 4 consecutive channels, layout [B][parts][Cout/4][2] -- a quarter of the partial traffic and
   * a third of the epilogue's shuffles; usable whenever every GroupNorm reading the tensor has a multiple of 4
   * channels per group (pass the same value as gran to advs_groupnorm_finalize_ex). */
  int32_t stats_gran;
  /* optional (out_mode 0, ADVS_BF16): "wide" storage for a tensor that a GroupNorm will read.  y stays a
   * round-to-nearest bf16 tensor (ties away from zero; GEMM operand / residual for every other consumer);
   * y_lo[B,H,W,Cout] int8 holds the next 8 mantissa bits: value ~= as_float((bits(y) << 16) + ((int)y_lo << 8)),
   * exact to 2^-15 relative (2^-16 on average).
   * advs_groupnorm_apply_wide reads the pair, so the normalised GEMM operand is rounded once instead of twice
   * (the reference normalises fp32 tensors, dm1:71-72, 83-84). */
  void* y_lo;
  /* sm100 path: 0 = every GEMM operand is bf16; 5 (bits 0 and 2) = segment 0 is an fp16 x fp16 GEMM: its activations
   * are a bounded tensor written as ADVS_F16 (advs_groupnorm_apply_wide / advs_stem_im2col_ex) and its weights were
   * packed as ADVS_F16.  The shortcut segments always stay bf16 x bf16; a segment whose two operands differ in
   * format cannot be expressed (the hardware rejects such an MMA). */
  int32_t operand_f16;
  /* out_mode 1: storage head dim of q / k / vt when it is larger than dh = Cout / (3*heads) (0 = dh): the epilogue
   * writes the dh real columns of each head into rows of qkv_dh_pad elements (q, k) / into the first dh of
   * qkv_dh_pad rows (vt); the caller zero-fills the buffers once.  See advs_attention_sm100_plan_ex. */
  int32_t qkv_dh_pad;
} advs_conv_params;

/* generic SIMT fp32-accumulate implementation: any dtype, any channel counts (multiple of 4) */
int advs_conv_simt(const advs_conv_params* p, void* stream);

/* sm_100a tcgen05/TMEM/TMA implementation, bf16 only. C and Cout multiples of 64.
 * The plan holds the TMA descriptors (they embed the tensor addresses), so it must be
 * re-created when any pointer in `p` changes.  ADVS_CONV_PLAN_BYTES bytes, 64-byte aligned,
 * caller-owned host memory. */
#define ADVS_CONV_PLAN_BYTES 2048
/* number of per-image partial rows the conv epilogue writes into stats_partial (0 = unsupported shape) */
int advs_conv_sm100_stats_parts(int B, int H, int W);
int advs_conv_sm100_plan(const advs_conv_params* p, void* plan_host);
int advs_conv_sm100_launch(const void* plan_host, void* stream);

/* hardware self-test used by the test-suite: D[i][j] = A[r0+i][j] through a row-shifted SWIZZLE_128B UMMA
 * descriptor (out: fp32 [128][64]); base_offset_mode selects the descriptor's base-offset convention */
int advs_selftest_umma_row_shift(int r0, int base_offset_mode, float* out, void* stream);

/* ---- K7: nearest 2x upsample (dm1:137) ---------------------------------------------------- */
int advs_upsample_nearest2x(const void* x, void* y, int B, int H, int W, int C, int dtype,
                            void* stream);

/* ---- K4: self-attention core softmax(q^T k) v (dm1:122-125) ------------------------------- */
/* q,k [B,heads,T,dh] (already scaled), vt [B,heads,dh,T], o NHWC [B,T,heads*dh]. */
size_t advs_attention_simt_workspace_bytes(int B, int heads, int T);
int advs_attention_simt(const void* q, const void* k, const void* vt, void* o, int B, int heads,
                        int T, int dh, void* workspace, size_t workspace_bytes, int dtype,
                        void* stream);
/* sm_100a flash-attention (tcgen05 QK^T and PV, online softmax in fp32), bf16, dh in {64,128,256},
 * T a multiple of 8 (a partial last key block is masked in-kernel).
 * _plan_ex: head dims below 64 (IDDM's nn.MultiheadAttention with 4 heads: 16 and 32, model/modules/attention.py:27)
 * run with dh = 64 storage -- q, k [B,heads,T,64] and vt [B,heads,64,T] zero beyond dh_valid (advs_conv_params.
 * qkv_dh_pad makes the qkv epilogue write them so) -- and o [B,T,heads*dh_valid] packed. */
#define ADVS_ATTN_PLAN_BYTES 1024
int advs_attention_sm100_plan(const void* q, const void* k, const void* vt, void* o, int B,
                              int heads, int T, int dh, void* plan_host);
int advs_attention_sm100_plan_ex(const void* q, const void* k, const void* vt, void* o, int B,
                                 int heads, int T, int dh, int dh_valid, void* plan_host);
int advs_attention_sm100_launch(const void* plan_host, void* stream);

/* ---- K8: DDIM / DDPM reverse-step updates (dm1:449-472, dm1:356-395) ---------------------- */
/* coef rows (fp32, device): [sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma, pad..]
 * 8 floats per step.  x' = sqrt(a_prev)*clamp((x - sqrt(1-a_t)*eps)/sqrt(a_t)) + dir*eps + sigma*z
 * evaluated with exactly the reference's fp32 operation order (no FMA contraction).
 * step_dev: device int32 holding the row to use (advanced by `advance` after use, so a captured
 * CUDA graph of one step can be replayed).  noise may be NULL (sigma*z := 0). */
int advs_ddim_step(const float* x, const float* eps, const float* noise, float* x_out,
                   size_t n_elems, const float* coef, int32_t* step_dev, int advance,
                   int clip_denoised, void* stream);
/* DDPM ancestral step. coef rows: [sqrt_recip_acp, sqrt_recipm1_acp, post_mean_coef1,
 * post_mean_coef2, exp(0.5*post_log_var)*(t!=0), pad..] 8 floats per step. */
int advs_ddpm_step(const float* x, const float* eps, const float* noise, float* x_out,
                   size_t n_elems, const float* coef, int32_t* step_dev, int advance,
                   int clip_denoised, void* stream);
/* copy row *step_dev of table[rows][row_floats] to each of the `reps` rows of dst[reps][row_floats]
 * (per-step timestep-embedding select: ddim_sample feeds the same t to every image, dm1:446) */
int advs_select_row(const float* table, int row_floats, const int32_t* step_dev, float* dst,
                    int reps, void* stream);

/* ---- K9: shadow mask + compositing (dm2:552-570, dm2:615-654, ts:147-174, ts:224-266) ------ */
/* mask[b,h,w] = (sqrt((w - cx_b)^2 + (h - cy_b)^2) <= r_b) ? 1 : 0, centers[b] = (c0,c1) used as
 * (x,y) exactly like dm2:567-569; fp32 arithmetic in the reference's order. */
int advs_shadow_disk_mask(const float* centers, const float* radii, int B, int H, int W,
                          float* mask, void* stream);
/* cv2.GaussianBlur(mask,(5,5),0): separable [1,4,6,4,1]/16, BORDER_REFLECT_101 (ts:147-153) */
int advs_gaussian_blur5(const float* src, float* dst, int B, int H, int W, void* stream);
/* combined = shadow_mask * feature_mask (feature_mask [B,Cm,H,W], Cm = 1 or C, broadcast);
 * shadowed = img*(1-m) + m*(img*(1-intensity))                                   (dm2:642-645)
 * out      = clamp(img*(1-m) + adv*m, 0, 1), adv = `adv` if given else shadowed   (dm2:650-653)
 * img/adv/out NCHW fp32 [B,C,H,W].  shadowed_out / out may be NULL to skip that output.
 * one_minus_intensity = (float)(1.0 - intensity) evaluated by the host in double like Python does. */
int advs_shadow_composite(const float* img, const float* shadow_mask, const float* feature_mask,
                          int Cm, const float* adv, float one_minus_intensity, float* shadowed_out,
                          float* out, int B, int C, int H, int W, void* stream);
/* fused last step of the shadow sampler: g = clip(x_final, 0, 1) (main.py:135 mapping) is used as
 * `adv` of advs_shadow_composite; mask is built in-kernel from (centers, radii). */
int advs_shadow_composite_generated(const float* img, const float* x_final, const float* centers,
                                    const float* radii, const float* feature_mask, int Cm,
                                    float* out, int B, int C, int H, int W, void* stream);
/* same with the mask flavour selectable: blur = 1 applies the 5x5 Gaussian to the disk mask in-kernel before the
 * feature mask multiplies it (tools/train_shadow.py:244-247, ddim2/test.py:851-854); blur = 0 is dm2:634-642. */
int advs_shadow_composite_generated_ex(const float* img, const float* x_final, const float* centers,
                                       const float* radii, const float* feature_mask, int Cm, int blur,
                                       float* out, int B, int C, int H, int W, void* stream);
/* The sampler's fused tail: the LAST DDIM update (advs_ddim_step with noise = NULL; x_out may alias x) and the
 * generated-shadow composite above in one pass over the images -- the north star's "update + compositing become
 * one fused elementwise kernel".  Reference: dm1:457-472 followed by dm2:634-653 (blur = 0) or ts:244-265
 * (blur = 1).  Bit-identical to running the two entry points one after the other. */
int advs_ddim_step_composite(const float* x, const float* eps, float* x_out, const float* coef, int32_t* step_dev,
                             int advance, int clip_denoised, const float* img, const float* centers,
                             const float* radii, const float* feature_mask, int Cm, int blur, float* out, int B,
                             int C, int H, int W, void* stream);

/* ---- IDDM class-conditional UNet + CFG DDIM (model/networks/unet.py:17-128, model/modules/{conv,block,attention}.py,
 *      model/samples/ddim.py:48-100): the bandwidth ops not shared with the diff_model path ------------------ */
/* nn.MaxPool2d(2) over NHWC (block.py:25) */
int advs_maxpool2x2(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);
/* nn.Upsample(scale_factor=2, bilinear, align_corners=True) (block.py:56) into channels [y_coff, y_coff+C) of
 * y [B,2H,2W,y_cstride]; advs_copy_channels places the skip tensor in the same concat buffer (block.py:73) */
int advs_upsample_bilinear2x(const void* x, void* y, int B, int H, int W, int C, int y_cstride, int y_coff, int dtype,
                             void* stream);
int advs_copy_channels(const void* x, void* y, size_t npix, int C, int y_cstride, int y_coff, int dtype, void* stream);
/* nn.LayerNorm([C]) per token (attention.py:29,31) */
int advs_layernorm(const void* x, const float* gamma, const float* beta, void* y, size_t rows, int C, float eps,
                   int dtype, void* stream);
/* y = act(x*scale + shift (+ residual)) (+ emb[b,c]); act: 0 none, 1 SiLU, 2 GELU(erf).  DoubleConv's GroupNorm(1,C)
 * + activation + residual form (conv.py:40-66) and the "x + emb" of Down/UpBlock (block.py:44-46, 76-78) */
int advs_groupnorm_apply_ex(const void* x, int B, int HW, int C, const float* scale_shift, const void* residual,
                            const float* emb, int emb_stride, int act, void* y, int dtype, void* stream);
/* 16-bit mode variants: x may carry an int8 mantissa extension x_lo (advs_conv_params.y_lo; NULL = none); y_dtype =
 * ADVS_F16 writes a (bounded) normalised tensor as an fp16 GEMM operand.  advs_layernorm_f16out: bf16 in, fp16 out. */
int advs_groupnorm_apply_ex16(const void* x, const void* x_lo, int B, int HW, int C, const float* scale_shift,
                              const void* residual, const float* emb, int emb_stride, int act, void* y, int y_dtype,
                              void* stream);
int advs_layernorm_f16out(const void* x, const float* gamma, const float* beta, void* y, size_t rows, int C, float eps,
                          void* stream);
int advs_activation(const void* x, void* y, size_t n, int act, int dtype, void* stream);
/* BaseNet.pos_encoding: out[i] = [sin(t_i f_j) | cos(t_i f_j)] (+ label_emb[labels[i]] when labels != NULL);
 * inv_freq[half] is the reference's 1/10000^(arange(0,C,2)/C) table (base.py:63), evaluated by the host */
int advs_pos_encoding(const int64_t* t, int nt, const float* inv_freq, int half, const int64_t* labels,
                      const float* label_emb, float* out, void* stream);
/* same; only rows [0, n_labeled) get the label embedding: the conditional and the unconditional prediction of
 * classifier-free guidance (ddim.py:81-85) as ONE forward over 2n rows */
int advs_pos_encoding_ex(const int64_t* t, int nt, const float* inv_freq, int half, const int64_t* labels,
                         const float* label_emb, int n_labeled, float* out, void* stream);
/* torch.lerp(uncond, cond, w): classifier-free guidance (ddim.py:89) */
int advs_cfg_lerp(const float* uncond, const float* cond, float w, float* out, size_t n, void* stream);
/* ((x + 1) * 0.5 * 255).type(uint8) without clamp (ddim.py:97-99) */
int advs_to_uint8(const float* x, uint8_t* out, size_t n, void* stream);

/* ---- attack-success decision (ASR_fast.py:101-126) --------------------------------------- */
/* flags[b] = argmax_c logits[b,c] != labels[b] (first max wins, like torch.max);
 * counts[0] += number of successes, counts[1] += B. */
int advs_success_flags(const float* logits, const int64_t* labels, int B, int classes,
                       uint8_t* flags, int64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADVSHADOW_B200_H */
