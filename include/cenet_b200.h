/*
 * cenet_b200.h -- C ABI of libcenet_b200.so: hand-written sm_100a kernels for the CENet forward hot path.
 *
 * The reference (xmindflow/cenet) is pure PyTorch-eager: it owns no native code, so there is no reference
 * FFI to mirror symbol-for-symbol.  Each entry point below instead replaces the ATen/cuDNN/cuBLAS call
 * sequence of one reference site (cited as file:line relative to /root/reference/src/networks/cenet/ and
 * /root/reference/src/utils/).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.
 *
 * Conventions:
 *   - plain pointers + sizes only; the caller owns ALL device memory, including workspaces;
 *   - kernels never allocate, never synchronise, and enqueue on the stream passed last;
 *   - return 0 on success, negative on error; cenet_last_error() gives the message (thread-local);
 *   - activations are channels-last: an image batch [B,H,W,C] is also the row-major matrix [B*H*W, C];
 *   - dtype codes: CENET_F32 / CENET_BF16; "fp32 vectors" (bias, gains, folded BN) are always float.
 */
#ifndef CENET_B200_H
#define CENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cenet_stream_t; /* cudaStream_t */

enum { CENET_F32 = 0, CENET_BF16 = 1 };
enum { CENET_ACT_NONE = 0, CENET_ACT_GELU = 1, CENET_ACT_RELU = 2, CENET_ACT_LEAKY = 3, CENET_ACT_SILU = 4,
       CENET_ACT_SIGMOID = 5, CENET_ACT_GELU_GRAD = 6 /* d gelu(v)/dv: `mul_act` of training dgrad GEMMs */ };
enum { CENET_GEMM_AUTO = -1, CENET_GEMM_SIMT = 0, CENET_GEMM_TCGEN05 = 1, CENET_GEMM_MMA = 2 };

/* ---- library ---------------------------------------------------------------------------------------- */
const char* cenet_last_error(void);
int cenet_abi_version(void);
/* number of kernels launched by this library in this process (bench.py's `gpu_launches`) */
long long cenet_launch_count(void);

/* ---- GEMM / implicit-GEMM convolution with fused epilogue ---------------------------------------------
 * C[M,N] = epilogue( A[M,K] * W[N,K]^T ).   Replaces every nn.Linear / 1x1 Conv2d / dense Conv2d of the
 * path (pvtv2.py:41-45,90-106,164-165,68; cfam.py:150-157,301-303; dseb.py:164; unet.py:201-214;
 * blocks.py:209-214; nlb.py:107-143) together with the bias / BatchNorm(eval, folded) / activation /
 * residual / gating ops that follow them.
 *   epilogue, in order:  v = alpha*acc;  v *= row_scale[m];  v += bias[n] (or bias[m]);
 *                        if(!act_after_res) v = act(v);      v *= mul_act(mul[m,n]);
 *                        v += res1[m,n]*(res1_cscale ? res1_cscale[n] : res1_scale);  v += res2[m,n];
 *                        if(act_after_res) v = act(v).
 * conv != 0: A is an NHWC image [Bimg,H,W,Cin]; M = Bimg*Ho*Wo; K = KH*KW*Cin ordered (kh,kw,cin);
 *            W is [N, KH*KW*Cin] (the reference's [Cout,Cin,KH,KW] weight permuted once at pack time).
 * batch > 1: z = zo*batch_inner + zi; pointer offsets are zo*bs_outer + zi*bs_inner elements.
 * impl: CENET_GEMM_TCGEN05 needs bf16 A and W, K-major W, K % 8 == 0, 16-byte aligned rows, batch == 1.
 *       CENET_GEMM_MMA (mma.sync, bf16 A and W): batched problems and the transposed operand layouts (a_mmajor / w_nmajor) --
 *       the materialised attention of the training path.  CENET_GEMM_AUTO: tcgen05 if eligible, else mma.sync if bf16, else CUDA cores.
 */
typedef struct {
  int M, N, K;
  int batch, batch_inner;
  const void* A; int a_dtype; long long lda, a_bs_outer, a_bs_inner;
  int conv, Bimg, H, W, Cin, KH, KW, stride, pad, Ho, Wo;
  const void* Wt; int w_dtype; long long ldw, w_bs_outer, w_bs_inner; int w_nmajor;
  void* C; int c_dtype; long long ldc, c_bs_outer, c_bs_inner;
  float alpha; const float* bias; int bias_per_row; const float* row_scale;
  int act; float slope; int act_after_res;
  const void* res1; int res1_dtype; long long ldr1; const float* res1_cscale; float res1_scale;
  const void* res2; int res2_dtype; long long ldr2;
  const void* mul; int mul_dtype; long long ldmul; int mul_act;
  int impl;
  /* training extensions; all-zero == inference behaviour */
  int a_mmajor;                       /* A is stored [K, M] (element (m,k) at A[k*lda + m]); CUDA-core path only */
  int rs_div;                         /* row_scale index is m / rs_div (0 -> 1) */
  const float* post_row_scale; int post_rs_div; /* v *= post_row_scale[m / post_rs_div] after bias/act/mul, before the residuals */
  const float* k_scale; int k_scale_div; long long k_scale_bs; /* A[m,k] *= k_scale[(z*k_scale_bs + k) / k_scale_div]; CUDA-core path */
  /* optional split-K workspace (fp32, caller-owned): problems with few output tiles and a long contraction (the
   * spatial-reduction convs: M = B*49 patches, K = 4096) are split over blockIdx.z into k-slices whose fp32 partial tiles
   * land here and are reduced (+ bias, cast) in fixed order by a second kernel.  NULL: never split. */
  float* split_ws; long long split_ws_elems;
} cenet_gemm_args;
int cenet_gemm(const cenet_gemm_args* a, cenet_stream_t s);

/* ---- normalisation / row reductions --------------------------------------------------------------------
 * LayerNorm over the last dim (pvtv2.py:146-147,189,320; eps 1e-6 / 1e-5). */
int cenet_layernorm(const void* x, int x_dtype, void* y, int y_dtype, const float* gamma, const float* beta,
                    long long rows, int C, float eps, cenet_stream_t s);
/* in-place softmax over rows of length n (materialised attention path; nlb.py:128, multihead_diffattn.py:108) */
int cenet_softmax_rows(void* x, int dtype, long long rows, int n, long long ld, cenet_stream_t s);
/* per-row [max, mean, std] over C channels -> stats[rows,3] fp32 (SRM, cfam.py:94-97; unbiased std) */
int cenet_row_stats(const void* x, int dtype, long long rows, int C, long long ld, int unbiased, float* stats,
                    cenet_stream_t s);

/* Mix-FFN tail in one tcgen05 kernel (mixffn_tc.cu): t[B*H*W, C] (fp32, in place) += fc2(GELU(dwconv3x3(h) + dw_bias)) + b2, with
 * h [B,H,W,Ch] bf16 the fc1 output, w9c [9][Ch] / dw_bias [Ch] fp32 the depthwise filter, w2 [C][Ch] bf16, b2 [C] fp32 (nullable).
 * The GELU'd depthwise result goes from registers into the shared-memory A operand of the MMA and never reaches global memory.
 * pvtv2.py:40-47,364-370.  C in {64,128}, Ch % 64 == 0, W % 4 == 0, W <= 128: cenet_mixffn_tail_supported() answers 1. */
int cenet_mixffn_tail_supported(int H, int W, int Ch, int C);
int cenet_mixffn_tail(const void* h, void* t, const float* w9c, const float* dw_bias, const void* w2, const float* b2, int B, int H,
                      int W, int Ch, int C, cenet_stream_t s);
/* ---- depthwise 3x3 family -----------------------------------------------------------------------------
 * y[b,h,w,c] = act( (sum_taps w[tap,c]*x[b,h+dh*dil,w+dw*dil,c] + bias[c]) * scale[c] + shift[c] )
 * Mix-FFN DWConv+GELU (pvtv2.py:364-370,42-43), CFAM Mlp dwconv+GELU (cfam.py:151-152), SepConvBN depthwise
 * + BN + ReLU with dilation (blocks.py:169-178), EUCB nearest-x2 + dw3x3 + BN + LeakyReLU (blocks.py:303-311;
 * up2 = 1: x is [B,H/2,W/2,C]).  w is [9,C] fp32 (tap-major).  ldx/ldy: channel pitch (>= C) so that channel
 * slices of a wider tensor can be processed in place. */
int cenet_dwconv3x3(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, const float* w9c,
                    const float* bias, const float* scale, const float* shift, int B, int H, int W, int C, int dil,
                    int up2, int act, float slope, cenet_stream_t s);

/* ---- layout ------------------------------------------------------------------------------------------- */
/* y_nchw[b, coff+c, h, w] = x_nhwc[b,h,w,c]   (torch.cat([dec,skip],1) of dseb.py:156 when called twice) */
int cenet_nhwc_to_nchw(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, int B, int HW, int C,
                       int Ctot, int coff, cenet_stream_t s);
/* y_nhwc[b,hw,c] (pitch ldy) = x_nchw[b,c,hw] */
int cenet_nchw_to_nhwc(const void* x, int x_dtype, void* y, int y_dtype, long long ldy, int B, int HW, int C,
                       cenet_stream_t s);
/* out[m, (kh,kw,ci)] = x_nhwc patch, zero padded, row pitch Kpad >= KH*KW*Cin (strided / non-overlapping convs:
 * patch embeds pvtv2.py:164-165 and the SR conv pvtv2.py:68 feed cenet_gemm through this) */
int cenet_im2col(const void* x, int x_dtype, void* out, int o_dtype, int B, int H, int W, int Cin, int KH, int KW,
                 int stride, int pad, int Ho, int Wo, int Kpad, cenet_stream_t s);
/* bilinear x2, align_corners=True, NHWC (UpConv, blocks.py:210) */
int cenet_upsample2x_ac(const void* x, int x_dtype, void* y, int y_dtype, int B, int H, int W, int C,
                        cenet_stream_t s);
/* y[b,h/2,w/2,coff+c] = wch[c] * max2x2(x[b,h,w,c])   (out.py:43,70: MaxPool2d then `self.w*`) */
int cenet_maxpool2_scale(const void* x, int x_dtype, void* y, int y_dtype, long long ldy, int coff, const float* wch,
                         int B, int H, int W, int C, cenet_stream_t s);
/* y = (x*scale[c]+shift[c]) * gate[b,c]   (BatchNorm(eval) then CCU gate, cfam.py:370,263-264) */
int cenet_affine_gate(const void* x, int x_dtype, void* y, int y_dtype, const float* scale, const float* shift,
                      const float* gate_bc, int B, int HW, int C, cenet_stream_t s);

/* ---- DSEB (dseb.py) -------------------------------------------------------------------------------------
 * FEA + gate combine on the NCHW buffer y[B,C2,H,W]:  z = 2*y + w[c]*edge(y) + gate*y  (dseb.py:63-76,157,118,162)
 * scales: up to 3 scale factors.  gate has the same flat layout as y (the `.view` reinterpretation). */
int cenet_fea_combine(const void* y, const void* gate, void* z, int dtype, const float* w_c, int B, int C2, int H,
                      int W, const float* scales, int nscales, cenet_stream_t s);
/* CENetOrg skip enhancer (cenet_org/decoders.py:112-143), same NCHW planes as cenet_fea_combine:
 * mode 1 (DoGEdge):  z = y + w[c] * | up(down_s0(y)) - up(down_s1(y)) |     (exactly two scale factors; gate unused, may be NULL)
 * mode 2 (combine):  z = y + gate * y                                       (scales unused) */
int cenet_dog_combine(const void* y, const void* gate, void* z, int dtype, const float* w_c, int B, int C2, int H, int W,
                      const float* scales, int nscales, int mode, cenet_stream_t s);
/* P[:, 2i] -= lambda * P[:, 2i+1] over contiguous maps of `map_elems` elements (multihead_diffattn.py:115-116) */
int cenet_diff_combine(void* P, int dtype, long long npairs, long long map_elems, float lambda, cenet_stream_t s);
/* y = x * rsqrt(mean_seg(x^2)+eps) * mult over segments of `seg` channels (rms_norm.py:15-22 and the
 * `*(1-lambda_init)` of multihead_diffattn.py:123) */
int cenet_rmsnorm_seg(const void* x, int x_dtype, void* y, int y_dtype, long long rows, int C, int seg, float eps,
                      float mult, cenet_stream_t s);
/* Fused differential flash attention (multihead_diffattn.py:92-124 without the N x N maps).
 * qkv: [B, N, 3E] bf16 rows (q | k | v), heads h, hd = E/(2h).  out: [B, N, E] bf16 =
 * RMSNorm_{2hd}( softmax(q_{2i}k_{2i}^T/sqrt(hd)) v_i - lambda*softmax(q_{2i+1}k_{2i+1}^T/sqrt(hd)) v_i ) * mult */
int cenet_diffattn_flash(const void* qkv, void* out, int B, int N, int E, int heads, float lambda, float eps,
                         float mult, float* kmax_ws, cenet_stream_t s);
/* kmax_ws: NULL, or B*2*heads floats of workspace.  When given, a pre-pass stores max_n|k_n| per softmax map and the
 * kernel uses the Cauchy-Schwarz bound |q_r|*max|k| as a fixed softmax shift (no running max / rescale) for every warp
 * whose bound stays inside the safe exponent range, and the online-max path otherwise.  Same result up to rounding. */

/* Same kernel for head dims that are not MMA friendly (Synapse 14x14 level: hd = 20): the host zero-pads the q/k heads to
 * hd_pad and the value heads to dv_pad when packing q_proj/k_proj/v_proj/out_proj.  Row layouts:
 * qkv [B,N, 2h*hd_pad | 2h*hd_pad | h*dv_pad], out [B,N, h*dv_pad].  Supported (hd_pad, dv_pad): (8,16) (16,32) (32,48)
 * (32,64) (64,128). */
int cenet_diffattn_flash_padded(const void* qkv, void* out, int B, int N, int heads, int hd_pad, int dv_pad, int hd_real,
                                float lambda, float eps, float mult, float* kmax_ws, cenet_stream_t s);

/* ---- encoder attention (pvtv2.py:88-105): softmax(q k^T * scale) v, any number of reduced keys, head_dim 64 --- */
int cenet_sr_attention(const void* q, int q_dtype, const void* kv, int kv_dtype, void* out, int o_dtype, int B, int N,
                       int Nk, int C, int heads, float scale, cenet_stream_t s);

/* ---- non-local block core (nlb.py:116-137): out[b,n,:] = softmax_p(theta_n . phi_p * scale) g_p ------------
 * tpg: [B, N, 3C] bf16 rows (theta | phi | g); out [B,N,C] bf16.  C in {64,128}. */
int cenet_nonlocal_flash(const void* tpg, void* out, int B, int N, int C, float scale, cenet_stream_t s);

/* ---- CFAM statistics (cfam.py) ------------------------------------------------------------------------------
 * CCU (cfam.py:251-264): stats of (x*scale+shift) over HW per (b,c) -> per-channel 3->3->1 MLP -> optional
 * BN1d(eval affine: bn_scale/bn_shift, pass NULL when B == 1) -> sigmoid -> gate[B,C].
 * ws: fp32 workspace of B*nchunk*C*3 floats, nchunk = cenet_ccu_nchunk(HW). */
int cenet_ccu_nchunk(int HW);
int cenet_ccu_gate(const void* x, int x_dtype, const float* scale, const float* shift, const float* fc1_c33,
                   const float* fc2_c3, const float* bn_scale, const float* bn_shift, float* gate_bc, float* ws,
                   int B, int HW, int C, cenet_stream_t s);
/* SRM gate (cfam.py:97-100): u[B,H,W,3] -> sigmoid(BN(GELU(pwc(u)+dwc3x3(u)))) -> gate[B*H*W]
 * pw: 3 floats, dw: [3,3,3] (cin,kh,kw) floats, bn: scale, shift */
int cenet_srm_gate(const float* u, float* gate, const float* pw3, const float* dw27, float bn_scale, float bn_shift,
                   int B, int H, int W, cenet_stream_t s);
/* image-pooling branch of MultiOrderDWConv (cfam.py:209-218,231-232):
 *  step 1: pooled[b,7,7,r] = LeakyReLU(BN(conv1x1(AdaptiveAvgPool7(x[..., coff:coff+r]))))
 *  step 2: y[b,h,w,coff_y+c] = bilinear(H,W; align False) of bilinear(49x49; align True) of pooled */
int cenet_pool_branch(const void* x, int x_dtype, long long ldx, int coff, void* y, int y_dtype, long long ldy,
                      int coff_y, const float* w_rr, const float* bn_scale, const float* bn_shift, float slope,
                      float* pooled_ws, int B, int H, int W, int r, cenet_stream_t s);

/* ---- OutHead stem (out.py:41-44, unet.py:201-209): first res-block conv on the raw image -----------------------------
 * o1 = LeakyReLU(conv5x5(x; w1)+b1)  (BN1 folded into w1 [32, 25*Cin] fp32 with K ordered (kh,kw,ci), and b1),
 * r  = conv1x1(x; w3)+b3             (BN3 folded; w3 [32,Cin]); r may be NULL.  x [B,H,W,Cin], Cin <= 4; outputs
 * [B,H,W,32] of dtype o_dtype. */
int cenet_stem5x5(const void* x, int x_dtype, const float* w1, const float* b1, const float* w3, const float* b3,
                  void* o1, void* r, int o_dtype, int B, int H, int W, int Cin, float slope, cenet_stream_t s);

/* ---- head (out.py:74 + metrics_eval.py:52) -------------------------------------------------------------------
 * y: [B,h,w,ncls] fp32 (NHWC, pixel pitch ldy >= ncls; 0 = ncls: the inference plan pads the class dimension to a multiple of 8 so
 * that the producing 1x1 conv stays on the vector epilogue) -> logits [B,ncls,2h,2w] fp32 (bilinear x2, align_corners=False) and / or
 * labels [B,2h,2w] int64 = argmax over classes, lowest index on ties.  Either output may be NULL. */
int cenet_head_upsample_argmax(const float* y, float* logits_nchw, long long* labels, int B, int h, int w, int ncls, int ldy,
                               cenet_stream_t s);

/* ---- tcgen05 flash attention, head width 64 or 128, or ONE head of width 192..1024 (multiple of 64; lse must be NULL)
 *      (nlb.py:116-137; pvtv2.py:98-103) ------------------------------------------------------------------------------------
 * o[b, n, h*D + :] = softmax_k( q[b,n,h*D+:] . k[b,k,h*D+:] * scale ) v[b,k,h*D+:]   for n < Nq, k < Nk, bf16 in / out, fp32
 * accumulation in TMEM; nothing Nq x Nk is written to HBM.  ld*: row pitches, b*: per-image strides (elements, multiples
 * of 8); pointers 16-byte aligned.  lse (nullable): fp32 [B, heads, Nq] natural log-sum-exp of the scaled scores. */
typedef struct {
  const void *q, *k, *v; void* o; float* lse;
  long long ldq, ldk, ldv, ldo;
  long long bq, bk, bv, bo;
  int B, heads, Nq, Nk, D;
  float scale;
  int lse_base2;                  /* 1: lse in log2 units (what cenet_flash_bwd reads); 0: natural log */
} cenet_attn_tc_args;
int cenet_attn_tc(const cenet_attn_tc_args* a, cenet_stream_t s);

/* ---- per-volume evaluation tail (utils/metrics_eval.py:53-71, utils_synapse.py:69-84; medpy.metric.binary.dc) -------------
 * pred_patch [D,ph,pw] int64 = the network's label maps at the patch size.  For every voxel (d,y,x) of the original volume:
 * pred_out[d,y,x] = pred_patch[d, iy[y], ix[x]]  (iy/ix: the nearest-neighbour source index tables of scipy's
 * `zoom(order=0)`, built on the host; -1 = outside -> label 0) and, when `label` is given ([D,H,W]; label_kind 0 float32 / 1 int64 / 2 uint8),
 * counts[0..ncls) += (pred == c && label == c), counts[ncls..2ncls) += (pred == c), counts[2ncls..3ncls) += (label == c).
 * counts is int64 [3*ncls], zeroed by the call; exact integers, deterministic.  pred_out / label may be NULL. */
int cenet_volume_labels_counts(const long long* pred_patch, int ph, int pw, const int* iy, const int* ix, const void* label,
                               int label_kind, long long* pred_out, long long* counts, int D, int H, int W, int ncls,
                               cenet_stream_t s);

/* ---- fused Dice + CE (utils/core.py:57-80,176-188) ------------------------------------------------------------
 * logits [B,ncls,H,W] fp32, labels [B,H,W] int64.  ws: (3*ncls+1)*nblk + 3*ncls+3 floats, nblk =
 * cenet_loss_nblocks(B*H*W).  loss_out[0] = w_dice*dice + w_ce*ce; loss_out[1+i] = per-class dice score.
 * dlogits (nullable): d loss / d logits, same shape as logits.  Deterministic (no float atomics). */
int cenet_loss_nblocks(long long npix);
int cenet_dice_ce(const float* logits, const long long* labels, float* loss_out, float* dlogits, float* ws, int B,
                  int ncls, int HW, float w_dice, float w_ce, float grad_scale, cenet_stream_t s);
/* Same three passes with the third term of Criterion (utils/core.py:161-188): BoundaryDoULoss (core.py:83-131), the loss the
 * ACDC / Synapse scripts select (`--loss_type boundary`).  loss_out[0] = w_dice*dice + w_ce*ce + w_boundary*boundary_dou.
 * The per-class boundary pixel counts C_c and class sizes S_c are exact integers (tot[4c+3], tot[4c+2] after the call, tot =
 * ws + nblk*(4*ncls+1)).  ws: (4*ncls+1)*nblk + 5*ncls+4 floats. */
int cenet_seg_loss(const float* logits, const long long* labels, float* loss_out, float* dlogits, float* ws, int B,
                   int ncls, int H, int W, float w_dice, float w_ce, float w_boundary, float grad_scale, cenet_stream_t s);

/* ================================================================================================================
 * TRAINING entry points: train-mode forward pieces and the hand-written backward of every op on the path
 * (the reference gets these from autograd; main_acdc.py:234-265 `loss.backward(); optimizer.step()`).
 * Conventions: `ld*` row pitches in elements; pointers are pre-offset to the first element of a channel slice;
 * `acc` / `acc_*` != 0 means "add to the destination" (gradient fan-in); `ws`/`ws_elems` is caller-owned fp32 scratch for
 * the two-stage deterministic reductions (no float atomics anywhere).  Parameter gradients are fp32 in the reference's
 * parameter layout.
 * ================================================================================================================ */

/* depthwise 3x3 forward that also stores the pre-activation z (contiguous [B,H,W,C], dtype of y) for the backward */
int cenet_dwconv3x3_train(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, void* zout,
                          const float* w9c, const float* bias, int B, int H, int W, int C, int dil, int act, int z_dtype,
                          float slope, cenet_stream_t s);
/* weight gradient of nn.Linear / 1x1 / dense conv:  dw[n, ci, t] = sum_m rs[m/rs_div] dy[m,n] x[m, t*Cin + ci], K = T*Cin
 * (T = KH*KW taps of an im2col'ed conv, 1 otherwise); dbias[n] = sum_m (rs) dy[m,n] (nullable; bias_unscaled: without rs).
 * bf16 operands: mma.sync tensor-core kernel with ldmatrix.trans staging, split over m; fp32 x fp32: CUDA cores. */
int cenet_gemm_wgrad(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, long long M, int N,
                     int K, int T, const float* row_scale, int rs_div, float* dw, float* dbias, int bias_unscaled, float* ws,
                     long long ws_elems, cenet_stream_t s);
/* The same GEMM with the reduction of its split partials DEFERRED, so that one launch (cenet_wgrad_reduce_batch) can reduce
 * the partials of a whole gradient bucket -- autograd runs one AccumulateGrad per parameter; here ~150 finalize launches
 * per step become ~6.  bf16 operands run on the tcgen05 kernel: TMA-fed MN-major operands, accumulator in TMEM, the bias
 * gradient as one more N=16 MMA per k-step against a constant ones tile.  rs_binary != 0 declares row_scale to be a DropPath mask
 * (every entry 0 or one common value c, timm drop_path: bernoulli(keep)/keep per sample): dropped samples are skipped and the
 * split count no longer grows with the batch.
 * On return *n_partials == 0: dw / dbias hold the final result.  Otherwise ws holds S = *n_partials partial matrices
 * [N][K] (S*N*K floats) followed, when *bias_partials != 0, by S partial bias rows [N]; dbias is final when
 * *bias_partials == 0. */
int cenet_gemm_wgrad_partial(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, long long M,
                             int N, int K, int T, const float* row_scale, int rs_div, int rs_binary, float* dw, float* dbias,
                             int bias_unscaled, float* ws, long long ws_elems, int* n_partials, int* bias_partials,
                             cenet_stream_t s);
/* one reduction job: dst[n, ci, t] = sum_{z<S} src[z*stride + n*src_ld + t*Cin + ci]  (K = T*Cin; a bias row is N=1, K=len, T=1;
 * src_ld = 0 means K; src_ld > K addresses an N x K block inside wider partial rows: the diagonal blocks of a block-diagonal
 * GEMM, for which cenet_gemm_wgrad_partial is called with dw = NULL so that it always leaves partials) */
typedef struct {
  const float* src;
  float* dst;
  long long stride;
  int S, N, K, T;
  int blk0;          /* first block of this job = running sum of cenet_wgrad_reduce_blocks over the preceding jobs */
  int src_ld;
} cenet_wgrad_job;
/* host-only: the tile / split plan the tcgen05 weight-gradient kernel would use; out[9] = {bn, S, per_group, parts, chunks per
 * split, chunks per group, groups, rows per group, total chunks} */
int cenet_wgrad_plan_query(long long M, int N, int K, int has_rs, int rs_div, int binary, long long max_partials, int* out);
/* 256-thread blocks job j needs (host-side helper, no launch) */
int cenet_wgrad_reduce_blocks(const cenet_wgrad_job* j);
/* jobs: DEVICE array; fixed summation order (bit-reproducible, no float atomics) */
int cenet_wgrad_reduce_batch(const cenet_wgrad_job* jobs, int njobs, int nblocks, cenet_stream_t s);
/* out[c] = sum_r rs[r / rs_div] * x[r, c] (rs nullable): bias gradients, and the weight gradient of a 1x1 conv over a ONE-channel
 * image (rs = the image; unet.py:205-207 with input_channels = 1).  Two-stage fixed-order reduction through ws. */
int cenet_colsum(const void* x, int dtype, long long ld, long long rows, int C, const float* row_scale, int rs_div, float* out,
                 float* ws, long long ws_elems, cenet_stream_t s);
/* DropPath masks of one step (timm drop_path via pvtv2.py:123,146-147): out[r, b] = bernoulli(keep[r]) / keep[r], r = branch row
 * (2 per encoder block), b = sample; counter-based generator, *counter (device) advances by one per launch */
int cenet_droppath_mask(float* out, const float* keep, int n, int B, unsigned long long seed, unsigned long long* counter,
                        cenet_stream_t s);
/* out[m, c] = x[m, c] * rs[m]  (contiguous [rows, C]; the per-pixel SRM gate folded into d(fc2 output), cfam.py:157) */
int cenet_row_scale(const void* x, int dtype, const float* rs, void* out, long long rows, int C, cenet_stream_t s);
/* dx[m, 0:N) (+)= sum_{k<K} dy[m, k] * w[k*ldw + n]: input gradient of a layer with a tiny output width (K <= 16 classes of the
 * segmentation head, unet.py:357-381): dy fp32 [rows, K] contiguous, w fp32, dx [rows, N] (pitch ldx), N % 8 == 0 */
int cenet_smallk_dgrad(const float* dy, int K, const float* w, long long ldw, void* dx, int dx_dtype, long long ldx, long long rows,
                       int N, int acc, cenet_stream_t s);
/* weight gradient of a dense stride-1 "same" conv WITHOUT im2col: x is the NHWC image [B,H,W,Cin] (pitch ldx), dy [B*H*W, N] (pitch
 * ldy); dw in the reference layout [N, Cin, k, k].  The X operand tile is gathered tap by tap inside the kernel. */
int cenet_conv_wgrad(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, int B, int H, int W,
                     int Cin, int ksize, int N, float* dw, float* ws, long long ws_elems, cenet_stream_t s);
/* LayerNorm backward (pvtv2.py:146-147,189,320): dx (+)= ..., dgamma, dbeta; statistics are recomputed from x.  C in {64,128,320,512} */
/* n_partials (nullable): when given, d(gamma) / d(beta) are NOT finalised: ws holds *n_partials partial rows [2][C] (d(gamma) then
 * d(beta)) for cenet_wgrad_reduce_batch -- the 53 LayerNorm parameter gradients of a step ride in the per-bucket batched reduction */
int cenet_layernorm_bwd(const void* dy, const void* x, int dtype, const float* gamma, float eps, long long rows, int C, void* dx,
                        int acc, float* dgamma, float* dbeta, float* ws, long long ws_elems, int* n_partials, cenet_stream_t s);
/* train-mode BatchNorm statistics of x[rows, C]: mean, rstd (biased var), scale = gamma*rstd, shift = beta - mean*scale; updates
 * running_mean / running_var (unbiased) with `momentum` and increments num_batches_tracked (all nullable) */
/* `frozen` != 0 (here and in cenet_bn_bwd / cenet_ccu_mlp_* / cenet_srm_*): eval-mode normalisation inside a gradient-enabled pass
 * -- the running statistics ARE the statistics, nothing is updated, and the backward has no batch terms (main_acdc.py:226 runs
 * `net.eval()` under grad mode). */
int cenet_bn_stats(const void* x, int x_dtype, long long ldx, long long rows, int C, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps, int frozen,
                   float* scale, float* shift, float* mean, float* rstd, float* ws, long long ws_elems, cenet_stream_t s);
/* out = act( a*sa+ta [+ (sb ? b*sb+tb : b)] )  -- BatchNorm apply (+ residual branch of UnetResBlock, unet.py:201-214) */
int cenet_affine_act(const void* a, int a_dtype, long long lda, const float* sa, const float* ta, const void* b, int b_dtype,
                     long long ldb, const float* sb, const float* tb, void* out, int o_dtype, long long ldo, long long rows, int C,
                     int act, float slope, cenet_stream_t s);
/* backward of y = act(BN_train(a) [+ other]): g = dy*act'(y) (act in NONE/RELU/LEAKY; y nullable for NONE);
 * da (+)= gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)); dgamma, dbeta; optional dres (+)= g for the residual operand.
 * dy and y share (ldy); a and da share (lda). */
int cenet_bn_bwd(const void* dy, int dy_dtype, const void* y, int y_dtype, long long ldy, const void* a, int a_dtype, long long lda,
                 const float* mean, const float* rstd, const float* gamma, long long rows, int C, int act, float slope, void* da,
                 int da_dtype, int acc_da, float* dgamma, float* dbeta, void* dres, int dres_dtype, long long lddres, int acc_dres, int frozen,
                 float* ws, long long ws_elems, cenet_stream_t s);
/* filter / bias gradient of the depthwise 3x3 family: dw [C,1,3,3], dbias [C] (nullable); x as in cenet_dwconv3x3 (up2) */
int cenet_dwconv3x3_wgrad(const void* x, int x_dtype, long long ldx, const void* dz, int dz_dtype, long long ldz, int B, int H, int W,
                          int C, int dil, int up2, float* dw, float* dbias, float* ws, long long ws_elems, cenet_stream_t s);
/* y[b,h,w,c] (+)= sum of the 2x2 block of x[b,2h..,2w..,c]  (adjoint of nearest x2, blocks.py:304) */
int cenet_sumpool2(const void* x, int x_dtype, void* y, int y_dtype, int B, int Ho, int Wo, int C, int acc, cenet_stream_t s);
/* adjoint of cenet_im2col: dx[b,h,w,ci] (+)= sum over the taps that read it */
int cenet_col2im(const void* dcol, int c_dtype, void* dx, int x_dtype, int B, int H, int W, int Cin, int k, int stride, int pad, int Ho,
                 int Wo, int Kpad, int acc, cenet_stream_t s);
/* flash attention with saved log2-sum-exp (bf16): O[:, m*dv..] = softmax(scale Q_m K_m^T) V_{m/vdiv}; lse [B,maps,Nq].
 * (dqk, dv) in {(8,16),(16,32),(32,64),(64,64),(128,128),(80,160)}.  Serves pvtv2.py:88-105, nlb.py:116-137, multihead_diffattn.py:92-116. */
int cenet_flash_fwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O, long long ldo,
                    float* lse, int B, int maps, int Nq, int Nk, int dqk, int dv, int vdiv, float scale, cenet_stream_t s);
/* training forward of the differential attention on the tcgen05 / TMEM / TMA kernel (diffattn_tc.cu): qkv rows
 * [q: 2h x hd | k: 2h x hd | v: h x 2hd] (bf16, [B*N, 3E]); om [B*N, 2E] = the 2h per-map outputs softmax(q_m k_m^T / sqrt(hd)) v_{m/2},
 * lse [B, 2h, N] in log2 units -- the operands cenet_flash_bwd and cenet_diff_rmsnorm_fwd expect.  head_dim in {8,16,32,64};
 * kmax_ws (nullable) [B*2h] floats enables the fixed softmax shift.  multihead_diffattn.py:92-113. */
int cenet_diffattn_fwd_train(const void* qkv, void* om, float* lse, int B, int N, int E, int heads, float* kmax_ws, cenet_stream_t s);
/* its backward: delta = rowsum(dO*O) (workspace [B,maps,Nq]); dQ, dK, dV written with the layouts of Q, K, V (no atomics).
 * ws (nullable): fp32 scratch; with <= 64 keys (the encoder's SR attention: 49) the dK / dV kernel splits the QUERIES over CTAs and
 * a fixed-order reduction adds the partials -- 2*148 CTAs instead of B*heads. */
int cenet_flash_bwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, const void* O,
                    const void* dO, long long ldo, const float* lse, float* delta, void* dQ, void* dK, void* dV, int B, int maps,
                    int Nq, int Nk, int dqk, int dv, int vdiv, float scale, float* ws, long long ws_elems, cenet_stream_t s);
/* materialised attention backward: dP <- P * (dP - rowsum(P*dP)) in place, rows of length n */
int cenet_softmax_bwd_rows(const void* P, void* dP, int dtype, long long rows, int n, cenet_stream_t s);
/* lambda = exp(q1.k1) - exp(q2.k2) + lambda_init on the device (multihead_diffattn.py:110-112) and its backward */
int cenet_lambda_fwd(const float* q1, const float* k1, const float* q2, const float* k2, int hd, float lambda_init, float* lam,
                     cenet_stream_t s);
int cenet_lambda_bwd(const float* dlam, const float* q1, const float* k1, const float* q2, const float* k2, int hd, float* g_q1,
                     float* g_k1, float* g_q2, float* g_k2, cenet_stream_t s);
/* o[r, h*seg..] = mult * RMSNorm_seg( Om[r, 2h*seg..] - lam * Om[r, (2h+1)*seg..] )  (multihead_diffattn.py:115-123) + backward */
int cenet_diff_rmsnorm_fwd(const void* Om, int dtype, const float* lam, void* o, long long rows, int heads, int seg, float eps,
                           float mult, cenet_stream_t s);
int cenet_diff_rmsnorm_bwd(const void* dO, const void* Om, int dtype, const float* lam, void* dOm, float* dlam, long long rows,
                           int heads, int seg, float eps, float mult, float* ws, long long ws_elems, cenet_stream_t s);
/* backward of cenet_fea_combine: dy (+)=, dgate =, dw[c]; mats [nscales][2][nmax][nmax] = per-axis operators Up_s Down_s,
 * bands [nscales][2 axes][2: rows, columns][nmax][2] = [lo, hi) of the non-zeros of every row / column of those operators;
 * bit s of ident_mask: scale factor s is 1.0 (identity operator: its residual and gradient term are exactly zero, skipped) */
int cenet_fea_bwd(const void* y, const void* gate, const void* dz, int dtype, const float* w, void* dy, int acc_dy, void* dgate,
                  float* dw, int B, int C2, int H, int W, const float* mats, const int* bands, int nmax, int nscales,
                  int ident_mask, float* ws, long long ws_elems, cenet_stream_t s);
/* out_nhwc[b,hw,c] (+)= x_nchw[b, coff+c, hw] */
int cenet_nchw_to_nhwc_slice(const void* x, int dtype, void* out, int B, int HW, int C, int Ctot, int coff, int acc, cenet_stream_t s);
/* dst (+)= src */
int cenet_add(void* dst, const void* src, int dtype, long long n, int acc, cenet_stream_t s);
/* CCU in train mode (cfam.py:251-264): stats u[B,C,3] = (max, mean, biased std) + arg max; the per-channel MLP with batch-statistics
 * BatchNorm1d over B (gamma == NULL: skipped, the reference's B == 1 case); and the three backward pieces */
int cenet_ccu_stats(const void* xb, int dtype, float* u, int* arg, int B, int HW, int C, float* ws, long long ws_elems, cenet_stream_t s);
int cenet_ccu_mlp_fwd(const float* u, const float* fc1, const float* fc2, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* nbt, float momentum, float eps, float* gate, float* save_bc8, int B, int C, int frozen,
                      cenet_stream_t s);
int cenet_ccu_dgate(const void* dx1, const void* xb, int dtype, float* dgate, int B, int HW, int C, float* ws, long long ws_elems,
                    cenet_stream_t s);
int cenet_ccu_mlp_bwd(const float* dgate, const float* u, const float* fc1, const float* fc2, const float* gamma, const float* beta,
                      const float* save_bc8, float* du, float* dfc1, float* dfc2, float* dgamma, float* dbeta, int B, int C, int frozen,
                      cenet_stream_t s);
int cenet_ccu_apply_bwd(const void* dx1, const void* xb, int dtype, const float* gate, const float* u, const int* arg, const float* du,
                        void* dxb, int acc, int B, int HW, int C, cenet_stream_t s);
/* SRM in train mode (cfam.py:93-101): per-pixel (max, mean, unbiased std) + arg max; gate with batch-statistics BN(1); backward */
int cenet_row_stats_arg(const void* x, int dtype, float* u, int* arg, long long rows, int C, cenet_stream_t s);
int cenet_srm_fwd(const float* u, const float* pw, const float* dw, const float* gamma, const float* beta, float* running_mean,
                  float* running_var, long long* nbt, float momentum, float eps, float* gm, float* save_m2, float* st, int B, int H,
                  int W, int frozen, float* ws, long long ws_elems, cenet_stream_t s);
int cenet_row_dot(const void* a, const void* b, int dtype, float* out, long long rows, int C, cenet_stream_t s);
int cenet_srm_bwd(const float* dgm, const float* u, const float* gm, float* save_m2, const float* st, const float* pw, const float* dw,
                  const float* gamma, const float* beta, float* du, float* dpw, float* ddw, float* dgamma, float* dbeta, int B, int H,
                  int W, int frozen, float* ws, long long ws_elems, cenet_stream_t s);
int cenet_srm_apply_bwd(const void* dh3, const void* h2, const void* z, int dtype, const float* gm, const float* u, const int* arg,
                        const float* du, void* dz, long long rows, int C, cenet_stream_t s);
/* SiLU(g)*SiLU(v) (cfam.py:303-304) and backward */
int cenet_silu_mul_fwd(const void* g, const void* v, void* out, int dtype, long long n, cenet_stream_t s);
int cenet_silu_mul_bwd(const void* dout, const void* g, const void* v, void* dg, void* dv, int dtype, long long n, cenet_stream_t s);
/* out = x + ls[c] * (y ? (1-w) y + w pz : pz), pz = s ? p*s[c]+t[c] : p   (layer scale + non-local mix, cfam.py:369,373; nlb.py:145-148)
 * backward: dy (+)= , dp = d(pz), dls[c], dw (scalar); d(x) is the incoming gradient itself (the caller aliases the buffers) */
int cenet_ls_combine_fwd(const void* x, const void* y, const void* p, int dtype, const float* s, const float* t, const float* ls,
                         const float* w, void* out, long long rows, int C, cenet_stream_t st);
int cenet_ls_combine_bwd(const void* dout, const void* y, const void* p, int dtype, const float* s, const float* t, const float* ls,
                         const float* w, void* dy, int acc_dy, void* dp, float* dls, float* dw, long long rows, int C, float* ws,
                         long long ws_elems, cenet_stream_t st);
/* sparse separable resampling with CSR tap tables per axis (start[No+1], idx[nnz], weight[nnz]): adaptive average pooling, bilinear
 * up-sampling and their adjoints */
int cenet_resample(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, int B, int Hi, int Wi, int Ho, int Wo,
                   int C, const int* h_start, const int* h_idx, const float* h_w, const int* w_start, const int* w_idx, const float* w_w,
                   int acc, cenet_stream_t s);
/* backward of cenet_maxpool2_scale: drb (first arg max of each window), dw[c] */
int cenet_maxpool2_scale_bwd(const void* dz, int dtype, long long lddz, const void* rb, int rb_dtype, const float* w, void* drb, float* dw,
                             int B, int H, int W, int C, float* ws, long long ws_elems, cenet_stream_t s);
/* adjoint of the head's bilinear x2: dlogits [B,ncls,2h,2w] fp32 -> dyh [B,h,w,ncls] fp32 */
int cenet_head_upsample_bwd(const float* dlogits, float* dyh, int B, int h, int w, int ncls, cenet_stream_t s);
/* Re-pack of the fp32 master weights into the compute layouts (bf16 [N,K] / transposed / tap-permuted / flipped conv
 * filters, zero padded) in ONE launch: dst[i] = map[i] ? src[map[i]-1] : 0, cast to dst_dtype.  The index map is built once
 * by the host from the torch expressions the reference applies implicitly (`.weight` used as-is by F.linear / F.conv2d,
 * pvtv2.py:41-45, blocks.py:209-214); n % 4 == 0, map and dst 16-byte aligned. */
int cenet_gather_cast(const float* src, const int* map, void* dst, int dst_dtype, long long n, cenet_stream_t s);
/* AdamW over the flat fp32 parameter buffer; hyper (device) = [lr, beta1, beta2, eps, weight_decay, step] (utils/core.py:16-18) */
int cenet_adamw(float* p, const float* g, float* m, float* v, long long n, const float* hyper, cenet_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* CENET_B200_H */
