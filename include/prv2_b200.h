/* prv2_b200.h -- C ABI of the B200-native (sm_100a) PatchRefinerV2 inference hot path.
 *
 * The reference (zhyever/PatchRefinerV2) is pure Python: its "FFI" for this path is the set of
 * PyTorch / torchvision / OpenCV calls made inside the estimator model's infer forward.  Each
 * entry point below replaces one such call site; the citation is the reference file:line whose
 * arithmetic the kernel reproduces.  All pointers are DEVICE pointers unless stated; all sizes are
 * explicit; nothing allocates (one documented exception: the rN blend's per-geometry column tables, see
 * prv2_blend_raw); every function is asynchronous on `stream` and returns 0 on
 * success or a negative PRV2_E* code (prv2_last_error() gives the message).  No torch types.
 *
 * Activation tensors ("act") are channels-last 16-bit planes: ONE bf16 plane (lo == NULL, one-pass mode) or a (hi, lo) pair of
 * FP16 planes whose sum carries ~22 mantissa bits ("x3" / fp32-class precision mode).  prv2_bf16 is the raw 16-bit storage type
 * of both.
 */
#ifndef PRV2_B200_H_
#define PRV2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* prv2_stream_t;           /* cudaStream_t */
typedef uint16_t prv2_bf16;            /* raw bfloat16 bits */

#define PRV2_OK 0
#define PRV2_EINVAL (-1)               /* bad argument */
#define PRV2_ECUDA (-2)                /* CUDA runtime / driver error */
#define PRV2_EUNSUPPORTED (-3)         /* shape outside what the kernels implement */

/* ABI version: bumped whenever a signature or the GemmDesc layout changes; the Python binding refuses any other value. */
#define PRV2_ABI_VERSION 206
int prv2_version(void);
/* sha256 of the CUDA sources + flags this library was compiled from (stamped by build.py with -DPRV2_BUILD_DIGEST);
 * the binding compares it with the digest of the sources it sits next to, so a stale .so is an error, not a silent mismatch. */
const char* prv2_build_digest(void);
const char* prv2_last_error(void);
/* Device properties the host uses for grid sizing: out[0]=SM count, out[1]=cc major, out[2]=cc minor,
 * out[3]=max dynamic smem per block (opt-in).  Host pointer. */
int prv2_device_info(int32_t* out4);

/* ------------------------------------------------------------------------------------------
 * (1) geometry: crop + resize, ROI gather            [HBM-bound gathers]
 * ------------------------------------------------------------------------------------------ */

/* baseline_pretrain.py:272-280 (regular_tile) / :167-175 (random_tile) + external/depth_anything/
 * transform.py:127-129: out[p] = bilinear(align_corners=True)(image[:, y0:y1, x0:x1]) -> [ph,pw].
 * image [3,H,W] fp32; bboxs [P,4] int32 rows (x0,y0,x1,y1); out [P,3,ph,pw] fp32.  Bit-exact with
 * ATen's CPU kernel (fma(lx0,a,lx1*b) ordering). */
int prv2_crop_resize(const float* image, int H, int W, const int32_t* bboxs, int P,
                     float* out, int ph, int pw, prv2_stream_t stream);

/* patchrefiner.py:199-217 -> torchvision.ops.roi_align(feat.repeat(P), rois, (h,w), h/ph,
 * aligned=True) in the one-sample-per-bin regime.  feat is ONE channels-last map [h,w,C] (no
 * repeat is materialised); rois [P,4] fp32 rows (x1,y1,x2,y2) = bboxs_feat[:,1:]; out [P,h,w,C].
 * f32 variant: bit-exact with torchvision's CPU kernel (no FMA contraction). */
int prv2_roi_gather_f32(const float* feat, int h, int w, int C, const float* rois, int P,
                        float spatial_scale, float* out, prv2_stream_t stream);
/* act variant: in/out are bf16 (hi[,lo]) planes with channel pitch in_cs / out_cs (elements). */
int prv2_roi_gather_act(const prv2_bf16* feat_hi, const prv2_bf16* feat_lo, int h, int w, int C, int in_cs,
                        const float* rois, int P, float spatial_scale,
                        prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (3) consistency-aware-inference blend               [HBM-bound, fused]
 * ------------------------------------------------------------------------------------------ */

/* One regular-grid stage of the process-resolution canvas (baseline_pretrain.py:249-270). */
typedef struct {
  int32_t off_h, off_w;      /* canvas offset of patch (0,0) of this stage           */
  int32_t n_h, n_w;          /* patches per column / row                              */
  int32_t first;             /* index of the stage's first patch in `preds`           */
} prv2_grid_stage;

/* Sequential-exact canvas blend: RunningAverageMap.__init__/update (estimator/models/utils.py:
 * 24-36) applied in the reference's patch order (baseline_pretrain.py:347-373): stage 0 assigns
 * (pred, mask); stages 1..n_stages-1 update where mask>0.  preds [n_patches,ph,pw] fp32 in
 * reference order; mask [ph,pw]; avg,cnt [Hc,Wc] outputs (cnt may be NULL).  Bit-exact. */
int prv2_blend_canvas(const float* preds, const float* mask, int ph, int pw,
                      const prv2_grid_stage* stages /*host*/, int n_stages, int Hc, int Wc,
                      float* avg, float* cnt, prv2_stream_t stream);

/* rN stage (patchrefiner.py:385-392; utils.py:38-43; baseline_pretrain.py:204-229): A0 =
 * nearest(avg_c), C0 = bilinear_ac(cnt_c) to [H,W], then every random patch k (draw order) with
 * raw bbox origin (y0,x0) in `starts` [n,2] int32 (device), prediction preds[k] [ph,pw] nearest-
 * resized to [rh,rw], weight rmask [rh,rw] (already +1e-3): sequential update.  out/out_cnt [H,W]
 * (out_cnt may be NULL).  n may be 0 (pure resize).
 * ALLOCATION NOTE (the one exception to "nothing allocates"): the first call for a new (Wc, W, pw, rw) geometry on a
 * device cudaMalloc()s <= 64 KB of column tables that the library keeps for the life of the process (64 geometries per
 * process; beyond that the table-free generic kernel runs).  Call once per geometry before capturing a CUDA graph. */
int prv2_blend_raw(const float* avg_c, const float* cnt_c, int Hc, int Wc,
                   const float* preds, const int32_t* starts, int n, int ph, int pw,
                   const float* rmask, int rh, int rw, int H, int W,
                   float* out, float* out_cnt, const void* prep /*NULL or prv2_blend_raw_prepare output*/, prv2_stream_t stream);

/* Optional one-time preparation of the random-patch weight map for the three *_raw entry points (same results; with it, and with
 * W, Wc multiples of 4 and 16-byte aligned canvases, prv2_blend_raw / prv2_blend_finalize_raw run the segment kernel, which
 * stages the canvas rows in shared memory by bulk copies: 41 instead of 60 us for the r32 stage of a 2160x3840 frame):
 * `prep` (device, 16-byte aligned, prv2_blend_raw_prep_bytes(rh, rw) bytes) receives four column-shifted,
 * zero-padded copies of rmask [rh, rw] and of the nearest-source-column table of a [*, pw] prediction (baseline_pretrain.py:210;
 * as u16 and as i32), so that the four weights / source columns of any aligned group of four output pixels are ONE 16-byte
 * vector.  The weight map depends only on (rh, rw): prepare it once per geometry and pass it to every frame's calls. */
int64_t prv2_blend_raw_prep_bytes(int rh, int rw);
int prv2_blend_raw_prepare(const float* rmask, int rh, int rw, int pw, void* prep, prv2_stream_t stream);

/* Patch-sharded form (multi-GPU): each rank adds ITS patches into packed partial
 * sums, one NCCL sum-reduce combines them, `finalize` normalises.  own[k]!=0 marks patches this
 * rank owns (device uint8).  num_c [Hc,Wc] += mask*pred for stages>=1; m1 [Hc,Wc] = pred for
 * stage 0 (disjoint); num_r [H,W] += rmask*nearest(pred) for random patches. */
int prv2_blend_partial_canvas(const float* preds, const uint8_t* own, const float* mask, int ph, int pw,
                              const prv2_grid_stage* stages /*host*/, int n_stages, int Hc, int Wc,
                              float* num_c, float* m1, prv2_stream_t stream);
int prv2_blend_partial_raw(const float* preds, const uint8_t* own, const int32_t* starts, int n, int ph, int pw,
                           const float* rmask, int rh, int rw, int H, int W,
                           float* num_r, const void* prep, prv2_stream_t stream);
/* avg = cnt>cnt0 ? (m1*cnt0 + num_c)/cnt : m1, cnt recomputed locally in reference order. */
int prv2_blend_finalize_canvas(const float* num_c, const float* m1, const float* mask, int ph, int pw,
                               const prv2_grid_stage* stages /*host*/, int n_stages, int Hc, int Wc,
                               float* avg, float* cnt, prv2_stream_t stream);
/* out = (A0*C0 + num_r)/(C0 + cnt_r) with cnt_r recomputed locally in draw order. */
int prv2_blend_finalize_raw(const float* avg_c, const float* cnt_c, int Hc, int Wc, const float* num_r,
                            const int32_t* starts, int n, const float* rmask, int rh, int rw, int H, int W,
                            float* out, float* out_cnt, const void* prep, prv2_stream_t stream);
/* Test / A-B hook: bit 0 routes every blend entry point through the any-alignment generic kernels instead of the
 * aligned fast paths; bit 1 switches the rN stage's segment kernel off (the table kernel runs); bits 8-15 force the segment
 * kernel's warps per CTA (0 = automatic).  Every path produces the same bits; tests assert it.  The value 0x10000 changes nothing
 * and returns how many *_raw calls the segment kernel has taken so far (tests use it to know which path they exercised). */
int prv2_debug_blend_generic(int on);

/* ------------------------------------------------------------------------------------------
 * (2) per-patch network: tcgen05 implicit-GEMM (linear layers AND convolutions)
 * ------------------------------------------------------------------------------------------ */

#define PRV2_MAX_SRC 12
#define PRV2_MAX_SEG 128

/* One channels-last bf16 source tensor [N,H,W,C] with channel pitch cs (elements, multiple of 8). */
typedef struct {
  const prv2_bf16* ptr;
  int32_t C;                 /* channels used from this source (K extent)             */
  int32_t cs;                /* channel pitch of the allocation                        */
} prv2_src;

/* One K-segment: all channels of source `src` seen through spatial offset (dh,dw).  The packed
 * weight matrix holds the segments back to back, each padded to a multiple of 64 columns.
 * taps_h == 3 makes the segment a vertical tap GROUP: the three offsets (dh-1,dw), (dh,dw), (dh+1,dw)
 * of a 3x3 conv share ONE activation fetch (a tile with a one-row halo above and below); its weight
 * columns are interleaved per 64-channel block: [block c: tap dh-1 | tap dh | tap dh+1], c = 0, 1, ...
 * A source is used either by groups or by single-tap segments, not both. */
typedef struct {
  int16_t src, dh, dw, taps_h;   /* taps_h: 0 or 1 = single tap, 3 = vertical group */
} prv2_seg;

#define PRV2_ACT_NONE 0
#define PRV2_ACT_RELU 1
#define PRV2_ACT_GELU 2       /* exact erf GELU (torch.nn.GELU default)                 */
#define PRV2_ACT_GELU_TANH 3  /* tanh-form GELU (|diff| <= 1e-3 abs vs erf); one-pass bf16 mode only */
#define PRV2_ACT_SIGMOID_GATE 4 /* EPI_STORE only: out = res * sigmoid(acc+bias)  (GatedConvUnit, bi_directional_fusion_model.py:70-78) */
#define PRV2_ACT_IDENTITY 5   /* LN epilogue only: LayerNorm without an activation (bi_directional_fusion_model.py:448-463); act 0 there means GELU */

#define PRV2_EPI_STORE 0      /* act(acc+bias) [+ residual act] -> out (and optional relu copy) */
#define PRV2_EPI_LN_GELU 1    /* channels-first LayerNorm over Cout of (acc+bias), then act = GELU | GELU_TANH (convs.py:21-29,64-75) | RELU (GatedConvUnit.fusion_conv) */
#define PRV2_EPI_RESID_F32 2  /* x_f32[m,n] += gamma[n]*(acc+bias[n])  (block.py:105-106, layer_scale.py:27) */
#define PRV2_EPI_F32 3        /* out_f32[m,n] = acc+bias                                 */
#define PRV2_EPI_SHUFFLE 4    /* ConvTranspose2d k==stride: n=(ky,kx,co) scattered to [N,H*k,W*k,Cout] (dpt.py:62-73) */
#define PRV2_EPI_HEAD 5       /* relu(acc+bias) . w2 + b2 -> sigmoid * max_depth -> f32 (dpt.py:109-114,190) */

typedef struct {
  /* problem: D[m, n] = sum_seg sum_c A_src[n_img, h+dh, w+dw, c] * Wt[n, kcol] */
  int32_t N, H, W;           /* output pixel grid (== input grid; stride-2 convs are fed phase-split sources) */
  int32_t Cout;              /* real output channels                                   */
  int32_t tile_w, tile_h;    /* spatial shape of the 128-pixel M tile (tile_w*tile_h==128) */
  int32_t block_n;           /* UMMA N (multiple of 16, <=256)                          */
  int32_t n_src, n_seg;
  prv2_src src[PRV2_MAX_SRC];
  prv2_seg seg[PRV2_MAX_SEG];
  const prv2_bf16* weight;   /* [Cout_pad, Ktot] bf16, K-major; Ktot = sum_seg ceil64(C) */
  int32_t Cout_pad, Ktot;
  /* epilogue */
  int32_t epi, act;
  const float* bias;         /* [Cout] or NULL                                          */
  const float* gamma;        /* LN weight / LayerScale gamma / head w2                   */
  const float* beta;         /* LN bias / head b2 (1 element)                            */
  float eps, head_scale;
  prv2_bf16* out_hi; prv2_bf16* out_lo; int32_t out_cs;       /* main act output         */
  prv2_bf16* relu_hi; prv2_bf16* relu_lo; int32_t relu_cs;    /* optional relu(out) copy */
  const prv2_bf16* res_hi; const prv2_bf16* res_lo; int32_t res_cs;  /* optional residual act */
  const prv2_bf16* res2_hi; const prv2_bf16* res2_lo; int32_t res2_cs;
  float* out_f32; int32_t out_f32_ld;                          /* RESID_F32 / F32 / HEAD  */
  int32_t shuffle_k;         /* EPI_SHUFFLE: kernel==stride                              */
  int32_t row_map_period, row_map_extra, row_map_offset;       /* out row = m + (m/period)*extra + offset (0 period = identity) */
  /* fp32-class ("x3") mode: sources and weights are (hi, lo) pairs of FP16 planes (f16 = 1; bf16 planes when 0).  FP16 has a
   * narrow exponent, so the host scales each layer's weights by a power of two that centres them in the fp16 range (the lo plane
   * stays normal) and passes the inverse here: every accumulator is multiplied by acc_scale before bias / activation (0 = 1). */
  float acc_scale;
  int32_t f16;
} prv2_gemm_desc;

/* attention.py:44-46 / mlp.py:30-32 (nn.Linear), patch_embed.py:69-82, dpt.py:48-80,116-150,
 * util/blocks.py:57-80,123-148, fusion_model.py:84-122, convs.py:31-75: every dense contraction
 * on the path, as one persistent TMA + tcgen05.mma (TMEM accumulator) kernel.  Host pointer. */
int prv2_umma_gemm(const prv2_gemm_desc* desc, prv2_stream_t stream);

/* attention.py:49-62: softmax(q k^T / sqrt(64)) v, all heads; qkv [B, T, 3*D] act (hi[,lo]),
 * out [B, T, D] act.  head_dim is 64 for every DINOv2 size. */
int prv2_attention(const prv2_bf16* qkv_hi, const prv2_bf16* qkv_lo, int B, int T, int heads,
                   prv2_bf16* out_hi, prv2_bf16* out_lo, prv2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2) per-patch network: hand-written pointwise / normalisation / resampling kernels
 * ------------------------------------------------------------------------------------------ */

/* nn.LayerNorm(eps) over the last dim of an fp32 matrix x [rows, D] -> act.  If drop_period>0 the
 * first row of every `drop_period` rows (the class token, dinov2.py:311) is skipped and the output
 * is compacted (block.py:56,68; dinov2.py:309-312). */
int prv2_layernorm(const float* x, int rows, int D, const float* w, const float* b, float eps,
                   int drop_period, prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream);
/* Channels-first LayerNorm + exact GELU of a conv output held as fp32 rows [pixels, D] (SingleConvCNNLN, convs.py:64-75) for
 * channel counts the GEMM's fused LN epilogue cannot hold in one N tile (D > 256; BiDirectionalFusion's 512-channel level). */
int prv2_layernorm_gelu(const float* x, int rows, int D, const float* w, const float* b, float eps,
                        prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream);

/* dpt.py:183 + patch_embed.py:69-82 im2col: crops [B,3,H,W] fp32 in [0,1] -> (x-mean)/std ->
 * rows (b,ty,tx), cols c*196+ky*14+kx, zero padded to Kp columns. */
int prv2_patchify(const float* crops, int B, int H, int W, prv2_bf16* out_hi, prv2_bf16* out_lo, int Kp,
                  prv2_stream_t stream);

/* dinov2.py:218-219: x[b,0]=cls+pos[0]; x[b,1+t]=emb[b,t]+pos[1+t].  pos is the already
 * interpolated table [T+1, D]. */
int prv2_assemble_tokens(const float* emb, const float* cls, const float* pos, int B, int T, int D,
                         float* x, prv2_stream_t stream);

/* F.interpolate(bilinear, align_corners=True) on channels-last act tensors
 * (util/blocks.py:144, dpt.py:146, fusion_model.py:16). relu!=0 applies ReLU after. */
int prv2_resize_bilinear_act(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int h, int w, int C, int in_cs,
                             prv2_bf16* out_hi, prv2_bf16* out_lo, int oh, int ow, int out_cs, int relu,
                             prv2_stream_t stream);

/* fusion_model.py:94-96,16-17: the reference concatenates the two depth maps [N,1,H,W] fp32, resized (bilinear ac) to the
 * level [oh,ow], to a feature tensor that feeds a 3x3 conv.  Here their 3x3 neighbourhoods are laid out once per level as
 * an 18-channel im2col tensor: out[n,y,x,(r*3+s)*2+d] = resized(pred_d)[y+r-1, x+s-1] (0 outside = the conv's zero
 * padding), channels 18..23 zero; the conv consumes it as one 1x1 K segment. */
int prv2_depth_taps(const float* pred1, const float* pred2, int N, int H, int W,
                    prv2_bf16* out_hi, prv2_bf16* out_lo, int oh, int ow, int out_cs, prv2_stream_t stream);

/* fusion_model.py:113-118: offset = conv3x3(feat, w[1,C,3,3], no bias); out = clamp(base+offset, 0).
 * The channel contraction runs through prv2_umma_gemm as a 1x1 conv with 9 outputs (one per tap):
 * taps [N,H,W,ld] fp32, taps[p, r*3+s] = sum_c feat[p,c] * w[0,c,r,s].  This stencil then forms
 * out[p] = clamp(base[p] + sum_{r,s} taps[p + (r-1, s-1), r*3+s], 0) with zero padding (base may
 * be NULL -> offset only, no clamp).  out fp32 [N,H,W]. */
int prv2_tap_stencil(const float* taps, int N, int H, int W, int ld, const float* base, float* out, prv2_stream_t stream);
/* The whole final conv fused: out[n,y,x] = clamp(base + sum_{r,s,c} feat[n,y+r-1,x+s-1,c] * w9c[(r*3+s)*C + c], 0) (no clamp when
 * base is NULL); feat is a channels-last act (hi [+lo]), w9c fp32 [9,C] (final_conv.weight [1,C,3,3] permuted), out fp32 [N,H,W].
 * Replaces FusionUnet.final_conv (estimator/models/blocks/fusion_model.py:113-118) in one HBM-bound pass. */
int prv2_final_conv3x3(const prv2_bf16* feat_hi, const prv2_bf16* feat_lo, int N, int H, int W, int C, int cs, const float* w9c,
                       const float* base, float* out, prv2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ZoeDepth metric-bins head, per-pixel part (the head's 1x1 convolutions run on prv2_umma_gemm)    [HBM-bound, fp32]
 * ------------------------------------------------------------------------------------------ */

/* AttractorLayerUnnormed.forward (external/zoedepth/models/layers/attractor.py:139-208) after its 1x1-conv MLP: a_raw [B,h,w,a_ld]
 * holds the MLP output (pre-softplus; the first n_attractors columns are used), b_prev [B,hp,wp,n_bins] the previous level's bin
 * centres (prev_is_raw != 0: they are the seed regressor's PRE-softplus output, localbins_layers.py:71-96).  b_out [B,h,w,n_bins] =
 * bilinear_ac(b_prev) + mean_i (or sum_i) inv_attractor(softplus(a_i) - bilinear_ac(b_prev)), inv_attractor(dx) = dx / (1 + alpha dx^2)
 * (attractor.py:45-57; the reference calls it with its default alpha = 300 whatever the config says, :194-197). */
int prv2_zoe_attractor(const float* a_raw, int a_ld, int n_attractors, const float* b_prev, int hp, int wp, int prev_is_raw,
                       float* b_out, int B, int h, int w, int n_bins, float alpha, int mean, prv2_stream_t stream);
/* ConditionalLogBinomial tail + LogBinomial + expectation (layers/dist_layers.py:29-122, zoedepth_v1.py:212-219): pt_raw [B,H,W,pt_ld]
 * = the 4 pre-softplus outputs of the conditional MLP; centers [B,hb,wb,n_bins] the final bin centres; depth [B,1,H,W] =
 * sum_k softmax((log C(K-1,k) + k log p + (K-1-k) log(1-p)) / T)_k * bilinear_ac(centers)_k. */
int prv2_zoe_logbinomial_depth(const float* pt_raw, int pt_ld, const float* centers, int hb, int wb, float* depth,
                               int B, int H, int W, int n_bins, float min_temp, float max_temp, prv2_stream_t stream);

/* Depthwise k x k convolution (k in {3, 5}, stride in {1, 2}, zero padding k/2) on channels-last acts with BatchNorm folded into
 * w [k*k, C] fp32 (tap-major) and bias [C] (may be NULL); relu != 0 applies ReLU.  Output [N, (H-1)/stride+1, (W-1)/stride+1, C].
 * The dw_start / dw_mid stages of timm's UniversalInvertedResidual (MobileNetV4), the encoder that
 * estimator/models/blocks/lightweight_refiner.py:259-262 creates through timm.create_model(features_only=True). */
int prv2_dwconv(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int H, int W, int C, int in_cs, const float* w, const float* bias,
                int k, int stride, int relu, prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream);

/* lightweight_refiner.py:293-298: out[n,y,x,0..2] = (crops[n,c,y,x] - mean[c]) / std[c], out[..,3] = depth[n,0,y,x] (depth may be NULL:
 * coarse_condition=False), channels 4.. zero; mean3 / std3 are HOST arrays of three floats. */
int prv2_encoder_input(const float* crops, const float* depth, int N, int H, int W, const float* mean3, const float* std3,
                       prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream);

/* Multi-GPU exchange (the reference has no multi-GPU inference: tester.py:62-69 runs one frame per process): in-place
 * ncclAllReduce(sum, float32) of the packed partial buffer [num_canvas (Hc*Wc) | m1_canvas (Hc*Wc) | num_raw (H*W)] written by
 * prv2_blend_partial_*, on the caller's ncclComm_t and stream.  NCCL is resolved from the process at run time (dlopen of
 * libnccl.so.2): no link-time dependency.  PRV2_ECUDA when NCCL is absent or the collective fails. */
int prv2_reduce_canvas(void* nccl_comm /*ncclComm_t*/, float* packed, int64_t count, prv2_stream_t stream);

/* act <-> fp32 helpers (layout changes at the API edge and for tests). */
int prv2_nchw_f32_to_act(const float* in, int N, int C, int H, int W,
                         prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream);
int prv2_act_to_nchw_f32(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int C, int H, int W, int in_cs,
                         float* out, prv2_stream_t stream);
/* 2x2 space-to-depth phase split for the stride-2 conv (dpt.py:75-80): out[p] [N,ceil(H/2),ceil(W/2),C], p=(py,px); zero where an odd
 * H / W leaves a phase one row / column short (that row / column is the conv's padding). */
int prv2_phase_split(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int H, int W, int C, int in_cs,
                     prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream);
/* fp32 [rows, cols] -> bf16 hi[,lo] with pitch (weights packing helper). */
int prv2_split_f32(const float* in, int64_t n, prv2_bf16* hi, prv2_bf16* lo, prv2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* PRV2_B200_H_ */
