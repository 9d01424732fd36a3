/* tta_b200.h -- C ABI of libtta_b200.so: the sm_100a kernels of the test-time-adaptation step.
 *
 * This is the drop-in boundary underneath the reference's (pure Python) plugin surface.  The
 * reference has no FFI of its own; the Python classes that mirror its registry interface
 * (multimodal_tta_b200/{unet_b200,tent,evaluation,sliding_window}.py) bind these entry points with
 * ctypes (multimodal_tta_b200/_lib.py) -- INTEGRATION.md shows the binding.  Each declaration
 * cites the reference code (or the third-party op the reference reaches) it replaces; paths are
 * relative to the reference repository (zhm1205/Multimodal_TTA).
 *
 * Conventions
 *   - plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer unless
 *     the parameter is documented as HOST; the caller owns all buffers; nothing is allocated.
 *   - every call is asynchronous on `stream` and may be captured into a CUDA graph.
 *   - return value: 0 = OK, 1 = bad argument, 2 = CUDA error, 3 = unsupported; the message of the
 *     last failure on the calling thread is returned by tta_last_error().
 *   - layouts: channel-blocked [N][C8][D][H][W][8], C8 = ceil(C/8), pad
 *     channels zero.  "f32 view": fp32 tensor in that layout.  "planes view": two 16-bit planes
 *     hi/lo with x ~= hi + lo.  A view is (pointer, n_stride in ELEMENTS between samples, C8, dims):
 *     a channel slice of a wider buffer is the same n_stride with an offset pointer.
 *   - dtype tags of planes: 0 = fp16 hi/lo, 1 = bf16 hi/lo, 2 = one fp16 plane (lo unused, may be
 *     NULL): loss-scaled gradients in the backward pass.
 */
#ifndef TTA_B200_H
#define TTA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tta_stream_t; /* cudaStream_t */

#define TTA_F16 0
#define TTA_BF16 1
#define TTA_F16_HI 2

/* ---- library ------------------------------------------------------------------------------ */
const char* tta_last_error(void);
int tta_abi_version(void);
int tta_device_sm(void); /* compute capability major*10+minor of the current device, <0 on error */

/* ---- input staging: replaces the H2D'd NCDHW batch of src/core/trainers/seg_trainer.py:107 and
 * MONAI sliding_window_inference's window slicing + constant pad (not in tree).
 * vol: fp32 NCDHW [n_vol][C][Ds][Hs][Ws]; win: int32 [NB][4] = (volume, d0, h0, w0), origins may
 * lie outside the volume (zero fill); chan_scale: optional fp32 [NB][C] (missing-modality dropout).
 * Output: fp16 hi/lo planes [NB][C8][D][H][W][8]; wsplit = 2: the COMPACT layout [NB][D][H][W][4] (<= 4 channels,
 * C8 = 1, out_n_stride = D*H*W*4) read by tta_conv_tc with flags bit 15; wsplit = 1 stores every w-row parity-split
 * ([H][2][W/2][8]: even-w voxels first), the operand layout of a stride-2 tcgen05 conv (flags bit3). */
int tta_gather_pack(const float* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win,
                    const float* chan_scale, int NB, int D, int H, int W, uint16_t* hi, uint16_t* lo,
                    long long out_n_stride, int C8, int wsplit, tta_stream_t stream);
/* same, with the intensity policy applied on the fly: affine [n_vol][C][4] = {lo, hi, mu, 1/sd} from
 * tta_intensity_stats (NULL: identity).  Bit-identical to tta_intensity_apply followed by tta_gather_pack. */
int tta_gather_pack_norm(const float* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win,
                         const float* chan_scale, const float* affine, int NB, int D, int H, int W, uint16_t* hi,
                         uint16_t* lo, long long o_n_stride, int C8, int wsplit, tta_stream_t stream);
/* same, the volume staged as IEEE fp16 (half the host -> device bytes of an fp32 batch) */
int tta_gather_pack_norm_f16(const uint16_t* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win,
                             const float* chan_scale, const float* affine, int NB, int D, int H, int W, uint16_t* hi,
                             uint16_t* lo, long long o_n_stride, int C8, int wsplit, tta_stream_t stream);

/* ---- convolution (forward and input gradient): replaces nn.Conv3d / nn.ConvTranspose3d inside
 * monai.networks.blocks.Convolution reached from src/models/unet.py:56-66, and autograd's
 * conv backward-data in loss.backward() (src/core/trainers/seg_trainer.py:142).  No weight
 * gradient exists: TENT freezes conv weights.
 * mode 0: out[o] = sum_k in[s*o - p + k] * Wg[k]      (conv; also the dgrad of a transposed conv)
 * mode 1: out[o] = sum_{k,i: s*i - p + k = o} in[i] * Wg[k]   (transposed conv; dgrad of a conv)
 * K in {1,3}, stride in {1,2}, p = (K-1)/2. */

/* tcgen05/TMEM/TMA implicit GEMM.  wpacked: layout.pack_weights_tc blob for (mode,K,stride,dtype).
 * flags: bit0 force one d-plane per work item, bit1 no split-K (deterministic), bit2 non-persistent,
 * bit3 stride-2 conv input planes are w-parity-split ([N][C8][D][H][2][W/2][8]), bit4 no resident
 * weights, bit5 one d-plane per 128-row tile even when H <= 8 (default there: two planes per tile),
 * bit6 split-K factor rounded up (A/B switch), bit7 / bit11 kd-stacked convs with one / at most two halo planes per stage (A/B),
 bit12 CTA pairs: 2-CTA clusters with multicast weight blobs (opt-in; measured neutral), bit13 nine (kd, kh)
 * pipeline groups for ordinary planes (opt-in; measured slower),
 * bits 8..10 real Cout of a small-Cout transposed conv
 * packed for the dense-GEMM + col2im kernel (tta_conv_tc_t2s),
 * bit17 EXPERIMENT: two products instead of three (A_hi * [B_hi | B_lo] only; measured: logits 1.2e-5 -> 6.8e-4, agreement
 * below 99.99 %, 0.8 % faster -- not used),
 * bit16 `out` is ONE fp16 plane in the chunk layout (16 B per voxel-chunk; out_n_stride in elements): the loss-scaled
 * input gradient of a dgrad, consumed by tta_norm_bwd_* with the fp16-source bits; implies bit1, accumulate becomes a
 * 16-byte read-modify-write. */
int tta_conv_tc_supported(int mode, int K, int stride, int cin, int cout);
int tta_conv_tc_ntile(int mode, int K, int stride, int cout, int split);
/* 1 when this layer runs kd-stacked (stride-1, single small n-tile, resident weights): its packed
 * weight layout differs (layout.pack_weights_tc), so packer and kernel ask the same question. */
int tta_conv_tc_stacked(int mode, int K, int stride, int cin, int cout, int split);
/* 1 when a transposed stride-2 conv qualifies for the dense-GEMM + col2im kernel (Cout <= 4, Cin a
 * multiple of 16 up to 64, split planes): pack with layout.pack_weights_tc(t2s=True) and pass the real
 * Cout in flags bits 8..10 of tta_conv_tc. */
int tta_conv_tc_t2s(int mode, int K, int stride, int cin, int cout, int split);
/* 1 when a stride-2 conv reads a one-chunk input (<= 8 channels): taps are packed and issued in pairs. */
int tta_conv_tc_s2pair(int mode, int K, int stride, int cin);
/* 1 when a stride-2 conv may read a COMPACT <= 4-channel input ([N][D][H][W][4] 16-bit values per plane, in_n_stride
 * in elements; tta_conv_tc flags bit 15): weights packed with layout.pack_weights_tc(s2c4=True).  Producers of the
 * compact layout: tta_gather_pack_norm(_f16) with wsplit = 2, tta_norm_bwd_apply_c4 with dy_c4 = 1. */
int tta_conv_tc_s2c4(int mode, int K, int stride, int cin);
int tta_conv_tc_gmax(int mode, int K, int stride);
int tta_conv_tc_ngroups(int mode, int K, int stride);
long long tta_conv_tc_packed_bytes(int mode, int K, int stride, int cin, int cout);
int tta_conv_tc(const uint16_t* in_hi, const uint16_t* in_lo, long long in_n_stride, int in_dtype, int N,
                int C8in, int Di, int Hi, int Wi, const void* wpacked, const float* bias, float* out,
                long long out_n_stride, int C8out, int Do, int Ho, int Wo, int mode, int K, int stride,
                int accumulate, int flags, float* stats_workspace, int stats_c8, tta_stream_t stream);
/* launch shape for the same arguments: *ksplit = split-K factor, *grid = CTAs, *nbuf = TMEM
 * accumulator buffers (2 = the epilogue overlaps the next work item's MMAs; may be NULL).  With
 * stats_workspace != NULL (requires ksplit == 1, accumulate == 0) the epilogue also leaves per-CTA
 * partial sums of y and y^2 over the leading stats_c8 output chunks in the workspace, which
 * tta_norm_stats_finalize(splits = grid) / tta_norm_apply(partial, partial_splits = grid) consume:
 * the separate statistics pass over y (tta_norm_stats) disappears. */
int tta_conv_tc_query(int in_dtype, int N, int C8in, int Di, int Hi, int Wi, int C8out, int Do, int Ho, int Wo,
                      int mode, int K, int stride, int accumulate, int flags, int* ksplit, int* grid, int* nbuf);

/* input-gradient conv whose result is the COMPLETE gradient w.r.t. the outputs of up to two norm
 * layers (channel segments of the output view, e.g. the two halves of a skip concat): the epilogue
 * also leaves, per segment, per-CTA partial sums of dz = g*[gamma*xhat+beta > 0] and dz*xhat in
 * seg.partial ([N][c8_count][grid][16] floats, grid from tta_conv_tc_query), reduced by
 * tta_norm_bwd_finalize(splits = grid) -- autograd's norm backward reduction (loss.backward(),
 * src/core/trainers/seg_trainer.py:142) without a separate pass over g and y.  Requires one fp16
 * plane (in_dtype 2) and ksplit == 1.  segs: HOST pointer. */
typedef struct tta_norm_bwd_seg {
  int c8_begin, c8_count, relu, pad;
  const float* y;          /* f32 view of the norm layer's conv result, same spatial dims as `out` */
  long long y_n_stride;
  const float* mean;       /* [N][c8_count*8] */
  const float* rstd;
  const float* gamma;      /* [c8_count*8] */
  const float* beta;
  float* partial;
} tta_norm_bwd_seg;
int tta_conv_tc_bwd_norm(const uint16_t* in_hi, const uint16_t* in_lo, long long in_n_stride, int in_dtype, int N,
                         int C8in, int Di, int Hi, int Wi, const void* wpacked, float* out, long long out_n_stride,
                         int C8out, int Do, int Ho, int Wo, int mode, int K, int stride, int accumulate, int flags,
                         const tta_norm_bwd_seg* segs, int nsegs, tta_stream_t stream);

/* fp32 CUDA-core conv for any (mode, K, stride); Wp: fp32 [tap][C8in][C8out][8][8]. */
int tta_conv_simt(const uint16_t* in_hi, const uint16_t* in_lo, long long in_n_stride, int in_dtype, int N,
                  int C8in, int Di, int Hi, int Wi, const float* Wp, const float* bias, float* out,
                  long long out_n_stride, int C8out, int Do, int Ho, int Wo, int mode, int K, int stride,
                  int accumulate, tta_stream_t stream);

/* direct 3x3x3 stride-1 conv for cin, cout <= 4 (the UNet head); W_host: HOST fp32 [27][8][8]. */
int tta_conv_small_supported(int K, int stride, int cin, int cout);
int tta_conv_small(const uint16_t* in_hi, const uint16_t* in_lo, long long in_n_stride, int in_dtype, int N,
                   int cin, int D, int H, int W, const float* W_host, const float* bias, float* out,
                   long long out_n_stride, int cout, int accumulate, tta_stream_t stream);

/* ---- normalisation: replaces nn.InstanceNorm3d / nn.BatchNorm3d (+ nn.ReLU, the ResidualUnit add
 * `cx + res`, and torch.cat of SkipConnection via channel-slice views) of MONAI's ADN block,
 * selected by configs/_global_patches/brats.yaml:17 / src/models/unet.py:44, and their autograd
 * backward.  batch_mode 0: per-(n,c) statistics (InstanceNorm); 1: per-c over the batch
 * (BatchNorm in TENT mode, no running statistics). */
long long tta_norm_workspace_floats(int N, int C8, long long V);
/* per-block partial sums of y, y^2 into workspace; finalize != 0 also writes mean/rstd [N][C8*8] */
int tta_norm_stats(const float* y, long long y_n_stride, int N, int C8, long long V, int batch_mode, float eps,
                   float* mean, float* rstd, float* workspace, int finalize, tta_stream_t stream);
/* out = relu?(gamma*(y-mean)*rstd+beta) (+ residual) as planes.  res_kind 0 none, 1 f32 view
 * (res_a), 2 planes view (res_a = hi, res_b = lo).  partial != NULL: statistics are finalized from
 * the workspace inside this kernel and mean/rstd are written for the backward pass.
 * ws_hi/ws_lo != NULL: a second copy of the planes in the w-parity-split layout (row length W) for a
 * stride-2 tcgen05 consumer (skip tensors feed both the next level's strided conv and the concat). */
int tta_norm_apply(const float* y, long long y_n_stride, int N, int C8, long long V, const float* mean,
                   const float* rstd, const float* gamma, const float* beta, int relu, int res_kind,
                   const void* res_a, const void* res_b, long long res_n_stride, uint16_t* out_hi,
                   uint16_t* out_lo, long long out_n_stride, int out_dtype, const float* partial,
                   int partial_splits, int batch_mode, float eps, uint16_t* ws_hi, uint16_t* ws_lo,
                   long long ws_n_stride, int W, tta_stream_t stream);
/* mean/rstd from the partial sums in a workspace with `splits` slots per (n, chunk) */
int tta_norm_stats_finalize(const float* workspace, int N, int C8, int splits, long long V, int batch_mode,
                            float eps, float* mean, float* rstd, tta_stream_t stream);
/* Gradient sources of the three norm-backward entry points: g0 / g1 are fp32 chunk tensors, or -- when bit 1 / bit 2
 * of the `relu` argument is set (bit 0 stays the ReLU switch) -- ONE loss-scaled fp16 plane each, as a tcgen05 dgrad
 * writes it with tta_conv_tc flags bit 16 (n strides stay in elements).
 * partial sums of dz and dz*xhat (dz = (g0+g1)*[z>0]); finalize != 0 also writes sums[N][C][2] and
 * dgamma/dbeta[C] (accumulate_dgb != 0: adds to them -- a layer processed sample by sample, each call
 * with N = 1 and pointers advanced by one sample, so that the apply pass re-reads g and y from L2) */
int tta_norm_bwd_reduce(const float* g0, long long g0_n_stride, const float* g1, long long g1_n_stride,
                        const float* y, long long y_n_stride, int N, int C8, int Creal, long long V,
                        const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                        int batch_mode, float* sums, float* dgamma, float* dbeta, float* workspace, int finalize,
                        int accumulate_dgb, tta_stream_t stream);
/* sums / dgamma / dbeta from partial sums laid out [N][C8][splits][16] (partial points at the slots) */
int tta_norm_bwd_finalize(const float* partial, int N, int C8, int Creal, int splits, int batch_mode, float* sums,
                          float* dgamma, float* dbeta, tta_stream_t stream);
/* dy = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) as planes; optional aux planes = g0+g1
 * (feeds the shortcut conv's dgrad).  partial != NULL: reductions finalized here, dgamma/dbeta written.
 * dy_wsplit_w > 0: dy is stored w-parity-split with row length W = dy_wsplit_w (its dgrad is a
 * stride-2 tcgen05 conv). */
int tta_norm_bwd_apply(const float* g0, long long g0_n_stride, const float* g1, long long g1_n_stride,
                       const float* y, long long y_n_stride, int N, int C8, long long V, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, int relu, int batch_mode,
                       const float* sums, uint16_t* dy_hi, uint16_t* dy_lo, long long dy_n_stride,
                       uint16_t* aux_hi, uint16_t* aux_lo, long long aux_n_stride, int out_dtype,
                       const float* partial, int Creal, float* dgamma, float* dbeta, int dy_wsplit_w,
                       tta_stream_t stream);
/* small layers (2 <= V <= 4096 voxels per instance, InstanceNorm): statistics + apply, and backward
 * reduction + apply, as ONE launch each (a CTA owns a whole (n, chunk) slab).  Same math and outputs
 * as tta_norm_stats + tta_norm_apply / tta_norm_bwd_reduce + tta_norm_bwd_apply. */
int tta_norm_small_supported(int N, long long V, int batch_mode);
int tta_norm_fwd_small(const float* y, long long y_n_stride, int N, int C8, long long V, float eps, float* mean,
                       float* rstd, const float* gamma, const float* beta, int relu, int res_kind,
                       const void* res_a, const void* res_b, long long res_n_stride, uint16_t* out_hi,
                       uint16_t* out_lo, long long out_n_stride, int out_dtype, uint16_t* ws_hi, uint16_t* ws_lo,
                       long long ws_n_stride, int W, tta_stream_t stream);
int tta_norm_bwd_small(const float* g0, long long g0_n_stride, const float* g1, long long g1_n_stride,
                       const float* y, long long y_n_stride, int N, int C8, int Creal, long long V,
                       const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                       float* sums, float* dgamma, float* dbeta, uint16_t* dy_hi, uint16_t* dy_lo,
                       long long dy_n_stride, uint16_t* aux_hi, uint16_t* aux_lo, long long aux_n_stride,
                       int out_dtype, int dy_wsplit_w, float* workspace, tta_stream_t stream);
int tta_split_f32(const float* g0, long long g0_n_stride, const float* g1, long long g1_n_stride, int N, int C8,
                  long long V, uint16_t* hi, uint16_t* lo, long long out_n_stride, int out_dtype,
                  tta_stream_t stream);

/* ---- fused full-resolution tail (csrc/tta_head.cu): the last ADN(norm, ReLU) -> 3x3x3 conv(C->C,
 * identity shortcut folded) -> entropy of MONAI's top ResidualUnit(last_conv_only) as ONE kernel, and
 * its backward (conv dgrad + ReLU mask + norm-backward reduction) as ONE kernel.  Replaces, for
 * C <= 4 heads, tta_norm_apply + tta_conv_small + tta_head_entropy and tta_conv_small(dgrad) +
 * tta_norm_bwd_reduce; semantics as those (src/core/trainers/seg_trainer.py:105-145).
 * y: f32 view [N][1][D][H][W][8]; mean/rstd [N][8]; W_host: HOST fp32 [27][8][8] = Wg[tap][ci][co];
 * logits/dlogits: NCDHW fp32 (dlogits NULL = inference); dz: f32 view of the masked gradient that
 * tta_norm_bwd_apply then consumes as g0; sums/dgamma/dbeta as tta_norm_bwd_reduce(finalize=1). */
int tta_head_fused_supported(int K, int stride, int cin, int cout);
int tta_head_fused_tiles(int D, int H, int W);
long long tta_head_fused_workspace_floats(int N, int D, int H, int W);
int tta_head_fused_fwd(const float* y, long long y_n_stride, int y_cpv, int N, int C, int D, int H, int W, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, int relu, const float* W_host,
                       const float* bias, int mode, float inv_count, float grad_scale, const float* sample_w,
                       float* logits, float* dlogits, float* workspace, float* loss, tta_stream_t stream);
int tta_head_fused_bwd(const float* dlogits, int N, int C, int D, int H, int W, const float* W_host, const float* y,
                       long long y_n_stride, int y_cpv, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, int relu, int batch_mode, float* dz, long long dz_n_stride, float* sums,
                       float* dgamma, float* dbeta, float* workspace, tta_stream_t stream);
/* COMPACT layout of the full-resolution tail (y_cpv = 4): the <= 4-channel tensors of the head otherwise carry
 * five pad channels through every pass.  y is then fp32 [N][D][H][W][4] (written by tta_conv_tc with flags bit 14,
 * y_n_stride in floats), dz is FP16 [N][D][H][W][4] (dz_n_stride in 16-bit elements), and the norm-backward apply
 * of the head's norm layer is tta_norm_bwd_apply_c4 (dy leaves in the usual 8-channel operand layout). */
int tta_norm_bwd_apply_c4(const uint16_t* dz, long long dz_n_stride, const float* y, long long y_n_stride, int N,
                          int Creal, long long V, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, int batch_mode, const float* sums, uint16_t* dy_hi, uint16_t* dy_lo,
                          long long dy_n_stride, int out_dtype, int dy_wsplit_w, int dy_c4, tta_stream_t stream);

/* ---- fused head: logits + entropy loss + dlogits in one pass.  Replaces the loss + backward
 * entry of src/core/trainers/seg_trainer.py:141-142 with the TENT entropy (mode 0 softmax,
 * mode 1 Bernoulli/sigmoid -- the reference's heads are sigmoid, configs/_global_patches/
 * brats.yaml:12,59) and writes logits in the reference's NCDHW fp32 layout
 * (src/evaluation/seg_eval.py:300-302).  dz = sample_w[n]*inv_count*grad_scale*dH/dz. */
int tta_head_entropy_blocks(int N, long long V);
int tta_head_entropy(const float* y, long long y_n_stride, int N, int R, long long V, int mode, float inv_count,
                     float grad_scale, int dz_dtype, const float* sample_w, float* logits, uint16_t* dz_hi,
                     uint16_t* dz_lo, long long dz_n_stride, float* partial, float* loss, tta_stream_t stream);

/* ---- optimizer: torch.optim.Adam built by src/core/experiment_manager.py:199-237 on the flat
 * [gamma || beta] buffer; *step_dev is incremented on the device (graph replay safe). */
int tta_adam_step(float* p, const float* g, float* m, float* v, int n, float lr, float b1, float b2, float eps,
                  float gscale, int* step_dev, tta_stream_t stream);

/* ---- sliding-window Gaussian blending (MONAI sliding_window_inference(mode="gaussian")) */
int tta_sw_blend(const float* logits, int NB, int R, int D, int H, int W, const int* win, const float* sample_w,
                 const float* gd, const float* gh, const float* gw, float wmin, float* acc, float* wsum, int n_vol,
                 int Ds, int Hs, int Ws, tta_stream_t stream);
int tta_sw_normalise(const float* acc, const float* wsum, int n_vol, int R, long long Vs, float* out,
                     tta_stream_t stream);

/* ---- evaluation: sigmoid >= thr, label > 0.5, per-(b,r) {intersection, pred sum, gt sum}:
 * replaces src/evaluation/seg_eval.py:304-308 + the sums of _binary_dice_iou (:41-68). */
int tta_dice_counts(const float* logits, const float* label, int BR, long long V, float thr,
                    unsigned long long* counts, tta_stream_t stream);

/* ---- input intensity policy on the device: per-channel clip + (masked) z-score, replaces
 * `_normalize_img` of src/datasets/transforms.py:147-200 (configs/_global_patches/hecktor21.yaml:27-46).
 * vol [n_vol][C][V] fp32.  rules [C][8] = {clip_on, lo, hi, z_mode, mask_gt, eps, mean, std}, z_mode 0 none,
 * 1 masked z-score (falls back to all voxels when fewer than min_count pass the mask), 2 unmasked,
 * 3 fixed (x - mean) / std.  tta_intensity_stats writes affine [n_vol][C][4] = {lo, hi, mu, 1/sd};
 * tta_intensity_apply: out = (clamp(x, lo, hi) - mu) / sd (in place allowed).  workspace: zeroed once,
 * tta_intensity_workspace_bytes. */
long long tta_intensity_workspace_bytes(int n_vol, int C, long long V);
int tta_intensity_stats(const float* vol, int n_vol, int C, long long V, const float* rules, int min_count,
                        float* affine, void* workspace, tta_stream_t stream);
int tta_intensity_apply(const float* vol, float* out, int n_vol, int C, long long V, const float* affine,
                        tta_stream_t stream);

/* ---- multimodal model (MultimodalUNetDeepFusion, src/models/unet_multimodal_midfusion.py:139-267): modality mean
 * of M operand tensors (pseudo-shared bottleneck feature :216, fused skips :221-224; rep > 1 writes one copy per
 * modality for the batched fusion layer), its backward / gradient fan-in as a scaled sum of fp32 views, and
 * MONAI UpSample("nontrainable") = nn.Upsample(trilinear, align_corners=True) (:114-120) forward and adjoint.
 * Pointer arrays are HOST arrays of nsrc <= 8 device pointers. */
int tta_mean_planes(const uint16_t* const* in_hi, const uint16_t* const* in_lo, const long long* in_n_stride, int nsrc,
                    int N, int C8, long long V, float scale, uint16_t* out_hi, uint16_t* out_lo, long long out_n_stride,
                    int rep, tta_stream_t stream);
int tta_sum_f32(const float* const* src, const long long* src_n_stride, int nsrc, int rep, int N, int C8, long long V,
                float scale, float* out, long long out_n_stride, int accumulate, tta_stream_t stream);
int tta_upsample_fwd(const float* in, long long in_n_stride, int N, int C8, int Di, int Hi, int Wi, int Do, int Ho,
                     int Wo, uint16_t* out_hi, uint16_t* out_lo, long long out_n_stride, int out_dtype,
                     tta_stream_t stream);
int tta_upsample_bwd(const float* g, long long g_n_stride, int N, int C8, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                     uint16_t* dy_hi, uint16_t* dy_lo, long long dy_n_stride, int out_dtype, tta_stream_t stream);

/* ---- supervised step (loss.backward() of src/core/trainers/seg_trainer.py:142 for EVERY parameter):
 * weight / bias gradients of Conv3d and ConvTranspose3d from the operands the forward and backward passes hold.
 * x: forward operand planes (fp16 hi + lo) of the conv's input view; dy: the 16-bit gradient planes of its output
 * (dy_dtype TTA_F16_HI: one loss-scaled plane).  dw (+=) scale * dL/dW in the PARAMETER layout: layout 0 = Conv3d
 * [Cout][Cin][k^3] (couts >= co_split go to dw2: fused unit0 || shortcut convs), layout 1 = ConvTranspose3d
 * [Cin][Cout][k^3]; fp32 atomics, the caller zeroes dw / db once per step. */
int tta_conv_wgrad(const uint16_t* x_hi, const uint16_t* x_lo, long long x_n_stride, int Dx, int Hx, int Wx, int x_wsplit,
                   const uint16_t* dy_hi, const uint16_t* dy_lo, long long dy_n_stride, int dy_dtype, int Dy, int Hy,
                   int Wy, int dy_wsplit, int N, int mode, int K, int stride, int Cin, int Cout, float scale, float* dw,
                   int layout, int co_split, float* dw2, tta_stream_t stream);
/* The same weight gradient on the tensor cores (tcgen05, csrc/tta_wgrad_tc.cu) for the 3x3x3 layers whose output
 * gradient is ONE loss-scaled fp16 plane (dy_dtype TTA_F16_HI): conv stride 1 / 2 and transposed conv stride 2.
 * Operand layouts: stride 1 -- x and dy plain; stride-2 conv -- x w-parity-split (x_wsplit = 1: the copy its forward
 * conv reads), dy plain; transposed -- x plain, dy w-parity-split (dy_wsplit = 1: what its dgrad reads).
 * flags bit 0: use only the hi plane of x (one fp16 product instead of two). */
int tta_conv_wgrad_tc_supported(int mode, int K, int stride, int Cin, int Cout, int dy_dtype);
int tta_conv_wgrad_tc(const uint16_t* x_hi, const uint16_t* x_lo, long long x_n_stride, int Dx, int Hx, int Wx,
                      int x_wsplit, const uint16_t* dy_hi, long long dy_n_stride, int Dy, int Hy, int Wy, int dy_wsplit,
                      int N, int mode, int stride, int Cin, int Cout, float scale, float* dw, int layout, int co_split,
                      float* dw2, int flags, tta_stream_t stream);
/* dL/dlogits (NCDHW fp32, from the reference's own loss through autograd) -> the 16-bit gradient plane(s) of the
 * last conv's result, times the power-of-two loss scale */
int tta_pack_grad(const float* g, int N, int R, long long V, float scale, uint16_t* dy_hi, uint16_t* dy_lo,
                  long long dy_n_stride, int out_dtype, tta_stream_t stream);
/* packed operand blob <- flat parameter buffer after an optimizer step: map[e] = +-(1 + source index) (sign = hi / lo
 * plane of the 16-bit formats, 0 = padding), add[e] != 0 adds the folded identity shortcut; kind 0 fp32, 1 fp16, 2 bf16 */
int tta_repack_weights(const float* src, const int* map, const signed char* add, long long n, void* out, int kind,
                       tta_stream_t stream);
int tta_bias_grad(const uint16_t* dy_hi, const uint16_t* dy_lo, long long dy_n_stride, int dy_dtype, int N, int Cout,
                  long long V, float scale, float* db, int co_split, float* db2, tta_stream_t stream);

/* ---- the step as a C object (plan_create / tta_step / workspace_bytes).  Build the launch list once
 * per input shape by calling the ordinary tta_* entry points between tta_plan_begin(plan, section) and tta_plan_end():
 * while a recording is active on the calling thread those calls store themselves (arguments by value; host pointer
 * arguments must outlive the plan) instead of launching.  tta_plan_run enqueues a section on `stream`; tta_step =
 * sections 0 (forward), 1 (training head: logits, entropy loss, dlogits) and 3 (backward).  Section 2 is the
 * inference head; 4..7 are free for the caller (e.g. the input gather of a fixed staging buffer, Adam). */
typedef struct tta_plan tta_plan;
int tta_plan_create(tta_plan** out);
int tta_plan_destroy(tta_plan* plan);
int tta_plan_begin(tta_plan* plan, int section);
int tta_plan_end(void);
int tta_plan_num_launches(const tta_plan* plan, int section);
int tta_plan_run(const tta_plan* plan, int section, tta_stream_t stream);
int tta_step(const tta_plan* plan, tta_stream_t stream);
long long tta_workspace_bytes(int N, int C8, long long V);

#ifdef __cplusplus
}
#endif
#endif /* TTA_B200_H */
