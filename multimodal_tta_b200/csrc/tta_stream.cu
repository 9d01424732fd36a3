// HBM-streaming kernels of the TTA step that are not norms:
//   * window gather / input pack   (NCDHW fp32 volume -> split fp16 operand planes, zero pad,
//                                   optional per-(window,channel) modality scale)
//   * fused head: logits (NCDHW fp32) + per-voxel entropy loss + dlogits in ONE pass
//   * Adam on the flat [gamma || beta] buffer (single CTA, device-side step counter so the
//     whole step can be replayed from a CUDA graph)
//   * sliding-window Gaussian blend (deterministic gather form) + normalise
//   * sigmoid/threshold/Dice counts (reference: src/evaluation/seg_eval.py:41-68,304-308)
#include "tta_common.cuh"

namespace tta {

constexpr int kThreads = 256;

// Optional intensity policy folded into the gather (tta_intensity_stats: affine[n_vol][C][4] =
// {lo, hi, mu, 1/sd}): x -> (clamp(x, lo, hi) - mu) / sd, the same expression as tta_intensity_apply, so
// the fused and the two-pass paths give identical bits.  No table: identity.
__device__ __forceinline__ void load_affine(const float* affine, int vi, int C, int c, float& lo, float& hi,
                                            float& mu, float& inv) {
  lo = -INFINITY; hi = INFINITY; mu = 0.f; inv = 1.f;
  if (affine != nullptr && c < C) {
    const float* A = affine + ((long long)vi * C + c) * 4;
    lo = A[0]; hi = A[1]; mu = A[2]; inv = A[3];
  }
}

// ---------------------------------------------------------------- gather / pack
// win[b*4 + {0,1,2,3}] = {volume index, d0, h0, w0} of window b (origins may be negative or
// run past the volume: those voxels read 0 = MONAI's constant pad).
// first four of eight values -> (hi, lo) fp16, 8 bytes per plane (compact <= 4-channel layout)
__device__ __forceinline__ void store_split4_f16(uint16_t* hi, uint16_t* lo, long long off, const float (&x)[8]) {
  uint16_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = f32_to_u16<TTA_F16>(x[i]);
    l[i] = f32_to_u16<TTA_F16>(x[i] - u16_to_f32<TTA_F16>(h[i]));
  }
  *reinterpret_cast<uint2*>(hi + off) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
  *reinterpret_cast<uint2*>(lo + off) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

__device__ __forceinline__ float src_val(const float* p) { return *p; }
__device__ __forceinline__ float src_val(const __half* p) { return __half2float(*p); }

template <typename T>
__global__ void __launch_bounds__(kThreads)
gather_pack_kernel(const T* vol, int C, int Ds, int Hs, int Ws,
                   const int* win, const float* chan_scale, const float* affine, int D, int H,
                   int W, int C8, uint16_t* hi, uint16_t* lo,
                   long long o_ns, int wsplit) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int vi = win[b * 4 + 0], d0 = win[b * 4 + 1], h0 = win[b * 4 + 2], w0 = win[b * 4 + 3];
  const long long V = (long long)D * H * W;
  const long long Vs = (long long)Ds * Hs * Ws;
  const T* src = vol + (long long)vi * C * Vs;
  float sc[8], alo[8], ahi[8], amu[8], ainv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = chunk * 8 + i;
    sc[i] = (c < C) ? (chan_scale ? chan_scale[b * C + c] : 1.f) : 0.f;
    load_affine(affine, vi, C, c, alo[i], ahi[i], amu[i], ainv[i]);
  }
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V;
       v += (long long)gridDim.x * kThreads) {
    const int w = (int)(v % W);
    const int h = (int)((v / W) % H);
    const int d = (int)(v / ((long long)W * H));
    const int sd = d + d0, sh = h + h0, sw = w + w0;
    float x[8];
    const bool inside = sd >= 0 && sd < Ds && sh >= 0 && sh < Hs && sw >= 0 && sw < Ws;
    const long long so = ((long long)sd * Hs + sh) * Ws + sw;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = chunk * 8 + i;
      x[i] = (inside && c < C) ? (fminf(fmaxf(src_val(src + (long long)c * Vs + so), alo[i]), ahi[i]) - amu[i]) * ainv[i] * sc[i] : 0.f;
    }
    if (wsplit == 2) {   // compact layout [N][D][H][W][4] (<= 4 channels): 8 B per voxel and plane
      store_split4_f16(hi, lo, (long long)b * o_ns + v * 4, x);
      continue;
    }
    // wsplit: the first conv is a stride-2 tcgen05 conv -> w-parity-split rows (tta_common.cuh)
    const long long vo = wsplit ? v - w + (w & 1) * (W >> 1) + (w >> 1) : v;
    store_split8<TTA_F16>(hi, lo, (long long)b * o_ns + ((long long)chunk * V + vo) * 8, x);
  }
}

// Same, four consecutive-w voxels per thread (W % 4 == 0): one 128-bit load per real channel instead of
// four scalar ones, 32-bit index arithmetic with two divisions per FOUR voxels.  Measured in-stream on
// the 2 x 4 x 128^3 input: 87 us for the per-voxel kernel above (200 MB moved).
template <typename T>
__global__ void __launch_bounds__(kThreads)
gather_pack4_kernel(const T* vol, int C, int Ds, int Hs, int Ws,
                    const int* win, const float* chan_scale, const float* affine, int D, int H,
                    int W, int C8, uint16_t* hi, uint16_t* lo,
                    long long o_ns, int wsplit) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int vi = win[b * 4 + 0], d0 = win[b * 4 + 1], h0 = win[b * 4 + 2], w0 = win[b * 4 + 3];
  const long long V = (long long)D * H * W;
  const long long Vs = (long long)Ds * Hs * Ws;
  const T* src = vol + (long long)vi * C * Vs;
  float sc[8], alo[8], ahi[8], amu[8], ainv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = chunk * 8 + i;
    sc[i] = (c < C) ? (chan_scale ? chan_scale[b * C + c] : 1.f) : 0.f;
    load_affine(affine, vi, C, c, alo[i], ahi[i], amu[i], ainv[i]);
  }
  const unsigned W4 = (unsigned)W >> 2, G = (unsigned)(V >> 2);
  for (unsigned g = blockIdx.x * kThreads + threadIdx.x; g < G; g += gridDim.x * kThreads) {
    const unsigned row = g / W4, w4 = g - row * W4;
    const unsigned d = row / (unsigned)H, h = row - d * (unsigned)H;
    const int sd = (int)d + d0, sh = (int)h + h0, sw = (int)(w4 * 4) + w0;
    float x[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[k][i] = 0.f;
    if (sd >= 0 && sd < Ds && sh >= 0 && sh < Hs && sw > -4 && sw < Ws) {
      const T* rowp = src + ((long long)sd * Hs + sh) * Ws + sw;
      const bool whole = sw >= 0 && sw + 3 < Ws;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = chunk * 8 + i;
        if (c < C) {
          const T* p = rowp + (long long)c * Vs;
          if (whole && (reinterpret_cast<unsigned long long>(p) & (4ull * sizeof(T) - 1ull)) == 0ull) {
            float vv[4];
            if (sizeof(T) == 4) {
              const float4 v = *reinterpret_cast<const float4*>(p);
              vv[0] = v.x; vv[1] = v.y; vv[2] = v.z; vv[3] = v.w;
            } else {   // fp16 staging: four values = one 64-bit load
              const uint2 v = *reinterpret_cast<const uint2*>(p);
              const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
              const float2 bq = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
              vv[0] = a.x; vv[1] = a.y; vv[2] = bq.x; vv[3] = bq.y;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k][i] = (fminf(fmaxf(vv[k], alo[i]), ahi[i]) - amu[i]) * ainv[i] * sc[i];
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (sw + k >= 0 && sw + k < Ws) x[k][i] = (fminf(fmaxf(src_val(p + k), alo[i]), ahi[i]) - amu[i]) * ainv[i] * sc[i];
          }
        }
      }
    }
    if (wsplit == 2) {   // compact layout: four voxels x 4 channels = 32 contiguous bytes per plane
      const long long o4 = (long long)b * o_ns + ((long long)row * W + (long long)w4 * 4) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) store_split4_f16(hi, lo, o4 + k * 4, x[k]);
      continue;
    }
    const long long rowbase = (long long)b * o_ns + ((long long)chunk * V + (long long)row * W) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int w = (int)(w4 * 4) + k;
      // wsplit: the first conv is a stride-2 tcgen05 conv -> w-parity-split rows (tta_common.cuh)
      const int wo = wsplit ? (w & 1) * (W >> 1) + (w >> 1) : w;
      store_split8<TTA_F16>(hi, lo, rowbase + (long long)wo * 8, x[k]);
    }
  }
}

// Compact output ([N][D][H][W][4], <= 4 channels), W % 4 == 0: one thread = four consecutive-w voxels of all (<= 4)
// channels: one 128-bit (fp32) / 64-bit (fp16) load per channel, two 32-byte stores; 16 live values instead of the
// 32 of the 8-channel kernel above -> more resident warps on a pass that is pure latency x bandwidth.
template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
gather_pack_c4_kernel(const T* vol, int C, int Ds, int Hs, int Ws, const int* win, const float* chan_scale,
                      const float* affine, int D, int H, int W, uint16_t* hi, uint16_t* lo, long long o_ns) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int vi = win[b * 4 + 0], d0 = win[b * 4 + 1], h0 = win[b * 4 + 2], w0 = win[b * 4 + 3];
  const long long V = (long long)D * H * W;
  const long long Vs = (long long)Ds * Hs * Ws;
  const T* src = vol + (long long)vi * C * Vs;
  float sc[4], alo[4], ahi[4], amu[4], ainv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sc[i] = (i < C) ? (chan_scale ? chan_scale[b * C + i] : 1.f) : 0.f;
    load_affine(affine, vi, C, i, alo[i], ahi[i], amu[i], ainv[i]);
  }
  const unsigned W4 = (unsigned)W >> 2, G = (unsigned)(V >> 2);
  for (unsigned g = blockIdx.x * kThreads + threadIdx.x; g < G; g += gridDim.x * kThreads) {
    const unsigned row = g / W4, w4 = g - row * W4;
    const unsigned d = row / (unsigned)H, h = row - d * (unsigned)H;
    const int sd = (int)d + d0, sh = (int)h + h0, sw = (int)(w4 * 4) + w0;
    float x[4][4];   // [voxel][channel]
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) x[k][i] = 0.f;
    if (sd >= 0 && sd < Ds && sh >= 0 && sh < Hs && sw > -4 && sw < Ws) {
      const T* rowp = src + ((long long)sd * Hs + sh) * Ws + sw;
      const bool whole = sw >= 0 && sw + 3 < Ws;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < C) {
          const T* p = rowp + (long long)i * Vs;
          float vv[4] = {0.f, 0.f, 0.f, 0.f};
          if (whole && (reinterpret_cast<unsigned long long>(p) & (4ull * sizeof(T) - 1ull)) == 0ull) {
            if (sizeof(T) == 4) {
              const float4 v = *reinterpret_cast<const float4*>(p);
              vv[0] = v.x; vv[1] = v.y; vv[2] = v.z; vv[3] = v.w;
            } else {
              const uint2 v = *reinterpret_cast<const uint2*>(p);
              const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
              const float2 bq = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
              vv[0] = a.x; vv[1] = a.y; vv[2] = bq.x; vv[3] = bq.y;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k][i] = (fminf(fmaxf(vv[k], alo[i]), ahi[i]) - amu[i]) * ainv[i] * sc[i];
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (sw + k >= 0 && sw + k < Ws) x[k][i] = (fminf(fmaxf(src_val(p + k), alo[i]), ahi[i]) - amu[i]) * ainv[i] * sc[i];
          }
        }
      }
    }
    uint32_t hw[8], lw[8];   // 4 voxels x 4 channels x 16 bit = 8 words per plane
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        const uint16_t h0_ = f32_to_u16<TTA_F16>(x[k][i]), h1_ = f32_to_u16<TTA_F16>(x[k][i + 1]);
        const uint16_t l0_ = f32_to_u16<TTA_F16>(x[k][i] - u16_to_f32<TTA_F16>(h0_));
        const uint16_t l1_ = f32_to_u16<TTA_F16>(x[k][i + 1] - u16_to_f32<TTA_F16>(h1_));
        hw[k * 2 + i / 2] = (uint32_t)h0_ | ((uint32_t)h1_ << 16);
        lw[k * 2 + i / 2] = (uint32_t)l0_ | ((uint32_t)l1_ << 16);
      }
    const long long o4 = (long long)b * o_ns + ((long long)row * W + (long long)w4 * 4) * 4;
    uint4* ph = reinterpret_cast<uint4*>(hi + o4);
    uint4* pl = reinterpret_cast<uint4*>(lo + o4);
    ph[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); ph[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    pl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); pl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  }
}

// ---------------------------------------------------------------- fused head
// One pass: read the last conv's fp32 result (chunk 0, R <= 8 real channels), write logits in
// the reference's NCDHW fp32 layout, the per-voxel entropy (block partial sums) and
// dlogits = sample_w[n] * inv_count * dH/dz as split bf16 planes for the first dgrad conv.
//   mode 0: softmax entropy   H = lse(z) - sum_c p_c z_c ;  dH/dz_k = -p_k (z_k - sum_c p_c z_c)
//   mode 1: Bernoulli entropy H = sum_c softplus(z_c) - p_c z_c ; dH/dz_c = -z_c p_c (1 - p_c)
template <int ODT>
__global__ void __launch_bounds__(kThreads)
head_entropy_kernel(const float* y, long long y_ns, int R, long long V, int mode,
                    float inv_count, float grad_scale, const float* sample_w,
                    float* logits, uint16_t* dz_hi,
                    uint16_t* dz_lo, long long dz_ns, float* partial) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  const float sw = sample_w ? sample_w[n] : 1.f;
  const float gs = sw * inv_count * grad_scale;  // grad_scale: power-of-two loss scale (fp16 backward)
  float hsum = 0.f;
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V;
       v += (long long)gridDim.x * kThreads) {
    float z[8], g[8];
    load_f32x8(y + (long long)n * y_ns + v * 8, z);
    if (logits) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < R) logits[((long long)n * R + c) * V + v] = z[c];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = 0.f;
    const float Hv = entropy_point<8>(z, R, mode, gs, g);
    hsum += Hv * sw;
    if (dz_hi) store_split8<ODT>(dz_hi, dz_lo, (long long)n * dz_ns + v * 8, g);
  }
  __shared__ float red[kThreads / 32];
  hsum = warp_sum(hsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void loss_finalize_kernel(const float* partial, int count, float inv_count,
                                     float* loss) {
  pdl_trigger();
  pdl_wait();
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += 32) s += (double)partial[i];
  s = warp_sum_d(s);
  if (threadIdx.x == 0) *loss = (float)(s * (double)inv_count);
}

// ---------------------------------------------------------------- Adam (torch.optim.Adam math)
__global__ void __launch_bounds__(1024)
adam_kernel(float* p, const float* g, float* m,
            float* v, int n, float lr, float b1, float b2, float eps, float gscale,
            int* step_dev) {
  pdl_trigger();
  pdl_wait();
  const int t = *step_dev + 1;
  __syncthreads();
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  const float step_size = (float)((double)lr / bc1);
  const float sq_bc2 = (float)sqrt(bc2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);  // exp_avg.lerp_(grad, 1-beta1)
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sq_bc2 + eps;
    p[i] -= step_size * (mi / denom);
  }
  if (threadIdx.x == 0) *step_dev = t;
}

// ---------------------------------------------------------------- sliding-window blend
// Deterministic gather form: one thread per volume voxel loops over the windows of this batch
// in order and accumulates  acc += w * logit,  wsum += w  with
// w = max(gd[d]*gh[h]*gw[w], wmin)  (MONAI gaussian importance map, clamped).
__global__ void __launch_bounds__(kThreads)
sw_blend_kernel(const float* logits, int NB, int R, int D, int H, int W,
                const int* win, const float* sample_w,
                const float* gd, const float* gh,
                const float* gw, float wmin, float* acc,
                float* wsum, int Ds, int Hs, int Ws) {
  pdl_trigger();
  pdl_wait();
  const long long Vs = (long long)Ds * Hs * Ws;
  const long long V = (long long)D * H * W;
  const int vi = blockIdx.y;
  for (long long s = (long long)blockIdx.x * kThreads + threadIdx.x; s < Vs;
       s += (long long)gridDim.x * kThreads) {
    const int sw_ = (int)(s % Ws);
    const int sh = (int)((s / Ws) % Hs);
    const int sd = (int)(s / ((long long)Ws * Hs));
    float wacc = 0.f;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
    bool touched = false;
    for (int b = 0; b < NB; ++b) {
      if (win[b * 4 + 0] != vi) continue;
      if (sample_w && sample_w[b] == 0.f) continue;
      const int d = sd - win[b * 4 + 1], h = sh - win[b * 4 + 2], w = sw_ - win[b * 4 + 3];
      if (d < 0 || d >= D || h < 0 || h >= H || w < 0 || w >= W) continue;
      const float wt = fmaxf(gd[d] * gh[h] * gw[w], wmin);
      const long long o = ((long long)d * H + h) * W + w;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < R) a[c] = fmaf(wt, logits[((long long)b * R + c) * V + o], a[c]);
      wacc += wt;
      touched = true;
    }
    if (touched) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < R) acc[((long long)vi * R + c) * Vs + s] += a[c];
      wsum[(long long)vi * Vs + s] += wacc;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
sw_normalise_kernel(const float* acc, const float* wsum, int R,
                    long long Vs, float* out) {
  pdl_trigger();
  pdl_wait();
  const int vi = blockIdx.y;
  for (long long s = (long long)blockIdx.x * kThreads + threadIdx.x; s < Vs;
       s += (long long)gridDim.x * kThreads) {
    const float w = wsum[(long long)vi * Vs + s];
    for (int c = 0; c < R; ++c) {
      const long long o = ((long long)vi * R + c) * Vs + s;
      out[o] = acc[o] / w;
    }
  }
}

// ---------------------------------------------------------------- Dice counts
// counts[(b*R + r)*3 + {0,1,2}] += {sum pred*gt, sum pred, sum gt} with
// pred = sigmoid(z) >= thr, gt = label > 0.5 (integer atomics: order-independent, exact).
// Four voxels per 128-bit load, four loads of each tensor in flight per thread (the scalar version waited for one
// 4-byte load at a time: 78 us for 50 MB); the sigmoid keeps the reference's arithmetic (1 / (1 + exp(-z)) in fp32
// compared against thr), so the counts stay bit exact.
__device__ __forceinline__ void dice_acc(float z, float y, float thr, unsigned int& inter, unsigned int& ps,
                                         unsigned int& gs) {
  const float p = 1.f / (1.f + expf(-z));
  const unsigned int pr = p >= thr, gt = y > 0.5f;
  inter += pr & gt;
  ps += pr;
  gs += gt;
}

__global__ void __launch_bounds__(kThreads)
dice_counts_kernel(const float* logits, const float* label, long long V,
                   float thr, unsigned long long* counts) {
  pdl_trigger();
  pdl_wait();
  const int br = blockIdx.y;
  const float* z = logits + (long long)br * V;
  const float* y = label + (long long)br * V;
  unsigned int inter = 0, ps = 0, gs = 0;
  const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x, nthr = (long long)gridDim.x * kThreads;
  // rows start 16-byte aligned when V % 4 == 0 and the bases are (cudaMalloc / torch allocations are)
  const bool vec = (V & 3) == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(label)) & 15) == 0;
  long long done = 0;
  if (vec) {
    const long long V4 = V >> 2;
    const float4* z4 = reinterpret_cast<const float4*>(z);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    long long v = tid;
    for (; v + 3 * nthr < V4; v += 4 * nthr) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = z4[v + u * nthr];
        b[u] = y4[v + u * nthr];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dice_acc(a[u].x, b[u].x, thr, inter, ps, gs);
        dice_acc(a[u].y, b[u].y, thr, inter, ps, gs);
        dice_acc(a[u].z, b[u].z, thr, inter, ps, gs);
        dice_acc(a[u].w, b[u].w, thr, inter, ps, gs);
      }
    }
    for (; v < V4; v += nthr) {
      const float4 a = z4[v], b = y4[v];
      dice_acc(a.x, b.x, thr, inter, ps, gs);
      dice_acc(a.y, b.y, thr, inter, ps, gs);
      dice_acc(a.z, b.z, thr, inter, ps, gs);
      dice_acc(a.w, b.w, thr, inter, ps, gs);
    }
    done = V;
  }
  for (long long v = done + tid; v < V; v += nthr) dice_acc(z[v], y[v], thr, inter, ps, gs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    ps += __shfl_xor_sync(0xffffffffu, ps, o);
    gs += __shfl_xor_sync(0xffffffffu, gs, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counts[br * 3 + 0], (unsigned long long)inter);
    atomicAdd(&counts[br * 3 + 1], (unsigned long long)ps);
    atomicAdd(&counts[br * 3 + 2], (unsigned long long)gs);
  }
}

static inline int xblocks(long long V, long long rows) {
  long long full = (V + kThreads - 1) / kThreads;
  long long want = (8LL * 148 + rows - 1) / rows;
  if (want > full) want = full;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace tta

using namespace tta;

extern "C" {

int tta_gather_pack_norm(const float* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win,
                         const float* chan_scale, const float* affine, int NB, int D, int H, int W, uint16_t* hi,
                         uint16_t* lo, long long o_ns, int C8, int wsplit, cudaStream_t stream);


int tta_gather_pack(const float* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win,
                    const float* chan_scale, int NB, int D, int H, int W, uint16_t* hi,
                    uint16_t* lo, long long o_ns, int C8, int wsplit, cudaStream_t stream) {
  return tta_gather_pack_norm(vol, n_vol, C, Ds, Hs, Ws, win, chan_scale, nullptr, NB, D, H, W, hi, lo, o_ns, C8,
                              wsplit, stream);
}

}  // extern "C"

template <typename T>
static int gather_impl(const T* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win, const float* chan_scale,
                       const float* affine, int NB, int D, int H, int W, uint16_t* hi, uint16_t* lo, long long o_ns, int C8,
                       int wsplit, cudaStream_t stream) {
  TTA_REQUIRE(vol && win && hi && lo, "tta_gather_pack: null pointer");
  TTA_REQUIRE(wsplit != 1 || W % 2 == 0, "tta_gather_pack: w-parity-split output needs an even W (got %d)", W);
  TTA_REQUIRE(wsplit >= 0 && wsplit <= 2 && (wsplit != 2 || (C <= 4 && C8 == 1)),
              "tta_gather_pack: layout %d (0 plain, 1 w-parity-split, 2 compact: <= 4 channels, one chunk)", wsplit);
  TTA_REQUIRE(NB > 0 && C > 0 && C8 * 8 >= C && n_vol > 0, "tta_gather_pack: bad shape");
  const long long V = (long long)D * H * W;
  if (wsplit == 2 && W % 4 == 0 && V / 4 < 0x7fffffffLL) {
    tta_launch(gather_pack_c4_kernel<T>, dim3(xblocks(V / 4, NB), NB), kThreads, 0, stream, tta_pdl_family(4), vol, C, Ds,
               Hs, Ws, win, chan_scale, affine, D, H, W, hi, lo, o_ns);
    return tta_check_launch("tta_gather_pack");
  }
  if (W % 4 == 0 && V / 4 < 0x7fffffffLL)
    tta_launch(gather_pack4_kernel<T>, dim3(xblocks(V / 4, (long long)NB * C8), C8, NB), kThreads, 0, stream,
               tta_pdl_family(4), vol, C, Ds, Hs, Ws, win, chan_scale, affine, D, H, W, C8, hi, lo, o_ns, wsplit);
  else
    tta_launch(gather_pack_kernel<T>, dim3(xblocks(V, (long long)NB * C8), C8, NB), kThreads, 0, stream, tta_pdl_family(4),
               vol, C, Ds, Hs, Ws, win, chan_scale, affine, D, H, W, C8, hi, lo, o_ns, wsplit);
  return tta_check_launch("tta_gather_pack");
}

extern "C" {

// same with the intensity policy applied on the fly: affine [n_vol][C][4] from tta_intensity_stats (or null)
int tta_gather_pack_norm(const float* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win,
                         const float* chan_scale, const float* affine, int NB, int D, int H, int W, uint16_t* hi,
                         uint16_t* lo, long long o_ns, int C8, int wsplit, cudaStream_t stream) {
  TTA_RECORDABLE(tta_gather_pack_norm(vol, n_vol, C, Ds, Hs, Ws, win, chan_scale, affine, NB, D, H, W, hi, lo, o_ns, C8, wsplit, s_));
  return gather_impl<float>(vol, n_vol, C, Ds, Hs, Ws, win, chan_scale, affine, NB, D, H, W, hi, lo, o_ns, C8, wsplit, stream);
}

// FP16 staging: the volume arrives as IEEE half (half the host -> device bytes of the fp32 batch); everything
// downstream is unchanged (the lo operand plane is exactly zero unless an affine rescales the values)
int tta_gather_pack_norm_f16(const uint16_t* vol, int n_vol, int C, int Ds, int Hs, int Ws, const int* win,
                             const float* chan_scale, const float* affine, int NB, int D, int H, int W, uint16_t* hi,
                             uint16_t* lo, long long o_ns, int C8, int wsplit, cudaStream_t stream) {
  TTA_RECORDABLE(tta_gather_pack_norm_f16(vol, n_vol, C, Ds, Hs, Ws, win, chan_scale, affine, NB, D, H, W, hi, lo, o_ns, C8, wsplit, s_));
  return gather_impl<__half>(reinterpret_cast<const __half*>(vol), n_vol, C, Ds, Hs, Ws, win, chan_scale, affine, NB, D,
                             H, W, hi, lo, o_ns, C8, wsplit, stream);
}

int tta_head_entropy_blocks(int N, long long V) { return xblocks(V, N); }

int tta_head_entropy(const float* y, long long y_ns, int N, int R, long long V, int mode,
                     float inv_count, float grad_scale, int dz_dtype, const float* sample_w, float* logits,
                     uint16_t* dz_hi, uint16_t* dz_lo, long long dz_ns, float* partial, float* loss,
                     cudaStream_t stream) {
  TTA_RECORDABLE(tta_head_entropy(y, y_ns, N, R, V, mode, inv_count, grad_scale, dz_dtype, sample_w, logits, dz_hi, dz_lo, dz_ns, partial, loss, s_));
  TTA_REQUIRE(y && partial && loss, "tta_head_entropy: null pointer");
  TTA_REQUIRE(R >= 1 && R <= 8, "tta_head_entropy: R=%d unsupported (1..8 region channels)", R);
  TTA_REQUIRE(mode == 0 || mode == 1, "tta_head_entropy: mode %d", mode);
  TTA_REQUIRE(!(mode == 0 && R < 2), "tta_head_entropy: softmax entropy is degenerate for R=1");
  const int xb = xblocks(V, N);
  TTA_REQUIRE(dz_dtype == TTA_BF16 || dz_dtype == TTA_F16_HI, "tta_head_entropy: dz dtype %d", dz_dtype);
  if (dz_dtype == TTA_BF16)
    tta_launch(head_entropy_kernel<TTA_BF16>, dim3(xb, N), kThreads, 0, stream, tta_pdl_family(4), 
        y, y_ns, R, V, mode, inv_count, grad_scale, sample_w, logits, dz_hi, dz_lo, dz_ns, partial);
  else
    tta_launch(head_entropy_kernel<TTA_F16_HI>, dim3(xb, N), kThreads, 0, stream, tta_pdl_family(4), 
        y, y_ns, R, V, mode, inv_count, grad_scale, sample_w, logits, dz_hi, dz_lo, dz_ns, partial);
  tta_launch(loss_finalize_kernel, 1, 32, 0, stream, tta_pdl_family(4), partial, xb * N, inv_count, loss);
  return tta_check_launch("tta_head_entropy");
}

int tta_adam_step(float* p, const float* g, float* m, float* v, int n, float lr, float b1,
                  float b2, float eps, float gscale, int* step_dev, cudaStream_t stream) {
  TTA_RECORDABLE(tta_adam_step(p, g, m, v, n, lr, b1, b2, eps, gscale, step_dev, s_));
  TTA_REQUIRE(p && g && m && v && step_dev, "tta_adam_step: null pointer");
  TTA_REQUIRE(n >= 0, "tta_adam_step: n=%d", n);
  if (n == 0) return TTA_OK;
  tta_launch(adam_kernel, 1, 1024, 0, stream, tta_pdl_family(4), p, g, m, v, n, lr, b1, b2, eps, gscale, step_dev);
  return tta_check_launch("tta_adam_step");
}

int tta_sw_blend(const float* logits, int NB, int R, int D, int H, int W, const int* win,
                 const float* sample_w, const float* gd, const float* gh, const float* gw,
                 float wmin, float* acc, float* wsum, int n_vol, int Ds, int Hs, int Ws,
                 cudaStream_t stream) {
  TTA_RECORDABLE(tta_sw_blend(logits, NB, R, D, H, W, win, sample_w, gd, gh, gw, wmin, acc, wsum, n_vol, Ds, Hs, Ws, s_));
  TTA_REQUIRE(logits && win && gd && gh && gw && acc && wsum, "tta_sw_blend: null pointer");
  TTA_REQUIRE(R >= 1 && R <= 8, "tta_sw_blend: R=%d unsupported", R);
  const long long Vs = (long long)Ds * Hs * Ws;
  tta_launch(sw_blend_kernel, dim3(xblocks(Vs, n_vol), n_vol), kThreads, 0, stream, tta_pdl_family(4), 
      logits, NB, R, D, H, W, win, sample_w, gd, gh, gw, wmin, acc, wsum, Ds, Hs, Ws);
  return tta_check_launch("tta_sw_blend");
}

int tta_sw_normalise(const float* acc, const float* wsum, int n_vol, int R, long long Vs, float* out,
                     cudaStream_t stream) {
  TTA_RECORDABLE(tta_sw_normalise(acc, wsum, n_vol, R, Vs, out, s_));
  TTA_REQUIRE(acc && wsum && out, "tta_sw_normalise: null pointer");
  tta_launch(sw_normalise_kernel, dim3(xblocks(Vs, n_vol), n_vol), kThreads, 0, stream, tta_pdl_family(4), acc, wsum, R, Vs, out);
  return tta_check_launch("tta_sw_normalise");
}

int tta_dice_counts(const float* logits, const float* label, int BR, long long V, float thr,
                    unsigned long long* counts, cudaStream_t stream) {
  TTA_RECORDABLE(tta_dice_counts(logits, label, BR, V, thr, counts, s_));
  TTA_REQUIRE(logits && label && counts, "tta_dice_counts: null pointer");
  tta_launch(dice_counts_kernel, dim3(xblocks(V, BR), BR), kThreads, 0, stream, tta_pdl_family(4), logits, label, V, thr, counts);
  return tta_check_launch("tta_dice_counts");
}

}  // extern "C"
