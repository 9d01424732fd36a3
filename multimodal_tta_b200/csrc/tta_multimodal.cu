// Streaming kernels of the multimodal model (MultimodalUNetDeepFusion,
// /root/reference/src/models/unet_multimodal_midfusion.py:139-267) on top of the UNet kernel vocabulary:
//
//   * tta_mean_planes   : modality mean of M operand tensors (bottleneck "pseudo-shared" feature :216, fused skips
//                         :221-224), optionally replicated `rep` times along the batch (one copy per modality for
//                         the batched fusion layer)
//   * tta_sum_f32       : scale * sum of K fp32 gradient views (the mean's backward; fan-in of replicated copies)
//   * tta_upsample_fwd  : nn.Upsample(scale_factor, mode="trilinear", align_corners=True) (MONAI UpSample
//                         "nontrainable", :114-120) from the fp32 result of the 1x1x1 pre-conv to operand planes
//   * tta_upsample_bwd  : its adjoint in gather form (deterministic, no atomics) from the fp32 gradient to the
//                         16-bit gradient plane(s) the pre-conv's dgrad reads
//
// All HBM-bound, one thread per voxel-chunk (8 channels), 256-bit accesses on fp32 / 128-bit on 16-bit planes.
// Index / weight arithmetic follows ATen's upsample_trilinear3d (area_pixel_compute_scale + source index in fp32):
//   scale = (in - 1) / (out - 1) (0 when out == 1), src = scale * o, i0 = (int)src, l1 = src - i0, l0 = 1 - l1,
//   i1 = i0 + (i0 < in - 1).
#include "tta_common.cuh"

namespace tta {

constexpr int kMmThreads = 256;
constexpr int kMaxSrc = 8;

struct PlaneSrcs {
  const uint16_t* hi[kMaxSrc];
  const uint16_t* lo[kMaxSrc];
  long long ns[kMaxSrc];
  int n;
};

struct F32Srcs {
  const float* p[kMaxSrc];
  long long ns[kMaxSrc];
  int n;
};

__global__ void __launch_bounds__(kMmThreads, 4)
mean_planes_kernel(const PlaneSrcs S, long long V, float scale, uint16_t* out_hi, uint16_t* out_lo, long long out_ns,
                   int rep) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, n = blockIdx.z;
  const long long slab = (long long)chunk * V * 8;
  for (long long v = (long long)blockIdx.x * kMmThreads + threadIdx.x; v < V; v += (long long)gridDim.x * kMmThreads) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k = 0; k < S.n; ++k) {   // fixed order: deterministic
      float x[8];
      load_split8<TTA_F16>(S.hi[k], S.lo[k], (long long)n * S.ns[k] + slab + v * 8, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += x[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= scale;
    for (int r = 0; r < rep; ++r)
      store_split8<TTA_F16>(out_hi, out_lo, ((long long)n * rep + r) * out_ns + slab + v * 8, acc);
  }
}

__global__ void __launch_bounds__(kMmThreads, 4)
sum_f32_kernel(const F32Srcs S, long long V, float scale, int rep, float* out, long long out_ns, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, n = blockIdx.z;
  const long long slab = (long long)chunk * V * 8;
  for (long long v = (long long)blockIdx.x * kMmThreads + threadIdx.x; v < V; v += (long long)gridDim.x * kMmThreads) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k = 0; k < S.n; ++k)
      for (int r = 0; r < rep; ++r) {
        float x[8];
        load_f32x8(S.p[k] + ((long long)n * rep + r) * S.ns[k] + slab + v * 8, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += x[i];
      }
    float* dst = out + (long long)n * out_ns + slab + v * 8;
    if (accumulate) {
      float o[8];
      load_f32x8(dst, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], scale, o[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= scale;
    }
    store_f32x8(dst, acc);
  }
}

struct UpGeom {
  int Di, Hi, Wi, Do, Ho, Wo;
  float sd, sh, sw;   // (in - 1) / (out - 1)
};

__device__ __forceinline__ void src_index(float scale, int o, int in, int& i0, int& i1, float& l0, float& l1) {
  const float s = scale * (float)o;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  l1 = s - (float)i0;
  l0 = 1.f - l1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
}

template <int ODT>
__global__ void __launch_bounds__(kMmThreads, 4)
upsample_fwd_kernel(const float* in, long long in_ns, const UpGeom G, uint16_t* out_hi, uint16_t* out_lo,
                    long long out_ns) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, n = blockIdx.z;
  const long long Vi = (long long)G.Di * G.Hi * G.Wi, Vo = (long long)G.Do * G.Ho * G.Wo;
  const float* ib = in + (long long)n * in_ns + (long long)chunk * Vi * 8;
  const long long ob = (long long)n * out_ns + (long long)chunk * Vo * 8;
  for (long long v = (long long)blockIdx.x * kMmThreads + threadIdx.x; v < Vo; v += (long long)gridDim.x * kMmThreads) {
    const int w = (int)(v % G.Wo), h = (int)((v / G.Wo) % G.Ho), d = (int)(v / ((long long)G.Wo * G.Ho));
    int d0, d1, h0, h1, w0, w1;
    float ld0, ld1, lh0, lh1, lw0, lw1;
    src_index(G.sd, d, G.Di, d0, d1, ld0, ld1);
    src_index(G.sh, h, G.Hi, h0, h1, lh0, lh1);
    src_index(G.sw, w, G.Wi, w0, w1, lw0, lw1);
    float a[8], b[8], acc[8];
    auto row = [&](int dd, int hh, float (&o)[8]) {   // lw0 * in[dd][hh][w0] + lw1 * in[dd][hh][w1]
      const float* p = ib + ((long long)dd * G.Hi + hh) * G.Wi * 8;
      load_f32x8(p + (long long)w0 * 8, a);
      load_f32x8(p + (long long)w1 * 8, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = lw0 * a[i] + lw1 * b[i];
    };
    float r00[8], r01[8], r10[8], r11[8];
    row(d0, h0, r00);
    row(d0, h1, r01);
    row(d1, h0, r10);
    row(d1, h1, r11);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      acc[i] = ld0 * (lh0 * r00[i] + lh1 * r01[i]) + ld1 * (lh0 * r10[i] + lh1 * r11[i]);
    store_split8<ODT>(out_hi, out_lo, ob + v * 8, acc);
  }
}

// weight with which output position o contributes to input position i along one axis (adjoint of src_index)
__device__ __forceinline__ float axis_weight(float scale, int o, int in, int i) {
  int i0, i1;
  float l0, l1;
  src_index(scale, o, in, i0, i1, l0, l1);
  return (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
}

// outputs o that can touch input i: scale * o in (i - 1, i + 1)
__device__ __forceinline__ void out_range(float scale, int i, int out, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
  lo = (int)floorf((float)(i - 1) / scale) - 1;   // one position of slack either side: fp32 rounding of the bounds;
  hi = (int)ceilf((float)(i + 1) / scale) + 1;    // axis_weight() is exactly 0 for outputs that do not touch i
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
}

template <int ODT>
__global__ void __launch_bounds__(kMmThreads, 3)
upsample_bwd_kernel(const float* g, long long g_ns, const UpGeom G, uint16_t* dy_hi, uint16_t* dy_lo, long long dy_ns) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, n = blockIdx.z;
  const long long Vi = (long long)G.Di * G.Hi * G.Wi, Vo = (long long)G.Do * G.Ho * G.Wo;
  const float* gb = g + (long long)n * g_ns + (long long)chunk * Vo * 8;
  const long long ob = (long long)n * dy_ns + (long long)chunk * Vi * 8;
  for (long long v = (long long)blockIdx.x * kMmThreads + threadIdx.x; v < Vi; v += (long long)gridDim.x * kMmThreads) {
    const int w = (int)(v % G.Wi), h = (int)((v / G.Wi) % G.Hi), d = (int)(v / ((long long)G.Wi * G.Hi));
    int dlo, dhi, hlo, hhi, wlo, whi;
    out_range(G.sd, d, G.Do, dlo, dhi);
    out_range(G.sh, h, G.Ho, hlo, hhi);
    out_range(G.sw, w, G.Wo, wlo, whi);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // per-axis weights once per voxel (the candidate ranges hold <= kUpMax outputs for scale factors >= 1.5; a longer
    // range falls back to evaluating the weight inside the loops): 3 x ~7 evaluations instead of ~7^3
    constexpr int kUpMax = 8;
    const bool tab = dhi - dlo < kUpMax && hhi - hlo < kUpMax && whi - wlo < kUpMax;
    float whv[kUpMax], wwv[kUpMax];
    if (tab) {
#pragma unroll
      for (int j = 0; j < kUpMax; ++j) {
        whv[j] = hlo + j <= hhi ? axis_weight(G.sh, hlo + j, G.Hi, h) : 0.f;
        wwv[j] = wlo + j <= whi ? axis_weight(G.sw, wlo + j, G.Wi, w) : 0.f;
      }
    }
    if (tab) {
#pragma unroll 1
      for (int jd = 0; jd <= dhi - dlo; ++jd) {
        const float wd = axis_weight(G.sd, dlo + jd, G.Di, d);
        if (wd == 0.f) continue;
#pragma unroll
        for (int jh = 0; jh < kUpMax; ++jh) {
          const float wh = whv[jh] * wd;
          if (wh == 0.f) continue;
          const float* p = gb + (((long long)(dlo + jd) * G.Ho + (hlo + jh)) * G.Wo + wlo) * 8;
#pragma unroll
          for (int jw = 0; jw < kUpMax; ++jw) {
            const float ww = wwv[jw] * wh;
            if (ww == 0.f) continue;
            float x[8];
            load_f32x8(p + (long long)jw * 8, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(ww, x[i], acc[i]);
          }
        }
      }
    } else {
      for (int od = dlo; od <= dhi; ++od) {
        const float wd = axis_weight(G.sd, od, G.Di, d);
        if (wd == 0.f) continue;
        for (int oh = hlo; oh <= hhi; ++oh) {
          const float wh = axis_weight(G.sh, oh, G.Hi, h) * wd;
          if (wh == 0.f) continue;
          const float* p = gb + ((long long)od * G.Ho + oh) * G.Wo * 8;
          for (int ow = wlo; ow <= whi; ++ow) {
            const float ww = axis_weight(G.sw, ow, G.Wi, w) * wh;
            if (ww == 0.f) continue;
            float x[8];
            load_f32x8(p + (long long)ow * 8, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(ww, x[i], acc[i]);
          }
        }
      }
    }
    store_split8<ODT>(dy_hi, dy_lo, ob + v * 8, acc);
  }
}

static inline int mm_xblocks(long long V, long long others) {
  long long full = (V + kMmThreads - 1) / kMmThreads;
  long long want = (8LL * 148 + others - 1) / (others > 0 ? others : 1);
  if (want > full) want = full;
  if (want < 1) want = 1;
  return (int)want;
}

static UpGeom make_up(int Di, int Hi, int Wi, int Do, int Ho, int Wo) {
  UpGeom G;
  G.Di = Di; G.Hi = Hi; G.Wi = Wi; G.Do = Do; G.Ho = Ho; G.Wo = Wo;
  G.sd = Do > 1 ? (float)(Di - 1) / (float)(Do - 1) : 0.f;
  G.sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  G.sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  return G;
}

}  // namespace tta

using namespace tta;

extern "C" {

// out[(n*rep + r)] = scale * sum_k in_k[n]   for r < rep.  in_*: HOST arrays of nsrc (<= 8) device pointers / n-strides
// (fp16 hi/lo operand planes, chunk layout [N][C8][V][8], strides in 16-bit elements).
int tta_mean_planes(const uint16_t* const* in_hi, const uint16_t* const* in_lo, const long long* in_ns, int nsrc, int N,
                    int C8, long long V, float scale, uint16_t* out_hi, uint16_t* out_lo, long long out_ns, int rep,
                    cudaStream_t stream) {
  TTA_RECORDABLE(tta_mean_planes(in_hi, in_lo, in_ns, nsrc, N, C8, V, scale, out_hi, out_lo, out_ns, rep, s_));
  TTA_REQUIRE(in_hi && in_lo && in_ns && out_hi && out_lo, "tta_mean_planes: null pointer");
  TTA_REQUIRE(nsrc >= 1 && nsrc <= kMaxSrc, "tta_mean_planes: %d sources (1..8)", nsrc);
  TTA_REQUIRE(N > 0 && C8 > 0 && V > 0 && rep >= 1, "tta_mean_planes: bad shape");
  PlaneSrcs S;
  S.n = nsrc;
  for (int k = 0; k < nsrc; ++k) {
    TTA_REQUIRE(in_hi[k] && in_lo[k], "tta_mean_planes: null source %d", k);
    S.hi[k] = in_hi[k]; S.lo[k] = in_lo[k]; S.ns[k] = in_ns[k];
  }
  tta_launch(mean_planes_kernel, dim3(mm_xblocks(V, (long long)N * C8), C8, N), kMmThreads, 0, stream, tta_pdl_family(2), S,
             V, scale, out_hi, out_lo, out_ns, rep);
  return tta_check_launch("tta_mean_planes");
}

// out[n] (=|+=) scale * sum_k sum_{r<rep} src_k[n*rep + r]   (fp32 chunk layout; src: HOST arrays of <= 8 entries)
int tta_sum_f32(const float* const* src, const long long* src_ns, int nsrc, int rep, int N, int C8, long long V,
                float scale, float* out, long long out_ns, int accumulate, cudaStream_t stream) {
  TTA_RECORDABLE(tta_sum_f32(src, src_ns, nsrc, rep, N, C8, V, scale, out, out_ns, accumulate, s_));
  TTA_REQUIRE(src && src_ns && out, "tta_sum_f32: null pointer");
  TTA_REQUIRE(nsrc >= 1 && nsrc <= kMaxSrc, "tta_sum_f32: %d sources (1..8)", nsrc);
  TTA_REQUIRE(N > 0 && C8 > 0 && V > 0 && rep >= 1, "tta_sum_f32: bad shape");
  F32Srcs S;
  S.n = nsrc;
  for (int k = 0; k < nsrc; ++k) {
    TTA_REQUIRE(src[k], "tta_sum_f32: null source %d", k);
    S.p[k] = src[k]; S.ns[k] = src_ns[k];
  }
  tta_launch(sum_f32_kernel, dim3(mm_xblocks(V, (long long)N * C8), C8, N), kMmThreads, 0, stream, tta_pdl_family(2), S, V,
             scale, rep, out, out_ns, accumulate);
  return tta_check_launch("tta_sum_f32");
}

int tta_upsample_fwd(const float* in, long long in_ns, int N, int C8, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                     uint16_t* out_hi, uint16_t* out_lo, long long out_ns, int out_dtype, cudaStream_t stream) {
  TTA_RECORDABLE(tta_upsample_fwd(in, in_ns, N, C8, Di, Hi, Wi, Do, Ho, Wo, out_hi, out_lo, out_ns, out_dtype, s_));
  TTA_REQUIRE(in && out_hi && out_lo, "tta_upsample_fwd: null pointer");
  TTA_REQUIRE(out_dtype == TTA_F16 || out_dtype == TTA_BF16, "tta_upsample_fwd: bad dtype");
  TTA_REQUIRE(N > 0 && C8 > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do >= Di && Ho >= Hi && Wo >= Wi, "tta_upsample_fwd: bad shape");
  const UpGeom G = make_up(Di, Hi, Wi, Do, Ho, Wo);
  const long long Vo = (long long)Do * Ho * Wo;
  const dim3 grid(mm_xblocks(Vo, (long long)N * C8), C8, N);
  if (out_dtype == TTA_F16)
    tta_launch(upsample_fwd_kernel<TTA_F16>, grid, kMmThreads, 0, stream, tta_pdl_family(2), in, in_ns, G, out_hi, out_lo, out_ns);
  else
    tta_launch(upsample_fwd_kernel<TTA_BF16>, grid, kMmThreads, 0, stream, tta_pdl_family(2), in, in_ns, G, out_hi, out_lo, out_ns);
  return tta_check_launch("tta_upsample_fwd");
}

// g: fp32 gradient w.r.t. the upsampled tensor (view [N][C8][Do][Ho][Wo][8]); dy: gradient w.r.t. its input as
// 16-bit plane(s) (dy_lo unused for TTA_F16_HI)
int tta_upsample_bwd(const float* g, long long g_ns, int N, int C8, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                     uint16_t* dy_hi, uint16_t* dy_lo, long long dy_ns, int out_dtype, cudaStream_t stream) {
  TTA_RECORDABLE(tta_upsample_bwd(g, g_ns, N, C8, Di, Hi, Wi, Do, Ho, Wo, dy_hi, dy_lo, dy_ns, out_dtype, s_));
  TTA_REQUIRE(g && dy_hi && (dy_lo || out_dtype == TTA_F16_HI), "tta_upsample_bwd: null pointer");
  TTA_REQUIRE(out_dtype >= 0 && out_dtype <= 2, "tta_upsample_bwd: bad dtype");
  TTA_REQUIRE(N > 0 && C8 > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do >= Di && Ho >= Hi && Wo >= Wi, "tta_upsample_bwd: bad shape");
  const UpGeom G = make_up(Di, Hi, Wi, Do, Ho, Wo);
  const long long Vi = (long long)Di * Hi * Wi;
  const dim3 grid(mm_xblocks(Vi, (long long)N * C8), C8, N);
  if (out_dtype == TTA_F16)
    tta_launch(upsample_bwd_kernel<TTA_F16>, grid, kMmThreads, 0, stream, tta_pdl_family(2), g, g_ns, G, dy_hi, dy_lo, dy_ns);
  else if (out_dtype == TTA_F16_HI)
    tta_launch(upsample_bwd_kernel<TTA_F16_HI>, grid, kMmThreads, 0, stream, tta_pdl_family(2), g, g_ns, G, dy_hi, dy_lo, dy_ns);
  else
    tta_launch(upsample_bwd_kernel<TTA_BF16>, grid, kMmThreads, 0, stream, tta_pdl_family(2), g, g_ns, G, dy_hi, dy_lo, dy_ns);
  return tta_check_launch("tta_upsample_bwd");
}

}  // extern "C"
