// Generic fp32 CUDA-core 3-D convolution / transposed convolution on the chunked layout.
//
// This is the exact-arithmetic companion of the tcgen05 kernel (tta_conv_tc.cu): it reads the
// same split 16-bit operand planes (x = hi + lo), multiplies in fp32 with fp32 weights, and
// writes the same fp32 result layout.  It serves (a) geometries the tensor-core kernel does not
// cover and (b) on-device cross-validation of the tensor-core kernel.  It is a product kernel
// (hand-written CUDA, no library call), not a fallback to another backend.
//
// Canonical gather semantics shared by both kernels (Wg[tap][ci][co], tap = (kd*K + kh)*K + kw):
//   mode 0 (conv)  : out[o][co] = sum_{k,ci} in[s*o - p + k][ci] * Wg[k][ci][co]
//   mode 1 (convT) : out[o][co] = sum_{k,ci,i : s*i - p + k = o} in[i][ci] * Wg[k][ci][co]
// with p = (K-1)/2.  nn.Conv3d forward  -> mode 0, Wg[k][ci][co] = w[co][ci][k]
//                    nn.ConvTranspose3d -> mode 1, Wg[k][ci][co] = w[ci][co][k]
//                    dgrad of Conv3d    -> mode 1, Wg[k][ci=cout][co=cin] = w[cout][cin][k]
//                    dgrad of ConvT3d   -> mode 0, Wg[k][ci=cout][co=cin] = w[cin][cout][k]
// Packed weight layout consumed here: Wp[tap][C8in][C8out][8 ci][8 co] fp32.
#include "tta_common.cuh"

namespace tta {

constexpr int kConvThreads = 128;
constexpr int kVox = 4;  // consecutive-w output voxels per thread

struct ConvGeom {
  int mode, K, stride;
  int C8in, Di, Hi, Wi;
  int C8out, Do, Ho, Wo;
  long long in_ns, out_ns;
};

template <int DT>
__global__ void __launch_bounds__(kConvThreads)
conv_simt_kernel(const uint16_t* in_hi, const uint16_t* in_lo,
                 const float* Wp, const float* bias,
                 float* out, ConvGeom g, int accumulate) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float wsm[];  // [taps][8 ci][8 co]
  const int co_chunk = blockIdx.y, n = blockIdx.z;
  const int K = g.K, taps = K * K * K, pad = (K - 1) / 2, s = g.stride;
  const int Wo4 = (g.Wo + kVox - 1) / kVox;
  const long long groups = (long long)g.Do * g.Ho * Wo4;
  const long long gi = (long long)blockIdx.x * kConvThreads + threadIdx.x;
  const bool active = gi < groups;
  const int w4 = active ? (int)(gi % Wo4) : 0;
  const int oh = active ? (int)((gi / Wo4) % g.Ho) : 0;
  const int od = active ? (int)(gi / ((long long)Wo4 * g.Ho)) : 0;
  const int ow0 = w4 * kVox;
  const long long Vi = (long long)g.Di * g.Hi * g.Wi;

  float acc[kVox][8];
#pragma unroll
  for (int j = 0; j < kVox; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;

  for (int cc = 0; cc < g.C8in; ++cc) {
    __syncthreads();
    for (int t = threadIdx.x; t < taps * 64; t += kConvThreads) {
      const int tap = t >> 6, e = t & 63;
      wsm[t] = Wp[(((long long)tap * g.C8in + cc) * g.C8out + co_chunk) * 64 + e];
    }
    __syncthreads();
    if (!active) continue;
    const long long in_base = (long long)n * g.in_ns + (long long)cc * Vi * 8;
    for (int kd = 0; kd < K; ++kd) {
      int id;
      if (g.mode == 0) {
        id = s * od - pad + kd;
      } else {
        const int t = od + pad - kd;
        if (t < 0 || (t % s) != 0) continue;
        id = t / s;
      }
      if (id < 0 || id >= g.Di) continue;
      for (int kh = 0; kh < K; ++kh) {
        int ih;
        if (g.mode == 0) {
          ih = s * oh - pad + kh;
        } else {
          const int t = oh + pad - kh;
          if (t < 0 || (t % s) != 0) continue;
          ih = t / s;
        }
        if (ih < 0 || ih >= g.Hi) continue;
        const long long row = in_base + ((long long)id * g.Hi + ih) * g.Wi * 8;
        for (int kw = 0; kw < K; ++kw) {
          const float* wt = wsm + ((kd * K + kh) * K + kw) * 64;
#pragma unroll
          for (int j = 0; j < kVox; ++j) {
            const int ow = ow0 + j;
            int iw;
            bool ok = ow < g.Wo;
            if (g.mode == 0) {
              iw = s * ow - pad + kw;
            } else {
              const int t = ow + pad - kw;
              ok = ok && t >= 0 && (t % s) == 0;
              iw = t / s;
            }
            ok = ok && iw >= 0 && iw < g.Wi;
            if (!ok) continue;
            float x[8];
            load_split8<DT>(in_hi, in_lo, row + (long long)iw * 8, x);
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
              const float4 wa = *reinterpret_cast<const float4*>(wt + ci * 8);
              const float4 wb = *reinterpret_cast<const float4*>(wt + ci * 8 + 4);
              acc[j][0] = fmaf(x[ci], wa.x, acc[j][0]);
              acc[j][1] = fmaf(x[ci], wa.y, acc[j][1]);
              acc[j][2] = fmaf(x[ci], wa.z, acc[j][2]);
              acc[j][3] = fmaf(x[ci], wa.w, acc[j][3]);
              acc[j][4] = fmaf(x[ci], wb.x, acc[j][4]);
              acc[j][5] = fmaf(x[ci], wb.y, acc[j][5]);
              acc[j][6] = fmaf(x[ci], wb.z, acc[j][6]);
              acc[j][7] = fmaf(x[ci], wb.w, acc[j][7]);
            }
          }
        }
      }
    }
  }
  if (!active) return;
  float b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = bias ? bias[co_chunk * 8 + i] : 0.f;
  const long long Vo = (long long)g.Do * g.Ho * g.Wo;
  float* ob = out + (long long)n * g.out_ns + (long long)co_chunk * Vo * 8 +
              (((long long)od * g.Ho + oh) * g.Wo) * 8;
#pragma unroll
  for (int j = 0; j < kVox; ++j) {
    const int ow = ow0 + j;
    if (ow >= g.Wo) continue;
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = acc[j][i] + b[i];
    if (accumulate) {
      float o[8];
      load_f32x8(ob + (long long)ow * 8, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] += o[i];
    }
    store_f32x8(ob + (long long)ow * 8, r);
  }
}

}  // namespace tta

using namespace tta;

extern "C" int tta_conv_simt(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns,
                             int in_dtype, int N, int C8in, int Di, int Hi, int Wi,
                             const float* Wp, const float* bias, float* out, long long out_ns,
                             int C8out, int Do, int Ho, int Wo, int mode, int K, int stride,
                             int accumulate, cudaStream_t stream) {
  TTA_RECORDABLE(tta_conv_simt(in_hi, in_lo, in_ns, in_dtype, N, C8in, Di, Hi, Wi, Wp, bias, out, out_ns, C8out, Do, Ho, Wo, mode, K, stride, accumulate, s_));
  TTA_REQUIRE(in_hi && (in_lo || in_dtype == TTA_F16_HI) && Wp && out, "tta_conv_simt: null pointer");
  TTA_REQUIRE(mode == 0 || mode == 1, "tta_conv_simt: mode %d", mode);
  TTA_REQUIRE(K == 1 || K == 3, "tta_conv_simt: kernel size %d unsupported (1 or 3)", K);
  TTA_REQUIRE(stride == 1 || stride == 2, "tta_conv_simt: stride %d unsupported (1 or 2)", stride);
  TTA_REQUIRE(in_dtype >= 0 && in_dtype <= 2, "tta_conv_simt: bad dtype");
  const int pad = (K - 1) / 2;
  if (mode == 0) {
    TTA_REQUIRE(Do == (Di + 2 * pad - K) / stride + 1 && Ho == (Hi + 2 * pad - K) / stride + 1 &&
                    Wo == (Wi + 2 * pad - K) / stride + 1,
                "tta_conv_simt: conv output dims (%d,%d,%d) inconsistent with input (%d,%d,%d)", Do,
                Ho, Wo, Di, Hi, Wi);
  } else {
    TTA_REQUIRE(Do == Di * stride && Ho == Hi * stride && Wo == Wi * stride,
                "tta_conv_simt: convT output dims (%d,%d,%d) must be stride*input (%d,%d,%d)", Do,
                Ho, Wo, Di, Hi, Wi);
  }
  ConvGeom g{mode, K, stride, C8in, Di, Hi, Wi, C8out, Do, Ho, Wo, in_ns, out_ns};
  const int Wo4 = (Wo + kVox - 1) / kVox;
  const long long groups = (long long)Do * Ho * Wo4;
  const dim3 grid((unsigned)((groups + kConvThreads - 1) / kConvThreads), C8out, N);
  const size_t smem = (size_t)K * K * K * 64 * sizeof(float);
  if (in_dtype == TTA_F16)
    tta_launch(conv_simt_kernel<TTA_F16>, grid, kConvThreads, smem, stream, tta_pdl_family(1), in_hi, in_lo, Wp, bias, out, g, accumulate);
  else if (in_dtype == TTA_F16_HI)
    tta_launch(conv_simt_kernel<TTA_F16_HI>, grid, kConvThreads, smem, stream, tta_pdl_family(1), in_hi, in_lo, Wp, bias, out, g, accumulate);
  else
    tta_launch(conv_simt_kernel<TTA_BF16>, grid, kConvThreads, smem, stream, tta_pdl_family(1), in_hi, in_lo, Wp, bias, out, g, accumulate);
  return tta_check_launch("tta_conv_simt");
}
