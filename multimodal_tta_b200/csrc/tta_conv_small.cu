// Small-channel 3x3x3 stride-1 convolution on CUDA cores (cin <= 8 and cout <= 8).
//
// The UNet head (conv 3->3 at full resolution, and its input gradient) has almost no arithmetic:
// 243 MACs per voxel.  As a tensor-core GEMM it would pad 3 channels to a 16-wide k-step and a
// 16-wide n-tile and become bound by the 4 KB A-tile reads (measured 320 us); as a direct
// convolution it is an HBM/L1 streaming problem (see the kernel comment for the access pattern).
//   out[o][co] = bias[co] + sum_{k, ci} in[o + k - 1][ci] * W[k][ci][co]     (zero padding)
// The transposed (input-gradient) form is the same kernel with the taps flipped on the host.
#include "tta_common.cuh"

namespace tta {

constexpr int kSmallThreads = 256;

// One thread = one output voxel, a warp = 32 consecutive voxels of the flattened (d, h, w) index.
// Per (kd, kh) input row every lane loads ITS OWN voxel once (a warp reads 512 contiguous bytes
// per plane -> 4 L1 wavefronts) and gets the w-1 / w+1 neighbours from the adjacent lanes by
// shuffle; only lanes at a warp or row edge issue an extra load.
// The 27*CIN*COUT weights travel as a kernel PARAMETER: with the tap loops fully unrolled every
// weight is a compile-time offset into the constant bank, i.e. a direct FFMA operand (no load).
template <int NW>
struct SmallW {
  float w[NW];
};

template <int DT, int CIN, int COUT>
__global__ void __launch_bounds__(kSmallThreads)
conv_small_kernel(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns,
                  const __grid_constant__ SmallW<27 * CIN * COUT> Wc, const float* bias,
                  float* out, long long out_ns, int D, int H, int Wd, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  const long long V = (long long)D * H * Wd;
  const long long gi0 = (long long)blockIdx.x * kSmallThreads + threadIdx.x;
  const bool active = gi0 < V;
  const long long gi = active ? gi0 : V - 1;  // inactive lanes still take part in the shuffles
  const int w = (int)(gi % Wd);
  const int h = (int)((gi / Wd) % H);
  const int d = (int)(gi / ((long long)Wd * H));
  const int lane = threadIdx.x & 31;
  const uint16_t* ph = in_hi + (long long)n * in_ns;
  const uint16_t* pl = in_lo ? in_lo + (long long)n * in_ns : nullptr;
  // the lane below / above holds voxel w-1 / w+1 of the same row unless we sit on a warp or row edge
  const bool left_by_shfl = lane > 0 && w > 0, right_by_shfl = lane < 31 && w < Wd - 1;

  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;

#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int id = d + kd - 1;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = h + kh - 1;
      const bool rok = id >= 0 && id < D && ih >= 0 && ih < H;
      const long long row = ((long long)(rok ? id : 0) * H + (rok ? ih : 0)) * Wd * 8;
      float xc[CIN], xl[CIN], xr[CIN];
      {
        float v[8];
        if (rok) {
          load_split8<DT>(ph, pl, row + (long long)w * 8, v);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) xc[c] = v[c];
      }
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        xl[c] = __shfl_up_sync(0xffffffffu, xc[c], 1);
        xr[c] = __shfl_down_sync(0xffffffffu, xc[c], 1);
      }
      if (!left_by_shfl) {
        float v[8];
        if (rok && w > 0) {
          load_split8<DT>(ph, pl, row + (long long)(w - 1) * 8, v);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) xl[c] = v[c];
      }
      if (!right_by_shfl) {
        float v[8];
        if (rok && w < Wd - 1) {
          load_split8<DT>(ph, pl, row + (long long)(w + 1) * 8, v);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) xr[c] = v[c];
      }
      constexpr int kStride = 3 * CIN * COUT;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          acc[co] = fmaf(xl[ci], Wc.w[(kd * 3 + kh) * kStride + (0 * CIN + ci) * COUT + co], acc[co]);
          acc[co] = fmaf(xc[ci], Wc.w[(kd * 3 + kh) * kStride + (1 * CIN + ci) * COUT + co], acc[co]);
          acc[co] = fmaf(xr[ci], Wc.w[(kd * 3 + kh) * kStride + (2 * CIN + ci) * COUT + co], acc[co]);
        }
    }
  }
  if (!active) return;
  float* ob = out + (long long)n * out_ns + gi * 8;
  float r[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) r[c] = (c < COUT) ? acc[c < COUT ? c : 0] + (bias ? bias[c] : 0.f) : 0.f;
  if (accumulate) {
    float o[8];
    load_f32x8(ob, o);
#pragma unroll
    for (int c = 0; c < 8; ++c) r[c] += o[c];
  }
  store_f32x8(ob, r);
}

template <int CIN, int COUT>
static int launch_small(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns, int in_dtype, int N,
                        const float* W, const float* bias, float* out, long long out_ns, int D, int H, int Wd,
                        int accumulate, cudaStream_t stream) {
  const long long groups = (long long)D * H * Wd;
  const dim3 grid((unsigned)((groups + kSmallThreads - 1) / kSmallThreads), N);
  SmallW<27 * CIN * COUT> Wc;  // W is a HOST pointer [27][8][8]; compact to [27][3 kw... ][CIN][COUT]
  for (int tap = 0; tap < 27; ++tap)
    for (int ci = 0; ci < CIN; ++ci)
      for (int co = 0; co < COUT; ++co) Wc.w[(tap * CIN + ci) * COUT + co] = W[(tap * 8 + ci) * 8 + co];
  if (in_dtype == TTA_F16)
    tta_launch(conv_small_kernel<TTA_F16, CIN, COUT>, grid, kSmallThreads, 0, stream, tta_pdl_family(32), in_hi, in_lo, in_ns, Wc, bias, out, out_ns, D, H, Wd, accumulate);
  else if (in_dtype == TTA_BF16)
    tta_launch(conv_small_kernel<TTA_BF16, CIN, COUT>, grid, kSmallThreads, 0, stream, tta_pdl_family(32), in_hi, in_lo, in_ns, Wc, bias, out, out_ns, D, H, Wd, accumulate);
  else
    tta_launch(conv_small_kernel<TTA_F16_HI, CIN, COUT>, grid, kSmallThreads, 0, stream, tta_pdl_family(32), in_hi, in_lo, in_ns, Wc, bias, out, out_ns, D, H, Wd, accumulate);
  return tta_check_launch("tta_conv_small");
}

}  // namespace tta

using namespace tta;

extern "C" {

// cin, cout <= 4: 27*16 weights = 1.7 KB of kernel parameters (the 4 KB limit rules out 8x8)
int tta_conv_small_supported(int K, int stride, int cin, int cout) {
  return K == 3 && stride == 1 && cin >= 1 && cin <= 4 && cout >= 1 && cout <= 4;
}

// W: HOST pointer, fp32 [27][8 ci][8 co] (zero padded; taps already flipped by the host for the
// transposed form) -- the weights are passed to the kernel by value as launch parameters.
// Input view must have exactly one channel chunk per sample slice (C8 = 1); output likewise.
int tta_conv_small(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns, int in_dtype, int N, int cin,
                   int D, int H, int Wd, const float* W, const float* bias, float* out, long long out_ns, int cout,
                   int accumulate, cudaStream_t stream) {
  TTA_RECORDABLE(tta_conv_small(in_hi, in_lo, in_ns, in_dtype, N, cin, D, H, Wd, W, bias, out, out_ns, cout, accumulate, s_));
  TTA_REQUIRE(in_hi && (in_lo || in_dtype == TTA_F16_HI) && W && out, "tta_conv_small: null pointer");
  TTA_REQUIRE(tta_conv_small_supported(3, 1, cin, cout), "tta_conv_small: cin=%d cout=%d unsupported", cin, cout);
  TTA_REQUIRE(in_dtype >= 0 && in_dtype <= 2, "tta_conv_small: bad dtype");
#define SMALL(CI, CO) \
  return launch_small<CI, CO>(in_hi, in_lo, in_ns, in_dtype, N, W, bias, out, out_ns, D, H, Wd, accumulate, stream)
  if (cin == 1 && cout == 1) SMALL(1, 1);
  if (cin == 2 && cout == 2) SMALL(2, 2);
  if (cin == 3 && cout == 3) SMALL(3, 3);
  SMALL(4, 4);
#undef SMALL
}

}  // extern "C"
