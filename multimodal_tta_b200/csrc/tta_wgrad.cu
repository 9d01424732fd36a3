// Weight and bias gradients of Conv3d / ConvTranspose3d (k in {1, 3}, stride in {1, 2}) for the SUPERVISED step
// (reference: loss.backward() in src/core/trainers/seg_trainer.py:142 for every parameter of the
// model, not only the norm affines TENT touches).  fp32 CUDA-core kernel -- exact arithmetic on the operands the
// forward / backward passes already hold:
//     conv      (mode 0): dW[co][ci][k] = sum_{n,o} x[n][s*o - p + k][ci] * dy[n][o][co]
//     transposed (mode 1): dW[ci][co][k] = sum_{n,i} x[n][i][ci] * dy[n][s*i - p + k][co]
// x = the forward operand planes (fp16 hi + lo, chunk layout, optionally w-parity-split rows), dy = the 16-bit
// gradient planes the dgrad convs read (one scaled fp16 plane or bf16 hi + lo).
//
// A CTA owns a (CI_T x CO_T) channel tile of dW for all taps and walks a range of spatial tiles: the "centre" tensor
// (dy for a conv, x for a transposed conv) and the halo of the other one are staged in shared memory as fp32; thread
// (kd, kh | ci group of 4 | co group of 4 | position slice) keeps 3 (kw) x 4 x 4 accumulators and per position does
// one 128-bit load of the centre values, three of the halo and 48 FMAs.  Partial sums leave through fp32 atomics
// once per CTA (a few thousand per layer).  This is the exact reference implementation of the row and the path for
// 1x1 convs, bf16x2 gradients and operand layouts the tensor-core kernel does not take (csrc/tta_wgrad_tc.cu runs the
// 3x3x3 layers with one scaled fp16 gradient plane); <= 4 x <= 4 channel stride-1 convs have their own per-voxel
// kernel below.
#include <cstring>

#include "tta_common.cuh"

namespace tta {

constexpr int kWgThreads = 288;

struct WgParams {
  const uint16_t* x_hi; const uint16_t* x_lo; long long x_ns; int Dx, Hx, Wx, x_wsplit;
  const uint16_t* dy_hi; const uint16_t* dy_lo; long long dy_ns; int dy_dtype, Dy, Hy, Wy, dy_wsplit;
  int N, stride, Cin, Cout, ci0, co0;   // ci0 / co0: first channel of this launch's channel tiles (set per block)
  int tiles_d, tiles_h, tiles_w, tiles_per_n, tiles_total, tiles_per_block;
  float scale;
  float* dw; int layout; int co_split; float* dw2;
};

__device__ __forceinline__ long long vox_index(int d, int h, int w, int H, int W, int wsplit) {
  const long long row = ((long long)d * H + h) * W;
  return row + (wsplit ? (w & 1) * (W >> 1) + (w >> 1) : w);
}

// MODE 0: centre = dy (CO_T channels), halo = x (CI_T);  MODE 1: centre = x (CI_T), halo = dy (CO_T)
template <int MODE, int K, int S, int CI_T, int CO_T, int TD, int TH, int TW>
__global__ void __launch_bounds__(kWgThreads)
wgrad_kernel(const WgParams P) {
  constexpr int NKK = K * K;                       // (kd, kh) pairs
  constexpr int CG = CI_T / 4, OG = CO_T / 4;
  constexpr int PS = kWgThreads / (NKK * CG * OG); // position slices
  static_assert(PS >= 1 && NKK * CG * OG * PS == kWgThreads, "thread mapping");
  constexpr int HD = S * (TD - 1) + K, HH = S * (TH - 1) + K, HW = S * (TW - 1) + K;
  constexpr int NC = TD * TH * TW, NH = HD * HH * HW;
  constexpr int CC = MODE == 0 ? CO_T : CI_T;      // centre channels
  constexpr int HC = MODE == 0 ? CI_T : CO_T;      // halo channels
  extern __shared__ __align__(16) float smem[];
  float* cen = smem;                               // [NC][CC]
  float* hal = smem + NC * CC;                     // [NH][HC]
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int ps = tid % PS;
  const int og = (tid / PS) % OG;
  const int cg = (tid / (PS * OG)) % CG;
  const int kk = tid / (PS * OG * CG);             // kd * K + kh
  const int kd = kk / K, kh = kk % K;
  const int ci0 = blockIdx.y * CI_T, co0 = blockIdx.z * CO_T;
  const int pad = (K - 1) / 2;
  // centre / halo tensor descriptions
  const uint16_t* c_hi = MODE == 0 ? P.dy_hi : P.x_hi;
  const uint16_t* c_lo = MODE == 0 ? P.dy_lo : P.x_lo;
  const uint16_t* h_hi = MODE == 0 ? P.x_hi : P.dy_hi;
  const uint16_t* h_lo = MODE == 0 ? P.x_lo : P.dy_lo;
  const long long c_ns = MODE == 0 ? P.dy_ns : P.x_ns, h_ns = MODE == 0 ? P.x_ns : P.dy_ns;
  const int cD = MODE == 0 ? P.Dy : P.Dx, cH = MODE == 0 ? P.Hy : P.Hx, cW = MODE == 0 ? P.Wy : P.Wx;
  const int hD = MODE == 0 ? P.Dx : P.Dy, hH = MODE == 0 ? P.Hx : P.Hy, hW = MODE == 0 ? P.Wx : P.Wy;
  const int c_ws = MODE == 0 ? P.dy_wsplit : P.x_wsplit, h_ws = MODE == 0 ? P.x_wsplit : P.dy_wsplit;
  const int c_dt = MODE == 0 ? P.dy_dtype : TTA_F16, h_dt = MODE == 0 ? TTA_F16 : P.dy_dtype;
  const int c_c0 = MODE == 0 ? co0 : ci0, h_c0 = MODE == 0 ? ci0 : co0;
  const int c_cmax = MODE == 0 ? (P.Cout + 7) / 8 * 8 : (P.Cin + 7) / 8 * 8;
  const int h_cmax = MODE == 0 ? (P.Cin + 7) / 8 * 8 : (P.Cout + 7) / 8 * 8;
  const long long cV = (long long)cD * cH * cW, hV = (long long)hD * hH * hW;

  float acc[K][4][4];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][i][j] = 0.f;

  auto load8 = [&](const uint16_t* hi, const uint16_t* lo, long long off, int dt, float (&v)[8]) {
    const U16x8 h = *reinterpret_cast<const U16x8*>(hi + off);
    if (dt == TTA_F16_HI) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = u16_to_f32<TTA_F16>(h.v[i]);
    } else {
      const U16x8 l = *reinterpret_cast<const U16x8*>(lo + off);
      if (dt == TTA_F16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = u16_to_f32<TTA_F16>(h.v[i]) + u16_to_f32<TTA_F16>(l.v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = u16_to_f32<TTA_BF16>(h.v[i]) + u16_to_f32<TTA_BF16>(l.v[i]);
      }
    }
  };

  const int t_begin = blockIdx.x * P.tiles_per_block;
  const int t_end = min(P.tiles_total, t_begin + P.tiles_per_block);
  for (int t = t_begin; t < t_end; ++t) {
    const int n = t / P.tiles_per_n;
    int r = t - n * P.tiles_per_n;
    const int tw = r % P.tiles_w; r /= P.tiles_w;
    const int th = r % P.tiles_h;
    const int td = r / P.tiles_h;
    const int d0 = td * TD, h0 = th * TH, w0 = tw * TW;           // centre origin
    const int hd0 = S * d0 - pad, hh0 = S * h0 - pad, hw0 = S * w0 - pad;   // halo origin
    __syncthreads();                                              // previous tile's readers are done
    // ---- stage the centre tile [NC][CC] and the halo tile [NH][HC] as fp32 (zero outside the tensors)
    for (int e = tid; e < NC * (CC / 8); e += kWgThreads) {
      const int ch = e % (CC / 8), p = e / (CC / 8);
      const int pw = p % TW, ph = (p / TW) % TH, pd = p / (TW * TH);
      const int d = d0 + pd, h = h0 + ph, w = w0 + pw;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      const int c = c_c0 + ch * 8;
      if (d < cD && h < cH && w < cW && c < c_cmax)
        load8(c_hi, c_lo, (long long)n * c_ns + ((long long)(c >> 3) * cV + vox_index(d, h, w, cH, cW, c_ws)) * 8, c_dt, v);
      float4* dst = reinterpret_cast<float4*>(cen + p * CC + ch * 8);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (int e = tid; e < NH * (HC / 8); e += kWgThreads) {
      const int ch = e % (HC / 8), p = e / (HC / 8);
      const int pw = p % HW, ph = (p / HW) % HH, pd = p / (HW * HH);
      const int d = hd0 + pd, h = hh0 + ph, w = hw0 + pw;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      const int c = h_c0 + ch * 8;
      if (d >= 0 && d < hD && h >= 0 && h < hH && w >= 0 && w < hW && c < h_cmax)
        load8(h_hi, h_lo, (long long)n * h_ns + ((long long)(c >> 3) * hV + vox_index(d, h, w, hH, hW, h_ws)) * 8, h_dt, v);
      float4* dst = reinterpret_cast<float4*>(hal + p * HC + ch * 8);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    // ---- accumulate: this thread's (kd, kh), ci group, co group over its slice of the centre positions
    for (int p = ps; p < NC; p += PS) {
      const int pw = p % TW, ph = (p / TW) % TH, pd = p / (TW * TH);
      const float4 cv = *reinterpret_cast<const float4*>(cen + p * CC + (MODE == 0 ? og : cg) * 4);
      const int hb = ((S * pd + kd) * HH + (S * ph + kh)) * HW + S * pw;
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const float4 hv = *reinterpret_cast<const float4*>(hal + (hb + kw) * HC + (MODE == 0 ? cg : og) * 4);
        const float xs[4] = {MODE == 0 ? hv.x : cv.x, MODE == 0 ? hv.y : cv.y, MODE == 0 ? hv.z : cv.z, MODE == 0 ? hv.w : cv.w};
        const float ys[4] = {MODE == 0 ? cv.x : hv.x, MODE == 0 ? cv.y : hv.y, MODE == 0 ? cv.z : hv.z, MODE == 0 ? cv.w : hv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[kw][i][j] = fmaf(xs[i], ys[j], acc[kw][i][j]);
      }
    }
  }
  // ---- flush: fp32 atomics into the weight gradient (reference parameter layout)
  constexpr int T = K * K * K;
#pragma unroll
  for (int kw = 0; kw < K; ++kw) {
    const int tap = (kd * K + kh) * K + kw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ci = ci0 + cg * 4 + i;
      if (ci >= P.Cin) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = co0 + og * 4 + j;
        if (co >= P.Cout) continue;
        const float v = acc[kw][i][j] * P.scale;
        if (v == 0.f) continue;
        float* dst;
        if (P.layout == 1) {                                   // ConvTranspose3d weight [ci][co][T]
          dst = P.dw + ((long long)ci * P.Cout + co) * T + tap;
        } else if (co < P.co_split) {                          // Conv3d weight [co][ci][T]
          dst = P.dw + ((long long)co * P.Cin + ci) * T + tap;
        } else {                                               // second half of a fused unit0 || shortcut conv
          dst = P.dw2 + ((long long)(co - P.co_split) * P.Cin + ci) * T + tap;
        }
        atomicAdd(dst, v);
      }
    }
  }
}

// db[c] += scale * sum_{n, v} dy[n][v][c]
__global__ void __launch_bounds__(256)
bias_grad_kernel(const uint16_t* dy_hi, const uint16_t* dy_lo, long long dy_ns, int dy_dtype, long long V, int Cout,
                 int co_split, float scale, float* db, float* db2) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, n = blockIdx.z;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const long long base = (long long)n * dy_ns + (long long)chunk * V * 8;
  for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < V; v += (long long)gridDim.x * 256) {
    float x[8];
    const U16x8 h = *reinterpret_cast<const U16x8*>(dy_hi + base + v * 8);
    if (dy_dtype == TTA_F16_HI) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = u16_to_f32<TTA_F16>(h.v[i]);
    } else {
      const U16x8 l = *reinterpret_cast<const U16x8*>(dy_lo + base + v * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x[i] = dy_dtype == TTA_F16 ? u16_to_f32<TTA_F16>(h.v[i]) + u16_to_f32<TTA_F16>(l.v[i])
                                   : u16_to_f32<TTA_BF16>(h.v[i]) + u16_to_f32<TTA_BF16>(l.v[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += x[i];
  }
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const int c = chunk * 8 + threadIdx.x;
    if (c < Cout && s != 0.f) atomicAdd(c < co_split ? db + c : db2 + (c - co_split), s * scale);
  }
}

template <int MODE, int K, int S, int CI_T, int CO_T, int TD, int TH, int TW>
static int launch_wgrad(WgParams P, int cD, int cH, int cW, cudaStream_t stream) {
  constexpr int HD = S * (TD - 1) + K, HH = S * (TH - 1) + K, HW = S * (TW - 1) + K;
  constexpr int CC = MODE == 0 ? CO_T : CI_T, HC = MODE == 0 ? CI_T : CO_T;
  constexpr size_t smem = (size_t)(TD * TH * TW * CC + HD * HH * HW * HC) * sizeof(float);
  static_assert(smem <= 200 * 1024, "wgrad tile does not fit shared memory");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad_kernel<MODE, K, S, CI_T, CO_T, TD, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return tta_check_launch("tta_conv_wgrad(cudaFuncSetAttribute)");
    configured = true;
  }
  P.tiles_d = (cD + TD - 1) / TD; P.tiles_h = (cH + TH - 1) / TH; P.tiles_w = (cW + TW - 1) / TW;
  P.tiles_per_n = P.tiles_d * P.tiles_h * P.tiles_w;
  P.tiles_total = P.tiles_per_n * P.N;
  const int nci = (P.Cin + CI_T - 1) / CI_T, nco = (P.Cout + CO_T - 1) / CO_T;
  int splits = (3 * 148 + nci * nco - 1) / (nci * nco);      // ~3 CTAs per SM over the whole launch
  if (splits > P.tiles_total) splits = P.tiles_total;
  if (splits < 1) splits = 1;
  P.tiles_per_block = (P.tiles_total + splits - 1) / splits;
  splits = (P.tiles_total + P.tiles_per_block - 1) / P.tiles_per_block;
  tta_launch(wgrad_kernel<MODE, K, S, CI_T, CO_T, TD, TH, TW>, dim3(splits, nci, nco), kWgThreads, smem, stream,
             tta_pdl_family(32), P);
  return tta_check_launch("tta_conv_wgrad");
}

template <int MODE, int K, int S, int TD, int TH, int TW>
static int dispatch_tiles(const WgParams& P, int cD, int cH, int cW, cudaStream_t stream) {
  const bool small_ci = P.Cin <= 8, small_co = P.Cout <= 8;
  if (small_ci && small_co) return launch_wgrad<MODE, K, S, 8, 8, TD, TH, TW>(P, cD, cH, cW, stream);
  if (small_ci) return launch_wgrad<MODE, K, S, 8, 32, TD, TH, TW>(P, cD, cH, cW, stream);
  if (small_co) return launch_wgrad<MODE, K, S, 16, 8, TD, TH, TW>(P, cD, cH, cW, stream);
  return launch_wgrad<MODE, K, S, 16, 32, TD, TH, TW>(P, cD, cH, cW, stream);
}


// ---------------------------------------------------------------- <= 4 x <= 4 channels, stride-1 conv (the UNet's
// full-resolution 3 -> 3 residual unit: 4.2 M voxels, 81 weights).  The generic kernel above tiles CHANNELS (16 x 32
// per CTA) and would carry 13 zero channels per real one; on the tensor cores the layer fills 3 of 64 rows.  Here a
// thread owns one (h, w) position of a 16 x 16 tile, a CTA one kd and a segment of d-planes: per plane the x tile
// (+ halo) is staged in shared memory as fp32, every thread reads its 9 neighbours x CI channels and does 9 * CI * CO
// FMAs with its own dy voxel into 9 * CI * CO register accumulators; one shuffle + shared-memory reduction and
// 9 * CI * CO atomics per CTA.
constexpr int kWsTile = 16, kWsThreads = kWsTile * kWsTile;

struct WsParams {
  const uint16_t* x_hi; const uint16_t* x_lo; long long x_ns;
  const uint16_t* dy_hi; const uint16_t* dy_lo; long long dy_ns;
  int D, H, W, N, Cin, Cout, dseg, nseg, tiles_h, tiles_w;
  float scale;
  float* dw;
};

template <int CI, int CO, int DT>
__global__ void __launch_bounds__(kWsThreads)
wgrad_small_kernel(const WsParams P) {
  constexpr int HT = kWsTile + 2, PW = HT + 1;               // halo tile, padded rows
  __shared__ float xs[CI][HT][PW];
  __shared__ float red[kWsThreads / 32][9 * CI * CO];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, tw = tid % kWsTile, th = tid / kWsTile;
  int b = blockIdx.x;
  const int tile_w = b % P.tiles_w; b /= P.tiles_w;
  const int tile_h = b % P.tiles_h; b /= P.tiles_h;
  const int seg = b % P.nseg, n = b / P.nseg;
  const int kd = blockIdx.y;
  const int h0 = tile_h * kWsTile, w0 = tile_w * kWsTile;
  const int h = h0 + th, w = w0 + tw;
  const bool inside = h < P.H && w < P.W;
  const long long plane = (long long)P.H * P.W;
  const uint16_t* xh = P.x_hi + (long long)n * P.x_ns;
  const uint16_t* xl = P.x_lo + (long long)n * P.x_ns;
  const uint16_t* gh = P.dy_hi + (long long)n * P.dy_ns;
  const uint16_t* gl = P.dy_lo ? P.dy_lo + (long long)n * P.dy_ns : nullptr;
  float acc[9][CI][CO];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < CI; ++i)
#pragma unroll
      for (int o = 0; o < CO; ++o) acc[t][i][o] = 0.f;
  const int d_begin = seg * P.dseg, d_end = min(P.D, d_begin + P.dseg);
  // software pipeline: the next plane's values (<= 2 halo positions and the thread's own dy voxel) are in flight in
  // registers while the current plane is consumed from shared memory
  constexpr int NE = (HT * HT + kWsThreads - 1) / kWsThreads;
  float xv[NE][CI], g[CO], gn[CO];
  auto fetch = [&](int d) {
    const int dx = d + kd - 1;
    const bool live = d < d_end && dx >= 0 && dx < P.D;     // block-uniform
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int e = tid + k * kWsThreads;
      const int ey = e / HT, ex = e % HT;
      const int hh = h0 + ey - 1, ww = w0 + ex - 1;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      if (live && e < HT * HT && hh >= 0 && hh < P.H && ww >= 0 && ww < P.W)
        load_split8<TTA_F16>(xh, xl, ((long long)dx * plane + (long long)hh * P.W + ww) * 8, v);
#pragma unroll
      for (int i = 0; i < CI; ++i) xv[k][i] = v[i];
    }
    float gv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gv[i] = 0.f;
    if (live && inside) load_split8<DT>(gh, gl, ((long long)d * plane + (long long)h * P.W + w) * 8, gv);
#pragma unroll
    for (int o = 0; o < CO; ++o) gn[o] = gv[o];
  };
  fetch(d_begin);
  for (int d = d_begin; d < d_end; ++d) {
    __syncthreads();                                         // previous plane's readers are done
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int e = tid + k * kWsThreads;
      if (e < HT * HT) {
#pragma unroll
        for (int i = 0; i < CI; ++i) xs[i][e / HT][e % HT] = xv[k][i];
      }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) g[o] = gn[o];
    __syncthreads();
    fetch(d + 1);                                            // zero padding planes arrive as zeros (g = 0: no contribution)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int i = 0; i < CI; ++i) {
          const float xval = xs[i][th + kh][tw + kw];
#pragma unroll
          for (int o = 0; o < CO; ++o) acc[kh * 3 + kw][i][o] = fmaf(xval, g[o], acc[kh * 3 + kw][i][o]);
        }
  }
  // ---- CTA reduction: warp shuffles, then one row per warp in shared memory
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < CI; ++i)
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        const float s = warp_sum(acc[t][i][o]);
        if (lane == 0) red[warp][(t * CI + i) * CO + o] = s;
      }
  __syncthreads();
  for (int e = tid; e < 9 * CI * CO; e += kWsThreads) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kWsThreads / 32; ++wv) s += red[wv][e];
    const int o = e % CO, i = (e / CO) % CI, t = e / (CO * CI);
    if (i < P.Cin && o < P.Cout && s != 0.f)
      atomicAdd(P.dw + ((long long)o * P.Cin + i) * 27 + kd * 9 + t, s * P.scale);   // Conv3d layout [Cout][Cin][27]
  }
}

template <int CI, int CO>
static int launch_wgrad_small(const WsParams& P, int dy_dtype, cudaStream_t stream) {
  const dim3 grid((unsigned)(P.N * P.nseg * P.tiles_h * P.tiles_w), 3);
  if (dy_dtype == TTA_F16_HI) tta_launch(wgrad_small_kernel<CI, CO, TTA_F16_HI>, grid, kWsThreads, 0, stream, tta_pdl_family(32), P);
  else if (dy_dtype == TTA_F16) tta_launch(wgrad_small_kernel<CI, CO, TTA_F16>, grid, kWsThreads, 0, stream, tta_pdl_family(32), P);
  else tta_launch(wgrad_small_kernel<CI, CO, TTA_BF16>, grid, kWsThreads, 0, stream, tta_pdl_family(32), P);
  return tta_check_launch("tta_conv_wgrad (small channels)");
}
}  // namespace tta

using namespace tta;

extern "C" {

// x: forward operand planes (fp16 hi + lo) of the conv's INPUT view [N][C8][Dx][Hx][Wx][8] (x_wsplit: rows stored
// w-parity-split); dy: gradient planes of the conv's OUTPUT [N][C8][Dy][Hy][Wy][8] (dy_dtype TTA_F16_HI: one plane,
// dy_lo unused; TTA_BF16 / TTA_F16: hi + lo).  dw (+=) scale * dL/dW in the PARAMETER layout: layout 0 = nn.Conv3d
// [Cout][Cin][k^3] (couts >= co_split go to dw2, a second Conv3d weight [Cout - co_split][Cin][k^3]: the fused
// unit0 || shortcut convs), layout 1 = nn.ConvTranspose3d [Cin][Cout][k^3].  The caller zeroes dw once per step.
int tta_conv_wgrad(const uint16_t* x_hi, const uint16_t* x_lo, long long x_ns, int Dx, int Hx, int Wx, int x_wsplit,
                   const uint16_t* dy_hi, const uint16_t* dy_lo, long long dy_ns, int dy_dtype, int Dy, int Hy, int Wy,
                   int dy_wsplit, int N, int mode, int K, int stride, int Cin, int Cout, float scale, float* dw,
                   int layout, int co_split, float* dw2, cudaStream_t stream) {
  TTA_RECORDABLE(tta_conv_wgrad(x_hi, x_lo, x_ns, Dx, Hx, Wx, x_wsplit, dy_hi, dy_lo, dy_ns, dy_dtype, Dy, Hy, Wy, dy_wsplit, N, mode, K, stride, Cin, Cout, scale, dw, layout, co_split, dw2, s_));
  TTA_REQUIRE(x_hi && x_lo && dy_hi && (dy_lo || dy_dtype == TTA_F16_HI) && dw, "tta_conv_wgrad: null pointer");
  TTA_REQUIRE((mode == 0 || mode == 1) && (K == 1 || K == 3) && (stride == 1 || stride == 2) && !(K == 1 && stride != 1),
              "tta_conv_wgrad: unsupported geometry mode=%d K=%d stride=%d", mode, K, stride);
  TTA_REQUIRE(dy_dtype >= 0 && dy_dtype <= 2 && N > 0 && Cin > 0 && Cout > 0, "tta_conv_wgrad: bad argument");
  TTA_REQUIRE(layout == 0 || layout == 1, "tta_conv_wgrad: layout %d", layout);
  if (co_split <= 0 || co_split > Cout) co_split = Cout;
  TTA_REQUIRE(co_split == Cout || (dw2 != nullptr && layout == 0), "tta_conv_wgrad: a split output needs dw2 and layout 0");
  if (mode == 0 && K == 3 && stride == 1 && Cin <= 4 && Cout <= 4 && !x_wsplit && !dy_wsplit && layout == 0 &&
      co_split == Cout) {
    // <= 4 x <= 4 channels at (typically) full resolution: one thread per voxel, no channel tiles
    WsParams S;
    memset(&S, 0, sizeof(S));
    S.x_hi = x_hi; S.x_lo = x_lo; S.x_ns = x_ns; S.dy_hi = dy_hi; S.dy_lo = dy_dtype == TTA_F16_HI ? nullptr : dy_lo;
    S.dy_ns = dy_ns; S.D = Dx; S.H = Hx; S.W = Wx; S.N = N; S.Cin = Cin; S.Cout = Cout; S.scale = scale; S.dw = dw;
    S.tiles_h = (Hx + kWsTile - 1) / kWsTile; S.tiles_w = (Wx + kWsTile - 1) / kWsTile;
    // d-segments: ~4 CTAs per SM over the grid, at least 8 planes each (the reduction costs about 4 planes' work)
    const long long cols = (long long)N * S.tiles_h * S.tiles_w * 3;
    int nseg = (int)((4LL * 148 + cols - 1) / cols);
    if (nseg > (Dx + 7) / 8) nseg = (Dx + 7) / 8;
    if (nseg < 1) nseg = 1;
    S.dseg = (Dx + nseg - 1) / nseg;
    S.nseg = (Dx + S.dseg - 1) / S.dseg;
    if (Cin <= 3 && Cout <= 3) return launch_wgrad_small<3, 3>(S, dy_dtype, stream);
    return launch_wgrad_small<4, 4>(S, dy_dtype, stream);
  }
  WgParams P;
  memset(&P, 0, sizeof(P));
  P.x_hi = x_hi; P.x_lo = x_lo; P.x_ns = x_ns; P.Dx = Dx; P.Hx = Hx; P.Wx = Wx; P.x_wsplit = x_wsplit;
  P.dy_hi = dy_hi; P.dy_lo = dy_lo; P.dy_ns = dy_ns; P.dy_dtype = dy_dtype; P.Dy = Dy; P.Hy = Hy; P.Wy = Wy;
  P.dy_wsplit = dy_wsplit;
  P.N = N; P.stride = stride; P.Cin = Cin; P.Cout = Cout; P.scale = scale;
  P.dw = dw; P.layout = layout; P.co_split = co_split; P.dw2 = dw2;
  // centre tensor: dy for a conv (mode 0), x for a transposed conv (mode 1)
  const int cD = mode == 0 ? Dy : Dx, cH = mode == 0 ? Hy : Hx, cW = mode == 0 ? Wy : Wx;
  if (mode == 0) {
    if (K == 1) return dispatch_tiles<0, 1, 1, 4, 4, 8>(P, cD, cH, cW, stream);
    if (stride == 1) return dispatch_tiles<0, 3, 1, 4, 4, 8>(P, cD, cH, cW, stream);
    return dispatch_tiles<0, 3, 2, 2, 4, 8>(P, cD, cH, cW, stream);
  }
  if (K == 1) return dispatch_tiles<1, 1, 1, 4, 4, 8>(P, cD, cH, cW, stream);
  if (stride == 1) return dispatch_tiles<1, 3, 1, 4, 4, 8>(P, cD, cH, cW, stream);
  return dispatch_tiles<1, 3, 2, 2, 2, 8>(P, cD, cH, cW, stream);
}

// db (+=) scale * sum over batch and voxels of dy (channels >= co_split go to db2)
int tta_bias_grad(const uint16_t* dy_hi, const uint16_t* dy_lo, long long dy_ns, int dy_dtype, int N, int Cout, long long V,
                  float scale, float* db, int co_split, float* db2, cudaStream_t stream) {
  TTA_RECORDABLE(tta_bias_grad(dy_hi, dy_lo, dy_ns, dy_dtype, N, Cout, V, scale, db, co_split, db2, s_));
  TTA_REQUIRE(dy_hi && (dy_lo || dy_dtype == TTA_F16_HI) && db && N > 0 && Cout > 0 && V > 0, "tta_bias_grad: bad argument");
  if (co_split <= 0 || co_split > Cout) co_split = Cout;
  TTA_REQUIRE(co_split == Cout || db2 != nullptr, "tta_bias_grad: a split output needs db2");
  const int C8 = (Cout + 7) / 8;
  long long xb = (V + 255) / 256;
  const long long want = (4LL * 148 + (long long)N * C8 - 1) / ((long long)N * C8);
  if (xb > want) xb = want;
  tta_launch(bias_grad_kernel, dim3((unsigned)xb, C8, N), 256, 0, stream, tta_pdl_family(32), dy_hi, dy_lo, dy_ns, dy_dtype, V,
             Cout, co_split, scale, db, db2);
  return tta_check_launch("tta_bias_grad");
}

}  // extern "C"

// ---------------------------------------------------------------- dL/dlogits -> gradient planes
// g: NCDHW fp32 (what autograd hands to the model's backward) -> the 16-bit gradient plane(s) of the last conv's
// result in the chunk layout, multiplied by the (power-of-two) loss scale
namespace tta {

template <int ODT>
__global__ void __launch_bounds__(256)
pack_grad_kernel(const float* g, int R, long long V, float scale, uint16_t* dy_hi, uint16_t* dy_lo, long long dy_ns) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, n = blockIdx.z;
  const float* gb = g + (long long)n * R * V;
  for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < V; v += (long long)gridDim.x * 256) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = chunk * 8 + i;
      x[i] = c < R ? gb[(long long)c * V + v] * scale : 0.f;
    }
    store_split8<ODT>(dy_hi, dy_lo, (long long)n * dy_ns + ((long long)chunk * V + v) * 8, x);
  }
}

}  // namespace tta

extern "C" int tta_pack_grad(const float* g, int N, int R, long long V, float scale, uint16_t* dy_hi, uint16_t* dy_lo,
                             long long dy_ns, int out_dtype, cudaStream_t stream) {
  TTA_RECORDABLE(tta_pack_grad(g, N, R, V, scale, dy_hi, dy_lo, dy_ns, out_dtype, s_));
  TTA_REQUIRE(g && dy_hi && (dy_lo || out_dtype == TTA_F16_HI) && N > 0 && R > 0 && V > 0 && out_dtype >= 0 && out_dtype <= 2,
              "tta_pack_grad: bad argument");
  const int C8 = (R + 7) / 8;
  long long xb = (V + 255) / 256;
  const long long want = (8LL * 148 + (long long)N * C8 - 1) / ((long long)N * C8);
  if (xb > want) xb = want;
  const dim3 grid((unsigned)xb, C8, N);
  if (out_dtype == TTA_F16) tta::tta_launch(tta::pack_grad_kernel<TTA_F16>, grid, 256, 0, stream, tta_pdl_family(32), g, R, V, scale, dy_hi, dy_lo, dy_ns);
  else if (out_dtype == TTA_F16_HI) tta::tta_launch(tta::pack_grad_kernel<TTA_F16_HI>, grid, 256, 0, stream, tta_pdl_family(32), g, R, V, scale, dy_hi, dy_lo, dy_ns);
  else tta::tta_launch(tta::pack_grad_kernel<TTA_BF16>, grid, 256, 0, stream, tta_pdl_family(32), g, R, V, scale, dy_hi, dy_lo, dy_ns);
  return tta_check_launch("tta_pack_grad");
}

// ---------------------------------------------------------------- device-side weight repack
// After an optimizer step the packed operand blobs of the conv kernels (fp16 hi / lo planes in the tcgen05 blob
// layouts, fp32 for the CUDA-core kernels, padded biases) are rebuilt FROM THE FLAT PARAMETER BUFFER by one gather
// per blob: map[e] = +-(1 + source index), sign = hi / lo plane, 0 = padding; add[e] != 0 adds the folded identity
// shortcut's 1.0 (layout.index_mode builds the maps once by running the host packers on index tensors).
namespace tta {

template <int KIND>   // 0: fp32, 1: fp16 planes, 2: bf16 planes
__global__ void __launch_bounds__(256)
repack_kernel(const float* src, const int* map, const signed char* add, long long n, void* out) {
  pdl_trigger();
  pdl_wait();
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    const int m = map[e];
    const int a = add ? (int)add[e] : 0;
    float v = 0.f;
    if (m != 0) v = src[(m > 0 ? m : -m) - 1];
    if (a) v += 1.f;
    if (KIND == 0) {
      reinterpret_cast<float*>(out)[e] = v;
    } else {
      constexpr int DT = KIND == 1 ? TTA_F16 : TTA_BF16;
      const uint16_t h = f32_to_u16<DT>(v);
      uint16_t r = h;
      if (m < 0) r = f32_to_u16<DT>(v - u16_to_f32<DT>(h));
      if (m == 0 && !a) r = 0;
      reinterpret_cast<uint16_t*>(out)[e] = r;
    }
  }
}

}  // namespace tta

extern "C" int tta_repack_weights(const float* src, const int* map, const signed char* add, long long n, void* out,
                                  int kind, cudaStream_t stream) {
  TTA_RECORDABLE(tta_repack_weights(src, map, add, n, out, kind, s_));
  TTA_REQUIRE(src && map && out && n > 0 && kind >= 0 && kind <= 2, "tta_repack_weights: bad argument");
  long long blocks = (n + 255) / 256;
  if (blocks > 8 * 148) blocks = 8 * 148;
  if (kind == 0) tta::tta_launch(tta::repack_kernel<0>, dim3((unsigned)blocks), 256, 0, stream, tta_pdl_family(32), src, map, add, n, out);
  else if (kind == 1) tta::tta_launch(tta::repack_kernel<1>, dim3((unsigned)blocks), 256, 0, stream, tta_pdl_family(32), src, map, add, n, out);
  else tta::tta_launch(tta::repack_kernel<2>, dim3((unsigned)blocks), 256, 0, stream, tta_pdl_family(32), src, map, add, n, out);
  return tta_check_launch("tta_repack_weights");
}
