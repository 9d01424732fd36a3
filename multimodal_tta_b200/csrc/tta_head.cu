// Fused top of the UNet at full resolution (HBM-bound: ~4 % of the FLOPs, ~30 % of the bytes).
//
// Forward, ONE kernel:   a = relu(gamma * (y - mean) * rstd + beta)          (norm apply, on the fly)
//                        z = bias + sum_{k,ci} a[o + k - 1][ci] * W[k][ci][co] (3x3x3 conv, <= 4 channels)
//                        logits <- z (NCDHW fp32), H(z) -> loss partials, dlogits <- dH/dz
// Backward, ONE kernel:  g[v][ci] = sum_{k,co} dlogits[v - k + 1][co] * W[k][ci][co]   (conv dgrad)
//                        dz = g * [gamma*xhat + beta > 0]   -> fp32 chunk layout (feeds tta_norm_bwd_apply)
//                        sum dz, sum dz*xhat per (n, c)     -> last block finalizes dgamma/dbeta
// The normalised activation, its split planes and the pre-mask gradient never touch HBM: compared
// with norm_apply + conv_small + head_entropy (+ conv_small dgrad + norm_bwd_reduce) this removes
// five full-resolution passes.
//
// Tiling: a CTA owns an 8(d) x 8(h) x 32(w) output tile; the 10 x 10 x 34 halo tile of the <= 4
// real channels is staged in shared memory (fp32, normalised on the way in), every thread slides
// along d over 8 outputs so each staged value feeds up to 3 x COUT x 3 FMAs.  The 27*CIN*COUT
// weights are kernel parameters (constant-bank FFMA operands, no loads).
//
// Reference semantics restated: MONAI UNet top layer = Convolution(convT) -> ADN(norm, ReLU) ->
// ResidualUnit(subunits=1, last_conv_only) reached from src/models/unet.py:56-66; entropy loss and
// TENT backward per SURVEY.md 8c-3; step shape src/core/trainers/seg_trainer.py:105-145.
#include "tta_common.cuh"
#include "tta_reduce.cuh"

namespace tta {

constexpr int kTD = 8, kTH = 8, kTW = 32;                    // output tile
constexpr int kHD = kTD + 2, kHH = kTH + 2, kHW = kTW + 2;   // halo tile
constexpr int kPitch = 36;                                   // padded smem row (floats)
constexpr int kPlane = kHH * kPitch, kVol = kHD * kPlane;    // floats per d-plane / per channel

template <int NW>
struct HeadW {
  float w[NW];  // [tap = (kd*3 + kh)*3 + kw][ci][co]
};

struct HeadGeom {
  int D, H, W, tiles_w, tiles_h, tiles_d;
};

__device__ __forceinline__ void tile_origin(const HeadGeom& G, int& d0, int& h0, int& w0) {
  int t = blockIdx.x;
  const int tw = t % G.tiles_w;
  t /= G.tiles_w;
  const int th = t % G.tiles_h;
  d0 = (t / G.tiles_h) * kTD;
  h0 = th * kTH;
  w0 = tw * kTW;
}

// ---------------------------------------------------------------- forward
template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 3)
head_fwd_kernel(const float* y, long long y_ns, const int cpv, const HeadGeom G, const float* mean,
                const float* rstd, const float* gamma, const float* beta,
                int relu, const __grid_constant__ HeadW<27 * CIN * COUT> Wc, const float* bias, int mode,
                float inv_count, float grad_scale, const float* sample_w, float* logits,
                float* dlogits, float* partial, unsigned int* counter,
                float* loss) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float tile[];  // [CIN][kHD][kHH][kPitch]
  const int n = blockIdx.y;
  int d0, h0, w0;
  tile_origin(G, d0, h0, w0);
  const long long V = (long long)G.D * G.H * G.W;
  const float* yb = y + (long long)n * y_ns;
  float mu[CIN], rs[CIN], ga[CIN], be[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    mu[c] = mean[n * 8 + c];
    rs[c] = rstd[n * 8 + c];
    ga[c] = gamma[c];
    be[c] = beta[c];
  }
  // ---- stage the normalised halo tile (zero padding applies to the ACTIVATION, so OOB -> 0).
  // A thread owns up to two fixed (h, w) positions of the 10 x 34 halo plane and walks the ten
  // d-planes: no index divisions in the loop, five independent 16-byte loads in flight per position.
  {
    const long long HW = (long long)G.H * G.W;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = threadIdx.x + k * kThreads;
      if (p < kHH * kHW) {  // warp-uniform for all but one warp
        const int hh = p / kHW, ww = p - hh * kHW;
        const int gh = h0 + hh - 1, gw = w0 + ww - 1;
        const bool hw_in = gh >= 0 && gh < G.H && gw >= 0 && gw < G.W;
        // cpv = floats per voxel of y: 8 (channel-chunk layout, 4 pad channels ride along in every sector) or
        // 4 (the compact layout the small-Cout transposed conv writes for this kernel: dense 16-byte voxels)
        const float* src = yb + ((long long)gh * G.W + gw) * cpv;
        float* dst = tile + hh * kPitch + ww;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 x[kHD / 2];
          bool in[kHD / 2];
#pragma unroll
          for (int j = 0; j < kHD / 2; ++j) {
            const int gd = d0 + half * (kHD / 2) + j - 1;
            in[j] = hw_in && gd >= 0 && gd < G.D;
            x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in[j]) x[j] = __ldcg(reinterpret_cast<const float4*>(src + gd * HW * cpv));
          }
#pragma unroll
          for (int j = 0; j < kHD / 2; ++j) {
            const float xv[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
              float z = (xv[c] - mu[c]) * rs[c];
              z = fmaf(z, ga[c], be[c]);
              if (relu) z = fmaxf(z, 0.f);
              dst[c * kVol + (half * (kHD / 2) + j) * kPlane] = in[j] ? z : 0.f;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- 3x3x3 conv, 8 outputs along d per thread.  Output plane o is complete once input plane
  // o + 2 has been consumed, so it is finished (bias, logits, entropy, dlogits) right there: only
  // three accumulator sets are ever live.
  const int hl = threadIdx.x >> 5, wl = threadIdx.x & 31;
  const float sw = sample_w ? sample_w[n] : 1.f;
  const float gs = sw * inv_count * grad_scale;
  const int h = h0 + hl, w = w0 + wl;
  const bool hw_ok = h < G.H && w < G.W;
  const long long HWo = (long long)G.H * G.W;
  const long long vbase = (long long)n * COUT * V + ((long long)d0 * G.H + h) * G.W + w;  // plane o: + o*HWo
  float* lg = logits + vbase;
  float* dl = dlogits + vbase;
  float hsum = 0.f;
  float bi[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) bi[c] = bias ? bias[c] : 0.f;
  float acc[kTD][COUT];
#pragma unroll
  for (int o = 0; o < kTD; ++o)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[o][c] = 0.f;
#pragma unroll
  for (int dd = 0; dd < kHD; ++dd) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float v = tile[ci * kVol + dd * kPlane + (hl + kh) * kPitch + wl + kw];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const int od = dd - kd;
            if (od >= 0 && od < kTD) {
#pragma unroll
              for (int co = 0; co < COUT; ++co)
                acc[od][co] = fmaf(v, Wc.w[(((kd * 3 + kh) * 3 + kw) * CIN + ci) * COUT + co], acc[od][co]);
            }
          }
        }
    if (dd >= 2) {
      const int o = dd - 2, d = d0 + o;
      if (d < G.D && hw_ok) {
        float z[COUT], g[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          z[c] = acc[o][c] + bi[c];
          g[c] = 0.f;
          if (logits) lg[c * V + o * HWo] = z[c];
        }
        hsum += entropy_point<COUT>(z, COUT, mode, gs, g) * sw;
        if (dlogits) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) dl[c * V + o * HWo] = g[c];
        }
      }
    }
    asm volatile("" ::: "memory");  // no hoisting of the next plane's staged values over this one
  }
  // ---- loss: per-block partial, the last block of the grid sums them in a fixed order (fp64)
  __shared__ float red[kThreads / 32];
  __shared__ unsigned int s_last;
  hsum = warp_sum(hsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hsum;
  __syncthreads();
  const unsigned int nblk = gridDim.x * gridDim.y;
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
    __threadfence();
    const unsigned int old = atomicAdd(counter, 1u);
    s_last = old == nblk - 1u;
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double redd[kThreads / 32];
  double s = 0.0;
  for (unsigned int i = threadIdx.x; i < nblk; i += kThreads) s += (double)__ldcg(partial + i);
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) redd[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += redd[i];
    *loss = (float)(t * (double)inv_count);
  }
}

// ---------------------------------------------------------------- backward
template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 3)
head_bwd_kernel(const float* dlogits, const HeadGeom G, const __grid_constant__ HeadW<27 * CIN * COUT> Wc,
                const float* y, long long y_ns, const int cpv, const float* mean,
                const float* rstd, const float* gamma, const float* beta,
                int relu, float* dz, long long dz_ns, float* partial,
                unsigned int* counters, int N, int batch_mode, float* sums,
                float* dgamma, float* dbeta) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float tile[];  // [COUT][kHD][kHH][kPitch]
  const int n = blockIdx.y;
  int d0, h0, w0;
  tile_origin(G, d0, h0, w0);
  const long long V = (long long)G.D * G.H * G.W;
  {
    const long long HW = (long long)G.H * G.W;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int p = threadIdx.x + k * kThreads;
      if (p < kHH * kHW) {
        const int hh = p / kHW, ww = p - hh * kHW;
        const int gh = h0 + hh - 1, gw = w0 + ww - 1;
        const bool hw_in = gh >= 0 && gh < G.H && gw >= 0 && gw < G.W;
        const float* src = dlogits + (long long)n * COUT * V + (long long)gh * G.W + gw;
        float* dst = tile + hh * kPitch + ww;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float x[kHD / 2][COUT];
#pragma unroll
          for (int j = 0; j < kHD / 2; ++j) {
            const int gd = d0 + half * (kHD / 2) + j - 1;
            const bool in = hw_in && gd >= 0 && gd < G.D;
#pragma unroll
            for (int c = 0; c < COUT; ++c) x[j][c] = in ? __ldcg(src + c * V + gd * HW) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < kHD / 2; ++j)
#pragma unroll
            for (int c = 0; c < COUT; ++c) dst[c * kVol + (half * (kHD / 2) + j) * kPlane] = x[j][c];
        }
      }
    }
  }
  __syncthreads();
  const int hl = threadIdx.x >> 5, wl = threadIdx.x & 31;
  float mu[CIN], rs[CIN], ga[CIN], be[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    mu[c] = mean[n * 8 + c];
    rs[c] = rstd[n * 8 + c];
    ga[c] = gamma[c];
    be[c] = beta[c];
  }
  float red16[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) red16[i] = 0.f;
  const int h = h0 + hl, w = w0 + wl;
  const bool hw_ok = h < G.H && w < G.W;
  const float* yb = y + (long long)n * y_ns;
  float* ob = dz + (long long)n * dz_ns;
  const long long HWo = (long long)G.H * G.W;
  const long long vbase = ((long long)d0 * G.H + h) * G.W + w;  // voxel of output plane o: + o*HWo
  float acc[kTD][CIN];
#pragma unroll
  for (int o = 0; o < kTD; ++o)
#pragma unroll
    for (int c = 0; c < CIN; ++c) acc[o][c] = 0.f;
  float4 yv[kTD];  // conv result of output plane o, requested one plane ahead of its use
#pragma unroll
  for (int dd = 0; dd < kHD; ++dd) {
    if (dd >= 1 && dd <= kTD) {
      const int o = dd - 1, d = d0 + o;
      yv[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d < G.D && hw_ok) yv[o] = __ldcg(reinterpret_cast<const float4*>(yb + (vbase + o * HWo) * cpv));
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          const float v = tile[co * kVol + dd * kPlane + (hl + 2 - kh) * kPitch + wl + 2 - kw];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const int od = dd + kd - 2;
            if (od >= 0 && od < kTD) {
#pragma unroll
              for (int ci = 0; ci < CIN; ++ci)
                acc[od][ci] = fmaf(v, Wc.w[(((kd * 3 + kh) * 3 + kw) * CIN + ci) * COUT + co], acc[od][ci]);
            }
          }
        }
    if (dd >= 2) {
      // output plane o = dd - 2 is complete: ReLU mask, norm-backward partial sums, masked gradient out
      const int o = dd - 2, d = d0 + o;
      if (d < G.D && hw_ok) {
        const long long v = vbase + o * HWo;
        const float xv[4] = {yv[o].x, yv[o].y, yv[o].z, yv[o].w};
        float r[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) r[c] = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float xh = (xv[c] - mu[c]) * rs[c];
          const float z = fmaf(xh, ga[c], be[c]);
          const float dzv = (relu && !(z > 0.f)) ? 0.f : acc[o][c];
          red16[c] += dzv;
          red16[8 + c] = fmaf(dzv, xh, red16[8 + c]);
          r[c] = dzv;
        }
        if (cpv == 8) {
          store_f32x8(ob + v * 8, r);
        } else {
          // compact layout: the masked gradient leaves as four fp16 values per voxel (8 B instead of 32 B);
          // its only consumer, tta_norm_bwd_apply_c4, rounds its own result to fp16 anyway
          const __half2 h01 = __floats2half2_rn(fminf(fmaxf(r[0], -65504.f), 65504.f), fminf(fmaxf(r[1], -65504.f), 65504.f));
          const __half2 h23 = __floats2half2_rn(fminf(fmaxf(r[2], -65504.f), 65504.f), fminf(fmaxf(r[3], -65504.f), 65504.f));
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned int*>(&h01);
          pk.y = *reinterpret_cast<const unsigned int*>(&h23);
          *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dz) + ((long long)n * dz_ns + v * 4)) = pk;
        }
      }
    }
    asm volatile("" ::: "memory");
  }
  const int splits = gridDim.x;
  block_reduce_store<16>(red16, partial + ((long long)n * splits + blockIdx.x) * 16);
  if (!last_block_of_chunk(counters, 0, (unsigned int)(N * splits))) return;
  norm_bwd_finalize_tail(partial, 1, 0, splits, N, batch_mode, CIN, sums, dgamma, dbeta);
}

static HeadGeom make_geom(int D, int H, int W) {
  HeadGeom G;
  G.D = D; G.H = H; G.W = W;
  G.tiles_w = (W + kTW - 1) / kTW;
  G.tiles_h = (H + kTH - 1) / kTH;
  G.tiles_d = (D + kTD - 1) / kTD;
  return G;
}

template <int CIN, int COUT>
static void fill_w(HeadW<27 * CIN * COUT>& Wc, const float* W_host) {
  for (int tap = 0; tap < 27; ++tap)
    for (int ci = 0; ci < CIN; ++ci)
      for (int co = 0; co < COUT; ++co) Wc.w[(tap * CIN + ci) * COUT + co] = W_host[(tap * 8 + ci) * 8 + co];
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

}  // namespace tta

using namespace tta;

extern "C" {

int tta_head_fused_supported(int K, int stride, int cin, int cout) {
  return K == 3 && stride == 1 && cin >= 1 && cin <= 4 && cin == cout;
}

// CTAs per sample (= per-sample loss partials, = reduction splits of the backward)
int tta_head_fused_tiles(int D, int H, int W) {
  const HeadGeom G = make_geom(D, H, W);
  return G.tiles_w * G.tiles_h * G.tiles_d;
}

// workspace (floats) shared by both kernels: [1024 counters][max(N*tiles loss partials, N*tiles*16)]
long long tta_head_fused_workspace_floats(int N, int D, int H, int W) {
  return 1024 + (long long)N * tta_head_fused_tiles(D, H, W) * 16;
}

// y: fp32 chunk view [N][1][D][H][W][8] (the transposed conv's result, C <= 4 real channels), or with
// y_cpv = 4 the compact view [N][D][H][W][4] written by tta_conv_tc(flags bit 14) (y_n_stride in floats);
// mean/rstd: [N][8]; W_host: HOST fp32 [27][8][8] = Wg[tap][ci][co] of the 3x3x3 conv (identity
// shortcut already folded into the centre tap); logits / dlogits: NCDHW fp32 [N][C][D][H][W]
// (dlogits may be NULL: inference).  loss <- mean entropy * 1 (same scaling as tta_head_entropy).
int tta_head_fused_fwd(const float* y, long long y_ns, int y_cpv, int N, int C, int D, int H, int W, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, int relu, const float* W_host,
                       const float* bias, int mode, float inv_count, float grad_scale, const float* sample_w,
                       float* logits, float* dlogits, float* workspace, float* loss, cudaStream_t stream) {
  TTA_RECORDABLE(tta_head_fused_fwd(y, y_ns, y_cpv, N, C, D, H, W, mean, rstd, gamma, beta, relu, W_host, bias, mode, inv_count, grad_scale, sample_w, logits, dlogits, workspace, loss, s_));
  TTA_REQUIRE(y_cpv == 8 || y_cpv == 4, "tta_head_fused_fwd: y_cpv %d (8 = chunk layout, 4 = compact)", y_cpv);
  TTA_REQUIRE(y && mean && rstd && gamma && beta && W_host && workspace && loss, "tta_head_fused_fwd: null pointer");
  TTA_REQUIRE(tta_head_fused_supported(3, 1, C, C), "tta_head_fused_fwd: %d channels unsupported (1..4)", C);
  TTA_REQUIRE(mode == 0 || mode == 1, "tta_head_fused_fwd: mode %d", mode);
  TTA_REQUIRE(!(mode == 0 && C < 2), "tta_head_fused_fwd: softmax entropy is degenerate for one channel");
  const HeadGeom G = make_geom(D, H, W);
  const dim3 grid(G.tiles_w * G.tiles_h * G.tiles_d, N);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace) + 1023;
#define HEAD_FWD(CC)                                                                                          \
  do {                                                                                                        \
    HeadW<27 * CC * CC> Wc;                                                                                   \
    fill_w<CC, CC>(Wc, W_host);                                                                               \
    const size_t smem = (size_t)CC * kVol * sizeof(float);                                                    \
    static bool configured = false;                                                                           \
    if (!configured) {                                                                                        \
      TTA_REQUIRE(set_smem(head_fwd_kernel<CC, CC>, smem), "tta_head_fused_fwd: cudaFuncSetAttribute failed"); \
      configured = true;                                                                                      \
    }                                                                                                         \
    tta_launch(head_fwd_kernel<CC, CC>, grid, kThreads, smem, stream, tta_pdl_family(16), y, y_ns, y_cpv, G, mean, rstd, gamma, beta, relu, Wc, bias, \
                                                              mode, inv_count, grad_scale, sample_w, logits,   \
                                                              dlogits, workspace + 1024, counter, loss);       \
  } while (0)
  if (C == 1) HEAD_FWD(1); else if (C == 2) HEAD_FWD(2); else if (C == 3) HEAD_FWD(3); else HEAD_FWD(4);
#undef HEAD_FWD
  return tta_check_launch("tta_head_fused_fwd");
}

// dz: fp32 chunk view [N][1][D][H][W][8] <- masked gradient w.r.t. the norm output (pad channels 0); with
// y_cpv = 4: FP16 [N][D][H][W][4] (dz_n_stride in 16-bit elements), consumed by tta_norm_bwd_apply_c4;
// sums [N][8][2], dgamma/dbeta [C]: finalized by the last block (as tta_norm_bwd_reduce, finalize=1);
// W_host: the SAME forward weights as tta_head_fused_fwd.
int tta_head_fused_bwd(const float* dlogits, int N, int C, int D, int H, int W, const float* W_host, const float* y,
                       long long y_ns, int y_cpv, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, int relu, int batch_mode, float* dz, long long dz_ns, float* sums,
                       float* dgamma, float* dbeta, float* workspace, cudaStream_t stream) {
  TTA_RECORDABLE(tta_head_fused_bwd(dlogits, N, C, D, H, W, W_host, y, y_ns, y_cpv, mean, rstd, gamma, beta, relu, batch_mode, dz, dz_ns, sums, dgamma, dbeta, workspace, s_));
  TTA_REQUIRE(y_cpv == 8 || y_cpv == 4, "tta_head_fused_bwd: y_cpv %d (8 = chunk layout, 4 = compact)", y_cpv);
  TTA_REQUIRE(dlogits && W_host && y && mean && rstd && gamma && beta && dz && sums && dgamma && dbeta && workspace,
              "tta_head_fused_bwd: null pointer");
  TTA_REQUIRE(tta_head_fused_supported(3, 1, C, C), "tta_head_fused_bwd: %d channels unsupported (1..4)", C);
  const HeadGeom G = make_geom(D, H, W);
  const dim3 grid(G.tiles_w * G.tiles_h * G.tiles_d, N);
  unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
#define HEAD_BWD(CC)                                                                                          \
  do {                                                                                                        \
    HeadW<27 * CC * CC> Wc;                                                                                   \
    fill_w<CC, CC>(Wc, W_host);                                                                               \
    const size_t smem = (size_t)CC * kVol * sizeof(float);                                                    \
    static bool configured = false;                                                                           \
    if (!configured) {                                                                                        \
      TTA_REQUIRE(set_smem(head_bwd_kernel<CC, CC>, smem), "tta_head_fused_bwd: cudaFuncSetAttribute failed"); \
      configured = true;                                                                                      \
    }                                                                                                         \
    tta_launch(head_bwd_kernel<CC, CC>, grid, kThreads, smem, stream, tta_pdl_family(16), dlogits, G, Wc, y, y_ns, y_cpv, mean, rstd, gamma, beta, \
                                                              relu, dz, dz_ns, workspace + 1024, counters, N,  \
                                                              batch_mode, sums, dgamma, dbeta);                \
  } while (0)
  if (C == 1) HEAD_BWD(1); else if (C == 2) HEAD_BWD(2); else if (C == 3) HEAD_BWD(3); else HEAD_BWD(4);
#undef HEAD_BWD
  return tta_check_launch("tta_head_fused_bwd");
}

}  // extern "C"
