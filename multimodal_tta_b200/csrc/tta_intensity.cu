// On-device input intensity policy: per-channel clip + (masked) z-score, the step immediately before
// the hot path.  Reference: /root/reference/src/datasets/transforms.py:147-200 (`_normalize_img`,
// branch (A) intensity_policy and branch (B) legacy mean/std), configured by
// configs/_global_patches/hecktor21.yaml:27-46.  The reference runs it on the CPU inside the dataset
// transform, one torch pass per channel with a host sync (`m.sum().item()`); here one streaming pass
// reduces every (volume, channel) slab at once and a second pass applies.
//
//   x  = clamp(x, lo, hi)                                   (rule.clip)
//   m  = x > mask_gt ; vals = x[m] if count(m) >= min_count else all of x     (rule.zscore.masked)
//   mu = mean(vals) ; sd = max(std(vals, unbiased=False), eps) ; x = (x - mu) / sd
//
// rules [C][8] floats per channel: {clip_on, lo, hi, z_mode, mask_gt, eps, mean, std}
//   z_mode 0: none, 1: masked z-score, 2: unmasked z-score, 3: fixed (x - mean) / std (legacy branch)
// affine [n_vol][C][4] floats written by the statistics pass: {lo, hi, mu, 1/sd}
#include "tta_reduce.cuh"

namespace tta {

__device__ __forceinline__ double block_sum_d(double v, double* red /* [8] */) {
  v = warp_sum_d(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) t += red[w];
  return t;
}

// grid (splits, C, n_vol).  partial [n_vol*C][splits][5] doubles: {count_m, sum_m, sq_m, sum_all, sq_all}
__global__ void __launch_bounds__(kThreads)
intensity_stats_kernel(const float* vol, int C, long long V, const float* rules, int min_count, int splits,
                       double* partial, unsigned int* counters, float* affine) {
  __shared__ double red[8];
  const int split = blockIdx.x, c = blockIdx.y, v = blockIdx.z;
  const float* R = rules + c * 8;
  const bool clip = R[0] != 0.f;
  const float lo = clip ? R[1] : -INFINITY, hi = clip ? R[2] : INFINITY;
  const int zmode = (int)R[3];
  const float mask_gt = R[4];
  const float* src = vol + ((long long)v * C + c) * V;
  const int slab = v * C + c;
  if (zmode == 1 || zmode == 2) {
    const long long per = (V + splits - 1) / splits;
    const long long i0 = (long long)split * per, i1 = min(V, i0 + per);
    double cm = 0.0, sm = 0.0, qm = 0.0, sa = 0.0, qa = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += kThreads) {
      const float x = fminf(fmaxf(src[i], lo), hi);
      const double xd = (double)x;
      sa += xd;
      qa += xd * xd;
      if (x > mask_gt) {
        cm += 1.0;
        sm += xd;
        qm += xd * xd;
      }
    }
    double tot[5] = {block_sum_d(cm, red), block_sum_d(sm, red), block_sum_d(qm, red), block_sum_d(sa, red),
                     block_sum_d(qa, red)};
    if (threadIdx.x == 0) {
      double* p = partial + ((long long)slab * splits + split) * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) p[k] = tot[k];
    }
  }
  // the last block of a slab finalizes (fixed summation order -> deterministic)
  if (!last_block_of_chunk(counters, slab, (unsigned int)splits)) return;
  if (threadIdx.x != 0) return;
  float mu = 0.f, inv = 1.f;
  if (zmode == 1 || zmode == 2) {
    double t[5] = {0, 0, 0, 0, 0};
    for (int s = 0; s < splits; ++s)
      for (int k = 0; k < 5; ++k) t[k] += partial[((long long)slab * splits + s) * 5 + k];
    const bool masked = zmode == 1 && t[0] >= (double)min_count;
    const double n = masked ? t[0] : (double)V;
    const double m = (masked ? t[1] : t[3]) / n;
    double var = (masked ? t[2] : t[4]) / n - m * m;
    if (var < 0.0) var = 0.0;
    double sd = sqrt(var);
    if (sd < (double)R[5]) sd = (double)R[5];
    mu = (float)m;
    inv = (float)(1.0 / sd);
  } else if (zmode == 3) {
    mu = R[6];
    inv = 1.f / R[7];
  }
  float* A = affine + (long long)slab * 4;
  A[0] = lo; A[1] = hi; A[2] = mu; A[3] = inv;
}

// out = (clamp(x, lo, hi) - mu) * inv_sd ; grid (xblocks, C, n_vol); in place allowed
__global__ void __launch_bounds__(kThreads)
intensity_apply_kernel(const float* vol, float* out, int C, long long V, const float* affine) {
  const int c = blockIdx.y, v = blockIdx.z;
  const float* A = affine + ((long long)v * C + c) * 4;
  const float lo = A[0], hi = A[1], mu = A[2], inv = A[3];
  const long long base = ((long long)v * C + c) * V;
  const bool vec = (V % 4 == 0) && ((reinterpret_cast<unsigned long long>(vol + base) & 15ull) == 0ull) &&
                   ((reinterpret_cast<unsigned long long>(out + base) & 15ull) == 0ull);
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(vol + base);
    float4* d4 = reinterpret_cast<float4*>(out + base);
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < V / 4; i += (long long)gridDim.x * kThreads) {
      float4 x = s4[i];
      x.x = (fminf(fmaxf(x.x, lo), hi) - mu) * inv;
      x.y = (fminf(fmaxf(x.y, lo), hi) - mu) * inv;
      x.z = (fminf(fmaxf(x.z, lo), hi) - mu) * inv;
      x.w = (fminf(fmaxf(x.w, lo), hi) - mu) * inv;
      d4[i] = x;
    }
  } else {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < V; i += (long long)gridDim.x * kThreads)
      out[base + i] = (fminf(fmaxf(vol[base + i], lo), hi) - mu) * inv;
  }
}

static inline int intensity_splits(int n_vol, int C, long long V) {
  long long want = (4LL * 148 + (long long)n_vol * C - 1) / ((long long)n_vol * C);
  const long long maxs = (V + 4095) / 4096;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return (int)want;
}

}  // namespace tta

using namespace tta;

extern "C" {

// workspace (bytes): [1024 block counters (zeroed once, self-resetting)][partials (doubles)]
long long tta_intensity_workspace_bytes(int n_vol, int C, long long V) {
  return 4096 + (long long)n_vol * C * intensity_splits(n_vol, C, V) * 5 * 8;
}

int tta_intensity_stats(const float* vol, int n_vol, int C, long long V, const float* rules, int min_count,
                        float* affine, void* workspace, cudaStream_t stream) {
  TTA_RECORDABLE(tta_intensity_stats(vol, n_vol, C, V, rules, min_count, affine, workspace, s_));
  TTA_REQUIRE(vol && rules && affine && workspace, "tta_intensity_stats: null pointer");
  TTA_REQUIRE(n_vol > 0 && C > 0 && V > 0 && n_vol * C <= 1024, "tta_intensity_stats: bad shape n_vol=%d C=%d V=%lld",
              n_vol, C, V);
  const int splits = intensity_splits(n_vol, C, V);
  intensity_stats_kernel<<<dim3(splits, C, n_vol), kThreads, 0, stream>>>(
      vol, C, V, rules, min_count, splits, reinterpret_cast<double*>(static_cast<char*>(workspace) + 4096),
      reinterpret_cast<unsigned int*>(workspace), affine);
  return tta_check_launch("tta_intensity_stats");
}

int tta_intensity_apply(const float* vol, float* out, int n_vol, int C, long long V, const float* affine,
                        cudaStream_t stream) {
  TTA_RECORDABLE(tta_intensity_apply(vol, out, n_vol, C, V, affine, s_));
  TTA_REQUIRE(vol && out && affine, "tta_intensity_apply: null pointer");
  TTA_REQUIRE(n_vol > 0 && C > 0 && V > 0, "tta_intensity_apply: bad shape");
  long long xb = (V / 4 + kThreads - 1) / kThreads;
  const long long want = (8LL * 148 + (long long)n_vol * C - 1) / ((long long)n_vol * C);
  if (xb > want) xb = want;
  if (xb < 1) xb = 1;
  intensity_apply_kernel<<<dim3((unsigned)xb, C, n_vol), kThreads, 0, stream>>>(vol, out, C, V, affine);
  return tta_check_launch("tta_intensity_apply");
}

}  // extern "C"
