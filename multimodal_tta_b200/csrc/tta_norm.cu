// Instance/Batch-norm statistics, fused normalise+affine+ReLU(+residual)+split, and the norm
// backward (affine gradients + input gradient) for the TTA step.  All HBM-streaming kernels:
// one thread handles one voxel-chunk (8 channels = 32 B fp32 / 16 B per 16-bit plane), a warp
// reads 1 KB contiguous, reductions are warp-shuffle -> smem -> per-block partials -> a tiny
// fp64 finalize kernel (deterministic, no atomics).
//
// Reference semantics restated:
//   nn.InstanceNorm3d(eps=1e-5): per-(n,c) mean / biased variance over D*H*W
//   nn.BatchNorm3d in TENT mode : per-c statistics over (n, D*H*W), no running stats
//   backward: dbeta = sum dz, dgamma = sum dz*xhat,
//             dy = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)),  dz = g * [z > 0]
#include "tta_common.cuh"
#include <cstdlib>

#include "tta_reduce.cuh"

namespace tta {

// ---------------------------------------------------------------- forward statistics
// partial[((n*C8 + chunk)*splits + split)*16 + {0..7: sum, 8..15: sumsq}]
__global__ void __launch_bounds__(kThreads, 4)
norm_stats_partial_kernel(const float* y, long long n_stride, int C8, long long V,
                          int splits, float* partial, unsigned int* counters,
                          int N, int batch_mode, float eps, float* mean,
                          float* rstd, int rev) {
  pdl_trigger();
  pdl_wait();
  // rev: blocks take the slabs / ranges in reverse launch order (what the producing conv wrote last is
  // still in L2); every block still sums its own range in ascending order -> same partial sums
  const int split = rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int chunk = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int n = rev ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const float* base = y + (long long)n * n_stride + (long long)chunk * V * 8;
  const long long per = (V + splits - 1) / splits;
  const long long v0 = (long long)split * per;
  const long long v1 = min(V, v0 + per);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 2
  for (long long v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float x[8];
    load_f32x8(base + v * 8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += x[i];
      acc[8 + i] = fmaf(x[i], x[i], acc[8 + i]);
    }
  }
  block_reduce_store<16>(acc, partial + ((long long)(n * C8 + chunk) * splits + split) * 16);
  if (counters == nullptr) return;
  if (!last_block_of_chunk(counters, chunk, (unsigned int)(N * splits))) return;
  // ---- finalize this chunk (all samples) in fp64
  const int C = C8 * 8;
  const double M = (double)V * (batch_mode ? N : 1);
  for (int nn = 0; nn < (batch_mode ? 1 : N); ++nn) {
    const int n0 = batch_mode ? 0 : nn, n1 = batch_mode ? N : nn + 1;
    const double s1 = reduce_partials_one(partial, C8, chunk, splits, n0, n1, threadIdx.x & 7);
    const double s2 = reduce_partials_one(partial, C8, chunk, splits, n0, n1, 8 + (threadIdx.x & 7));
    if (threadIdx.x < 8) {
      const double m = s1 / M;
      double var = s2 / M - m * m;
      if (var < 0.0) var = 0.0;
      const float mu = (float)m, rs = (float)(1.0 / sqrt(var + (double)eps));
      for (int n2 = batch_mode ? 0 : nn; n2 < (batch_mode ? N : nn + 1); ++n2) {
        mean[n2 * C + chunk * 8 + threadIdx.x] = mu;
        rstd[n2 * C + chunk * 8 + threadIdx.x] = rs;
      }
    }
  }
}

// one thread per (n, channel): combine splits (and n for batch mode) in fp64.
// mean/rstd are written per (n, c) in both modes so consumers are mode-agnostic.
__global__ void norm_stats_finalize_kernel(const float* partial, int N, int C8,
                                           int splits, long long V, int batch_mode, float eps,
                                           float* mean, float* rstd) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = C8 * 8;
  if (idx >= N * C) return;
  const int n = idx / C, c = idx % C, chunk = c >> 3, j = c & 7;
  double s = 0.0, q = 0.0;
  const int n0 = batch_mode ? 0 : n, n1 = batch_mode ? N : n + 1;
  for (int nn = n0; nn < n1; ++nn) {
    const float* p = partial + (long long)(nn * C8 + chunk) * splits * 16;
    for (int sp = 0; sp < splits; ++sp) {
      s += (double)p[sp * 16 + j];
      q += (double)p[sp * 16 + 8 + j];
    }
  }
  const double M = (double)V * (batch_mode ? N : 1);
  const double mu = s / M;
  double var = q / M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[idx] = (float)mu;
  rstd[idx] = (float)(1.0 / sqrt(var + (double)eps));
}

// one CTA per (channel chunk, normalisation group): fp64 reduction of the `splits` partial slots of both sums at
// once -- thread (v, j) walks slots j, j+16, ... of value v with four independent loads in flight (the launch is
// pure latency: ~150 slots of 64 B per group; one serial chain per sample and sum took 10 us)
// (the partials may come from norm_stats_partial_kernel or from the tcgen05 conv epilogue)
__global__ void __launch_bounds__(kThreads)
norm_stats_finalize_chunk_kernel(const float* partial, int N, int C8, int splits, long long V, int batch_mode,
                                 float eps, float* mean, float* rstd) {
  pdl_trigger();
  pdl_wait();
  __shared__ double red[16][17];
  const int chunk = blockIdx.x;
  const int C = C8 * 8;
  const double M = (double)V * (batch_mode ? N : 1);
  const int n0 = batch_mode ? 0 : (int)blockIdx.y, n1 = batch_mode ? N : (int)blockIdx.y + 1;
  const int v = threadIdx.x & 15, j = threadIdx.x >> 4;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int nn = n0; nn < n1; ++nn) {      // fixed order -> deterministic
    const float* p = partial + (long long)(nn * C8 + chunk) * splits * 16 + v;
    int sp = j;
    for (; sp + 48 < splits; sp += 64) {
      const float x0 = __ldcg(p + sp * 16), x1 = __ldcg(p + (sp + 16) * 16), x2 = __ldcg(p + (sp + 32) * 16),
                  x3 = __ldcg(p + (sp + 48) * 16);
      a0 += (double)x0; a1 += (double)x1; a2 += (double)x2; a3 += (double)x3;
    }
    for (; sp < splits; sp += 16) a0 += (double)__ldcg(p + sp * 16);
  }
  red[v][j] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.x < 8) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      s1 += red[threadIdx.x][jj];
      s2 += red[8 + threadIdx.x][jj];
    }
    const double m = s1 / M;
    double var = s2 / M - m * m;
    if (var < 0.0) var = 0.0;
    const float mu = (float)m, rs = (float)(1.0 / sqrt(var + (double)eps));
    for (int n2 = n0; n2 < n1; ++n2) {
      mean[n2 * C + chunk * 8 + threadIdx.x] = mu;
      rstd[n2 * C + chunk * 8 + threadIdx.x] = rs;
    }
  }
}

// one CTA per channel chunk: norm-backward sums / dgamma / dbeta from `splits` partial slots
// (produced by norm_bwd_partial_kernel or by the tcgen05 dgrad epilogue)
__global__ void __launch_bounds__(kThreads)
norm_bwd_finalize_chunk_kernel(const float* partial, int N, int C8, int Creal, int splits, int batch_mode,
                               float* sums, float* dgamma, float* dbeta) {
  pdl_trigger();
  pdl_wait();
  norm_bwd_finalize_tail(partial, C8, blockIdx.x, splits, N, batch_mode, Creal, sums, dgamma, dbeta);
}

// ---------------------------------------------------------------- forward apply
// out = relu?(gamma*(y-mean)*rstd + beta) (+ residual), written as split 16-bit planes.
// RES: 0 none, 1 fp32 view, 2 split-plane view (dtype ODT)
template <int RES, int ODT>
__global__ void __launch_bounds__(kThreads, 4)
norm_apply_kernel(const float* y, long long y_ns, int C8, long long V,
                  const float* mean, const float* rstd,
                  const float* gamma, const float* beta, int relu,
                  const float* res_f32, const uint16_t* res_hi,
                  const uint16_t* res_lo, long long res_ns,
                  uint16_t* out_hi, uint16_t* out_lo, long long out_ns,
                  const float* partial, int splits, int N, int batch_mode, float eps,
                  float* mean_w, float* rstd_w, uint16_t* ws_hi,
                  uint16_t* ws_lo, long long ws_ns, int Wd, int rev) {
  pdl_trigger();
  pdl_wait();
  // rev: walk the slabs and the voxels of a slab in the REVERSE of the order in which the preceding
  // statistics pass (or conv epilogue) touched them, so that this pass starts on what is still in L2
  const bool rv = rev && splits >= 0;
  const int chunk = rv ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int n = rv ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int bx = rv ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int C = C8 * 8;
  float mu[8], rs[8], ga[8], be[8];
  if (splits < 0) {
    // SMALL mode (one CTA owns the whole (n, chunk) slab, V <= 4096, per-instance statistics): the
    // statistics pass runs right here -- the slab (<= 128 KB) is then re-read from L1/L2 by the
    // apply loop below, and the separate tta_norm_stats launch disappears
    float acc[16], tot[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const float* ys = y + (long long)n * y_ns + (long long)chunk * V * 8;
    // gridDim.x > 1: a thread-block CLUSTER of gridDim.x CTAs owns the slab; every CTA reduces the
    // voxels it will also apply (same stride as the apply loop -> its re-read hits L1/L2) and the
    // per-CTA totals meet through distributed shared memory
#pragma unroll 4
    for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V; v += (long long)gridDim.x * kThreads) {
      float x[8];
      load_f32x8(ys + v * 8, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += x[i];
        acc[8 + i] = fmaf(x[i], x[i], acc[8 + i]);
      }
    }
    block_reduce_bcast<16>(acc, tot);
    if (gridDim.x > 1) cluster_sum16(tot);
    const double M = (double)V;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double m = (double)tot[i] / M;
      double var = (double)tot[8 + i] / M - m * m;
      if (var < 0.0) var = 0.0;
      mu[i] = (float)m;
      rs[i] = (float)(1.0 / sqrt(var + (double)eps));
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {  // keep them for the backward pass
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mean_w[n * C + chunk * 8 + i] = mu[i];
        rstd_w[n * C + chunk * 8 + i] = rs[i];
      }
    }
  } else if (partial != nullptr) {
    // statistics finalize fused here: every block reduces the per-block partial sums itself
    double tot[16];
    reduce_partials(partial, C8, chunk, splits, batch_mode ? 0 : n, batch_mode ? N : n + 1, tot);
    const double M = (double)V * (batch_mode ? N : 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double m = tot[i] / M;
      double var = tot[8 + i] / M - m * m;
      if (var < 0.0) var = 0.0;
      mu[i] = (float)m;
      rs[i] = (float)(1.0 / sqrt(var + (double)eps));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // keep them for the backward pass
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mean_w[n * C + chunk * 8 + i] = mu[i];
        rstd_w[n * C + chunk * 8 + i] = rs[i];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mu[i] = mean[n * C + chunk * 8 + i];
      rs[i] = rstd[n * C + chunk * 8 + i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ga[i] = gamma[chunk * 8 + i];
    be[i] = beta[chunk * 8 + i];
  }
  const long long slab = (long long)chunk * V * 8;
  const float* yb = y + (long long)n * y_ns + slab;
  const long long ob = (long long)n * out_ns + slab;
  const long long rb = (long long)n * res_ns + slab;
  const long long vstride = (long long)gridDim.x * kThreads, vfirst = (long long)bx * kThreads + threadIdx.x;
  const long long vlast_it = vfirst < V ? (V - 1 - vfirst) / vstride : -1;
  long long v = rv ? vfirst + vlast_it * vstride : vfirst;
  const long long vstep = rv ? -vstride : vstride;
#pragma unroll 2
  for (long long it = 0; it <= vlast_it; ++it, v += vstep) {
    float x[8];
    load_f32x8(yb + v * 8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float z = (x[i] - mu[i]) * rs[i];
      z = fmaf(z, ga[i], be[i]);
      x[i] = relu ? fmaxf(z, 0.f) : z;
    }
    if (RES == 1) {
      float r[8];
      load_f32x8(res_f32 + rb + v * 8, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] += r[i];
    } else if (RES == 2) {
      float r[8];
      load_split8<ODT>(res_hi, res_lo, rb + v * 8, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] += r[i];
    }
    store_split8<ODT>(out_hi, out_lo, ob + v * 8, x);
    // second copy in the w-parity-split layout for a stride-2 tcgen05 consumer
    if (ws_hi) store_split8<ODT>(ws_hi, ws_lo, (long long)n * ws_ns + slab + wsplit_index(v, Wd) * 8, x);
  }
}

// incoming gradient of a norm backward: fp32 chunks, or ONE loss-scaled fp16 plane (a dgrad that ran with
// tta_conv_tc flags bit 16) -- `g` is the base pointer, `off` the element offset of the voxel-chunk
__device__ __forceinline__ void load_grad8(const float* g, long long off, bool f16, float (&o)[8]) {
  if (f16) {
    const U16x8 h = *reinterpret_cast<const U16x8*>(reinterpret_cast<const uint16_t*>(g) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = u16_to_f32<TTA_F16>(h.v[i]);
  } else {
    load_f32x8(g + off, o);
  }
}

// ---------------------------------------------------------------- backward reductions
// partial[..][0..7] = sum dz, [8..15] = sum dz*xhat     (dz = (g0+g1) * [z>0])
__global__ void __launch_bounds__(kThreads, 4)
norm_bwd_partial_kernel(const float* g0, long long g0_ns,
                        const float* g1, long long g1_ns,
                        const float* y, long long y_ns, int C8, long long V,
                        const float* mean, const float* rstd,
                        const float* gamma, const float* beta, int relu_flags,
                        int splits, float* partial, unsigned int* counters, int N,
                        int batch_mode, int Creal, float* sums, float* dgamma,
                        float* dbeta, int acc_dgb, int rev) {
  pdl_trigger();
  pdl_wait();
  const int relu = relu_flags & 1;
  const bool g0h = (relu_flags & 2) != 0, g1h = (relu_flags & 4) != 0;
  const int split = rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int chunk = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int n = rev ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int C = C8 * 8;
  float mu[8], rs[8], ga[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean[n * C + chunk * 8 + i];
    rs[i] = rstd[n * C + chunk * 8 + i];
    ga[i] = gamma[chunk * 8 + i];
    be[i] = beta[chunk * 8 + i];
  }
  const long long slab = (long long)chunk * V * 8;
  const float* yb = y + (long long)n * y_ns + slab;
  const long long g0o = (long long)n * g0_ns + slab, g1o = (long long)n * g1_ns + slab;
  const long long per = (V + splits - 1) / splits;
  const long long v0 = (long long)split * per;
  const long long v1 = min(V, v0 + per);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 2
  for (long long v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float x[8], g[8];
    load_f32x8(yb + v * 8, x);
    load_grad8(g0, g0o + v * 8, g0h, g);
    if (g1) {
      float h[8];
      load_grad8(g1, g1o + v * 8, g1h, h);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += h[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (x[i] - mu[i]) * rs[i];
      const float z = fmaf(xh, ga[i], be[i]);
      const float dz = (relu && !(z > 0.f)) ? 0.f : g[i];
      acc[i] += dz;
      acc[8 + i] = fmaf(dz, xh, acc[8 + i]);
    }
  }
  block_reduce_store<16>(acc, partial + ((long long)(n * C8 + chunk) * splits + split) * 16);
  if (counters == nullptr) return;
  if (!last_block_of_chunk(counters, chunk, (unsigned int)(N * splits))) return;
  norm_bwd_finalize_tail(partial, C8, chunk, splits, N, batch_mode, Creal, sums, dgamma, dbeta, acc_dgb);
}

// sums[(n*C + c)*2 + {0,1}] = {sum dz, sum dz*xhat} over the normalisation group (per n for IN,
// over all n for BN, broadcast to every n);  dgamma[c], dbeta[c] = sums over n (IN) / the group (BN).
__global__ void norm_bwd_finalize_kernel(const float* partial, int N, int C8, int Creal,
                                         int splits, int batch_mode, float* sums,
                                         float* dgamma, float* dbeta) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = C8 * 8;
  if (c >= C) return;
  const int chunk = c >> 3, j = c & 7;
  double t1 = 0.0, t2 = 0.0;
  for (int n = 0; n < N; ++n) {
    const float* p = partial + (long long)(n * C8 + chunk) * splits * 16;
    double s1 = 0.0, s2 = 0.0;
    for (int sp = 0; sp < splits; ++sp) {
      s1 += (double)p[sp * 16 + j];
      s2 += (double)p[sp * 16 + 8 + j];
    }
    t1 += s1;
    t2 += s2;
    if (!batch_mode) {
      sums[(n * C + c) * 2 + 0] = (float)s1;
      sums[(n * C + c) * 2 + 1] = (float)s2;
    }
  }
  if (batch_mode) {
    for (int n = 0; n < N; ++n) {
      sums[(n * C + c) * 2 + 0] = (float)t1;
      sums[(n * C + c) * 2 + 1] = (float)t2;
    }
  }
  if (c < Creal) {
    dgamma[c] = (float)t2;
    dbeta[c] = (float)t1;
  }
}

// dy = gamma*rstd*(dz - S1/M - xhat*S2/M) -> split planes (ODT);  optionally also the summed
// incoming gradient itself as split planes (aux), which feeds the shortcut conv's dgrad.
template <int ODT>
__global__ void __launch_bounds__(kThreads, 4)
norm_bwd_apply_kernel(const float* g0, long long g0_ns, const float* g1,
                      long long g1_ns, const float* y, long long y_ns, int C8,
                      long long V, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, int relu_flags,
                      const float* sums, float inv_m, uint16_t* dy_hi,
                      uint16_t* dy_lo, long long dy_ns, uint16_t* aux_hi,
                      uint16_t* aux_lo, long long aux_ns, const float* partial,
                      int splits, int N, int batch_mode, int Creal, float* dgamma,
                      float* dbeta, int dy_wsplit_w, float* small_partial, unsigned int* small_counters,
                      float* sums_w, int rev) {
  pdl_trigger();
  pdl_wait();
  const int relu = relu_flags & 1;
  const bool g0h = (relu_flags & 2) != 0, g1h = (relu_flags & 4) != 0;
  // rev: reverse of the reduction pass's order -> the tail of g / y it just streamed is still in L2
  const bool rv = rev && splits >= 0;
  const int chunk = rv ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int n = rv ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int bx = rv ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int C = C8 * 8;
  float mu[8], rs[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = chunk * 8 + i;
    mu[i] = mean[n * C + c];
    rs[i] = rstd[n * C + c];
    ga[i] = gamma[c];
    be[i] = beta[c];
  }
  if (splits < 0) {
    // SMALL mode (one CTA owns the whole (n, chunk) slab): the reduction pass runs here, the slab is
    // re-read from L1/L2 below; the last CTA of the chunk (over n) finalizes dgamma / dbeta
    float acc[16], tot[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const long long sl = (long long)chunk * V * 8;
    const float* ys = y + (long long)n * y_ns + sl;
    const long long gs0 = (long long)n * g0_ns + sl, gs1 = (long long)n * g1_ns + sl;
#pragma unroll 2
    for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V; v += (long long)gridDim.x * kThreads) {
      float x[8], g[8];
      load_f32x8(ys + v * 8, x);
      load_grad8(g0, gs0 + v * 8, g0h, g);
      if (g1) {
        float h[8];
        load_grad8(g1, gs1 + v * 8, g1h, h);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += h[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (x[i] - mu[i]) * rs[i];
        const float z = fmaf(xh, ga[i], be[i]);
        const float dz = (relu && !(z > 0.f)) ? 0.f : g[i];
        acc[i] += dz;
        acc[8 + i] = fmaf(dz, xh, acc[8 + i]);
      }
    }
    if (gridDim.x == 1) {
      block_reduce_bcast<16>(acc, tot, small_partial + (long long)(n * C8 + chunk) * 16);
    } else {
      // cluster of gridDim.x CTAs per slab: totals through distributed shared memory; CTA 0 of the
      // cluster publishes them for the cross-sample dgamma / dbeta finalize
      block_reduce_bcast<16>(acc, tot);
      cluster_sum16(tot);
      if (blockIdx.x == 0 && threadIdx.x < 16) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) t = threadIdx.x == i ? tot[i] : t;
        small_partial[(long long)(n * C8 + chunk) * 16 + threadIdx.x] = t;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      m1[i] = tot[i] * inv_m;
      m2[i] = tot[8 + i] * inv_m;
    }
    if (blockIdx.x == 0 && last_block_of_chunk(small_counters, chunk, (unsigned int)N))
      norm_bwd_finalize_tail(small_partial, C8, chunk, 1, N, 0, Creal, sums_w, dgamma, dbeta);
  } else if (partial != nullptr) {
    // reduction finalize fused here (no separate kernel): S1 = sum dz, S2 = sum dz*xhat
    double tot[16];
    reduce_partials(partial, C8, chunk, splits, batch_mode ? 0 : n, batch_mode ? N : n + 1, tot);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      m1[i] = (float)tot[i] * inv_m;
      m2[i] = (float)tot[8 + i] * inv_m;
    }
    if (blockIdx.x == 0 && n == 0) {  // affine gradients: sums over ALL samples
      if (!batch_mode && N > 1) reduce_partials(partial, C8, chunk, splits, 0, N, tot);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = chunk * 8 + i;
          if (c < Creal) {
            dbeta[c] = (float)tot[i];
            dgamma[c] = (float)tot[8 + i];
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = chunk * 8 + i;
      m1[i] = sums[(n * C + c) * 2 + 0] * inv_m;
      m2[i] = sums[(n * C + c) * 2 + 1] * inv_m;
    }
  }
  const long long slab = (long long)chunk * V * 8;
  const float* yb = y + (long long)n * y_ns + slab;
  const long long g0o = (long long)n * g0_ns + slab, g1o = (long long)n * g1_ns + slab;
  const long long ob = (long long)n * dy_ns + slab;
  const long long ab = (long long)n * aux_ns + slab;
  const long long vstride = (long long)gridDim.x * kThreads, vfirst = (long long)bx * kThreads + threadIdx.x;
  const long long vlast_it = vfirst < V ? (V - 1 - vfirst) / vstride : -1;
  long long v = rv ? vfirst + vlast_it * vstride : vfirst;
  const long long vstep = rv ? -vstride : vstride;
#pragma unroll 2
  for (long long it = 0; it <= vlast_it; ++it, v += vstep) {
    float x[8], g[8];
    load_f32x8(yb + v * 8, x);
    load_grad8(g0, g0o + v * 8, g0h, g);
    if (g1) {
      float h[8];
      load_grad8(g1, g1o + v * 8, g1h, h);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += h[i];
    }
    if (aux_hi) store_split8<ODT>(aux_hi, aux_lo, ab + v * 8, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (x[i] - mu[i]) * rs[i];
      const float z = fmaf(xh, ga[i], be[i]);
      const float dz = (relu && !(z > 0.f)) ? 0.f : g[i];
      x[i] = ga[i] * rs[i] * (dz - m1[i] - xh * m2[i]);
    }
    // dy_wsplit_w = W > 0: the dgrad that consumes dy is a stride-2 tcgen05 conv (w-parity-split)
    store_split8<ODT>(dy_hi, dy_lo, ob + (dy_wsplit_w > 0 ? wsplit_index(v, dy_wsplit_w) : v) * 8, x);
  }
}


// Compact-layout variant for the fused full-resolution head (<= 4 channels): dz is the MASKED gradient as four
// fp16 values per voxel (tta_head_fused_bwd, y_cpv = 4), y the conv result as four floats per voxel
// (tta_conv_tc flags bit 14); dy leaves in the 8-channel chunk layout the dgrad conv's TMA boxes read.
template <int ODT>
__global__ void __launch_bounds__(kThreads, 4)
norm_bwd_apply_c4_kernel(const uint16_t* dz, long long dz_ns, const float* y, long long y_ns, long long V,
                         const float* mean, const float* rstd, const float* gamma, const float* beta,
                         const float* sums, float inv_m, int Creal, uint16_t* dy_hi, uint16_t* dy_lo, long long dy_ns,
                         int dy_wsplit_w, int dy_c4) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  float k0[4], k1[4], k2[4], mu[4], rs[4];   // dy = k0*dz - k1 - xhat*k2
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const bool live = c < Creal;
    mu[c] = live ? mean[n * 8 + c] : 0.f;
    rs[c] = live ? rstd[n * 8 + c] : 0.f;
    const float gr = live ? gamma[c] * rs[c] : 0.f;
    k0[c] = gr;
    k1[c] = live ? gr * sums[(n * 8 + c) * 2 + 0] * inv_m : 0.f;
    k2[c] = live ? gr * sums[(n * 8 + c) * 2 + 1] * inv_m : 0.f;
  }
  (void)beta;
  const uint16_t* dzb = dz + (long long)n * dz_ns;
  const float* yb = y + (long long)n * y_ns;
  const long long ob = (long long)n * dy_ns;
#pragma unroll 2
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V; v += (long long)gridDim.x * kThreads) {
    const uint2 pk = *reinterpret_cast<const uint2*>(dzb + v * 4);
    const float4 yv = *reinterpret_cast<const float4*>(yb + v * 4);
    const __half2 h01 = *reinterpret_cast<const __half2*>(&pk.x), h23 = *reinterpret_cast<const __half2*>(&pk.y);
    const float2 d01 = __half22float2(h01), d23 = __half22float2(h23);
    const float dzv[4] = {d01.x, d01.y, d23.x, d23.y}, xv[4] = {yv.x, yv.y, yv.z, yv.w};
    float x[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float xh = (xv[c] - mu[c]) * rs[c];
      x[c] = fmaf(k0[c], dzv[c], -k1[c]) - xh * k2[c];
      x[4 + c] = 0.f;
    }
    if (dy_c4) {   // compact gradient planes [N][V][4] for a GEOM_S2C4 dgrad (dy_ns in 16-bit elements)
      uint16_t h[4], l[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        h[c] = f32_to_u16<ODT>(x[c]);
        l[c] = ODT == TTA_F16_HI ? (uint16_t)0 : f32_to_u16<ODT>(x[c] - u16_to_f32<ODT>(h[c]));
      }
      *reinterpret_cast<uint2*>(dy_hi + ob + v * 4) =
          make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
      if (ODT != TTA_F16_HI)
        *reinterpret_cast<uint2*>(dy_lo + ob + v * 4) =
            make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
      continue;
    }
    store_split8<ODT>(dy_hi, dy_lo, ob + (dy_wsplit_w > 0 ? wsplit_index(v, dy_wsplit_w) : v) * 8, x);
  }
}

// plain fp32 (sum of up to two views) -> split planes; used where a gradient feeds a conv
// directly without a norm in between.
template <int ODT>
__global__ void __launch_bounds__(kThreads)
split_f32_kernel(const float* g0, long long g0_ns, const float* g1,
                 long long g1_ns, int C8, long long V, uint16_t* hi,
                 uint16_t* lo, long long o_ns) {
  pdl_trigger();
  pdl_wait();
  const int chunk = blockIdx.y, n = blockIdx.z;
  const long long slab = (long long)chunk * V * 8;
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V;
       v += (long long)gridDim.x * kThreads) {
    float g[8];
    load_f32x8(g0 + (long long)n * g0_ns + slab + v * 8, g);
    if (g1) {
      float h[8];
      load_f32x8(g1 + (long long)n * g1_ns + slab + v * 8, h);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += h[i];
    }
    store_split8<ODT>(hi, lo, (long long)n * o_ns + slab + v * 8, g);
  }
}

// apply passes walk the data in reverse of the pass before them (L2 reuse); TTA_NORM_FORWARD_ORDER=1
// restores the forward order (A/B)
// TTA_NORM_ORDER (A/B, ms per step on one box): 0 = every pass forward (2.321), 1 (default) = apply passes
// reversed (2.293), 2 = reduction passes reversed (they follow a conv that wrote front to back), the apply
// pass after a reduction forward again, the apply pass that follows a conv directly reversed (2.305)
static inline int norm_order() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TTA_NORM_ORDER");
    v = e ? atoi(e) : 1;
  }
  return v;
}
static inline int norm_rev_reduce() { return norm_order() == 2; }
static inline int norm_rev_apply(bool follows_reduction) {
  return norm_order() == 1 || (norm_order() == 2 && !follows_reduction);
}

static inline int pick_splits(int N, int C8, long long V) {
  // enough CTAs for two waves of 4 resident CTAs on 148 SMs, at least 512 voxel-chunks (two per thread) per CTA: the
  // small 8^3..32^3 layers are latency bound and want every SM streaming
  long long want = (8LL * 148 + (long long)N * C8 - 1) / ((long long)N * C8);
  long long maxs = (V + 511) / 512;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}
static inline int pick_xblocks(int N, int C8, long long V) {
  long long full = (V + kThreads - 1) / kThreads;
  long long want = (8LL * 148 + (long long)N * C8 - 1) / ((long long)N * C8);
  if (want > full) want = full;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace tta

using namespace tta;

extern "C" {

// Workspace (floats) needed by tta_norm_stats / tta_norm_bwd_reduce for a given shape.
// workspace = [1024 block counters (zero-initialised once, self-resetting)][per-block partials]
long long tta_norm_workspace_floats(int N, int C8, long long V) {
  return 1024 + (long long)N * C8 * pick_splits(N, C8, V) * 16;
}

// finalize = 0: only the per-block partial sums are produced; tta_norm_apply(partial = workspace)
// then finalizes them in its prologue (one launch less per layer).
int tta_norm_stats(const float* y, long long y_ns, int N, int C8, long long V, int batch_mode,
                   float eps, float* mean, float* rstd, float* workspace, int finalize,
                   cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_stats(y, y_ns, N, C8, V, batch_mode, eps, mean, rstd, workspace, finalize, s_));
  TTA_REQUIRE(y && workspace && (!finalize || (mean && rstd)), "tta_norm_stats: null pointer");
  TTA_REQUIRE(N > 0 && C8 > 0 && V > 0, "tta_norm_stats: empty shape N=%d C8=%d V=%lld", N, C8, V);
  TTA_REQUIRE(C8 <= 1024, "tta_norm_stats: more than 8192 channels unsupported");
  const int splits = pick_splits(N, C8, V);
  // finalize = 1: the last block of every chunk turns the partial sums into mean/rstd (single pass)
  tta_launch(norm_stats_partial_kernel, dim3(splits, C8, N), kThreads, 0, stream, tta_pdl_family(2), 
      y, y_ns, C8, V, splits, workspace + 1024, finalize ? reinterpret_cast<unsigned int*>(workspace) : nullptr, N,
      batch_mode, eps, mean, rstd, norm_rev_reduce());
  return tta_check_launch("tta_norm_stats");
}

// mean/rstd [N][C8*8] from partial sums in `workspace` laid out for `splits` slots per (n, chunk)
// (tta_norm_stats with finalize = 0, or the fused statistics of tta_conv_tc with splits = its grid)
int tta_norm_stats_finalize(const float* workspace, int N, int C8, int splits, long long V, int batch_mode,
                            float eps, float* mean, float* rstd, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_stats_finalize(workspace, N, C8, splits, V, batch_mode, eps, mean, rstd, s_));
  TTA_REQUIRE(workspace && mean && rstd && splits > 0, "tta_norm_stats_finalize: bad argument");
  tta_launch(norm_stats_finalize_chunk_kernel, dim3(C8, batch_mode ? 1 : N), kThreads, 0, stream, tta_pdl_family(2),
             workspace + 1024, N, C8, splits, V, batch_mode, eps, mean, rstd);
  return tta_check_launch("tta_norm_stats_finalize");
}

int tta_norm_apply(const float* y, long long y_ns, int N, int C8, long long V, const float* mean,
                   const float* rstd, const float* gamma, const float* beta, int relu,
                   int res_kind, const void* res_a, const void* res_b, long long res_ns,
                   uint16_t* out_hi, uint16_t* out_lo, long long out_ns, int out_dtype,
                   const float* partial, int partial_splits, int batch_mode, float eps, uint16_t* ws_hi,
                   uint16_t* ws_lo, long long ws_ns, int W, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_apply(y, y_ns, N, C8, V, mean, rstd, gamma, beta, relu, res_kind, res_a, res_b, res_ns, out_hi, out_lo, out_ns, out_dtype, partial, partial_splits, batch_mode, eps, ws_hi, ws_lo, ws_ns, W, s_));
  TTA_REQUIRE(y && mean && rstd && gamma && beta && out_hi && out_lo, "tta_norm_apply: null pointer");
  TTA_REQUIRE(!ws_hi || (ws_lo && W > 0 && W % 2 == 0 && V % W == 0),
              "tta_norm_apply: w-parity-split copy needs an even row length W=%d dividing V", W);
  // partial_splits > 0: the partial sums were produced by the conv epilogue (one slot per conv CTA)
  const int splits = partial_splits > 0 ? partial_splits : pick_splits(N, C8, V);
  TTA_REQUIRE(res_kind >= 0 && res_kind <= 2, "tta_norm_apply: res_kind %d", res_kind);
  TTA_REQUIRE(out_dtype == TTA_F16 || out_dtype == TTA_BF16, "tta_norm_apply: bad dtype");
  const dim3 grid(pick_xblocks(N, C8, V), C8, N);
#define LAUNCH(RES, DT)                                                                        \
  tta_launch(norm_apply_kernel<RES, DT>, grid, kThreads, 0, stream, tta_pdl_family(2),                                    \
      y, y_ns, C8, V, mean, rstd, gamma, beta, relu, (const float*)res_a,                      \
      (const uint16_t*)res_a, (const uint16_t*)res_b, res_ns, out_hi, out_lo, out_ns,                       \
      partial ? partial + 1024 : nullptr, splits, N,                                                        \
      batch_mode, eps, const_cast<float*>(mean), const_cast<float*>(rstd), ws_hi, ws_lo, ws_ns, W,                \
      norm_rev_apply(partial_splits <= 0))
  if (out_dtype == TTA_F16) {
    if (res_kind == 0) LAUNCH(0, TTA_F16); else if (res_kind == 1) LAUNCH(1, TTA_F16); else LAUNCH(2, TTA_F16);
  } else {
    if (res_kind == 0) LAUNCH(0, TTA_BF16); else if (res_kind == 1) LAUNCH(1, TTA_BF16); else LAUNCH(2, TTA_BF16);
  }
#undef LAUNCH
  return tta_check_launch("tta_norm_apply");
}

// sums [N][C8*8][2], dgamma/dbeta [Creal] from partial sums laid out [N][C8][splits][16] (NO
// 1024-float counter prefix: `partial` points at the slots themselves)
int tta_norm_bwd_finalize(const float* partial, int N, int C8, int Creal, int splits, int batch_mode, float* sums,
                          float* dgamma, float* dbeta, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_bwd_finalize(partial, N, C8, Creal, splits, batch_mode, sums, dgamma, dbeta, s_));
  TTA_REQUIRE(partial && sums && dgamma && dbeta && splits > 0, "tta_norm_bwd_finalize: bad argument");
  tta_launch(norm_bwd_finalize_chunk_kernel, C8, kThreads, 0, stream, tta_pdl_family(2), partial, N, C8, Creal, splits,
             batch_mode, sums, dgamma, dbeta);
  return tta_check_launch("tta_norm_bwd_finalize");
}

int tta_norm_bwd_reduce(const float* g0, long long g0_ns, const float* g1, long long g1_ns,
                        const float* y, long long y_ns, int N, int C8, int Creal, long long V,
                        const float* mean, const float* rstd, const float* gamma,
                        const float* beta, int relu, int batch_mode, float* sums, float* dgamma,
                        float* dbeta, float* workspace, int finalize, int accumulate_dgb, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_bwd_reduce(g0, g0_ns, g1, g1_ns, y, y_ns, N, C8, Creal, V, mean, rstd, gamma, beta, relu, batch_mode, sums, dgamma, dbeta, workspace, finalize, accumulate_dgb, s_));
  TTA_REQUIRE(g0 && y && mean && rstd && gamma && beta && workspace &&
                  (!finalize || (sums && dgamma && dbeta)),
              "tta_norm_bwd_reduce: null pointer");
  TTA_REQUIRE(C8 <= 1024, "tta_norm_bwd_reduce: more than 8192 channels unsupported");
  const int splits = pick_splits(N, C8, V);
  tta_launch(norm_bwd_partial_kernel, dim3(splits, C8, N), kThreads, 0, stream, tta_pdl_family(2), 
      g0, g0_ns, g1, g1_ns, y, y_ns, C8, V, mean, rstd, gamma, beta, relu, splits, workspace + 1024,
      finalize ? reinterpret_cast<unsigned int*>(workspace) : nullptr, N, batch_mode, Creal, sums, dgamma, dbeta,
      accumulate_dgb, norm_rev_reduce());
  return tta_check_launch("tta_norm_bwd_reduce");
}

int tta_norm_bwd_apply(const float* g0, long long g0_ns, const float* g1, long long g1_ns,
                       const float* y, long long y_ns, int N, int C8, long long V,
                       const float* mean, const float* rstd, const float* gamma, const float* beta,
                       int relu, int batch_mode, const float* sums, uint16_t* dy_hi,
                       uint16_t* dy_lo, long long dy_ns, uint16_t* aux_hi, uint16_t* aux_lo,
                       long long aux_ns, int out_dtype, const float* partial, int Creal, float* dgamma,
                       float* dbeta, int dy_wsplit_w, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_bwd_apply(g0, g0_ns, g1, g1_ns, y, y_ns, N, C8, V, mean, rstd, gamma, beta, relu, batch_mode, sums, dy_hi, dy_lo, dy_ns, aux_hi, aux_lo, aux_ns, out_dtype, partial, Creal, dgamma, dbeta, dy_wsplit_w, s_));
  TTA_REQUIRE(dy_wsplit_w == 0 || (dy_wsplit_w > 0 && dy_wsplit_w % 2 == 0 && V % dy_wsplit_w == 0),
              "tta_norm_bwd_apply: w-parity-split dy needs an even row length dividing V (got %d)", dy_wsplit_w);
  TTA_REQUIRE(g0 && y && mean && rstd && gamma && beta && (sums || (partial && dgamma && dbeta)) && dy_hi &&
                  (dy_lo || out_dtype == TTA_F16_HI),
              "tta_norm_bwd_apply: null pointer");
  const int splits = pick_splits(N, C8, V);
  const float inv_m = (float)(1.0 / ((double)V * (batch_mode ? N : 1)));
  const dim3 grid(pick_xblocks(N, C8, V), C8, N);
  TTA_REQUIRE(out_dtype >= 0 && out_dtype <= 2, "tta_norm_bwd_apply: bad dtype");
  if (out_dtype == TTA_F16)
    tta_launch(norm_bwd_apply_kernel<TTA_F16>, grid, kThreads, 0, stream, tta_pdl_family(2), 
        g0, g0_ns, g1, g1_ns, y, y_ns, C8, V, mean, rstd, gamma, beta, relu, sums, inv_m, dy_hi,
        dy_lo, dy_ns, aux_hi, aux_lo, aux_ns, partial ? partial + 1024 : nullptr, splits, N, batch_mode, Creal, dgamma,
        dbeta, dy_wsplit_w, nullptr, nullptr, nullptr, norm_rev_apply(true));
  else if (out_dtype == TTA_F16_HI)
    tta_launch(norm_bwd_apply_kernel<TTA_F16_HI>, grid, kThreads, 0, stream, tta_pdl_family(2), 
        g0, g0_ns, g1, g1_ns, y, y_ns, C8, V, mean, rstd, gamma, beta, relu, sums, inv_m, dy_hi,
        dy_lo, dy_ns, aux_hi, aux_lo, aux_ns, partial ? partial + 1024 : nullptr, splits, N, batch_mode, Creal, dgamma,
        dbeta, dy_wsplit_w, nullptr, nullptr, nullptr, norm_rev_apply(true));
  else
    tta_launch(norm_bwd_apply_kernel<TTA_BF16>, grid, kThreads, 0, stream, tta_pdl_family(2), 
        g0, g0_ns, g1, g1_ns, y, y_ns, C8, V, mean, rstd, gamma, beta, relu, sums, inv_m, dy_hi,
        dy_lo, dy_ns, aux_hi, aux_lo, aux_ns, partial ? partial + 1024 : nullptr, splits, N, batch_mode, Creal, dgamma,
        dbeta, dy_wsplit_w, nullptr, nullptr, nullptr, norm_rev_apply(true));
  return tta_check_launch("tta_norm_bwd_apply");
}

// ---- small layers (V <= 65536 voxels per instance, per-instance statistics): statistics + apply,
// and reduction + apply, as ONE launch each -- a CTA (V <= 512) or a thread-block cluster of up to 8
// CTAs owns a whole (n, chunk) slab, reduces it (cluster: totals exchanged through distributed shared
// memory), and re-reads it from L1/L2.  The 8^3 .. 32^3 levels are pure launch latency otherwise.
int tta_norm_small_supported(int N, long long V, int batch_mode) {
  return V >= 2 && V <= 65536 && !batch_mode && N >= 1;
}

// CTAs per (n, chunk) slab: 1 up to 512 voxels, else a thread-block cluster (<= 8, the portable
// maximum) with >= 512 voxels per CTA, grown until the launch has ~2 CTAs per SM
static inline int small_cluster(int N, int C8, long long V) {
  int cs = 1;
  while (cs < 8 && V / (cs * 2) >= 512 && (long long)N * C8 * cs < 2 * 148) cs *= 2;
  return cs;
}

int tta_norm_fwd_small(const float* y, long long y_ns, int N, int C8, long long V, float eps, float* mean,
                       float* rstd, const float* gamma, const float* beta, int relu, int res_kind,
                       const void* res_a, const void* res_b, long long res_ns, uint16_t* out_hi,
                       uint16_t* out_lo, long long out_ns, int out_dtype, uint16_t* ws_hi, uint16_t* ws_lo,
                       long long ws_ns, int W, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_fwd_small(y, y_ns, N, C8, V, eps, mean, rstd, gamma, beta, relu, res_kind, res_a, res_b, res_ns, out_hi, out_lo, out_ns, out_dtype, ws_hi, ws_lo, ws_ns, W, s_));
  TTA_REQUIRE(y && mean && rstd && gamma && beta && out_hi && out_lo, "tta_norm_fwd_small: null pointer");
  TTA_REQUIRE(tta_norm_small_supported(N, V, 0), "tta_norm_fwd_small: V=%lld unsupported (2..65536)", V);
  TTA_REQUIRE(res_kind >= 0 && res_kind <= 2, "tta_norm_fwd_small: res_kind %d", res_kind);
  TTA_REQUIRE(out_dtype == TTA_F16 || out_dtype == TTA_BF16, "tta_norm_fwd_small: bad dtype");
  TTA_REQUIRE(!ws_hi || (ws_lo && W > 0 && W % 2 == 0 && V % W == 0), "tta_norm_fwd_small: bad parity-split copy");
  const int cs = small_cluster(N, C8, V);
  const dim3 grid(cs, C8, N);
#define LAUNCH(RES, DT)                                                                              \
  tta_launch_cluster(norm_apply_kernel<RES, DT>, grid, kThreads, 0, stream, cs, y, y_ns, C8, V, mean, \
             rstd, gamma, beta, relu, (const float*)res_a, (const uint16_t*)res_a, (const uint16_t*)res_b, \
             res_ns, out_hi, out_lo, out_ns, (const float*)nullptr, -1, N, 0, eps, mean, rstd, ws_hi, ws_lo, \
             ws_ns, W, 0)
  if (out_dtype == TTA_F16) {
    if (res_kind == 0) LAUNCH(0, TTA_F16); else if (res_kind == 1) LAUNCH(1, TTA_F16); else LAUNCH(2, TTA_F16);
  } else {
    if (res_kind == 0) LAUNCH(0, TTA_BF16); else if (res_kind == 1) LAUNCH(1, TTA_BF16); else LAUNCH(2, TTA_BF16);
  }
#undef LAUNCH
  return tta_check_launch("tta_norm_fwd_small");
}

// workspace: as tta_norm_bwd_reduce ([1024 counters][partials], tta_norm_workspace_floats)
int tta_norm_bwd_small(const float* g0, long long g0_ns, const float* g1, long long g1_ns, const float* y,
                       long long y_ns, int N, int C8, int Creal, long long V, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, int relu, float* sums,
                       float* dgamma, float* dbeta, uint16_t* dy_hi, uint16_t* dy_lo, long long dy_ns,
                       uint16_t* aux_hi, uint16_t* aux_lo, long long aux_ns, int out_dtype, int dy_wsplit_w,
                       float* workspace, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_bwd_small(g0, g0_ns, g1, g1_ns, y, y_ns, N, C8, Creal, V, mean, rstd, gamma, beta, relu, sums, dgamma, dbeta, dy_hi, dy_lo, dy_ns, aux_hi, aux_lo, aux_ns, out_dtype, dy_wsplit_w, workspace, s_));
  TTA_REQUIRE(g0 && y && mean && rstd && gamma && beta && sums && dgamma && dbeta && dy_hi && workspace &&
                  (dy_lo || out_dtype == TTA_F16_HI),
              "tta_norm_bwd_small: null pointer");
  TTA_REQUIRE(tta_norm_small_supported(N, V, 0), "tta_norm_bwd_small: V=%lld unsupported (2..65536)", V);
  TTA_REQUIRE(out_dtype >= 0 && out_dtype <= 2, "tta_norm_bwd_small: bad dtype");
  TTA_REQUIRE(dy_wsplit_w == 0 || (dy_wsplit_w > 0 && dy_wsplit_w % 2 == 0 && V % dy_wsplit_w == 0),
              "tta_norm_bwd_small: bad parity-split row length %d", dy_wsplit_w);
  const float inv_m = (float)(1.0 / (double)V);
  const int cs = small_cluster(N, C8, V);
  const dim3 grid(cs, C8, N);
#define LAUNCH(DT)                                                                                          \
  tta_launch_cluster(norm_bwd_apply_kernel<DT>, grid, kThreads, 0, stream, cs, g0, g0_ns, g1, g1_ns, y, \
             y_ns, C8, V, mean, rstd, gamma, beta, relu, (const float*)nullptr, inv_m, dy_hi, dy_lo, dy_ns, aux_hi, \
             aux_lo, aux_ns, (const float*)nullptr, -1, N, 0, Creal, dgamma, dbeta, dy_wsplit_w, workspace + 1024,  \
             reinterpret_cast<unsigned int*>(workspace), sums, 0)
  if (out_dtype == TTA_F16) LAUNCH(TTA_F16); else if (out_dtype == TTA_F16_HI) LAUNCH(TTA_F16_HI); else LAUNCH(TTA_BF16);
#undef LAUNCH
  return tta_check_launch("tta_norm_bwd_small");
}

// Compact-layout norm backward apply of the fused head (see norm_bwd_apply_c4_kernel): dz fp16 [N][V][4]
// (n stride in 16-bit elements), y fp32 [N][V][4] (n stride in floats), mean/rstd [N][8], sums [N][8][2];
// dy: one chunk per voxel in the usual operand layout.  InstanceNorm and BatchNorm alike (sums / mean / rstd are
// already per (n, c)); inv_m = 1 / (voxels per normalisation group).
int tta_norm_bwd_apply_c4(const uint16_t* dz, long long dz_ns, const float* y, long long y_ns, int N, int Creal,
                          long long V, const float* mean, const float* rstd, const float* gamma, const float* beta,
                          int batch_mode, const float* sums, uint16_t* dy_hi, uint16_t* dy_lo, long long dy_ns,
                          int out_dtype, int dy_wsplit_w, int dy_c4, cudaStream_t stream) {
  TTA_RECORDABLE(tta_norm_bwd_apply_c4(dz, dz_ns, y, y_ns, N, Creal, V, mean, rstd, gamma, beta, batch_mode, sums, dy_hi, dy_lo, dy_ns, out_dtype, dy_wsplit_w, dy_c4, s_));
  TTA_REQUIRE(!(dy_c4 && dy_wsplit_w), "tta_norm_bwd_apply_c4: compact dy is never w-parity-split");
  TTA_REQUIRE(dz && y && mean && rstd && gamma && beta && sums && dy_hi && (dy_lo || out_dtype == TTA_F16_HI),
              "tta_norm_bwd_apply_c4: null pointer");
  TTA_REQUIRE(Creal >= 1 && Creal <= 4 && N > 0 && V > 0, "tta_norm_bwd_apply_c4: bad shape C=%d", Creal);
  TTA_REQUIRE(dy_wsplit_w == 0 || (dy_wsplit_w > 0 && dy_wsplit_w % 2 == 0 && V % dy_wsplit_w == 0),
              "tta_norm_bwd_apply_c4: bad parity-split row length %d", dy_wsplit_w);
  TTA_REQUIRE(out_dtype >= 0 && out_dtype <= 2, "tta_norm_bwd_apply_c4: bad dtype");
  const float inv_m = (float)(1.0 / ((double)V * (batch_mode ? N : 1)));
  long long xb = (V + kThreads - 1) / kThreads;
  const long long want = (8LL * 148 + N - 1) / N;
  if (xb > want) xb = want;
  const dim3 grid((unsigned)xb, N);
#define LAUNCH(DT)                                                                                                   \
  tta_launch(norm_bwd_apply_c4_kernel<DT>, grid, kThreads, 0, stream, tta_pdl_family(2), dz, dz_ns, y, y_ns, V, mean, \
             rstd, gamma, beta, sums, inv_m, Creal, dy_hi, dy_lo, dy_ns, dy_wsplit_w, dy_c4)
  if (out_dtype == TTA_F16) LAUNCH(TTA_F16); else if (out_dtype == TTA_F16_HI) LAUNCH(TTA_F16_HI); else LAUNCH(TTA_BF16);
#undef LAUNCH
  return tta_check_launch("tta_norm_bwd_apply_c4");
}

int tta_split_f32(const float* g0, long long g0_ns, const float* g1, long long g1_ns, int N, int C8,
                  long long V, uint16_t* hi, uint16_t* lo, long long o_ns, int out_dtype,
                  cudaStream_t stream) {
  TTA_RECORDABLE(tta_split_f32(g0, g0_ns, g1, g1_ns, N, C8, V, hi, lo, o_ns, out_dtype, s_));
  TTA_REQUIRE(g0 && hi && (lo || out_dtype == TTA_F16_HI), "tta_split_f32: null pointer");
  const dim3 grid(pick_xblocks(N, C8, V), C8, N);
  if (out_dtype == TTA_F16)
    tta_launch(split_f32_kernel<TTA_F16>, grid, kThreads, 0, stream, tta_pdl_family(2), g0, g0_ns, g1, g1_ns, C8, V, hi, lo, o_ns);
  else if (out_dtype == TTA_F16_HI)
    tta_launch(split_f32_kernel<TTA_F16_HI>, grid, kThreads, 0, stream, tta_pdl_family(2), g0, g0_ns, g1, g1_ns, C8, V, hi, lo, o_ns);
  else
    tta_launch(split_f32_kernel<TTA_BF16>, grid, kThreads, 0, stream, tta_pdl_family(2), g0, g0_ns, g1, g1_ns, C8, V, hi, lo, o_ns);
  return tta_check_launch("tta_split_f32");
}

}  // extern "C"
