// Block-level reduction helpers shared by the norm kernels (tta_norm.cu) and the fused head
// (tta_head.cu): warp-shuffle -> smem -> per-block partial sums in HBM -> fp64 finalize by the last
// block of a chunk (deterministic order, no atomics on the data, self-resetting counters).
// All of them assume 256-thread blocks.
#pragma once
#include "tta_common.cuh"

namespace tta {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- block reduction of 16 values
template <int NV>
__device__ __forceinline__ void block_reduce_store(float (&acc)[NV], float* dst) {
  __shared__ float red[kThreads / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// same reduction, totals broadcast to every thread through shared memory (fp32 in, fp32 out)
template <int NV>
__device__ __forceinline__ void block_reduce_bcast(float (&acc)[NV], float (&tot)[NV], float* gdst = nullptr) {
  __shared__ float red[kThreads / 32][NV];
  __shared__ float tots[NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
    tots[threadIdx.x] = s;
    if (gdst) gdst[threadIdx.x] = s;  // also publish the totals (one value per thread)
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) tot[i] = tots[i];
}

// ---------------------------------------------------------------- cluster reduction (DSMEM)
// tot[16] holds this CTA's block totals (identical in every thread); on return it holds the sum over
// all CTAs of the thread-block cluster, accumulated in rank order in fp64 -> every CTA gets the same
// bits.  Every thread of every CTA of the cluster must call it.
__device__ __forceinline__ void cluster_sum16(float (&tot)[16]) {
  __shared__ float cl_part[16];
  __shared__ float cl_tot[16];
  uint32_t nranks;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nranks));
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t = threadIdx.x == i ? tot[i] : t;
    cl_part[threadIdx.x] = t;
  }
  // release/acquire cluster barrier: every CTA's partials are visible cluster-wide afterwards
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 16) {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&cl_part[threadIdx.x]);
    double a = 0.0;
    for (uint32_t r = 0; r < nranks; ++r) {
      uint32_t remote;
      float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
      a += (double)v;
    }
    cl_tot[threadIdx.x] = (float)a;
  }
  // nobody leaves (or overwrites cl_part) while a peer may still be reading it
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) tot[i] = cl_tot[i];
}

// ---------------------------------------------------------------- fused finalize (consumer prologue)
// Sums the 16 per-block partial values of chunk `chunk` over splits (and over samples n0..n1)
// in fp64 with all 256 threads; result in tot[16] (shared).  Deterministic order.
__device__ __forceinline__ void reduce_partials(const float* partial, int C8, int chunk,
                                                int splits, int n0, int n1, double (&tot)[16]) {
  __shared__ double red[16][17];
  const int v = threadIdx.x & 15, j = threadIdx.x >> 4;  // 16 values x 16 split lanes
  double a = 0.0;
  for (int nn = n0; nn < n1; ++nn) {
    const float* p = partial + (long long)(nn * C8 + chunk) * splits * 16;
    for (int sp = j; sp < splits; sp += 16) a += (double)__ldcg(p + sp * 16 + v);
  }
  red[v][j] = a;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double t = 0.0;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) t += red[i][jj];
    tot[i] = t;
  }
  __syncthreads();
}
// same, but every thread only needs value `which` (0..15): avoids dynamic register indexing
__device__ __forceinline__ double reduce_partials_one(const float* partial, int C8, int chunk,
                                                      int splits, int n0, int n1, int which) {
  __shared__ double red1[16][17];
  const int v = threadIdx.x & 15, j = threadIdx.x >> 4;
  double a = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int nn = n0; nn < n1; ++nn) {
    // four independent loads in flight per thread: these tails are pure L2 latency (a serial chain of ~10
    // dependent 600 ns loads per call otherwise)
    const float* p = partial + (long long)(nn * C8 + chunk) * splits * 16 + v;
    int sp = j;
    for (; sp + 48 < splits; sp += 64) {
      const float x0 = __ldcg(p + sp * 16), x1 = __ldcg(p + (sp + 16) * 16), x2 = __ldcg(p + (sp + 32) * 16),
                  x3 = __ldcg(p + (sp + 48) * 16);
      a += (double)x0; a1 += (double)x1; a2 += (double)x2; a3 += (double)x3;
    }
    for (; sp < splits; sp += 16) a += (double)__ldcg(p + sp * 16);
  }
  a = (a + a1) + (a2 + a3);
  red1[v][j] = a;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) t += red1[which][jj];
  __syncthreads();
  return t;
}

// Single-pass reduction: after publishing its partial sums a block bumps the per-chunk counter;
// the block that observes the final count (all N*splits blocks of this chunk have published)
// finalizes in a fixed order -> deterministic, no second launch, counter self-resets.
__device__ __forceinline__ bool last_block_of_chunk(unsigned int* counters, int chunk, unsigned int total) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(&counters[chunk], 1u);
    s_last = (old == total - 1u) ? 1u : 0u;
    if (s_last) counters[chunk] = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0u;
}


// Tail of the norm-backward reduction, executed by the LAST block of a chunk: sums[(n*C + c)*2 +
// {0,1}] = {sum dz, sum dz*xhat} per normalisation group (per n for InstanceNorm, over all n for
// BatchNorm, broadcast to every n); dgamma/dbeta[c] = the sums over all samples.
__device__ __forceinline__ void norm_bwd_finalize_tail(const float* partial, int C8, int chunk,
                                                       int splits, int N, int batch_mode, int Creal,
                                                       float* sums, float* dgamma,
                                                       float* dbeta, int acc_dgb = 0) {
  const int C = C8 * 8;
  const int which = threadIdx.x & 15;
  const int cc = chunk * 8 + (which & 7);
  const double tall = reduce_partials_one(partial, C8, chunk, splits, 0, N, which);
  if (threadIdx.x < 16) {
    if (cc < Creal) {
      // acc_dgb: this launch covers a subset of the samples (per-sample backward, see engine.py):
      // the affine gradients of the launches add up in launch order (deterministic)
      float* dst = which < 8 ? dbeta + cc : dgamma + cc;
      *dst = acc_dgb ? *dst + (float)tall : (float)tall;
    }
    if (batch_mode)
      for (int nn = 0; nn < N; ++nn) sums[(nn * C + cc) * 2 + (which >> 3)] = (float)tall;
  }
  if (!batch_mode) {
    for (int nn = 0; nn < N; ++nn) {
      const double tn = N == 1 ? tall : reduce_partials_one(partial, C8, chunk, splits, nn, nn + 1, which);
      if (threadIdx.x < 16) sums[(nn * C + cc) * 2 + (which >> 3)] = (float)tn;
    }
  }
}

}  // namespace tta
