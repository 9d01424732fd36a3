// Shared device/host helpers for the B200 TTA kernels (sm_100a only).
//
// Data layout in HBM:
//   * channel-blocked "chunk" layout  [N][C8][D][H][W][8]   (C8 = ceil(C/8), pad channels are 0)
//   * conv RESULTS   are fp32   : one voxel-chunk = 32 B
//   * conv OPERANDS  are split  : two 16-bit planes hi/lo (fp16 in forward, bf16 in backward),
//                                 x ~= hi + lo, one voxel-chunk = 16 B per plane
//   A tensor "view" is (base pointer, n_stride in ELEMENTS, C8, D, H, W): a channel slice of a
//   concat buffer is just a pointer offset with the parent's n_stride -> torch.cat is free.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define TTA_OK 0
#define TTA_ERR_ARG 1
#define TTA_ERR_CUDA 2
#define TTA_ERR_UNSUPPORTED 3

// dtype tags of the 16-bit operand planes
#define TTA_F16 0
#define TTA_BF16 1
#define TTA_F16_HI 2  // single fp16 plane (no lo plane is written or read): scaled gradients in backward
#define TTA_PLAN_SECTIONS 8
extern "C" long long tta_norm_workspace_floats(int N, int C8, long long V);

void tta_set_error(const char* fmt, ...);
int tta_check_launch(const char* what);

// Programmatic dependent launch (PDL), OPT-IN with TTA_PDL=1: every kernel of the step is then
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may be scheduled
// while the previous kernel drains and run their input-independent prologue (barrier init, TMEM
// allocation, parameter loads); pdl_wait() blocks until the previous grid has completed and flushed.
// Rules: (1) no global-memory access other than constants before pdl_wait(); (2) every kernel
// executes pdl_wait() so completion stays transitive along the stream; (3) NO ld.global.nc on
// anything an earlier kernel writes -- the read-only path is only valid for data that is constant
// over the kernel's lifetime, and a PDL kernel is alive while its producer still runs (measured:
// 20-40 % gradient errors with `const T* __restrict__` inputs), hence no __restrict__/__ldg in
// this library.  Measured inside the whole-step CUDA graph (B200, 2x4x128^3): 3.06 ms with PDL
// vs 2.97 ms without -- the graph's kernel-to-kernel latency is already small and early-resident
// dependents cost more than the overlapped prologues save, so the default is OFF.
bool tta_pdl_enabled();
bool tta_pdl_family(int family_bit);  // TTA_PDL_OFF=<mask> switches PDL off per kernel family (debugging)

// ---- launch recorder (csrc/tta_plan.cu): while a tta_plan section is being recorded on the calling thread, every
// launching entry point stores itself (arguments by value) instead of launching; tta_plan_run / tta_step replay
// the stored calls on a stream.  Host pointer arguments (W_host, pointer arrays, segment tables) must outlive the plan.
#include <functional>
bool tta_recording();
void tta_record_push(std::function<int(cudaStream_t)> fn);
#define TTA_RECORDABLE(call_with_s_)                                                   \
  do {                                                                                 \
    if (tta_recording()) {                                                             \
      tta_record_push([=](cudaStream_t s_) -> int { return call_with_s_; });           \
      return TTA_OK;                                                                   \
    }                                                                                  \
  } while (0)

#define TTA_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      tta_set_error(__VA_ARGS__);           \
      return TTA_ERR_ARG;                   \
    }                                       \
  } while (0)

namespace tta {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute (pdl = false: the launch
// follows a memset or must not overlap its predecessor)
template <typename... KArgs, typename... Args>
inline cudaError_t tta_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && tta_pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, as thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t tta_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster_x > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct alignas(16) U16x8 {
  uint16_t v[8];
};
struct alignas(16) F32x4 {
  float v[4];
};

// One voxel-chunk of fp32 = 32 B = exactly one DRAM/L2 sector: moved with ONE 256-bit access
// (sm_100 LDG/STG.256).  Two 128-bit accesses would each touch half of every sector of the warp's
// 1 KB run, i.e. twice the L1/L2 transactions for the same bytes.
__device__ __forceinline__ void load_f32x8(const float* p, float (&o)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7])
               : "l"(p));
}
__device__ __forceinline__ void store_f32x8(float* p, const float (&o)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]),
               "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
               : "memory");
}

template <int DT>
__device__ __forceinline__ float u16_to_f32(uint16_t u) {
  if (DT == TTA_F16 || DT == TTA_F16_HI) return __half2float(__ushort_as_half(u));
  return __bfloat162float(__ushort_as_bfloat16(u));
}
template <int DT>
__device__ __forceinline__ uint16_t f32_to_u16(float f) {
  if (DT == TTA_F16 || DT == TTA_F16_HI) {
    // saturate instead of producing inf: fp16 max is 65504
    f = fminf(fmaxf(f, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(f));
  }
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}

// x -> (hi, lo) with hi = rn16(x), lo = rn16(x - hi)
template <int DT>
__device__ __forceinline__ void split8(const float (&x)[8], U16x8& hi, U16x8& lo) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint16_t h = f32_to_u16<DT>(x[i]);
    hi.v[i] = h;
    lo.v[i] = f32_to_u16<DT>(x[i] - u16_to_f32<DT>(h));
  }
}
template <int DT>
__device__ __forceinline__ void store_split8(uint16_t* hi, uint16_t* lo,
                                             long long off, const float (&x)[8]) {
  if (DT == TTA_F16_HI) {
    U16x8 h;
#pragma unroll
    for (int i = 0; i < 8; ++i) h.v[i] = f32_to_u16<DT>(x[i]);
    *reinterpret_cast<U16x8*>(hi + off) = h;
    return;
  }
  U16x8 h, l;
  split8<DT>(x, h, l);
  *reinterpret_cast<U16x8*>(hi + off) = h;
  *reinterpret_cast<U16x8*>(lo + off) = l;
}
template <int DT>
__device__ __forceinline__ void load_split8(const uint16_t* hi,
                                            const uint16_t* lo, long long off,
                                            float (&x)[8]) {
  const U16x8 h = *reinterpret_cast<const U16x8*>(hi + off);
  if (DT == TTA_F16_HI) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = u16_to_f32<DT>(h.v[i]);
    return;
  }
  const U16x8 l = *reinterpret_cast<const U16x8*>(lo + off);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = u16_to_f32<DT>(h.v[i]) + u16_to_f32<DT>(l.v[i]);
}

// w-parity-split ("wsplit") position of voxel v = (d*H + h)*W + w inside its [D][H][2][W/2] slab:
// the even-w voxels of a row come first, then the odd ones.  Operand planes that feed a stride-2
// tcgen05 conv are stored this way so that every parity class is a dense TMA row (W must be even).
__device__ __forceinline__ long long wsplit_index(long long v, int W) {
  const int w = (int)(v % W);
  return v - w + (w & 1) * (W >> 1) + (w >> 1);
}

// Per-voxel entropy of R logits and its gradient, returns H:
//   mode 0: softmax entropy   H = lse(z) - sum_c p_c z_c ;  dH/dz_k = -p_k (z_k - sum_c p_c z_c)
//   mode 1: Bernoulli entropy H = sum_c softplus(z_c) - p_c z_c ; dH/dz_c = -z_c p_c (1 - p_c)
// g[c] = gs * dH/dz_c for c < R (entries >= R are left untouched).
template <int RMAX>
__device__ __forceinline__ float entropy_point(const float (&z)[RMAX], int R, int mode, float gs, float (&g)[RMAX]) {
  float Hv = 0.f;
  if (mode == 1) {
    // hardware exp2/log2/rcp (MUFU): |error| ~1e-7 absolute on p and on the entropy term, far below
    // the fp16 gradient planes (5e-4) and averaged out of the loss mean -- ~5x fewer instructions
    // than expf/log1pf, which matters because the fused head is issue bound.
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      if (i < R) {
        const float zi = z[i];
        const float t = __expf(-fabsf(zi));          // e^-|z| in (0, 1]
        const float q = __fdividef(1.f, 1.f + t);    // sigmoid(|z|)
        const float p = zi >= 0.f ? q : t * q;       // sigmoid(z)
        const float sp = fmaxf(zi, 0.f) - __logf(q); // softplus(z) = max(z,0) + log(1 + e^-|z|)
        Hv += sp - p * zi;
        g[i] = -zi * (t * q * q) * gs;             // p (1 - p) = sigmoid(z) sigmoid(-z) = t q^2
      }
    }
  } else {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < RMAX; ++i)
      if (i < R) m = fmaxf(m, z[i]);
    float e[RMAX], S = 0.f;
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      e[i] = (i < R) ? expf(z[i] - m) : 0.f;
      S += e[i];
    }
    const float invS = 1.f / S;
    float pz = 0.f;
#pragma unroll
    for (int i = 0; i < RMAX; ++i)
      if (i < R) pz = fmaf(e[i] * invS, z[i], pz);
    Hv = (m + logf(S)) - pz;
#pragma unroll
    for (int i = 0; i < RMAX; ++i)
      if (i < R) g[i] = -(e[i] * invS) * (z[i] - pz) * gs;
  }
  return Hv;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace tta
