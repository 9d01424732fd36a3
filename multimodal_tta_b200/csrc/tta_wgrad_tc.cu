// Weight gradients of the 3x3x3 convs on the tensor cores (tcgen05 / TMEM / TMA) -- SURVEY.md 8f-4.
//
//   conv      : dW[co][ci][k] = sum_{n,o} x[n][s*o - 1 + k][ci] * dy[n][o][co]
//   transposed: dW[ci][co][k] = sum_{n,i} x[n][i][ci] * dy[n][2*i - 1 + k][co]
//
// GEMM view per tap: D_k[ci (M = 64 | 128)][co (N = 8 | 16 | 32)] += A[ci][v] * B[v][co] with K = voxels.  Both operands
// are CHANNEL-contiguous in HBM and in shared memory (a voxel-chunk = 8 channels = 16 B), i.e. MN-major UMMA operands
// (instruction-descriptor bits 15 / 16): in the no-swizzle canonical layout a core matrix is 8 K-rows of 16 B = eight
// consecutive-w voxels x 8 channels -- exactly the halo tile the forward conv stages -- LBO = pitch between 8-voxel
// K groups (the next h row), SBO = pitch between 8-channel M / N groups (the next chunk).  So, as in the forward
// kernel, every filter tap is only a different descriptor START ADDRESS into one loaded tile; stride-2 and transposed
// convs read the finer tensor through the (h, w)-parity sub-tiles of its w-parity-split copy (which the forward /
// dgrad stride-2 kernels need anyway).  x = fp16 hi + lo planes (two MMAs), dy = one loss-scaled fp16 plane.
//
// A CTA owns (kd, a ci tile, a co tile): nine tap accumulators of N columns each in TMEM (<= 288 of 512 columns),
// walks a range of tile_h(h) x 8(w) centre tiles (one pipeline stage per tile: TMA boxes of the x and dy tiles,
// 9 taps x tile_h/2 row pairs x {hi, lo} MMAs of K = 16 voxels), and flushes once with fp32 atomics into the parameter
// layout.  Roles: warp 0 TMA producer, warp 1 TMEM owner + MMA issue, warps 2..5 epilogue (one per TMEM lane quarter).
// M = 64 (Cin <= 64, and every stride-2 conv: its four parity sub-tiles of x must fit a stage) keeps rows
// 16 q .. 16 q + 15 in the first 16 lanes of TMEM lane quarter q.
#include <cstdlib>
#include <cstring>

#include "tta_common.cuh"
#include "tta_tc_common.cuh"

namespace tta {

constexpr int kWtThreads = 192;
constexpr int kWtMaxLoads = 12;
constexpr int kWtMaxStages = 6;

struct WtLoad {
  int map, smem_off, dw, dh;   // box origin = (w0 + dw, h0 + dh) in the map's own voxel space
  int is_dy, halo;             // which chunk coordinate; halo: plane = halo_d_mul * d + kd - 1 (else d)
};
struct WtTap {
  int a_off, a_lbo16, a_sbo16, b_off, b_lbo16, b_sbo16;  // byte offsets inside a stage; pitches in 16 B units
};
struct WtParams {
  CUtensorMap map[10];
  WtLoad ld[kWtMaxLoads];
  WtTap tap[9];
  int nloads, tx_bytes, x_lo_off, use_lo;                // x_lo_off: byte distance of the lo plane behind the hi plane
  int nstages, stage_bytes, n_cols, tmem_cols;
  int m64, rows2, tile_h;                                // rows2 = tile_h / 2 K = 16 steps per tap and tile
  int N, Dc, tiles_h, tiles_w, tiles_per_n, tiles_total, tiles_per_block;
  int halo_d_mul;
  int c8x_view, c8y_view, m_chunks, co_chunks;
  int Cin, Cout, layout, co_split, co_tile;
  int ci_tiles;
  float scale;
  float* dw;
  float* dw2;
  unsigned idesc;
};

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

template <int ROWS2, int USE_LO>
__device__ __forceinline__ void wt_mma_loop(const WtParams& P, uint32_t tmem_base, uint32_t smem_base, int t_begin,
                                            int t_end, uint64_t* bar_full, uint64_t* bar_empty, uint64_t* bar_done) {
  const uint32_t leader = elect_one();
  const uint32_t lo16 = (uint32_t)P.x_lo_off >> 4;
  // per-tap descriptor words relative to the stage base (16 B units): low word = start | LBO << 16, high = SBO | version
  uint32_t a_rel[9], b_rel[9], a_w1[9], b_w1[9], a_step[9], b_step[9];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const WtTap T = P.tap[tp];
    a_rel[tp] = ((uint32_t)T.a_off >> 4) | ((uint32_t)T.a_lbo16 << 16);
    b_rel[tp] = ((uint32_t)T.b_off >> 4) | ((uint32_t)T.b_lbo16 << 16);
    a_w1[tp] = (uint32_t)T.a_sbo16 | (1u << 14);
    b_w1[tp] = (uint32_t)T.b_sbo16 | (1u << 14);
    a_step[tp] = 2u * (uint32_t)T.a_lbo16;
    b_step[tp] = 2u * (uint32_t)T.b_lbo16;
  }
  const uint32_t ncols = (uint32_t)P.n_cols, idesc = P.idesc;
  int s = 0, ph = 0;
  for (int t = t_begin; t < t_end; ++t) {
    mbar_wait(smem_u32(&bar_full[s]), ph);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t stage16 = ((smem_base + s * P.stage_bytes) & 0x3FFFFu) >> 4;
    const uint32_t first = t == t_begin ? 0u : 1u;
#pragma unroll
    for (int r = 0; r < ROWS2; ++r) {   // row pairs of the centre tile: two 8-voxel K groups per MMA
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const uint32_t aw = a_rel[tp] + stage16 + (uint32_t)r * a_step[tp];
        const uint32_t bw = b_rel[tp] + stage16 + (uint32_t)r * b_step[tp];
        const uint64_t ad_hi = ((uint64_t)a_w1[tp] << 32) | aw, bd = ((uint64_t)b_w1[tp] << 32) | bw;
        if (leader) {
          umma_f16(tmem_base + (uint32_t)tp * ncols, ad_hi, bd, idesc, r == 0 ? first : 1u);
          if (USE_LO) umma_f16(tmem_base + (uint32_t)tp * ncols, ad_hi + lo16, bd, idesc, 1u);
        }
      }
    }
    __syncwarp();
    if (leader) umma_commit(smem_u32(&bar_empty[s]));
    if (++s == P.nstages) { s = 0; ph ^= 1; }
  }
  if (leader) umma_commit(smem_u32(bar_done));
  __syncwarp();
}

__global__ void __launch_bounds__(kWtThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WtParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[kWtMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kWtMaxStages];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.nstages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();
  const uint32_t tmem_base = tmem_base_smem;
  // dynamic shared memory is only guaranteed 16-byte aligned: stages start at the next 128-byte boundary
  const uint32_t smem_base = (smem_u32(smem) + 127u) & ~127u;
  const int kd = blockIdx.y / P.ci_tiles, cit = blockIdx.y % P.ci_tiles, cot = blockIdx.z;
  const int t_begin = blockIdx.x * P.tiles_per_block;
  const int t_end = min(P.tiles_total, t_begin + P.tiles_per_block);

  if (warp == 0) {
    // ===================== TMA producer: lane l issues box l of a stage =====================
    int s = 0, ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int n = t / P.tiles_per_n;
      int r = t - n * P.tiles_per_n;
      const int tw = r % P.tiles_w;
      r /= P.tiles_w;
      const int th = r % P.tiles_h;
      const int d = r / P.tiles_h;
      const int h0 = th * P.tile_h, w0 = tw * 8;
      if (lane == 0) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
        mbar_expect_tx(smem_u32(&bar_full[s]), (uint32_t)P.tx_bytes);
      }
      __syncwarp();
      if (lane < P.nloads) {
        const WtLoad& L = P.ld[lane];
        const int dd = L.halo ? P.halo_d_mul * d + kd - 1 : d;
        const int chunk = L.is_dy ? n * P.c8y_view + cot * P.co_chunks : n * P.c8x_view + cit * P.m_chunks;
        tma_load_4d(smem_base + s * P.stage_bytes + L.smem_off, &P.map[L.map], smem_u32(&bar_full[s]),
                    (w0 + L.dw) * 8, h0 + L.dh, dd, chunk);
      }
      if (++s == P.nstages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issue (warp-uniform code, one elected lane issues) =====================
    // A single thread issues ~150 small MMAs per stage: the loop is fully unrolled, every descriptor is one add away
    // from the previous one, and consecutive MMAs go to DIFFERENT tap accumulators (row pair outer, tap inner).
    if (P.rows2 == 8) {
      if (P.use_lo) wt_mma_loop<8, 1>(P, tmem_base, smem_base, t_begin, t_end, bar_full, bar_empty, &bar_done);
      else wt_mma_loop<8, 0>(P, tmem_base, smem_base, t_begin, t_end, bar_full, bar_empty, &bar_done);
    } else {
      if (P.use_lo) wt_mma_loop<4, 1>(P, tmem_base, smem_base, t_begin, t_end, bar_full, bar_empty, &bar_done);
      else wt_mma_loop<4, 0>(P, tmem_base, smem_base, t_begin, t_end, bar_full, bar_empty, &bar_done);
    }
  } else {
    // ===================== epilogue: TMEM -> fp32 atomics into the parameter layout =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int ci = P.m64 ? cit * 64 + q * 16 + lane : cit * 128 + q * 32 + lane;
    const bool row_live = ci < P.Cin && (!P.m64 || lane < 16);
    if (t_end > t_begin) {
      mbar_wait(smem_u32(&bar_done), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int tp = 0; tp < 9; ++tp) {
        const int tapi = kd * 9 + tp;
        for (int c8 = 0; c8 < P.n_cols / 8; ++c8) {
          uint32_t rr[8];
          tmem_ld8_nowait(tb + (uint32_t)(tp * P.n_cols + c8 * 8), rr);
          tmem_ld_wait();
          if (row_live) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int co = cot * P.co_tile + c8 * 8 + j;
              const float v = __uint_as_float(rr[j]) * P.scale;
              if (co < P.Cout && v != 0.f) {
                float* dst;
                if (P.layout == 1) dst = P.dw + ((long long)ci * P.Cout + co) * 27 + tapi;
                else if (co < P.co_split) dst = P.dw + ((long long)co * P.Cin + ci) * 27 + tapi;
                else dst = P.dw2 + ((long long)(co - P.co_split) * P.Cin + ci) * 27 + tapi;
                atomicAdd(dst, v);
              }
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                 : "memory");
  }
}

static int wt_round128(int x) { return (x + 127) / 128 * 128; }

}  // namespace tta

using namespace tta;

extern "C" {

// 1 when tta_conv_wgrad_tc covers this layer: 3x3x3, stride 1 or 2, conv or transposed, one scaled fp16 gradient plane
int tta_conv_wgrad_tc_supported(int mode, int K, int stride, int Cin, int Cout, int dy_dtype) {
  if (K != 3 || dy_dtype != TTA_F16_HI) return 0;
  if (!((mode == 0 && (stride == 1 || stride == 2)) || (mode == 1 && stride == 2))) return 0;
  // <= 4-channel tensors (network input, logits) would fill 4 of 64 rows / 3 of 8 columns of every MMA: measured
  // slower than the CUDA-core kernel (3 -> 3 at 128^3: 4.8 ms against 1.2 ms), so those layers stay there
  return Cin >= 8 && Cout >= 8 ? 1 : 0;
}

// Same contract as tta_conv_wgrad (dw += scale * dL/dW in the parameter layout).  Operand layouts: stride 1 -- x and dy
// plain; stride-2 conv -- x W-PARITY-SPLIT (x_wsplit must be 1), dy plain; transposed stride-2 -- x plain, dy
// w-parity-split (dy_wsplit must be 1).  flags bit 0: hi plane of x only (one product instead of two).
int tta_conv_wgrad_tc(const uint16_t* x_hi, const uint16_t* x_lo, long long x_ns, int Dx, int Hx, int Wx, int x_wsplit,
                      const uint16_t* dy_hi, long long dy_ns, int Dy, int Hy, int Wy, int dy_wsplit, int N, int mode,
                      int stride, int Cin, int Cout, float scale, float* dw, int layout, int co_split, float* dw2,
                      int flags, cudaStream_t stream) {
  TTA_RECORDABLE(tta_conv_wgrad_tc(x_hi, x_lo, x_ns, Dx, Hx, Wx, x_wsplit, dy_hi, dy_ns, Dy, Hy, Wy, dy_wsplit, N, mode, stride, Cin, Cout, scale, dw, layout, co_split, dw2, flags, s_));
  TTA_REQUIRE(x_hi && dy_hi && dw && N > 0, "tta_conv_wgrad_tc: null pointer");
  const bool use_lo = !(flags & 1);
  TTA_REQUIRE(x_lo || !use_lo, "tta_conv_wgrad_tc: the lo plane of x is missing");
  TTA_REQUIRE(tta_conv_wgrad_tc_supported(mode, 3, stride, Cin, Cout, TTA_F16_HI), "tta_conv_wgrad_tc: unsupported layer");
  const bool strided = stride == 2;
  const bool tr = mode == 1;
  TTA_REQUIRE(!strided || (tr ? (dy_wsplit && !x_wsplit) : (x_wsplit && !dy_wsplit)),
              "tta_conv_wgrad_tc: the finer operand of a stride-2 layer must be stored w-parity-split");
  TTA_REQUIRE(strided || (!x_wsplit && !dy_wsplit), "tta_conv_wgrad_tc: stride-1 operands are plain");
  if (co_split <= 0 || co_split > Cout) co_split = Cout;
  TTA_REQUIRE(co_split == Cout || (dw2 != nullptr && layout == 0), "tta_conv_wgrad_tc: a split output needs dw2 and layout 0");
  EncodeTiledFn enc = get_encode();
  TTA_REQUIRE(enc != nullptr, "tta_conv_wgrad_tc: cuTensorMapEncodeTiled entry point not found");
  const long long Vx = (long long)Dx * Hx * Wx, Vy = (long long)Dy * Hy * Wy;
  TTA_REQUIRE(x_ns % (Vx * 8) == 0 && dy_ns % (Vy * 8) == 0, "tta_conv_wgrad_tc: n strides must be whole chunks");
  const int hWd = tr ? Wy : Wx;   // the halo'd (finer for stride 2) operand
  TTA_REQUIRE(!strided || hWd % 2 == 0, "tta_conv_wgrad_tc: w-parity-split rows need an even W");

  WtParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.Cin = Cin; P.Cout = Cout; P.layout = layout; P.co_split = co_split; P.scale = scale; P.dw = dw; P.dw2 = dw2;
  P.use_lo = use_lo ? 1 : 0;
  P.c8x_view = (int)(x_ns / (Vx * 8)); P.c8y_view = (int)(dy_ns / (Vy * 8));
  const int C8x = (Cin + 7) / 8, C8y = (Cout + 7) / 8;
  // M = 64 for Cin <= 64 and for every stride-2 conv (four parity sub-tiles of x, hi + lo, must fit a stage)
  P.m64 = (Cin <= 64 || (strided && !tr)) ? 1 : 0;
  if (getenv("TTA_WT_M128") && !(strided && !tr)) P.m64 = 0;   // A/B experiment: M = 128 tiles also for Cin <= 64
  const int M = P.m64 ? 64 : 128;
  P.m_chunks = M / 8;
  P.ci_tiles = (C8x + P.m_chunks - 1) / P.m_chunks;
  P.co_tile = Cout > 16 ? 32 : (Cout > 8 || !P.m64 ? 16 : 8);   // N % 16 == 0 for M = 128
  P.co_chunks = P.co_tile / 8;
  const int co_tiles = (Cout + P.co_tile - 1) / P.co_tile;
  P.n_cols = P.co_tile;
  P.tmem_cols = 32;
  while (P.tmem_cols < 9 * P.n_cols) P.tmem_cols *= 2;
  P.idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((unsigned)(P.n_cols >> 3) << 17) | ((unsigned)(M >> 4) << 24);
  // chunks a box of each operand really carries (the rest of the M / N groups reads stale shared memory: those
  // rows / columns are discarded by the epilogue)
  const int xc = C8x < P.m_chunks ? C8x : P.m_chunks;
  const int yc = C8y < P.co_chunks ? C8y : P.co_chunks;
  // centre space: dy for a conv, x for a transposed conv
  const int Dc = tr ? Dx : Dy, Hc = tr ? Hx : Hy, Wc = tr ? Wx : Wy;
  P.Dc = Dc;
  P.halo_d_mul = strided ? 2 : 1;
  const int hW = tr ? Wy : Wx, hH = tr ? Hy : Hx, hD = tr ? Dy : Dx;
  const int hchunks = tr ? yc : xc, cchunks = tr ? xc : yc;
  const int hgroups = tr ? P.co_chunks : P.m_chunks;     // groups the descriptor of the halo'd operand spans
  const int cgroups = tr ? P.m_chunks : P.co_chunks;
  const int hplanes = tr ? 1 : (use_lo ? 2 : 1), cplanes = tr ? (use_lo ? 2 : 1) : 1;
  // tile height: 16 rows unless the stage would not allow two stages
  auto stage_bytes_for = [&](int th) {
    long long hv = strided ? (long long)(th * 8 + th * 9 + (th + 1) * 8 + (th + 1) * 9) : (long long)(th + 2) * 10;
    long long hb = hv * 16 * hgroups, cb = (long long)th * 8 * 16 * cgroups;
    return (long long)hplanes * (hb + 512) + (long long)cplanes * (cb + 128) + 1024;
  };
  const long long smem_max = 227 * 1024 - 4096;   // static shared memory (barriers) + alignment slack come on top
  P.tile_h = 16;
  if (Hc <= 8 || 2 * stage_bytes_for(16) > smem_max) P.tile_h = 8;
  TTA_REQUIRE(stage_bytes_for(P.tile_h) <= smem_max, "tta_conv_wgrad_tc: a stage does not fit shared memory");
  P.rows2 = P.tile_h / 2;
  const int TH = P.tile_h;
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const cuuint64_t ncx = (cuuint64_t)((long long)(N - 1) * P.c8x_view + C8x),
                   ncy = (cuuint64_t)((long long)(N - 1) * P.c8y_view + C8y);
  bool ok = true;
  // plain tile of an operand: rows of W*8 16-bit values
  auto enc_plain = [&](CUtensorMap* m, const uint16_t* base, int W_, int H_, int D_, cuuint64_t nc_ext, int bw, int bh, int bc) {
    cuuint64_t gdim[4] = {(cuuint64_t)W_ * 8, (cuuint64_t)H_, (cuuint64_t)D_, nc_ext};
    cuuint64_t gstr[3] = {(cuuint64_t)16 * W_, (cuuint64_t)16 * W_ * H_, (cuuint64_t)16 * W_ * H_ * D_};
    cuuint32_t box[4] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, 1, (cuuint32_t)bc};
    ok = ok && enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  // (h, w)-parity class of a w-parity-split operand ([D][H][2][W/2][8]): a dense run of W/2 voxels per row
  auto enc_par = [&](CUtensorMap* m, const uint16_t* base, int W_, int H_, int D_, cuuint64_t nc_ext, int par_h, int par_w,
                     int bw, int bh, int bc) {
    const uint16_t* ptr = base + ((long long)par_h * W_ + (long long)par_w * (W_ / 2)) * 8;
    cuuint64_t gdim[4] = {(cuuint64_t)(W_ / 2) * 8, (cuuint64_t)((H_ - par_h + 1) / 2), (cuuint64_t)D_, nc_ext};
    cuuint64_t gstr[3] = {(cuuint64_t)32 * W_, (cuuint64_t)16 * W_ * H_, (cuuint64_t)16 * W_ * H_ * D_};
    cuuint32_t box[4] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, 1, (cuuint32_t)bc};
    ok = ok && enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  // ---- stage layout: [halo'd operand: plane 0 sub-tiles | plane 1 sub-tiles][centre operand: plane 0 | plane 1].
  // Every sub-tile region is sized for ALL the chunk groups the MMA descriptor spans, so that reads of absent chunks
  // stay inside the stage.
  const uint16_t* hbase[2] = {tr ? dy_hi : x_hi, tr ? nullptr : x_lo};
  const uint16_t* cbase[2] = {tr ? x_hi : dy_hi, tr ? x_lo : nullptr};
  const cuuint64_t hnc = tr ? ncy : ncx, cnc = tr ? ncx : ncy;
  const int nsub = strided ? 4 : 1;
  int sub_hx[4], sub_wx[4], sub_off[4];
  int off = 0, nl = 0, mapi = 0, tx = 0;
  int plane_off[2] = {0, 0};
  for (int pl = 0; pl < hplanes; ++pl) {
    plane_off[pl] = off;
    for (int m = 0; m < nsub; ++m) {
      const int phh = m >> 1, pw = m & 1;
      const int hx = strided ? TH + phh : TH + 2, wx = strided ? 8 + pw : 10;
      sub_hx[m] = hx; sub_wx[m] = wx;
      if (pl == 0) sub_off[m] = off - plane_off[0];
      TTA_REQUIRE(nl < kWtMaxLoads && mapi < 10, "tta_conv_wgrad_tc: too many TMA boxes per stage");
      if (strided) enc_par(&P.map[mapi], hbase[pl], hW, hH, hD, hnc, phh, pw, wx, hx, hchunks);
      else enc_plain(&P.map[mapi], hbase[pl], hW, hH, hD, hnc, wx, hx, hchunks);
      P.ld[nl].map = mapi; P.ld[nl].smem_off = off;
      P.ld[nl].dw = strided ? (pw ? -1 : 0) : -1;
      P.ld[nl].dh = strided ? (phh ? -1 : 0) : -1;
      P.ld[nl].is_dy = tr ? 1 : 0; P.ld[nl].halo = 1;
      tx += hx * wx * 16 * hchunks;
      off += wt_round128(hx * wx * 16 * hgroups);
      ++nl; ++mapi;
    }
  }
  int cen_off[2] = {0, 0};
  for (int pl = 0; pl < cplanes; ++pl) {
    cen_off[pl] = off;
    TTA_REQUIRE(nl < kWtMaxLoads && mapi < 10, "tta_conv_wgrad_tc: too many TMA boxes per stage");
    enc_plain(&P.map[mapi], cbase[pl], Wc, Hc, Dc, cnc, 8, TH, cchunks);
    P.ld[nl].map = mapi; P.ld[nl].smem_off = off; P.ld[nl].dw = 0; P.ld[nl].dh = 0;
    P.ld[nl].is_dy = tr ? 0 : 1; P.ld[nl].halo = 0;
    tx += TH * 8 * 16 * cchunks;
    off += wt_round128(TH * 8 * 16 * cgroups);
    ++nl; ++mapi;
  }
  TTA_REQUIRE(ok, "tta_conv_wgrad_tc: cuTensorMapEncodeTiled failed");
  P.nloads = nl;
  P.tx_bytes = tx;
  P.stage_bytes = (off + 1023) / 1024 * 1024;
  P.x_lo_off = use_lo ? (tr ? cen_off[1] - cen_off[0] : plane_off[1] - plane_off[0]) : 0;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      int m = 0, rh = kh, rw = kw;
      if (strided) {   // fine index 2*c - 1 + k: k = 1 -> parity 0 (class index c); k = 0 / 2 -> parity 1 (class index c - 1 / c)
        m = (kh != 1 ? 2 : 0) + (kw != 1 ? 1 : 0);
        rh = kh == 2 ? 1 : 0;
        rw = kw == 2 ? 1 : 0;
      }
      const int wx = sub_wx[m], hx = sub_hx[m];
      const int h_off = plane_off[0] + sub_off[m] + (rh * wx + rw) * 16;   // tap = start-address shift inside the sub-tile
      WtTap& T = P.tap[kh * 3 + kw];
      if (!tr) {
        T.a_off = h_off; T.a_lbo16 = wx; T.a_sbo16 = hx * wx;
        T.b_off = cen_off[0]; T.b_lbo16 = 8; T.b_sbo16 = TH * 8;
      } else {
        T.a_off = cen_off[0]; T.a_lbo16 = 8; T.a_sbo16 = TH * 8;
        T.b_off = h_off; T.b_lbo16 = wx; T.b_sbo16 = hx * wx;
      }
    }
  P.tiles_h = (Hc + TH - 1) / TH;
  P.tiles_w = (Wc + 7) / 8;
  P.tiles_per_n = Dc * P.tiles_h * P.tiles_w;
  P.tiles_total = N * P.tiles_per_n;
  const int gy = 3 * P.ci_tiles, gz = co_tiles;
  // ~2 CTAs per SM over the whole grid (one resident at a time: the second wave hides the first one's flush), at
  // least 2 tiles per CTA so that the pipeline has something to overlap
  // (every CTA flushes 9 x M x N atomics: where the (kd, ci, co) tiles alone fill the GPU, one CTA per tile)
  int gx = gy * gz >= 148 ? 1 : (2 * 148 + gy * gz - 1) / (gy * gz);
  if (gx > (P.tiles_total + 1) / 2) gx = (P.tiles_total + 1) / 2;
  if (gx < 1) gx = 1;
  P.tiles_per_block = (P.tiles_total + gx - 1) / gx;
  gx = (P.tiles_total + P.tiles_per_block - 1) / P.tiles_per_block;
  int nst = (int)((smem_max) / P.stage_bytes);
  if (nst > kWtMaxStages) nst = kWtMaxStages;
  if (nst > P.tiles_per_block) nst = P.tiles_per_block;
  TTA_REQUIRE(nst >= 1, "tta_conv_wgrad_tc: no stage fits");
  P.nstages = nst;
  const size_t smem_bytes = (size_t)nst * P.stage_bytes + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048) != cudaSuccess) {
      tta_set_error("tta_conv_wgrad_tc: cudaFuncSetAttribute failed");
      return TTA_ERR_CUDA;
    }
    attr_set = true;
  }
  tta_launch(wgrad_tc_kernel, dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), kWtThreads, smem_bytes, stream,
             tta_pdl_family(32), P);
  return tta_check_launch("tta_conv_wgrad_tc");
}

}  // extern "C"
