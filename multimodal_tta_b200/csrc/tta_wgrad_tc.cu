// Weight gradients of the 3x3x3 convs on the tensor cores (tcgen05 / TMEM / TMA) -- SURVEY.md 8f-4.
//
//   conv      : dW[co][ci][k] = sum_{n,o} x[n][s*o - 1 + k][ci] * dy[n][o][co]
//   transposed: dW[ci][co][k] = sum_{n,i} x[n][i][ci] * dy[n][2*i - 1 + k][co]
//
// GEMM view per tap: D_k[ci][co] += A[ci][v] * B[v][co] with K = voxels, A = x, B = dy.  Both operands are
// CHANNEL-contiguous in HBM and in shared memory (a voxel-chunk = 8 channels = 16 B), i.e. MN-major UMMA operands
// (instruction-descriptor bits 15 / 16): in the no-swizzle canonical layout a core matrix is 8 K-rows of 16 B = eight
// consecutive-w voxels x 8 channels -- exactly the tile the forward conv stages -- LBO = pitch between 8-voxel K groups
// (the next h row), SBO = pitch between 8-channel M / N groups (the next chunk).  One of the two operands carries the
// filter taps (the "halo'd" one: x for a conv, dy for a transposed conv; stride 2 reads it through the (h, w)-parity
// classes of its w-parity-split copy, which the forward / dgrad stride-2 kernels need anyway), the other is the centre
// tile.  x = fp16 hi + lo planes (two MMAs), dy = one loss-scaled fp16 plane.
//
// A small-shape MMA costs ~40 cycles whatever M and N are (measured: M = 64 / 128 and N = 8 / 32 within 10 %), so the
// kernel makes every MMA as large as the layer allows:
//   * few halo'd chunks (Cin <= 40 for a conv, Cout <= 80 transposed): the TAPS are packed into the M (N) dimension.
//     tpk = 3: the tile is loaded as three w-shifted boxes stacked as chunk groups, one accumulator per kh (kh = a
//     start-address shift / parity class); tpk = 9 (one chunk: network input, logits): nine shifted boxes, ONE
//     accumulator;
//   * otherwise tpk = 1: one halo tile, taps = start-address shifts, one accumulator per tap; with more than 48 centre
//     channels a CTA owns ONE kh (3 accumulators of up to 128 columns) instead of all nine taps.
// A CTA owns (kd [, kh], a ci tile, a co tile), walks a range of tile_h(h) x 8(w) centre tiles (one pipeline stage per
// tile), and flushes its TMEM accumulators once with fp32 atomics into the parameter layout.  Roles: warp 0 TMA
// producer, warp 1 TMEM owner + MMA issue, warps 2..5 epilogue (one per TMEM lane quarter).  M = 64 keeps rows
// 16 q .. 16 q + 15 in the first 16 lanes of TMEM lane quarter q.
#include <cstdlib>
#include <cstring>

#include "tta_common.cuh"
#include "tta_tc_common.cuh"

namespace tta {

constexpr int kWtThreads = 192;
constexpr int kWtMaxLoads = 24;
constexpr int kWtMaxStages = 6;

struct WtLoad {
  int map, smem_off, dw, dh;   // box origin = (w0 + dw, h0 + dh) in the map's own voxel space
  short is_dy, halo;           // which chunk coordinate; halo: plane = halo_d_mul * d + kd - 1 (else d)
  short ph, pad;               // ph: -1 always, 0 / 1 only for CTAs whose kh has this row parity class
};
struct WtAcc {
  int a_off, a_lbo16, a_sbo16, b_off, b_lbo16, b_sbo16;  // byte offsets inside a stage; pitches in 16 B units
  int tap;                                               // kh * 3 + kw of pack index 0
};
struct WtParams {
  CUtensorMap map[10];
  WtLoad ld[kWtMaxLoads];
  WtAcc acc[9];                   // [kh][kw] (tpk = 1), [kh] (tpk = 3), [0] (tpk = 9)
  int nloads, tx_bytes[3], x_lo_off, use_lo;      // tx_bytes[kh] when a CTA owns one kh, else [0]
  int nstages, stage_bytes, n_cols, tmem_cols;
  int m64, rows2, tile_h;
  int tpk, khs, nkh, acc_per_kh, nacc;            // nacc: accumulators per CTA
  int pack_rows, hcg;                             // packed taps live in the rows (conv) or columns (transposed)
  int N, tiles_h, tiles_w, tiles_per_n, tiles_total, tiles_per_block;
  int halo_d_mul, s2;
  int c8x_view, c8y_view, x_tile_chunks, y_tile_chunks;
  int Cin, Cout, layout, co_split;
  int ci_tiles;
  float scale;
  float* dw;
  float* dw2;
  unsigned idesc;
  int debug;   // TTA_WT_DEBUG (timing experiments): 1 = no MMAs, 2 = no TMA loads
};

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// A single thread issues the MMAs of a stage: the loop is fully unrolled, every descriptor is one add away from a
// per-accumulator constant, and consecutive MMAs go to DIFFERENT accumulators (row pair outer, accumulator inner).
template <int ROWS2, int USE_LO>
__device__ __forceinline__ void wt_mma_loop(const WtParams& P, uint32_t tmem_base, uint32_t smem_base, int acc_first,
                                            int t_begin, int t_end, uint64_t* bar_full, uint64_t* bar_empty,
                                            uint64_t* bar_done) {
  const uint32_t leader = elect_one();
  const uint32_t lo16 = (uint32_t)P.x_lo_off >> 4;
  const int nacc = P.nacc;
  uint32_t a_rel[9], b_rel[9], a_w1[9], b_w1[9], a_step[9], b_step[9];
#pragma unroll
  for (int a = 0; a < 9; ++a) {
    const WtAcc T = P.acc[a < nacc ? acc_first + a : acc_first];
    a_rel[a] = ((uint32_t)T.a_off >> 4) | ((uint32_t)T.a_lbo16 << 16);
    b_rel[a] = ((uint32_t)T.b_off >> 4) | ((uint32_t)T.b_lbo16 << 16);
    a_w1[a] = (uint32_t)T.a_sbo16 | (1u << 14);
    b_w1[a] = (uint32_t)T.b_sbo16 | (1u << 14);
    a_step[a] = 2u * (uint32_t)T.a_lbo16;
    b_step[a] = 2u * (uint32_t)T.b_lbo16;
  }
  const uint32_t ncols = (uint32_t)P.n_cols, idesc = P.idesc;
  int s = 0, ph = 0;
  for (int t = t_begin; t < t_end; ++t) {
    if (P.debug & 8) {
      if (leader) mbar_wait(smem_u32(&bar_full[s]), ph);
      __syncwarp();
    } else {
      mbar_wait(smem_u32(&bar_full[s]), ph);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t stage16 = ((smem_base + s * P.stage_bytes) & 0x3FFFFu) >> 4;
    const uint32_t first = t == t_begin ? 0u : 1u;
#pragma unroll
    for (int r = 0; r < ROWS2; ++r) {   // row pairs of the centre tile: two 8-voxel K groups per MMA
#pragma unroll
      for (int a = 0; a < 9; ++a) {
        if (a < nacc) {
          const uint32_t aw = a_rel[a] + stage16 + (uint32_t)r * a_step[a];
          const uint32_t bw = b_rel[a] + stage16 + (uint32_t)r * b_step[a];
          const uint64_t ad_hi = ((uint64_t)a_w1[a] << 32) | aw, bd = ((uint64_t)b_w1[a] << 32) | bw;
          if (leader && (P.debug & 3) != 1) {
            umma_f16(tmem_base + (uint32_t)a * ncols, ad_hi, bd, idesc, r == 0 ? first : 1u);
            if (USE_LO) umma_f16(tmem_base + (uint32_t)a * ncols, ad_hi + lo16, bd, idesc, 1u);
          }
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
  int by = blockIdx.y;
  const int cit = by % P.ci_tiles; by /= P.ci_tiles;
  const int khi = by % P.nkh, kd = by / P.nkh;   // khi: the kh this CTA owns (nkh = 3), 0 when it owns all three
  const int cot = blockIdx.z;
  const int acc_first = P.nkh == 3 ? khi * P.acc_per_kh : 0;
  const int t_begin = blockIdx.x * P.tiles_per_block;
  const int t_end = min(P.tiles_total, t_begin + P.tiles_per_block);

  if (warp == 0) {
    // ===================== TMA producer: lane l issues box l of a stage =====================
    const int my_ph = P.nkh == 3 ? (khi != 1 ? 1 : 0) : -1;    // row parity class of this CTA's taps (stride 2)
    bool mine = false;
    WtLoad L = P.ld[0];
    if (lane < P.nloads) {
      L = P.ld[lane];
      mine = L.ph < 0 || my_ph < 0 || L.ph == my_ph;
    }
    const uint32_t tx = (uint32_t)P.tx_bytes[P.nkh == 3 ? khi : 0];
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
        if ((P.debug & 3) == 2) mbar_arrive(smem_u32(&bar_full[s])); else mbar_expect_tx(smem_u32(&bar_full[s]), tx);
      }
      __syncwarp();
      if (mine && (P.debug & 3) != 2) {
        const int dd = L.halo ? P.halo_d_mul * d + kd - 1 : d;
        const int chunk = L.is_dy ? n * P.c8y_view + cot * P.y_tile_chunks : n * P.c8x_view + cit * P.x_tile_chunks;
        tma_load_4d(smem_base + s * P.stage_bytes + L.smem_off, &P.map[L.map], smem_u32(&bar_full[s]),
                    (w0 + L.dw) * 8, h0 + L.dh, dd, chunk);
      }
      if (++s == P.nstages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issue (warp-uniform code, one elected lane issues) =====================
    if (P.rows2 == 8) {
      if (P.use_lo) wt_mma_loop<8, 1>(P, tmem_base, smem_base, acc_first, t_begin, t_end, bar_full, bar_empty, &bar_done);
      else wt_mma_loop<8, 0>(P, tmem_base, smem_base, acc_first, t_begin, t_end, bar_full, bar_empty, &bar_done);
    } else {
      if (P.use_lo) wt_mma_loop<4, 1>(P, tmem_base, smem_base, acc_first, t_begin, t_end, bar_full, bar_empty, &bar_done);
      else wt_mma_loop<4, 0>(P, tmem_base, smem_base, acc_first, t_begin, t_end, bar_full, bar_empty, &bar_done);
    }
  } else {
    // ===================== epilogue: TMEM -> fp32 atomics into the parameter layout =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = P.m64 ? q * 16 + lane : q * 32 + lane;
    const bool lane_ok = !P.m64 || lane < 16;
    // rows: conv with packed taps -> (pack index, chunk, channel); otherwise ci = tile base + row
    int ci, row_pk = 0;
    if (P.pack_rows) {
      const int grp = row >> 3;
      row_pk = grp / P.hcg;
      ci = cit * P.x_tile_chunks * 8 + (grp - row_pk * P.hcg) * 8 + (row & 7);
    } else {
      ci = cit * P.x_tile_chunks * 8 + row;
    }
    const bool row_live = lane_ok && ci < P.Cin && row_pk < P.tpk;
    if (t_end > t_begin) {
      // the four epilogue warps idle for the whole main loop: one lane each polls, with back-off, so that the
      // producer's and the MMA thread's barrier traffic is not competing with 128 spinning threads
      if (P.debug & 4) {
        if (lane == 0) {
          uint32_t spins = 0;
          while (!mbar_try(smem_u32(&bar_done), 0)) {
            __nanosleep(512);
            if (++spins > (1u << 22)) asm volatile("trap;");
          }
        }
        __syncwarp();
      } else {
        mbar_wait(smem_u32(&bar_done), 0);
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int a = 0; a < P.nacc; ++a) {
        const int tap0 = P.acc[acc_first + a].tap;
        for (int c8 = 0; c8 < P.n_cols / 8; ++c8) {
          uint32_t rr[8];
          tmem_ld8_nowait(tb + (uint32_t)(a * P.n_cols + c8 * 8), rr);
          tmem_ld_wait();
          int col_pk = 0, co0;
          if (P.pack_rows) {
            co0 = cot * P.y_tile_chunks * 8 + c8 * 8;
          } else {
            col_pk = c8 / P.hcg;
            co0 = cot * P.y_tile_chunks * 8 + (c8 - col_pk * P.hcg) * 8;
          }
          if (row_live && col_pk < P.tpk) {
            const int tapi = kd * 9 + tap0 + row_pk + col_pk;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int co = co0 + j;
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

static int wt_round(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace tta

using namespace tta;

extern "C" {

// 1 when tta_conv_wgrad_tc covers this layer: 3x3x3, stride 1 or 2, conv or transposed, one scaled fp16 gradient plane
int tta_conv_wgrad_tc_supported(int mode, int K, int stride, int Cin, int Cout, int dy_dtype) {
  if (K != 3 || dy_dtype != TTA_F16_HI) return 0;
  if (!((mode == 0 && (stride == 1 || stride == 2)) || (mode == 1 && stride == 2))) return 0;
  // <= 4 x <= 4 channels (the full-resolution 3 -> 3 residual unit) fill 3 of 64 rows and 3 of 8 columns of every MMA
  // and re-read the tile once per tap: measured 2.3 .. 4.0 ms against 0.1 ms of tta_conv_wgrad's per-voxel kernel
  return (Cin > 4 || Cout > 4) ? 1 : 0;
}

// Same contract as tta_conv_wgrad (dw += scale * dL/dW in the parameter layout).  Operand layouts: stride 1 -- x and dy
// plain; stride-2 conv -- x W-PARITY-SPLIT (x_wsplit must be 1), dy plain; transposed stride-2 -- x plain, dy
// w-parity-split (dy_wsplit must be 1).  flags bit 0: hi plane of x only (one product instead of two); bit 1: no tap
// packing (A/B and test switch).
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
  TTA_REQUIRE(!strided || (tr ? Wy : Wx) % 2 == 0, "tta_conv_wgrad_tc: w-parity-split rows need an even W");

  WtParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.Cin = Cin; P.Cout = Cout; P.layout = layout; P.co_split = co_split; P.scale = scale; P.dw = dw; P.dw2 = dw2;
  P.use_lo = use_lo ? 1 : 0;
  P.s2 = strided ? 1 : 0;
  { const char* e = getenv("TTA_WT_DEBUG"); P.debug = e ? atoi(e) : 0; }
  P.c8x_view = (int)(x_ns / (Vx * 8)); P.c8y_view = (int)(dy_ns / (Vy * 8));
  const int C8x = (Cin + 7) / 8, C8y = (Cout + 7) / 8;
  const bool nopack = (flags & 2) != 0;

  // ---- shape of the MMAs.  hc: chunks of the halo'd operand (x for a conv, dy transposed)
  int tpk = 1, M, n_cols, khs, co_tiles, ci_tiles;
  int hcg;                       // chunks per pack group of the halo'd operand inside one CTA tile
  int x_tile_chunks, y_tile_chunks;
  if (!tr) {
    const int hc = C8x;
    if (!nopack && hc == 1) tpk = 9;
    else if (!nopack && hc <= 5) tpk = 3;
    if (tpk > 1) {
      M = tpk * hc <= 8 ? 64 : 128;
      hcg = hc;
      x_tile_chunks = hc;
      ci_tiles = 1;
    } else {
      M = (C8x <= 8 || strided) ? 64 : 128;      // stride 2: four parity sub-tiles of x (hi + lo) per stage
      hcg = M / 8;
      x_tile_chunks = M / 8;
      ci_tiles = (C8x + x_tile_chunks - 1) / x_tile_chunks;
    }
    const int nr = M == 128 ? 16 : 8;            // N granularity
    const int cout_r = wt_round(Cout, nr);
    if (tpk == 9) { khs = 3; n_cols = cout_r < 256 ? cout_r : 256; }
    else if (tpk == 3) { khs = 3; n_cols = cout_r <= 160 ? cout_r : 128; }
    else if (cout_r <= 48) { khs = 3; n_cols = cout_r; }
    else { khs = 1; n_cols = cout_r <= 160 ? cout_r : 128; }
    y_tile_chunks = n_cols / 8;
    co_tiles = (Cout + n_cols - 1) / n_cols;
  } else {
    const int hc = C8y;
    M = C8x <= 8 ? 64 : 128;
    x_tile_chunks = M / 8;
    ci_tiles = (C8x + x_tile_chunks - 1) / x_tile_chunks;
    const int nr = M == 128 ? 16 : 8;
    if (!nopack && hc <= 3) { tpk = 9; khs = 3; n_cols = wt_round(9 * hc * 8, nr); hcg = hc; }
    else if (!nopack && hc <= 10) { tpk = 3; n_cols = wt_round(3 * hc * 8, nr); khs = 3 * n_cols <= 512 ? 3 : 1; hcg = hc; }
    else { tpk = 1; khs = 1; n_cols = C8y <= 16 ? wt_round(C8y * 8, nr) : 128; hcg = n_cols / 8; }
    y_tile_chunks = tpk > 1 ? hc : n_cols / 8;
    co_tiles = tpk > 1 ? 1 : (C8y + y_tile_chunks - 1) / y_tile_chunks;
  }
  P.tpk = tpk; P.khs = khs; P.nkh = 3 / khs; P.m64 = M == 64 ? 1 : 0; P.n_cols = n_cols;
  P.acc_per_kh = tpk == 1 ? 3 : 1;
  P.nacc = tpk == 9 ? 1 : P.acc_per_kh * khs;
  P.pack_rows = tr ? 0 : 1;
  P.hcg = hcg;
  P.x_tile_chunks = x_tile_chunks; P.y_tile_chunks = y_tile_chunks; P.ci_tiles = ci_tiles;
  TTA_REQUIRE(P.nacc * n_cols <= 512, "tta_conv_wgrad_tc: %d accumulator columns exceed TMEM", P.nacc * n_cols);
  P.tmem_cols = 32;
  while (P.tmem_cols < P.nacc * n_cols) P.tmem_cols *= 2;
  P.idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((unsigned)(n_cols >> 3) << 17) | ((unsigned)(M >> 4) << 24);
  // chunk groups the descriptors span (absent chunks read stale shared memory inside the stage: discarded rows / columns)
  const int a_groups = M / 8, b_groups = n_cols / 8;
  const int h_groups = tr ? b_groups : a_groups, c_groups = tr ? a_groups : b_groups;
  // chunks a box really carries
  const int x_box = C8x < x_tile_chunks ? C8x : x_tile_chunks, y_box = C8y < y_tile_chunks ? C8y : y_tile_chunks;
  const int h_box = tr ? y_box : x_box, c_box = tr ? x_box : y_box;
  // centre space: dy for a conv, x for a transposed conv
  const int Dc = tr ? Dx : Dy, Hc = tr ? Hx : Hy, Wc = tr ? Wx : Wy;
  P.halo_d_mul = strided ? 2 : 1;
  const int hW = tr ? Wy : Wx, hH = tr ? Hy : Hx, hD = tr ? Dy : Dx;
  const int hplanes = tr ? 1 : (use_lo ? 2 : 1), cplanes = tr ? (use_lo ? 2 : 1) : 1;

  // ---- halo'd-operand regions of one plane for tile height th: {rows, w extent, bytes}
  struct Region { int rows, wx, bytes, off; };
  auto regions_for = [&](int th, Region* R) -> int {   // returns the number of regions; fills rows / wx / bytes
    int nr_ = 0;
    if (tpk == 1) {
      if (!strided) { R[0] = {th + 2, 10, 0, 0}; nr_ = 1; }
      else for (int m = 0; m < 4; ++m) R[nr_++] = {th + (m >> 1), 8 + (m & 1), 0, 0};   // class (ph, pw) = (m >> 1, m & 1)
    } else if (tpk == 3) {
      if (!strided) { R[0] = {th + 2, 8, 0, 0}; nr_ = 1; }
      else { R[0] = {th, 8, 0, 0}; R[1] = {th + 1, 8, 0, 0}; nr_ = 2; }                  // row parity class 0 / 1
    } else {
      R[0] = {th, 8, 0, 0}; nr_ = 1;
    }
    for (int i = 0; i < nr_; ++i) R[i].bytes = wt_round(R[i].rows * R[i].wx * 16 * h_groups, 128);
    return nr_;
  };
  auto stage_bytes_for = [&](int th) {
    Region R[4];
    const int nr_ = regions_for(th, R);
    long long b = 0;
    for (int i = 0; i < nr_; ++i) b += R[i].bytes;
    return (long long)hplanes * b + (long long)cplanes * wt_round(th * 8 * 16 * c_groups, 128);
  };
  const long long smem_max = 227 * 1024 - 4096;   // static shared memory (barriers) + alignment slack come on top
  P.tile_h = 16;
  if (Hc <= 8 || 2 * wt_round((int)stage_bytes_for(16), 1024) > smem_max) P.tile_h = 8;
  TTA_REQUIRE(wt_round((int)stage_bytes_for(P.tile_h), 1024) <= smem_max, "tta_conv_wgrad_tc: a stage does not fit shared memory");
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

  // ---- stage layout: [halo'd operand plane 0 regions | plane 1 regions][centre plane 0 | plane 1]
  const uint16_t* hbase[2] = {tr ? dy_hi : x_hi, tr ? nullptr : x_lo};
  const uint16_t* cbase[2] = {tr ? x_hi : dy_hi, tr ? x_lo : nullptr};
  const cuuint64_t hnc = tr ? ncy : ncx, cnc = tr ? ncx : ncy;
  Region R[4];
  const int nreg = regions_for(TH, R);
  int off = 0, nl = 0, mapi = 0;
  int tx[3] = {0, 0, 0};
  int plane_off[2] = {0, 0};
  auto add_load = [&](int map, int smem_off, int dw_, int dh_, int is_dy, int halo, int ph, int bytes) -> bool {
    if (nl >= kWtMaxLoads) return false;
    P.ld[nl].map = map; P.ld[nl].smem_off = smem_off; P.ld[nl].dw = dw_; P.ld[nl].dh = dh_;
    P.ld[nl].is_dy = (short)is_dy; P.ld[nl].halo = (short)halo; P.ld[nl].ph = (short)ph; P.ld[nl].pad = 0;
    ++nl;
    // bytes a CTA that owns kh = 0, 1, 2 (or all of them: index 0 is then the total) waits for
    for (int kh = 0; kh < 3; ++kh) {
      const bool needed = khs == 3 || ph < 0 || ph == (kh != 1 ? 1 : 0);
      if (needed && (khs == 1 || kh == 0)) tx[kh] += bytes;
    }
    return true;
  };
  for (int pl = 0; pl < hplanes; ++pl) {
    plane_off[pl] = off;
    for (int i = 0; i < nreg; ++i) {
      if (pl == 0) R[i].off = off;
      const int rows = R[i].rows, wx = R[i].wx;
      const int box_bytes = rows * wx * 16 * h_box;
      const int is_dy = tr ? 1 : 0;
      if (tpk == 1) {
        TTA_REQUIRE(mapi < 10, "tta_conv_wgrad_tc: too many tensor maps");
        if (!strided) {
          enc_plain(&P.map[mapi], hbase[pl], hW, hH, hD, hnc, wx, rows, h_box);
          TTA_REQUIRE(add_load(mapi, off, -1, -1, is_dy, 1, -1, box_bytes), "tta_conv_wgrad_tc: too many TMA boxes");
        } else {
          const int phh = i >> 1, pw = i & 1;
          enc_par(&P.map[mapi], hbase[pl], hW, hH, hD, hnc, phh, pw, wx, rows, h_box);
          TTA_REQUIRE(add_load(mapi, off, pw ? -1 : 0, phh ? -1 : 0, is_dy, 1, phh, box_bytes), "tta_conv_wgrad_tc: too many TMA boxes");
        }
        ++mapi;
      } else if (tpk == 3) {
        // three w-shifted boxes stacked as chunk groups: group = kw * h_box + chunk, pitch = rows * 128 B
        if (!strided) {
          TTA_REQUIRE(mapi < 10, "tta_conv_wgrad_tc: too many tensor maps");
          enc_plain(&P.map[mapi], hbase[pl], hW, hH, hD, hnc, 8, rows, h_box);
          for (int kw = 0; kw < 3; ++kw)
            TTA_REQUIRE(add_load(mapi, off + kw * box_bytes, kw - 1, -1, is_dy, 1, -1, box_bytes), "tta_conv_wgrad_tc: too many TMA boxes");
          ++mapi;
        } else {
          const int phh = i;   // region i = row parity class i
          TTA_REQUIRE(mapi + 1 < 10, "tta_conv_wgrad_tc: too many tensor maps");
          enc_par(&P.map[mapi], hbase[pl], hW, hH, hD, hnc, phh, 0, 8, rows, h_box);       // w parity 0: kw = 1
          enc_par(&P.map[mapi + 1], hbase[pl], hW, hH, hD, hnc, phh, 1, 8, rows, h_box);   // w parity 1: kw = 0 (index - 1), 2
          for (int kw = 0; kw < 3; ++kw)
            TTA_REQUIRE(add_load(kw == 1 ? mapi : mapi + 1, off + kw * box_bytes, kw == 0 ? -1 : 0, phh ? -1 : 0, is_dy, 1, phh,
                                 box_bytes), "tta_conv_wgrad_tc: too many TMA boxes");
          mapi += 2;
        }
      } else {
        // nine shifted boxes of exactly the centre tile's shape: group = (kh * 3 + kw) * h_box + chunk
        if (!strided) {
          TTA_REQUIRE(mapi < 10, "tta_conv_wgrad_tc: too many tensor maps");
          enc_plain(&P.map[mapi], hbase[pl], hW, hH, hD, hnc, 8, rows, h_box);
          for (int t9 = 0; t9 < 9; ++t9)
            TTA_REQUIRE(add_load(mapi, off + t9 * box_bytes, t9 % 3 - 1, t9 / 3 - 1, is_dy, 1, -1, box_bytes), "tta_conv_wgrad_tc: too many TMA boxes");
          ++mapi;
        } else {
          TTA_REQUIRE(mapi + 3 < 10, "tta_conv_wgrad_tc: too many tensor maps");
          for (int m = 0; m < 4; ++m) enc_par(&P.map[mapi + m], hbase[pl], hW, hH, hD, hnc, m >> 1, m & 1, 8, rows, h_box);
          for (int t9 = 0; t9 < 9; ++t9) {
            const int kh = t9 / 3, kw = t9 % 3;
            const int m = (kh != 1 ? 2 : 0) + (kw != 1 ? 1 : 0);
            TTA_REQUIRE(add_load(mapi + m, off + t9 * box_bytes, kw == 0 ? -1 : 0, kh == 0 ? -1 : 0, is_dy, 1, -1, box_bytes),
                        "tta_conv_wgrad_tc: too many TMA boxes");
          }
          mapi += 4;
        }
      }
      off += R[i].bytes;
    }
  }
  int cen_off[2] = {0, 0};
  for (int pl = 0; pl < cplanes; ++pl) {
    cen_off[pl] = off;
    TTA_REQUIRE(mapi < 10, "tta_conv_wgrad_tc: too many tensor maps");
    enc_plain(&P.map[mapi], cbase[pl], Wc, Hc, Dc, cnc, 8, TH, c_box);
    TTA_REQUIRE(add_load(mapi, off, 0, 0, tr ? 0 : 1, 0, -1, TH * 8 * 16 * c_box), "tta_conv_wgrad_tc: too many TMA boxes");
    off += wt_round(TH * 8 * 16 * c_groups, 128);
    ++mapi;
  }
  TTA_REQUIRE(ok, "tta_conv_wgrad_tc: cuTensorMapEncodeTiled failed");
  P.nloads = nl;
  for (int kh = 0; kh < 3; ++kh) P.tx_bytes[kh] = tx[kh];
  P.stage_bytes = wt_round(off, 1024);
  P.x_lo_off = use_lo ? (tr ? cen_off[1] - cen_off[0] : plane_off[1] - plane_off[0]) : 0;

  // ---- accumulator table: descriptor start / pitches of the halo'd operand per (kh [, kw])
  auto set_acc = [&](int idx, int h_off, int h_lbo, int h_sbo, int tap) {
    WtAcc& T = P.acc[idx];
    if (!tr) {
      T.a_off = h_off; T.a_lbo16 = h_lbo; T.a_sbo16 = h_sbo;
      T.b_off = cen_off[0]; T.b_lbo16 = 8; T.b_sbo16 = TH * 8;
    } else {
      T.a_off = cen_off[0]; T.a_lbo16 = 8; T.a_sbo16 = TH * 8;
      T.b_off = h_off; T.b_lbo16 = h_lbo; T.b_sbo16 = h_sbo;
    }
    T.tap = tap;
  };
  if (tpk == 1) {
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        int m = 0, rh = kh, rw = kw;
        if (strided) {   // fine index 2*c - 1 + k: k = 1 -> parity 0 (class index c); k = 0 / 2 -> parity 1 (c - 1 / c)
          m = (kh != 1 ? 2 : 0) + (kw != 1 ? 1 : 0);
          rh = kh == 2 ? 1 : 0;
          rw = kw == 2 ? 1 : 0;
        }
        set_acc(kh * 3 + kw, R[m].off + (rh * R[m].wx + rw) * 16, R[m].wx, R[m].rows * R[m].wx, kh * 3 + kw);
      }
  } else if (tpk == 3) {
    for (int kh = 0; kh < 3; ++kh) {
      const int m = strided ? (kh != 1 ? 1 : 0) : 0;
      const int rh = strided ? (kh == 2 ? 1 : 0) : kh;
      set_acc(kh, R[m].off + rh * 8 * 16, 8, R[m].rows * 8, kh * 3);
    }
  } else {
    set_acc(0, R[0].off, 8, TH * 8, 0);
  }

  P.tiles_h = (Hc + TH - 1) / TH;
  P.tiles_w = (Wc + 7) / 8;
  P.tiles_per_n = Dc * P.tiles_h * P.tiles_w;
  P.tiles_total = N * P.tiles_per_n;
  const int gy = 3 * P.nkh * P.ci_tiles, gz = co_tiles;
  // every CTA flushes nacc x M x N atomics into dW, so the tile ranges are as long as filling the GPU once allows
  // (ncu, profiles/wgrad_tc_r2.md: the 8^3 / 16^3 layers were bound by their flush -- 432 CTAs x 49 152 scattered
  // atomics for 512 -> 512 -- so: ONE wave of CTAs, at least four tiles each)
  int gx = 148 / (gy * gz);
  if (gx > (P.tiles_total + 3) / 4) gx = (P.tiles_total + 3) / 4;
  if (gx < 1) gx = 1;
  P.tiles_per_block = (P.tiles_total + gx - 1) / gx;
  gx = (P.tiles_total + P.tiles_per_block - 1) / P.tiles_per_block;
  int nst = (int)(smem_max / P.stage_bytes);
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
