// tcgen05 / TMEM / TMA implicit-GEMM 3-D convolution for sm_100a (forward AND input-gradient).
//
// GEMM view per work item:  D[128 voxels x NT couts] += A[128 voxels x 16 cin] * B[16 cin x NT]
// for every (16-channel block, tap).  M = 128 rows = a 16(h) x 8(w) patch of one d-plane of the
// "tile space"; one work item owns up to 8 such accumulators in TMEM (TD consecutive d-planes for
// convs, the 8 output-parity classes for stride-2 transposed convs).
//
//  * A operand: the input HALO tile is loaded once per (channel block, tap group) by TMA from the
//    channel-blocked layout [N][C8][D][H][W][8] into smem as [kchunk][d][h][w][8ch], NO swizzle.
//    In the UMMA K-major no-swizzle canonical layout a core matrix is 8 rows x 16 B contiguous --
//    exactly 8 consecutive-w voxels x 8 channels -- so every filter tap is just a different
//    descriptor START ADDRESS into the same halo tile (SBO = halo row pitch, LBO = k-chunk pitch):
//    27 taps reuse one smem tile, no im2col materialisation, zero padding comes from TMA OOB fill.
//    Stride-2 convs read 4 (h,w)-parity sub-tiles through strided tensor maps; stride-2 transposed
//    convs are 8 output-parity sub-convolutions over the same halo tile.
//  * B operand: weights pre-packed on the host per (n-tile, channel block, tap group) as
//    [tap][kchunk 2][hi NT rows | lo NT rows][8ch] blobs: ONE cp.async.bulk per stage.
//  * Precision: operands are split 16-bit planes (x = hi + lo; fp16 forward, bf16 backward).
//    Per k-step TWO MMAs:  A_hi x [B_hi | B_lo]  (N = 2*NT: the stacked B reads A_hi once) and
//    A_lo x B_hi (N = NT); the epilogue adds the two column halves.  That is hi*hi + hi*lo + lo*hi
//    in fp32 -- ~fp32-equivalent products -- for 2/3 of the shared-memory operand traffic of three
//    separate MMAs (small-N MMAs are bound by the 4 KB A-tile read, not by tensor math).
//  * Persistent CTAs: grid = min(work items, #SMs), 352 threads: warps 0 and 10 = TMA producers
//    (alternating ring positions), warp 1 = TMEM alloc + MMA issue (warp-uniform code, one elected lane
//    issues), warps 2..9 = epilogue (two per TMEM lane quarter).  The smem stage ring (full/empty
//    mbarriers) runs across work items; TMEM accumulators are double buffered when they fit, so the
//    epilogue of item i overlaps the main loop of item i+1.
//  * Geometry variants (issue_group<GEOM>): S1 / S1T / K1 / S2 / T2 as above; S1P / S1TP = two d-planes
//    per 128-row tile on planes with H <= 8; S1K / S1TK = kd-stacked MMAs for small n-tiles with
//    resident weights; S2 with a one-chunk input issues tap PAIRS as the two k-chunks of one MMA.
//    conv_t2s_kernel (below) is a separate kernel: transposed stride-2 conv with <= 4 output channels as
//    one dense GEMM per input plane + a shared-memory col2im.
#include <cuda.h>

#include <algorithm>
#include <type_traits>
#include <cstdio>
#include <cstdlib>

#include "tta_common.cuh"
#include "tta_tc_common.cuh"

namespace tta {

constexpr int kTcThreads = 352;  // TMA producer, MMA issuer, 8 epilogue warps, second TMA producer
constexpr int kMaxGroups = 9;
constexpr int kMaxLoads = 4;
constexpr int kMaxAcc = 8;
constexpr int kMaxStages = 8;

enum { GEOM_NONE = 0, GEOM_S1, GEOM_K1, GEOM_S1T, GEOM_S2, GEOM_T2, GEOM_S1P, GEOM_S1TP, GEOM_S1K, GEOM_S1TK, GEOM_S2C4 };

struct TcLoad {
  int map, dw, dh, dd, smem_off, bytes, chunk_pitch, pad;  // bytes / pitch are per k-chunk (8 channels)
};
struct TcGroup {
  int nloads, nmma, tx_bytes, pad;  // tx_bytes: A bytes of both planes for ONE k-chunk
  TcLoad ld[kMaxLoads];
};
struct TcParams {
  CUtensorMap amap[8];  // stride-2: [parity class 0..3][hi, lo]; otherwise [0] = hi, [1] = lo
  TcGroup grp[kMaxGroups];
  int ngroups, ncblk, nstages, ntile, n_ntiles, nacc, td, nbuf;
  int tiles_w, tiles_h, tiles_d, d_mul;
  int a_plane_bytes, b_entry_bytes, b_blob_bytes, stage_bytes;
  int c8_view, c8in, single_chunk, tmem_cols;
  int out_mul, Do, Ho, Wo, C8out, accumulate, idesc_n, idesc_2n, idesc0, pad_i;
  int ksplit, cb_per_split, work_items, b_off;  // b_off: byte offset of the B blob inside a stage
  int s2_rows;  // stride-2 input stored w-parity-split: sub-tiles are whole-row 4-D TMA boxes
  int b_nblob, b_cb_bytes, b_g_bytes;  // resident weights: blobs to copy; bytes per channel block / per group
  // CTA PAIRS (cluster of 2, opt-in): the two CTAs work on neighbouring tiles of the same (n-tile, split)
  // and each fetches HALF of every weight blob, multicast into both; cl2 = 2 when active
  int cl2, pair_items;
  int pps;      // kd-stacked convs: halo planes per pipeline stage
  int out_f16;  // flags bit 16: the result is ONE fp16 plane (16 B per voxel-chunk): loss-scaled input gradients
  int s2pair;   // GEOM_S2 with a one-chunk input: tap pairs share one K = 16 MMA (see issue_group)
  int t2_jh16;  // GEOM_T2: offset (16 B units) of the h+1 halo rows inside a k-chunk: 9 = next row, or a second box
  int pl2;      // small-plane tiles (H <= 8): the 128 rows are 2 d-planes x 8 h x 8 w (GEOM_S1P / GEOM_S1TP)
  // fused norm statistics: per-CTA partial sums of y and y^2 over the leading stats_c8 chunks of the
  // output, layout [n][chunk][cta][16] (0..7 sum, 8..15 sum of squares) = what tta_norm_apply
  // finalizes with splits = gridDim.x
  int stats_c8, n_batch;
  float* stats;
  // fused norm-BACKWARD reductions (STATS == 2): the gradient this dgrad writes is the complete
  // dL/d(act) of up to two norm layers (channel segments of a concat buffer); the epilogue turns it
  // into dz = g * [gamma*xhat + beta > 0] and leaves per-CTA partial sums of dz and dz*xhat, layout
  // [n][chunk - c8_begin][cta][16], which tta_norm_bwd_finalize reduces (tta_norm_bwd_reduce's pass
  // over g and y disappears)
  struct BwdSeg {
    int c8_begin, c8_end, relu, pad;
    const float* y;
    long long y_ns;
    const float* mean;
    const float* rstd;
    const float* gamma;
    const float* beta;
    float* partial;
  } seg[2];
  int nseg;
  // resident weights: all (channel block, group) blobs of the single n-tile are copied into shared
  // memory ONCE per CTA (persistent), stages then carry only the A operand
  int b_res, b_res_off, pad_res;
  // x / d == __umulhi(x, magic(d)) for every work-item index (host checks work_items * d < 2^32)
  unsigned mg_per_tile, mg_tiles, mg_ksplit, mg_tiles_w, mg_tiles_h;
  int lbo16[4];  // k-chunk pitch (16 B units, 128 B aligned) per A sub-tile
  signed char acc_pd[kMaxAcc], acc_qd[kMaxAcc], acc_qh[kMaxAcc], acc_qw[kMaxAcc];
  long long out_ns;
  const uint8_t* wpacked;
  const float* bias;
  float* out;
};

// One k-step (16 channels) of one tap for one accumulator:
//   D[0:2NT] (+)= A_hi * [B_hi | B_lo]   and   D[0:NT] += A_lo * B_hi
// descriptor words: w0 = (addr >> 4) | lbo16 << 16 ; w1 = sbo16 | version(1) << 14.
// Called by ALL lanes of the MMA warp with warp-uniform arguments (so the descriptor arithmetic
// stays in the uniform datapath); only the elected lane issues the two tcgen05.mma.
// SPLIT = 0 (single-plane operands, e.g. scaled-fp16 gradients): one MMA, D[0:NT] (+)= A * B.
template <int SPLIT>
__device__ __forceinline__ void mma_pair(uint32_t leader, uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t a_w1,
                                         uint32_t b_w0, uint32_t b_w1, uint32_t idesc_2n, uint32_t idesc_n,
                                         uint32_t accumulate) {
  const uint64_t ad_hi = ((uint64_t)a_w1 << 32) | a_hi, ad_lo = ((uint64_t)a_w1 << 32) | a_lo;
  const uint64_t bd = ((uint64_t)b_w1 << 32) | b_w0;
  if (leader) {
    if (SPLIT) {
      umma_f16(d, ad_hi, bd, idesc_2n, accumulate);
      if (idesc_n != 0u) umma_f16(d, ad_lo, bd, idesc_n, 1u);   // idesc_n == 0: flags bit 17, two products (no A_lo * B_hi)
    } else {
      umma_f16(d, ad_hi, bd, idesc_n, accumulate);
    }
  }
}

// All MMAs of pipeline group g for one smem stage.  Tap geometry and TD are compile-time, so per
// MMA the warp spends a couple of uniform integer adds; `first` = first stage of the work item
// (the first MMA into each accumulator overwrites instead of accumulating).
template <int GEOM, int TD, int SPLIT>
__device__ __forceinline__ void issue_group(const TcParams& P, uint32_t leader, int g, uint32_t stage,
                                            uint32_t bsrc, uint32_t tmem_acc0, bool first) {
  // shared-memory addresses carry the CTA's rank in the cluster in bits 24+ (pair mode); the matrix
  // descriptors take the 18-bit offset inside the CTA's own window
  stage &= 0x3FFFFu;
  bsrc &= 0x3FFFFu;
  const uint32_t nt = P.ntile;
  const uint32_t acc_cols = SPLIT ? 2u * nt : nt;
  const uint32_t a_hi0 = stage >> 4, a_lo0 = (stage + P.a_plane_bytes) >> 4;
  const uint32_t b_w0 = (bsrc >> 4) | (acc_cols << 16);  // bsrc: this (cblk, group) blob; lbo16 = B rows per k-chunk
  const uint32_t b_ent = (uint32_t)P.b_entry_bytes >> 4;                          // entry pitch, 16 B units
  const uint32_t b_w1 = 8u | (1u << 14);                                          // sbo16 = 8 (128 B)
  const uint32_t i2n = P.idesc_2n, in_ = P.idesc_n;
  if (GEOM == GEOM_S1 || GEOM == GEOM_S1T) {
    const uint32_t a_w1 = 10u | (1u << 14);
    const uint32_t lbo = (uint32_t)P.lbo16[0] << 16;
    const uint32_t ah = a_hi0 | lbo, al = a_lo0 | lbo;   // address field never carries into lbo
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int rh = GEOM == GEOM_S1 ? kh : 2 - kh, rw = GEOM == GEOM_S1 ? kw : 2 - kw;
        const uint32_t bo = b_w0 + (uint32_t)(kh * 3 + kw) * b_ent;
        const uint32_t accum = (first && g == 0 && kh == 0 && kw == 0) ? 0u : 1u;
#pragma unroll
        for (int p = 0; p < TD; ++p) {
          const uint32_t ao = (uint32_t)(rh * 10 + rw + p * 180);
          mma_pair<SPLIT>(leader, tmem_acc0 + p * acc_cols, ah + ao, al + ao, a_w1, bo, b_w1, i2n, in_, accum);
        }
      }
  } else if (GEOM == GEOM_S1K || GEOM == GEOM_S1TK) {
    // kd-STACKED stride-1 conv for small n-tiles.  An M = 128, K = 16 MMA reads a 4 KB A tile (32 cycles
    // of shared-memory operand bandwidth) but needs only N/2 tensor cycles, so one MMA per (plane, tap)
    // with N = acc_cols < 128 leaves the tensor pipe mostly idle.  Input plane j of the (TD + 2)-plane
    // halo feeds output planes j-2 .. j through the three kd taps of a (kh, kw) pair: ONE MMA over the
    // stack of their weights (slot s -> accumulator j - 2 + s, accumulators contiguous in TMEM,
    // N = k * acc_cols <= 192).  A stage holds P.pps consecutive halo planes (each loaded once per channel
    // block instead of once per kd); weights are resident, entry (kh, kw) =
    // [kchunk][slot 0..2][hi NT | lo NT][8].  With split planes A_hi and A_lo both multiply the whole
    // [B_hi | B_lo] stack (as GEOM_T2 does; the extra lo*lo term is below fp32 resolution).
    const uint32_t a_w1 = 10u | (1u << 14);
    const uint32_t lbo = (uint32_t)P.lbo16[0] << 16;
    const uint32_t lbo_b = (3u * acc_cols) << 16;                      // rows of all three slots per k-chunk
    const uint32_t idesc_1 = P.idesc0 | ((acc_cols >> 3) << 17);
    for (int jj = 0; jj < P.pps; ++jj) {
      const int j = g * P.pps + jj;
      const uint32_t ah = (a_hi0 + (uint32_t)jj * 180u) | lbo, al = (a_lo0 + (uint32_t)jj * 180u) | lbo;
      const int s_lo = j < 2 ? 2 - j : 0, s_hi = (TD + 1 - j) < 2 ? (TD + 1 - j) : 2;
      const uint32_t k = (uint32_t)(s_hi - s_lo + 1);
      const uint32_t d0 = tmem_acc0 + (uint32_t)(j - 2 + s_lo) * acc_cols;
      const uint32_t sbase = (bsrc >> 4) + (uint32_t)s_lo * acc_cols;    // 16-byte rows
      const uint32_t idesc_k = P.idesc0 | (((k * acc_cols) >> 3) << 17);
      const uint32_t idesc_k1 = P.idesc0 | ((((k - 1u) * acc_cols) >> 3) << 17);
      const bool fresh = first && j < TD;   // first channel block: accumulator j (the top slot) starts here
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int rh = GEOM == GEOM_S1K ? kh : 2 - kh, rw = GEOM == GEOM_S1K ? kw : 2 - kw;
          const uint32_t ao = (uint32_t)(rh * 10 + rw);
          const uint32_t bw0 = (sbase + (uint32_t)(kh * 3 + kw) * b_ent) | lbo_b;
          const uint64_t bd = ((uint64_t)b_w1 << 32) | bw0;
          const uint64_t ad_hi = ((uint64_t)a_w1 << 32) | (ah + ao), ad_lo = ((uint64_t)a_w1 << 32) | (al + ao);
          if (fresh && kh == 0 && kw == 0) {
            // the older accumulators of the stack accumulate, the fresh one is overwritten
            const uint64_t bd_top = ((uint64_t)b_w1 << 32) | ((bw0 + (k - 1u) * acc_cols));
            if (leader) {
              if (k > 1u) {
                umma_f16(d0, ad_hi, bd, idesc_k1, 1u);
                if (SPLIT) umma_f16(d0, ad_lo, bd, idesc_k1, 1u);
              }
              umma_f16(d0 + (k - 1u) * acc_cols, ad_hi, bd_top, idesc_1, 0u);
              if (SPLIT) umma_f16(d0 + (k - 1u) * acc_cols, ad_lo, bd_top, idesc_1, 1u);
            }
          } else if (leader) {
            umma_f16(d0, ad_hi, bd, idesc_k, 1u);
            if (SPLIT) umma_f16(d0, ad_lo, bd, idesc_k, 1u);
          }
        }
    }
  } else if (GEOM == GEOM_S1P || GEOM == GEOM_S1TP) {
    // Small planes (H <= 8, the 8^3 level): a 16 x 8 tile of ONE plane would be half empty, so the
    // 128 rows are 2 d-planes x 8 h x 8 w.  The descriptor needs ONE stride between 8-row groups,
    // i.e. a plane pitch of exactly 8 halo rows: every (kd, kh) pair is its own pipeline group that
    // loads an [2*TD planes][8 rows][10 w] box at (d + kd - 1, h + kh - 1); only kw is a start-address
    // shift.  Accumulator p = planes 2p, 2p + 1; g = kd*3 + kh; weight entries of the group = [kw].
    const uint32_t a_w1 = 10u | (1u << 14);
    const uint32_t lbo = (uint32_t)P.lbo16[0] << 16;
    const uint32_t ah = a_hi0 | lbo, al = a_lo0 | lbo;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int rw = GEOM == GEOM_S1P ? kw : 2 - kw;
      const uint32_t bo = b_w0 + (uint32_t)kw * b_ent;
      const uint32_t accum = (first && kw == 0) ? 0u : 1u;
#pragma unroll
      for (int p = 0; p < TD; ++p) {
        const uint32_t ao = (uint32_t)(rw + p * 160);
        mma_pair<SPLIT>(leader, tmem_acc0 + p * acc_cols, ah + ao, al + ao, a_w1, bo, b_w1, i2n, in_, accum);
      }
    }
  } else if (GEOM == GEOM_K1) {
    const uint32_t a_w1 = 8u | (1u << 14);
    const uint32_t lbo = (uint32_t)P.lbo16[0] << 16;
    const uint32_t ah = a_hi0 | lbo, al = a_lo0 | lbo;
#pragma unroll
    for (int p = 0; p < TD; ++p) {
      const uint32_t ao = (uint32_t)(p * 128);
      mma_pair<SPLIT>(leader, tmem_acc0 + p * acc_cols, ah + ao, al + ao, a_w1, b_w0, b_w1, i2n, in_, first ? 0u : 1u);
    }
  } else if (GEOM == GEOM_S2C4) {
    // Stride-2 conv over a COMPACT <= 4-channel input ([N][D][H][W][4] 16-bit values, 8 B per voxel: the network
    // input, the gradient of the head norm).  Two consecutive-w input voxels are one 16-byte K row, and the rows of
    // 8 consecutive OUTPUT voxels (input voxels 2w-1, 2w+1, ...) are 16 B apart -- a dense core matrix without any
    // parity splitting.  k-chunk 0 = taps (kw 0, kw 1), k-chunk 1 = (kw 2, a zero-weight fourth tap) = the same row
    // 16 B further (LBO = 16 B: overlapping windows of ONE loaded halo row).  A group = kd; its three (kh) boxes
    // [16 rows][18 voxels] (input rows 2h+kh-1 through an h-parity tensor map) are three K = 16 MMAs: 9 per tile
    // instead of 15 tap pairs, on a quarter of the operand bytes of the 8-channel chunk layout.
    const uint32_t a_w1 = 9u | (1u << 14);                 // SBO = 144 B: next output row h
    const uint32_t lbo = 1u << 16;                         // LBO = 16 B
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const uint32_t ao = (uint32_t)(kh * 144);            // 2304-byte boxes
      const uint32_t bo = b_w0 + (uint32_t)kh * b_ent;
      const uint32_t accum = (first && g == 0 && kh == 0) ? 0u : 1u;
      mma_pair<SPLIT>(leader, tmem_acc0, (a_hi0 + ao) | lbo, (a_lo0 + ao) | lbo, a_w1, bo, b_w1, i2n, in_, accum);
    }
  } else if (GEOM == GEOM_S2) {
    // parity sub-tiles [ph][pw] at fixed 128-aligned offsets: 16x8, 16x9, 17x8, 17x9 voxels
    constexpr int off16[4] = {0, 4096 / 16, 8704 / 16, 13056 / 16};
    if (P.s2pair) {
      // ONE-chunk inputs (stem: 4 channels; dgrad of the head convT: 3): half of every K = 16 MMA
      // would multiply the all-zero second k-chunk.  Two taps of the same parity class differ only by
      // an address offset (one voxel, or one sub-tile row), so the SECOND k-chunk of the descriptor is
      // pointed at the second tap's window (LBO = that offset) and the weights of the pair are packed as
      // the two k-chunks of one entry: 5 MMAs per kd instead of 9.  Entries (layout.s2_pairs):
      // {class m, rh, rw, offset of the partner in 16 B units (0 = single tap (1,1), -1 = one row)}
      constexpr int kPair[5][4] = {{3, 0, 0, 1}, {3, 1, 0, 1}, {2, 0, 0, -1}, {1, 0, 0, 1}, {0, 0, 0, 0}};
#pragma unroll
      for (int e = 0; e < 5; ++e) {
        const int m = kPair[e][0], wx = (m & 1) ? 9 : 8;
        const uint32_t a_w1 = (uint32_t)wx | (1u << 14);
        const uint32_t dl = kPair[e][3] == 0 ? (uint32_t)P.lbo16[m] : (kPair[e][3] < 0 ? (uint32_t)wx : (uint32_t)kPair[e][3]);
        const uint32_t ao = (uint32_t)(off16[m] + kPair[e][1] * wx + kPair[e][2]);
        const uint32_t bo = b_w0 + (uint32_t)e * b_ent;
        const uint32_t accum = (first && g == 0 && e == 0) ? 0u : 1u;
        mma_pair<SPLIT>(leader, tmem_acc0, (a_hi0 + ao) | (dl << 16), (a_lo0 + ao) | (dl << 16), a_w1, bo, b_w1, i2n, in_,
                        accum);
      }
      return;
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ph = kh != 1, pw = kw != 1, rh = kh == 2, rw = kw == 2, m = ph * 2 + pw;
        const int wx = pw ? 9 : 8;
        const uint32_t a_w1 = (uint32_t)wx | (1u << 14);
        const uint32_t lbo = (uint32_t)P.lbo16[m] << 16;
        const uint32_t ao = (uint32_t)(off16[m] + rh * wx + rw);
        const uint32_t bo = b_w0 + (uint32_t)(kh * 3 + kw) * b_ent;
        const uint32_t accum = (first && g == 0 && kh == 0 && kw == 0) ? 0u : 1u;
        mma_pair<SPLIT>(leader, tmem_acc0, (a_hi0 + ao) | lbo, (a_lo0 + ao) | lbo, a_w1, bo, b_w1, i2n, in_, accum);
      }
  } else {
    // GEOM_T2: group 0 = input plane d0 (kd = 1 -> even, kd = 2 -> odd out planes), group 1 = d0+1
    // (kd = 0).  Output-parity accumulator a = qd*4 + qh*2 + qw.  Taps that read the SAME shifted A
    // tile feed different parity accumulators, so they are issued as ONE MMA over a stack of their
    // weights (N = k * acc_cols, accumulators contiguous in TMEM): 14 A-tile reads per 16-channel
    // block instead of 27.  The 128x16 A tile read (4 KB at 128 B/clk) is what bounds these small-N
    // MMAs.  With split planes both A_hi and A_lo multiply the whole [B_hi | B_lo] stack (the extra
    // lo*lo term is below fp32 resolution).  Stack tables mirror layout.t2_stacks().
    const uint32_t a_w1 = 9u | (1u << 14);
    const uint32_t lbo = (uint32_t)P.lbo16[0] << 16;
    const uint32_t ah = a_hi0 | lbo, al = a_lo0 | lbo;
    const uint32_t maxk = 256u / acc_cols;  // accumulators one MMA may span (N <= 256)
    constexpr int kG0[9][4] = {{0, 8, 0, 0}, {2, 2, 1, 0}, {6, 2, 1, 0}, {1, 1, 0, 1}, {3, 1, 0, 1},
                               {5, 1, 0, 1}, {7, 1, 0, 1}, {3, 1, 1, 1}, {7, 1, 1, 1}};  // {acc0, k, jh, jw}
    constexpr int kG1[5][4] = {{4, 4, 0, 0}, {6, 2, 1, 0}, {5, 1, 0, 1}, {7, 1, 0, 1}, {7, 1, 1, 1}};
    auto stack = [&](int acc0, int k, int jh, int jw, uint32_t tap_off, bool overwrite) {
      const uint32_t ao = (uint32_t)jh * (uint32_t)P.t2_jh16 + (uint32_t)jw;  // two-plane tiles: the jh = 1 rows are a second box
      const uint32_t lbo_b = ((uint32_t)k * acc_cols) << 16;          // rows of the whole stack per k-chunk
      const uint32_t sbase = (bsrc >> 4) + tap_off * b_ent;
      const uint32_t kk = (uint32_t)k > maxk ? maxk : (uint32_t)k;
      for (uint32_t part = 0; part * kk < (uint32_t)k; ++part) {
        const uint32_t n = kk * acc_cols;
        const uint32_t idesc = P.idesc0 | ((n >> 3) << 17);
        const uint32_t d = tmem_acc0 + ((uint32_t)acc0 + part * kk) * acc_cols;
        const uint32_t bw0 = (sbase + part * kk * acc_cols) | lbo_b;   // 16-byte rows: kk accumulators further
        const uint64_t bd = ((uint64_t)b_w1 << 32) | bw0;
        const uint64_t ad_hi = ((uint64_t)a_w1 << 32) | (ah + ao), ad_lo = ((uint64_t)a_w1 << 32) | (al + ao);
        if (leader) {
          umma_f16(d, ad_hi, bd, idesc, overwrite ? 0u : 1u);
          if (SPLIT) umma_f16(d, ad_lo, bd, idesc, 1u);
        }
      }
    };
    if (g == 0) {
      uint32_t off = 0;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        stack(kG0[i][0], kG0[i][1], kG0[i][2], kG0[i][3], off, first && i == 0);
        off += (uint32_t)kG0[i][1];
      }
    } else {
      uint32_t off = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        stack(kG1[i][0], kG1[i][1], kG1[i][2], kG1[i][3], off, false);
        off += (uint32_t)kG1[i][1];
      }
    }
  }
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (halving butterfly): on
// return lane l holds the warp total of value (l >> 1); both lanes of a pair hold the same number.
__device__ __forceinline__ float warp_reduce16(const float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? a8[i] : a8[i + 4], keep = b3 ? a8[i + 4] : a8[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? a4[i] : a4[i + 2], keep = b2 ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? a2[0] : a2[1], keep = b1 ? a2[1] : a2[0];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

}  // namespace tta
// HOST-side description of one fused norm-backward segment (include/tta_b200.h: tta_norm_bwd_seg)
struct tta_norm_bwd_seg {
  int c8_begin, c8_count, relu, pad;
  const float* y;
  long long y_n_stride;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* beta;
  float* partial;
};
namespace tta {

struct WorkItem {
  int n, nt, ks, w0, h0, d0, cb0, nit;
};
// item = ((n * tiles + tile) * n_ntiles + nt) * ksplit + ks : neighbouring CTAs share the A tile in L2
// x / d through the host-made reciprocal (0 encodes d == 1)
__device__ __forceinline__ int fast_div(int x, unsigned mg) { return mg ? (int)__umulhi((unsigned)x, mg) : x; }

__device__ __forceinline__ WorkItem decode_item(const TcParams& P, int item) {
  // divisions by launch constants as multiply-high with host-made reciprocals: every role decodes
  // every item, and seven hardware-emulated integer divisions (~100 dependent cycles each) were a
  // measurable part of the per-item critical path of the MMA-light layers
  WorkItem w;
  const int per_tile = P.n_ntiles * P.ksplit;
  int t = fast_div(item, P.mg_per_tile);
  const int ntks = item - t * per_tile;
  const int tiles = P.tiles_w * P.tiles_h * P.tiles_d;
  w.n = fast_div(t, P.mg_tiles);
  t -= w.n * tiles;
  w.nt = fast_div(ntks, P.mg_ksplit);
  w.ks = ntks - w.nt * P.ksplit;
  const int row = fast_div(t, P.mg_tiles_w);      // t / tiles_w
  const int tw = t - row * P.tiles_w;
  const int tdi = fast_div(row, P.mg_tiles_h);    // t / (tiles_w * tiles_h)
  const int th = row - tdi * P.tiles_h;
  w.w0 = tw * 8;
  w.h0 = th * 16;
  w.d0 = tdi * P.td;
  w.cb0 = w.ks * P.cb_per_split;
  const int cb1 = min(P.ncblk, w.cb0 + P.cb_per_split);
  w.nit = (cb1 - w.cb0) * P.ngroups;
  return w;
}

template <int GEOM, int TD, int SPLIT, int STATS>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_bres;
  __shared__ TcGroup grp_s[kMaxGroups];  // per-lane indexed by the producer (constant bank would serialise)
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[128];
  __shared__ float stat_s[STATS ? 8 : 1][STATS ? 16 * 16 : 1];
  __shared__ float ncst_s[STATS == 2 ? 4 : 1][STATS == 2 ? 128 : 1];  // mean, rstd, gamma, beta of the n-tile's channels  // [epilogue warp][chunk of the n-tile][16]: fused norm statistics

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();  // the next kernel's CTAs may be scheduled (they block in their own pdl_wait)

  for (int i = threadIdx.x; i < (int)(sizeof(TcGroup) * kMaxGroups / 4); i += kTcThreads)
    reinterpret_cast<int*>(grp_s)[i] = reinterpret_cast<const int*>(P.grp)[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.nstages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), P.cl2 == 2 ? 2 : 1);   // pair mode: this CTA's and the peer's MMAs release a stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_acc_full[b]), 1);
      mbar_init(smem_u32(&bar_acc_empty[b]), 8);  // one arrival per epilogue warp
    }
    mbar_init(smem_u32(&bar_bres), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (P.single_chunk) {
    // 1-chunk inputs (stem, head): the second k-chunk of every A tile is never loaded -> zero it once
    const int n16 = (P.nstages * P.stage_bytes) >> 4;
    for (int i = threadIdx.x; i < n16; i += kTcThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (P.cl2) cluster_sync_all();   // the peer's barriers exist before anything is multicast into them
  // pair mode: CTA r of pair p takes tile 2*tpair + r of every (tile pair, n-tile, split) work unit
  const int it_first = P.cl2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int it_stride = P.cl2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int it_count = P.cl2 ? P.pair_items : P.work_items;
  const int pair_r = (int)(blockIdx.x & 1);
  auto item_of = [&](int q) -> int {
    if (!P.cl2) return q;
    const int per_tile = P.n_ntiles * P.ksplit;
    const int tpair = fast_div(q, P.mg_per_tile);
    return (2 * tpair + pair_r) * per_tile + (q - tpair * per_tile);
  };
  // everything above (barriers, TMEM allocation, smem clear) overlapped the previous kernel's tail;
  // activations / gradients written by it are only touched from here on
  pdl_wait();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t acc_cols = SPLIT ? 2u * P.ntile : (uint32_t)P.ntile;
  const uint32_t buf_cols = (uint32_t)P.nacc * acc_cols;

  if (warp == 0 || warp == 10) {
    // ===================== TMA producers (warp 0: even ring positions, warp 10: odd) =====================
    // Two warps alternate over the stage ring: filling a stage is a serial chain of address
    // arithmetic + per-copy operand broadcasts, and for layers with few MMAs per stage one warp
    // could not keep the tensor pipe fed.
    // (parity waits only order a waiter that is at most one phase ahead: with a single stage the
    // idle producer would skip ahead and see a stale "completed" parity, so one stage = one producer)
    const bool two_prod = P.nstages >= 2;
    const int my_par = warp == 0 ? 0 : 1;
    // The whole warp walks the stage ring; every lane owns ONE tensor copy of a stage
    // (load x k-chunk x plane, <= 16 per stage) and issues it itself, so a stage costs one pass of
    // uniform control flow instead of a serial per-copy loop in a single thread (measured on the
    // stride-2 stem: the MMA warp sat 59 % of the time on the full barrier while lane 0 spent
    // ~300 cycles per copy on index arithmetic and constant-bank lookups).
    if (P.b_res && lane == 0 && warp == 0) {
      // weights never change: one copy per CTA lifetime instead of one per stage
      const int nblob = P.b_nblob;
      mbar_expect_tx(smem_u32(&bar_bres), (uint32_t)(nblob * P.b_blob_bytes));
      for (int b = 0; b < nblob; ++b)
        bulk_load(smem_base + P.b_res_off + b * P.b_blob_bytes, P.wpacked + (long long)b * P.b_blob_bytes,
                  (uint32_t)P.b_blob_bytes, smem_u32(&bar_bres));
    }
    constexpr int kPlanes = SPLIT ? 2 : 1;
    const int my_plane = lane % kPlanes;
    const int my_kc = (lane / kPlanes) & 1;
    const int my_l = lane / (kPlanes * 2);
    // stage-ring position (s, phase) continues across work items; (g, cb) walk the groups of an
    // item -- all advanced by compare-and-wrap: an integer division costs ~100 dependent cycles and
    // this loop is the critical path of every layer with few MMAs per stage
    int s = 0, ph = 0, par = 0;
    for (int q = (two_prod || warp == 0) ? it_first : it_count; q < it_count; q += it_stride) {
      const int item = item_of(q);
      const WorkItem wi = decode_item(P, item);
      int g = 0, cb = wi.cb0;
      for (int it = 0; it < wi.nit; ++it, par ^= 1) {
        if (two_prod && par != my_par) {  // the other producer's stage: only advance the counters
          if (++g == P.ngroups) { g = 0; ++cb; }
          if (++s == P.nstages) { s = 0; ph ^= 1; }
          continue;
        }
        if (lane == 0) mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
        __syncwarp();
        const TcGroup& G = grp_s[g];
        const uint32_t full = smem_u32(&bar_full[s]);
        const uint32_t stage = smem_base + s * P.stage_bytes;
        const int c4 = wi.n * P.c8_view + cb * 2;
        const int nkc = (cb * 2 + 1 < P.c8in || !P.single_chunk) ? 2 : 1;
        const uint32_t bbytes = P.b_res ? 0u : (uint32_t)G.nmma * P.b_entry_bytes;
        if (lane == 0) {
          mbar_expect_tx(full, (uint32_t)((SPLIT ? G.tx_bytes : G.tx_bytes / 2) * nkc) + bbytes);
          if (!P.b_res) {
            // every (nt, cb, g) blob reserves gmax entries; only this group's nmma entries are copied
            const long long blob = ((long long)wi.nt * P.ncblk + cb) * P.ngroups + g;
            if (P.cl2 == 2) {   // this CTA's half of the blob, delivered to both CTAs of the pair
              const uint32_t hb = bbytes >> 1;
              bulk_load_mc2(stage + P.b_off + pair_r * hb, P.wpacked + blob * P.b_blob_bytes + pair_r * hb, hb, full);
            } else {
              bulk_load(stage + P.b_off, P.wpacked + blob * P.b_blob_bytes, bbytes, full);
            }
          }
        }
        __syncwarp();
        if (my_l < G.nloads && my_kc < nkc) {
          const TcLoad& L = G.ld[my_l];
          const int cw = wi.w0 + L.dw, chh = wi.h0 + L.dh, cd = wi.d0 * P.d_mul + L.dd;
          const uint32_t dst = stage + L.smem_off + my_kc * L.chunk_pitch + my_plane * P.a_plane_bytes;
          if (GEOM == GEOM_S2C4) {
            // map = h parity of the input row; coordinates (element of the W*4 row, h / 2, d, n)
            // (the box starts at the EVEN voxel 2*w0 - 2: TMA box rows and UMMA operand rows must start on 16 bytes)
            tma_load_4d(dst, &P.amap[L.map * 2 + my_plane], full, (2 * wi.w0 - 2) * 4, wi.h0 + L.dh, cd, wi.n);
          } else if (GEOM == GEOM_S2) {
            const CUtensorMap* m = &P.amap[L.map * 2 + my_plane];
            if (P.s2_rows) tma_load_4d(dst, m, full, cw * 8, chh, cd, c4 + my_kc);
            else tma_load_5d(dst, m, full, 0, cw, chh, cd, c4 + my_kc);
          } else {
            tma_load_4d(dst, &P.amap[my_plane], full, cw * 8, chh, cd, c4 + my_kc);
          }
        }
        if (++g == P.ngroups) { g = 0; ++cb; }
        if (++s == P.nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs uniform code, one elected lane issues) =====
    const uint32_t leader = elect_one();
    uint32_t local = 0;
    int s = 0, ph = 0;
    if (P.b_res) mbar_wait(smem_u32(&bar_bres), 0);  // resident weights have landed
    for (int q = it_first; q < it_count; q += it_stride, ++local) {
      const int item = item_of(q);
      const WorkItem wi = decode_item(P, item);
      const uint32_t buf = P.nbuf == 2 ? (local & 1u) : 0u, use = P.nbuf == 2 ? (local >> 1) : local;
      mbar_wait(smem_u32(&bar_acc_empty[buf]), (use & 1u) ^ 1u);  // epilogue drained this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc0 = tmem_base + buf * buf_cols;
      int g = 0;
      uint32_t bres = smem_base + P.b_res_off + (uint32_t)wi.cb0 * (uint32_t)P.b_cb_bytes;
      for (int it = 0; it < wi.nit; ++it) {
        mbar_wait(smem_u32(&bar_full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage = smem_base + s * P.stage_bytes;
        const uint32_t bsrc = P.b_res ? bres + (uint32_t)g * (uint32_t)P.b_g_bytes : stage + P.b_off;
        constexpr bool kStacked = GEOM == GEOM_S1K || GEOM == GEOM_S1TK;
        issue_group<GEOM, TD, SPLIT>(P, leader, g, stage, bsrc, acc0, kStacked ? it < P.ngroups : it == 0);
        __syncwarp();
        if (leader) {  // frees the smem stage when these MMAs retire (pair mode: in both CTAs -- the peer's
                       // half-blob copy lands in this CTA's stage too)
          if (P.cl2 == 2) umma_commit_mc2(smem_u32(&bar_empty[s])); else umma_commit(smem_u32(&bar_empty[s]));
        }
        if (++g == P.ngroups) { g = 0; bres += P.b_cb_bytes; }
        if (++s == P.nstages) { s = 0; ph ^= 1; }
      }
      if (leader) umma_commit(smem_u32(&bar_acc_full[buf]));  // accumulators of this item complete
      __syncwarp();
    }
  } else if (warp < 10) {
    // ===================== epilogue (8 warps: two per TMEM lane quarter) =====================
    // Small-K layers (stem, every single-product dgrad) do little MMA work per 128 x NT tile, so the
    // TMEM -> registers -> HBM drain is their critical path: measured on the stride-2 stem, four
    // epilogue warps (one per scheduler, nothing to hide latency with) were busy 80 % of the time
    // while the MMA warp idled.  Two warps per lane quarter split the (accumulator, 16-column)
    // units of an item between them.
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;       // which of the quarter's two warps
    const int row = q * 32 + lane;
    const int hh = P.pl2 ? (row >> 3) & 7 : row >> 3, ww = row & 7;
    const int dr = P.pl2 ? row >> 6 : 0;    // small-plane tiles: rows 64..127 are the second d-plane
    const int et = threadIdx.x - 64;        // 0..255
    const int ew = warp - 2;                // 0..7: statistics slot
    const long long Vo = (long long)P.Do * P.Ho * P.Wo;
    const int nchunks = P.ntile >> 3;
    const int n16 = P.ntile >> 4;
    const int nunits = P.nacc * n16;
    const bool rmw = P.accumulate && P.ksplit == 1;
    uint32_t local = 0;
    // ---- fused norm statistics (forward convs that feed a norm, no split-K): every warp adds its
    // 32 rows into a private smem slot; when the CTA moves on to another (n, n-tile) the eight slots
    // are summed in a fixed order into this CTA's own global slot -> deterministic, no atomics
    constexpr bool do_stats = STATS != 0;
    constexpr int kSlots = STATS ? 8 : 1;
    int st_n = -1, st_nt = -1;
    auto stats_flush = [&]() {
      if (!STATS) return;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int t = et; t < nchunks * 16; t += 256) {
        const int gchunk = st_nt * nchunks + (t >> 4);
        float tot = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < kSlots; ++w8) {
          tot += stat_s[w8][t];
          stat_s[w8][t] = 0.f;
        }
        if (STATS == 1) {
          if (gchunk < P.stats_c8)
            P.stats[(((long long)st_n * P.stats_c8 + gchunk) * gridDim.x + blockIdx.x) * 16 + (t & 15)] += tot;
        } else {
          for (int sg = 0; sg < P.nseg; ++sg) {
            const TcParams::BwdSeg& S = P.seg[sg];
            if (gchunk >= S.c8_begin && gchunk < S.c8_end)
              S.partial[(((long long)st_n * (S.c8_end - S.c8_begin) + (gchunk - S.c8_begin)) * gridDim.x + blockIdx.x) * 16 +
                        (t & 15)] += tot;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    if (do_stats) {
      for (int t = et; t < kSlots * 256; t += 256) (&stat_s[0][0])[t] = 0.f;
      if (STATS == 1) {
        for (int t = et; t < P.n_batch * P.stats_c8 * 16; t += 256)
          P.stats[((long long)(t >> 4) * gridDim.x + blockIdx.x) * 16 + (t & 15)] = 0.f;
      } else {
        for (int sg = 0; sg < P.nseg; ++sg) {
          const TcParams::BwdSeg& S = P.seg[sg];
          for (int t = et; t < P.n_batch * (S.c8_end - S.c8_begin) * 16; t += 256)
            S.partial[((long long)(t >> 4) * gridDim.x + blockIdx.x) * 16 + (t & 15)] = 0.f;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    for (int q = it_first; q < it_count; q += it_stride, ++local) {
      const int item = item_of(q);
      const WorkItem wi = decode_item(P, item);
      const uint32_t buf = P.nbuf == 2 ? (local & 1u) : 0u, use = P.nbuf == 2 ? (local >> 1) : local;
      if (do_stats && (wi.n != st_n || wi.nt != st_nt)) {
        if (st_n >= 0) stats_flush();
        st_n = wi.n;
        st_nt = wi.nt;
      }
      // bias of this n-tile -> smem (only split 0 adds it)
      asm volatile("bar.sync 1, 256;" ::: "memory");  // previous item's readers are done with bias_s
      if (et < P.ntile) {
        const int c = wi.nt * P.ntile + et;
        bias_s[et] = (P.bias != nullptr && wi.ks == 0 && c < P.C8out * 8) ? P.bias[c] : 0.f;
        if (STATS == 2) {
          float mu = 0.f, rs = 0.f, ga = 0.f, be = 0.f;
          for (int sg = 0; sg < P.nseg; ++sg) {
            const TcParams::BwdSeg& S = P.seg[sg];
            if ((c >> 3) >= S.c8_begin && (c >> 3) < S.c8_end) {
              const int cl = c - S.c8_begin * 8, Cn = (S.c8_end - S.c8_begin) * 8;
              mu = S.mean[wi.n * Cn + cl];
              rs = S.rstd[wi.n * Cn + cl];
              ga = S.gamma[cl];
              be = S.beta[cl];
            }
          }
          ncst_s[0][STATS == 2 ? et : 0] = mu;
          ncst_s[STATS == 2 ? 1 : 0][STATS == 2 ? et : 0] = rs;
          ncst_s[STATS == 2 ? 2 : 0][STATS == 2 ? et : 0] = ga;
          ncst_s[STATS == 2 ? 3 : 0][STATS == 2 ? et : 0] = be;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(smem_u32(&bar_acc_full[buf]), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tbase = tmem_base + buf * buf_cols + ((uint32_t)(q * 32) << 16);
      // units = (accumulator, 16-column block); this warp takes every second one, two per round so
      // that the read-modify-write variant (gradient fan-in) has four 32-byte loads in flight before
      // it touches TMEM
      for (int u0 = half; u0 < nunits; u0 += 4) {
        float4 old[2][2][2];
        float yv[STATS == 2 ? 2 : 1][STATS == 2 ? 2 : 1][8];  // conv results of the norm layer (backward statistics)
        bool ylive[2][2];
        float* obase[2];
        bool valid[2];
        int c16s[2];
        uint32_t tacc[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int u = u0 + 2 * k;
          const int acc = u < nunits ? u / n16 : 0;
          c16s[k] = u < nunits ? u - acc * n16 : 0;
          const int od = P.out_mul * (wi.d0 + P.acc_pd[acc] + dr) + P.acc_qd[acc];
          const int oh = P.out_mul * (wi.h0 + hh) + P.acc_qh[acc];
          const int ow = P.out_mul * (wi.w0 + ww) + P.acc_qw[acc];
          valid[k] = u < nunits && od < P.Do && oh < P.Ho && ow < P.Wo;
          const long long vox = ((long long)od * P.Ho + oh) * P.Wo + ow;
          obase[k] = P.out + (long long)wi.n * P.out_ns + vox * 8;
          tacc[k] = tbase + acc * acc_cols;
          if (STATS == 2) {
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              const int co_chunk = wi.nt * nchunks + c16s[k] * 2 + hlf;
              const float* ysrc = nullptr;
              for (int sg = 0; sg < P.nseg; ++sg) {
                const TcParams::BwdSeg& S = P.seg[sg];
                if (co_chunk >= S.c8_begin && co_chunk < S.c8_end)
                  ysrc = S.y + (long long)wi.n * S.y_ns + ((long long)(co_chunk - S.c8_begin) * Vo + vox) * 8;
              }
              ylive[k][hlf] = valid[k] && ysrc != nullptr;
              if (ylive[k][hlf]) {
                load_f32x8(ysrc, yv[k][hlf]);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) yv[k][hlf][i] = 0.f;
              }
            }
          }
          if (rmw) {
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              const int co_chunk = wi.nt * nchunks + c16s[k] * 2 + hlf;
              if (valid[k] && co_chunk < P.C8out) {
                const float* src = obase[k] + (long long)co_chunk * Vo * 8;
                float ov[8];
                if (P.out_f16) {
                  const U16x8 h = *reinterpret_cast<const U16x8*>(reinterpret_cast<const uint16_t*>(P.out) + (src - P.out));
#pragma unroll
                  for (int i = 0; i < 8; ++i) ov[i] = u16_to_f32<TTA_F16>(h.v[i]);
                } else {
                  load_f32x8(src, ov);  // one 256-bit load per 32-byte voxel-chunk
                }
                old[k][hlf][0] = make_float4(ov[0], ov[1], ov[2], ov[3]);
                old[k][hlf][1] = make_float4(ov[4], ov[5], ov[6], ov[7]);
              } else {
                old[k][hlf][0] = old[k][hlf][1] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (u0 + 2 * k >= nunits) break;  // warp-uniform
          const int c16 = c16s[k];
          uint32_t ra[16], rb[16];
          tmem_ld16_nowait(tacc[k] + c16 * 16, ra);            // hi*hi + lo*hi columns
          if (SPLIT) {
            tmem_ld16_nowait(tacc[k] + P.ntile + c16 * 16, rb);  // hi*lo columns
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) rb[i] = 0u;
          }
          tmem_ld_wait();
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            const int co_chunk = wi.nt * nchunks + c16 * 2 + hlf;
            const bool live = valid[k] && co_chunk < P.C8out;
            const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[c16 * 16 + hlf * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[c16 * 16 + hlf * 8 + 4]);
            const int o = hlf * 8;
            float4 r0 = make_float4(__uint_as_float(ra[o + 0]) + __uint_as_float(rb[o + 0]) + b0.x,
                                    __uint_as_float(ra[o + 1]) + __uint_as_float(rb[o + 1]) + b0.y,
                                    __uint_as_float(ra[o + 2]) + __uint_as_float(rb[o + 2]) + b0.z,
                                    __uint_as_float(ra[o + 3]) + __uint_as_float(rb[o + 3]) + b0.w);
            float4 r1 = make_float4(__uint_as_float(ra[o + 4]) + __uint_as_float(rb[o + 4]) + b1.x,
                                    __uint_as_float(ra[o + 5]) + __uint_as_float(rb[o + 5]) + b1.y,
                                    __uint_as_float(ra[o + 6]) + __uint_as_float(rb[o + 6]) + b1.z,
                                    __uint_as_float(ra[o + 7]) + __uint_as_float(rb[o + 7]) + b1.w);
            if (live) {
              float* dst = obase[k] + (long long)co_chunk * Vo * 8;
              if (P.ksplit > 1) {
                // split-K partial sums meet in HBM (destination pre-zeroed unless accumulating)
                atomicAdd(reinterpret_cast<float4*>(dst), r0);
                atomicAdd(reinterpret_cast<float4*>(dst + 4), r1);
              } else {
                if (rmw) {
                  const float4 o0 = old[k][hlf][0], o1 = old[k][hlf][1];
                  r0.x += o0.x; r0.y += o0.y; r0.z += o0.z; r0.w += o0.w;
                  r1.x += o1.x; r1.y += o1.y; r1.z += o1.z; r1.w += o1.w;
                }
                const float rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                if (P.out_f16) {
                  U16x8 h;
#pragma unroll
                  for (int i = 0; i < 8; ++i) h.v[i] = f32_to_u16<TTA_F16>(rv[i]);
                  *reinterpret_cast<U16x8*>(reinterpret_cast<uint16_t*>(P.out) + (dst - P.out)) = h;
                } else {
                  store_f32x8(dst, rv);  // whole 32-byte sector in one store
                }
              }
            }
            if (STATS == 2) {
              // norm backward: dz = g * relu'(gamma*xhat + beta); sums of dz and dz*xhat (warp-uniform branch)
              bool in_seg = false;
              int relu = 0;
              for (int sg = 0; sg < P.nseg; ++sg)
                if (co_chunk >= P.seg[sg].c8_begin && co_chunk < P.seg[sg].c8_end) {
                  in_seg = true;
                  relu = P.seg[sg].relu;
                }
              if (in_seg) {
                const float gv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                const bool yl = ylive[STATS == 2 ? k : 0][STATS == 2 ? hlf : 0];
                float sv[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int cc = c16 * 16 + hlf * 8 + i;
                  const float xh = (yv[STATS == 2 ? k : 0][STATS == 2 ? hlf : 0][i] - ncst_s[0][STATS == 2 ? cc : 0]) *
                                   ncst_s[STATS == 2 ? 1 : 0][STATS == 2 ? cc : 0];
                  const float z = fmaf(xh, ncst_s[STATS == 2 ? 2 : 0][STATS == 2 ? cc : 0], ncst_s[STATS == 2 ? 3 : 0][STATS == 2 ? cc : 0]);
                  const float dzv = (!yl || (relu && !(z > 0.f))) ? 0.f : gv[i];
                  sv[i] = dzv;
                  sv[8 + i] = dzv * xh;
                }
                const float tot = warp_reduce16(sv, lane);
                if (!(lane & 1)) stat_s[STATS ? ew : 0][(c16 * 2 + hlf) * 16 + (lane >> 1)] += tot;
              }
            }
            if (STATS == 1 && co_chunk < P.stats_c8) {  // warp-uniform: the whole warp reduces
              float sv[16];
              sv[0] = live ? r0.x : 0.f; sv[1] = live ? r0.y : 0.f; sv[2] = live ? r0.z : 0.f; sv[3] = live ? r0.w : 0.f;
              sv[4] = live ? r1.x : 0.f; sv[5] = live ? r1.y : 0.f; sv[6] = live ? r1.z : 0.f; sv[7] = live ? r1.w : 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) sv[8 + i] = sv[i] * sv[i];
              const float tot = warp_reduce16(sv, lane);
              if (!(lane & 1)) stat_s[STATS ? ew : 0][(c16 * 2 + hlf) * 16 + (lane >> 1)] += tot;
            }
          }
        }
      }
      // this warp's TMEM reads are complete: hand the accumulator buffer back to the MMA thread
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[buf]));
    }
    if (do_stats && st_n >= 0) stats_flush();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
  }
  if (P.cl2) cluster_sync_all();   // the peer may still arrive on this CTA's barriers until it is done too
}

// =====================================================================================================
// conv_t2s_kernel: transposed stride-2 3x3x3 conv with a TINY output-channel count (the UNet's head
// convT 64 -> 3) as ONE dense GEMM per input plane + a shared-memory col2im.
//
// The generic GEOM_T2 path needs 28 small-N MMAs per 16-channel block and tile (every shifted A tile x
// every parity class) and pads 3 output channels to 16: measured 156 .. 228 us for 5.4 GFLOP.  Here
//     P[v, (tap, co)] = sum_ci X[v, ci] * W[ci, tap, co]          (M = 128 input voxels, N = 27*Cout)
// is computed for every INPUT voxel of a 16(h) x 8(w) tile with 2 MMAs per 16-channel block
// (A_hi x [B_hi | B_lo], A_lo x B_hi), the 27*Cout partial products of a voxel go TMEM -> registers ->
// shared memory, and every output voxel gathers its 1 .. 8 contributions from the P rows of its input
// voxel and of the +1 neighbours in h, w (same tile: the tile's last row / column is halo) and d (next
// plane: a CTA walks a d-segment and keeps the P tiles of two consecutive planes):
//     out[2d+qd, 2h+qh, 2w+qw] = b + sum over axes { q = 0: (shift 0, k = 1);  q = 1: (0, 2), (1, 0) }.
// Roles: warp 0 TMA producer (one box [Cin/8 chunks][16 h][8 w] per plane and operand plane), warp 1
// MMA issue, warps 2..5 epilogue; TMEM accumulators and the P tiles are double buffered.
constexpr int kT2sThreads = 320;  // TMA producer, MMA issuer, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kT2sStages = 3;  // at most

struct T2sParams {
  CUtensorMap amap[2];
  int N, C8in, c8_view, nks;             // nks = Cin / 16 k-steps
  int D, H, W;                           // input dims (output = 2x)
  int tiles_h, tiles_w, dseg, seg_len, work_items;
  int np;                                // padded GEMM N = round16(27 * Cout)
  int stage_bytes, plane_bytes;          // plane_bytes = C8in * 2048 (one operand plane of a stage)
  int w_bytes, tmem_cols, nstages;
  unsigned idesc_2n, idesc_n;
  long long out_ns;
  const uint8_t* wpacked;
  const float* bias;
  float* out;
  float* stats;                          // [n][1][grid][16] or nullptr
  int cpv;                               // floats per output voxel: 8 (chunk layout) or 4 (compact, flags bit 14)
};

template <int COUT>
__global__ void __launch_bounds__(kT2sThreads, 1)
conv_t2s_kernel(const __grid_constant__ T2sParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[kT2sStages];
  __shared__ __align__(8) uint64_t bar_empty[kT2sStages];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red_s[8][8];
  constexpr int NC = 27 * COUT;          // real GEMM columns
  constexpr int PROW = NC | 1;           // odd row pitch of the P tiles: conflict-free column access

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.nstages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_acc_full[b]), 1);
      mbar_init(smem_u32(&bar_acc_empty[b]), 8);
    }
    mbar_init(smem_u32(&bar_w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t w_smem = smem_base + P.nstages * P.stage_bytes;
  float* p_tiles = reinterpret_cast<float*>(smem + P.nstages * P.stage_bytes + P.w_bytes);  // [2][128][PROW]
  const int per_n = P.dseg * P.tiles_h * P.tiles_w;

  auto decode = [&](int item, int& n, int& ds, int& de, int& h0, int& w0) {
    n = item / per_n;
    int t = item - n * per_n;
    const int sg = t / (P.tiles_h * P.tiles_w);
    t -= sg * P.tiles_h * P.tiles_w;
    const int th = t / P.tiles_w, tw = t - th * P.tiles_w;
    ds = sg * P.seg_len;
    de = min(P.D, ds + P.seg_len);
    h0 = th * 15;
    w0 = tw * 7;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bar_w), (uint32_t)P.w_bytes);
      bulk_load(w_smem, P.wpacked, (uint32_t)P.w_bytes, smem_u32(&bar_w));
      int s = 0, ph = 0;
      for (int item = blockIdx.x; item < P.work_items; item += gridDim.x) {
        int n, ds, de, h0, w0;
        decode(item, n, ds, de, h0, w0);
        const int last = min(de, P.D - 1);   // plane de is the +1 halo of the segment (absent past the volume)
        for (int p = ds; p <= last; ++p) {
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
          const uint32_t full = smem_u32(&bar_full[s]);
          const uint32_t stage = smem_base + s * P.stage_bytes;
          mbar_expect_tx(full, 2u * (uint32_t)P.plane_bytes);
          tma_load_4d(stage, &P.amap[0], full, w0 * 8, h0, p, n * P.c8_view);
          tma_load_4d(stage + P.plane_bytes, &P.amap[1], full, w0 * 8, h0, p, n * P.c8_view);
          if (++s == P.nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t leader = elect_one();
    mbar_wait(smem_u32(&bar_w), 0);
    int s = 0, ph = 0;
    uint32_t step = 0;
    const uint32_t a_w1 = 8u | (1u << 14), b_w1 = 8u | (1u << 14);           // SBO = 128 B
    const uint32_t a_lbo = (2048u >> 4) << 16, b_lbo = ((2u * (uint32_t)P.np * 16u) >> 4) << 16;
    for (int item = blockIdx.x; item < P.work_items; item += gridDim.x) {
      int n, ds, de, h0, w0;
      decode(item, n, ds, de, h0, w0);
      const int last = min(de, P.D - 1);
      for (int p = ds; p <= last; ++p, ++step) {
        const uint32_t buf = step & 1u, use = step >> 1;
        mbar_wait(smem_u32(&bar_acc_empty[buf]), (use & 1u) ^ 1u);
        mbar_wait(smem_u32(&bar_full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage = smem_base + s * P.stage_bytes;
        const uint32_t d = tmem_base + buf * 2u * (uint32_t)P.np;
        for (int ks = 0; ks < P.nks; ++ks) {
          const uint32_t a_hi = ((stage + ks * 4096) >> 4) | a_lbo;
          const uint32_t a_lo = ((stage + P.plane_bytes + ks * 4096) >> 4) | a_lbo;
          const uint32_t b0 = ((w_smem + ks * (2 * 2 * P.np * 16)) >> 4) | b_lbo;
          const uint64_t ad_hi = ((uint64_t)a_w1 << 32) | a_hi, ad_lo = ((uint64_t)a_w1 << 32) | a_lo;
          const uint64_t bd = ((uint64_t)b_w1 << 32) | b0;
          if (leader) {
            umma_f16(d, ad_hi, bd, P.idesc_2n, ks == 0 ? 0u : 1u);   // [hi*hi | hi*lo]
            umma_f16(d, ad_lo, bd, P.idesc_n, 1u);                   // += lo*hi
          }
        }
        __syncwarp();
        if (leader) {
          umma_commit(smem_u32(&bar_empty[s]));
          umma_commit(smem_u32(&bar_acc_full[buf]));
        }
        __syncwarp();
        if (++s == P.nstages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> P tile (smem) -> col2im gather -> HBM =====================
    // two warps per TMEM lane quarter: both own the same 32 input voxels; warp `half` copies every second
    // 16-column block of the P rows and gathers the output planes of parity qd = half
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;                 // 0..7
    const int row = q * 32 + lane;           // input voxel of the tile = TMEM lane
    const int hh = row >> 3, ww = row & 7;
    const int et = threadIdx.x - 64;         // 0..255
    const int Do = 2 * P.D, Ho = 2 * P.H, Wo = 2 * P.W;
    float bias_r[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias_r[c] = P.bias ? P.bias[c] : 0.f;
    float s1[COUT], s2[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) s1[c] = s2[c] = 0.f;
    int st_n = -1;
    auto stats_flush = [&]() {
      // fixed-order reduction of the 128 epilogue threads into this CTA's slot of sample st_n
      float v[2 * COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        v[c] = warp_sum(s1[c]);
        v[COUT + c] = warp_sum(s2[c]);
        s1[c] = s2[c] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 2 * COUT; ++i) red_s[ew][i] = v[i];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et < 2 * COUT) {
        float tot = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) tot += red_s[w8][et];
        const int k = et < COUT ? et : 8 + (et - COUT);
        P.stats[(((long long)st_n * gridDim.x) + blockIdx.x) * 16 + k] += tot;
      }
    };
    if (P.stats) {
      for (int t = et; t < P.N * 16; t += 256) P.stats[((long long)(t >> 4) * gridDim.x + blockIdx.x) * 16 + (t & 15)] = 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    uint32_t step = 0;
    for (int item = blockIdx.x; item < P.work_items; item += gridDim.x) {
      int n, ds, de, h0, w0;
      decode(item, n, ds, de, h0, w0);
      if (P.stats && n != st_n) {
        if (st_n >= 0) stats_flush();
        st_n = n;
      }
      const bool mine = hh < 15 && ww < 7 && h0 + hh < P.H && w0 + ww < P.W;
      const int ih = h0 + hh, iw = w0 + ww;
      for (int p = ds; p <= de; ++p) {
        const bool have = p < P.D;             // plane p exists (p == D: the halo past the volume is zero)
        float* cur = p_tiles + (size_t)(p & 1) * 128 * PROW;
        if (have) {
          const uint32_t buf = step & 1u, use = step >> 1;
          ++step;
          mbar_wait(smem_u32(&bar_acc_full[buf]), use & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tb = tmem_base + buf * 2u * (uint32_t)P.np + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int c16 = half; c16 < (NC + 15) / 16; c16 += 2) {
            uint32_t ra[16], rb[16];
            tmem_ld16_nowait(tb + c16 * 16, ra);
            tmem_ld16_nowait(tb + P.np + c16 * 16, rb);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c16 * 16 + i < NC) cur[row * PROW + c16 * 16 + i] = __uint_as_float(ra[i]) + __uint_as_float(rb[i]);
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[buf]));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // P tile of plane p complete
        if (p > ds && mine) {
          const int d = p - 1;
          const float* lo_t = p_tiles + (size_t)(d & 1) * 128 * PROW;   // P of plane d
          const float* hi_t = cur;                                      // P of plane d + 1 (jd = 1)
          // the output-plane parity this warp gathers is a compile-time constant inside the lambda
          auto gather = [&](auto QD) {
            constexpr int qd = decltype(QD)::value;
#pragma unroll
          for (int qh = 0; qh < 2; ++qh) {
              float o[2][COUT];
#pragma unroll
              for (int qw = 0; qw < 2; ++qw) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) o[qw][c] = bias_r[c];
#pragma unroll
                for (int ad = 0; ad <= qd; ++ad)
#pragma unroll
                  for (int ah = 0; ah <= qh; ++ah)
#pragma unroll
                    for (int aw = 0; aw <= qw; ++aw) {
                      // shift j = a, tap k: q = 0 -> k = 1; q = 1 -> (j = 0, k = 2), (j = 1, k = 0)
                      const int kd = qd ? (ad ? 0 : 2) : 1, kh = qh ? (ah ? 0 : 2) : 1, kw = qw ? (aw ? 0 : 2) : 1;
                      if (ad && !have) continue;
                      const float* src = (ad ? hi_t : lo_t) + (row + ah * 8 + aw) * PROW + (kd * 9 + kh * 3 + kw) * COUT;
#pragma unroll
                      for (int c = 0; c < COUT; ++c) o[qw][c] += src[c];
                    }
              }
              const int od = 2 * d + qd, oh = 2 * ih + qh, ow = 2 * iw;
              float* dst = P.out + (long long)n * P.out_ns + (((long long)od * Ho + oh) * Wo + ow) * P.cpv;
#pragma unroll
              for (int qw = 0; qw < 2; ++qw) {
                float rv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) rv[c] = 0.f;
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                  rv[c] = o[qw][c];
                  s1[c] += o[qw][c];
                  s2[c] = fmaf(o[qw][c], o[qw][c], s2[c]);
                }
                if (P.cpv == 8) store_f32x8(dst + qw * 8, rv);
                else *reinterpret_cast<float4*>(dst + qw * 4) = make_float4(rv[0], rv[1], rv[2], rv[3]);
              }
            }
          };
          if (half == 0) gather(std::integral_constant<int, 0>{}); else gather(std::integral_constant<int, 1>{});
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // gathers done before the older P tile is overwritten
      }
    }
    if (P.stats && st_n >= 0) stats_flush();
    (void)Do;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------- host side
static int geom_of(int mode, int K, int stride) {
  if (K == 1 && stride == 1) return GEOM_K1;  // 1x1: conv and its dgrad are both plain GEMMs
  if (K != 3) return GEOM_NONE;
  if (mode == 0) return stride == 1 ? GEOM_S1 : (stride == 2 ? GEOM_S2 : GEOM_NONE);
  return stride == 1 ? GEOM_S1T : (stride == 2 ? GEOM_T2 : GEOM_NONE);
}

// n-tile: couts per work item.  Each accumulator takes 2*NT TMEM columns (stacked hi|lo halves).
static int ntile_of(int geom, int cout, int split) {
  const int c16 = (cout + 15) / 16 * 16;
  // T2: 8 parity accumulators x (split ? 2 : 1)*NT <= 512 TMEM columns
  const int cap = geom == GEOM_T2 ? (split ? 32 : 64) : 128;
  return c16 < cap ? c16 : cap;
}

static int round128(int x) { return (x + 127) / 128 * 128; }

// transposed stride-2 conv with <= 4 output channels: dense GEMM + col2im (conv_t2s_kernel)
static bool t2s_of(int geom, int cin, int cout, int split) {
  return geom == GEOM_T2 && split && cin % 16 == 0 && cin >= 16 && cin <= 64 && cout >= 1 && cout <= 4;
}

// stride-2 conv over a one-chunk input (<= 8 channels): taps paired into the two k-chunks of one MMA
static bool s2pair_of(int geom, int cin) { return geom == GEOM_S2 && cin <= 8; }

// kd-stacked stride-1 convs (GEOM_S1K / GEOM_S1TK, see issue_group): single n-tile, three accumulators
// within one MMA (3 * acc_cols <= 256) and all weights resident in shared memory.  A function of the
// layer alone, because the PACKED WEIGHT LAYOUT depends on it (layout.pack_weights_tc asks).
static bool stacked_of(int geom, int cin, int cout, int split) {
  if (geom != GEOM_S1 && geom != GEOM_S1T) return false;
  const int cout_pad = (cout + 7) / 8 * 8;
  const int nt = ntile_of(geom, cout_pad, split);
  if ((cout_pad + nt - 1) / nt != 1) return false;
  const int acc_cols = split ? 2 * nt : nt;
  if (3 * acc_cols > 256) return false;
  const int ncblk = (cin + 15) / 16;
  return (long long)ncblk * 27 * 2 * acc_cols * 16 <= 112 * 1024;
}

static int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace tta

using namespace tta;

extern "C" {

int tta_conv_tc_supported(int mode, int K, int stride, int cin, int cout) {
  (void)cin;
  (void)cout;
  return geom_of(mode, K, stride) != GEOM_NONE;
}

int tta_conv_tc_ntile(int mode, int K, int stride, int cout, int split) {
  return ntile_of(geom_of(mode, K, stride), cout, split);
}

int tta_conv_tc_stacked(int mode, int K, int stride, int cin, int cout, int split) {
  return stacked_of(geom_of(mode, K, stride), cin, cout, split) ? 1 : 0;
}

int tta_conv_tc_s2pair(int mode, int K, int stride, int cin) { return s2pair_of(geom_of(mode, K, stride), cin) ? 1 : 0; }

// 1 when a stride-2 conv may read a COMPACT <= 4-channel input ([N][D][H][W][4], flags bit 15): pack its weights
// with layout.pack_weights_tc(s2c4=True)
int tta_conv_tc_s2c4(int mode, int K, int stride, int cin) { return geom_of(mode, K, stride) == GEOM_S2 && cin <= 4 ? 1 : 0; }

int tta_conv_tc_t2s(int mode, int K, int stride, int cin, int cout, int split) {
  return t2s_of(geom_of(mode, K, stride), cin, cout, split) ? 1 : 0;
}

int tta_conv_tc_gmax(int mode, int K, int stride) {
  const int g = geom_of(mode, K, stride);
  return g == GEOM_K1 ? 1 : (g == GEOM_T2 ? 18 : 9);
}

int tta_conv_tc_ngroups(int mode, int K, int stride) {
  const int g = geom_of(mode, K, stride);
  return g == GEOM_K1 ? 1 : (g == GEOM_T2 ? 2 : 3);
}

// in: split planes view [N][C8in (pitch from in_ns)][Di][Hi][Wi][8];  out: fp32 view; wpacked from
// layout.pack_weights_tc.  flags bit0: force TD=1, bit1: no split-K (deterministic), bit2: one
// work item per CTA (non-persistent; testing), bit4: no resident weights (testing), bit3: the input planes of a stride-2 conv are stored
// w-parity-split ([N][C8][D][H][2][W/2][8]: even-w voxels of a row first, then the odd ones).
}  // extern "C"

// ---- transposed stride-2 conv with <= 4 output channels: dense GEMM + col2im (conv_t2s_kernel)

static int conv_t2s_impl(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns, int N, int C8in, int Di, int Hi,
                         int Wi, const void* wpacked, const float* bias, float* out, long long out_ns, int cout,
                         int Do, int Ho, int Wo, int accumulate, float* stats_ws, int stats_c8, int cpv, int* q_ksplit,
                         int* q_grid, int* q_nbuf, cudaStream_t stream) {
  const bool query = q_ksplit != nullptr;
  TTA_REQUIRE(Do == 2 * Di && Ho == 2 * Hi && Wo == 2 * Wi, "tta_conv_tc: transposed s2 output dims");
  TTA_REQUIRE(!accumulate, "tta_conv_tc: the small-Cout transposed conv does not accumulate");
  const long long Vi = (long long)Di * Hi * Wi;
  TTA_REQUIRE(in_ns % (Vi * 8) == 0, "tta_conv_tc: n_stride must be a whole number of channel chunks");
  T2sParams P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.C8in = C8in; P.c8_view = (int)(in_ns / (Vi * 8)); P.nks = C8in / 2;
  P.D = Di; P.H = Hi; P.W = Wi;
  P.tiles_h = (Hi + 14) / 15; P.tiles_w = (Wi + 6) / 7;
  P.np = (27 * cout + 15) / 16 * 16;
  P.plane_bytes = C8in * 2048;
  P.stage_bytes = 2 * P.plane_bytes;
  P.w_bytes = P.nks * 2 * 2 * P.np * 16;
  const int p_bytes = 2 * 128 * ((27 * cout) | 1) * 4;
  P.nstages = (227 * 1024 - 2048 - P.w_bytes - p_bytes) / P.stage_bytes;
  if (P.nstages > kT2sStages) P.nstages = kT2sStages;
  TTA_REQUIRE(P.nstages >= 2, "tta_conv_tc: small-Cout transposed conv does not fit shared memory (Cin %d, Cout %d)",
              C8in * 8, cout);
  P.tmem_cols = 32;
  while (P.tmem_cols < 4 * P.np) P.tmem_cols *= 2;   // two buffers of [main np | corr np]
  const int idesc0 = (1 << 4) | ((128 >> 4) << 24);  // fp16 operands, fp32 accumulate, M = 128
  P.idesc_n = (unsigned)(idesc0 | ((P.np >> 3) << 17));
  P.idesc_2n = (unsigned)(idesc0 | (((2 * P.np) >> 3) << 17));
  {
    // d-segments: every CTA walks a segment plane by plane (+1 halo plane); pick the count that
    // minimises (waves of CTAs) x (planes per segment + halo + fill)
    const long long cols = (long long)N * P.tiles_h * P.tiles_w;
    int best_len = Di;
    double best_cost = 1e30;
    for (int k = 1; k <= Di; ++k) {
      const int len = (Di + k - 1) / k, real = (Di + len - 1) / len;
      const long long waves = (cols * real + num_sms() - 1) / num_sms();
      const double cost = (double)waves * (len + 2.5);
      if (cost < best_cost) { best_cost = cost; best_len = len; }
    }
    P.seg_len = best_len;
    P.dseg = (Di + P.seg_len - 1) / P.seg_len;
    P.work_items = (int)(cols * P.dseg);
  }
  const int grid_x = P.work_items < num_sms() ? P.work_items : num_sms();
  if (getenv("TTA_TC_DEBUG"))
    fprintf(stderr, "tta_conv_tc: t2s N %d Cin %d Cout %d in %dx%dx%d | np %d stages %d x %d B | tiles %dx%d dseg %d x %d "
            "items %d waves %.2f\n", N, C8in * 8, cout, Di, Hi, Wi, P.np, P.nstages, P.stage_bytes, P.tiles_h, P.tiles_w,
            P.dseg, P.seg_len, P.work_items, (double)P.work_items / num_sms());
  if (query) {
    *q_ksplit = 1;
    *q_grid = grid_x;
    if (q_nbuf) *q_nbuf = 2;
    return TTA_OK;
  }
  TTA_REQUIRE(in_hi && in_lo && wpacked && out, "tta_conv_tc: null pointer");
  EncodeTiledFn enc = get_encode();
  TTA_REQUIRE(enc != nullptr, "tta_conv_tc: cuTensorMapEncodeTiled entry point not found");
  P.out_ns = out_ns; P.wpacked = (const uint8_t*)wpacked; P.bias = bias; P.out = out;
  P.cpv = cpv;
  P.stats = nullptr;
  if (stats_ws != nullptr && stats_c8 > 0) {
    TTA_REQUIRE(stats_c8 == 1, "tta_conv_tc: stats_c8 %d > 1 chunk", stats_c8);
    P.stats = stats_ws + 1024;
  }
  const cuuint64_t nc_extent = (cuuint64_t)((long long)(N - 1) * P.c8_view + C8in);
  const cuuint32_t es[4] = {1, 1, 1, 1};
  bool ok = true;
  for (int pl = 0; pl < 2; ++pl) {
    cuuint64_t gdim[4] = {(cuuint64_t)Wi * 8, (cuuint64_t)Hi, (cuuint64_t)Di, nc_extent};
    cuuint64_t gstr[3] = {(cuuint64_t)16 * Wi, (cuuint64_t)16 * Wi * Hi, (cuuint64_t)16 * Vi};
    cuuint32_t box[4] = {64, 16, 1, (cuuint32_t)C8in};
    ok = ok && enc(&P.amap[pl], CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)(pl ? in_lo : in_hi), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  TTA_REQUIRE(ok, "tta_conv_tc: cuTensorMapEncodeTiled failed (t2s, dims %d,%d,%d)", Di, Hi, Wi);
  const size_t smem = (size_t)P.nstages * P.stage_bytes + P.w_bytes + p_bytes + 1024;
#define TTA_T2S_LAUNCH(CO)                                                                                      \
  do {                                                                                                          \
    static bool configured = false;                                                                             \
    if (!configured) {                                                                                          \
      if (cudaFuncSetAttribute(conv_t2s_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                               227 * 1024 - 1024) != cudaSuccess)                                               \
        return tta_check_launch("tta_conv_tc(t2s cudaFuncSetAttribute)");                                       \
      configured = true;                                                                                        \
    }                                                                                                           \
    tta_launch(conv_t2s_kernel<CO>, dim3(grid_x), kT2sThreads, smem, stream, tta_pdl_family(8), P);            \
  } while (0)
  switch (cout) {
    case 1: TTA_T2S_LAUNCH(1); break;
    case 2: TTA_T2S_LAUNCH(2); break;
    case 3: TTA_T2S_LAUNCH(3); break;
    default: TTA_T2S_LAUNCH(4); break;
  }
#undef TTA_T2S_LAUNCH
  return tta_check_launch("tta_conv_tc(t2s)");
}

static int conv_tc_impl(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns, int in_dtype, int N, int C8in,
                        int Di, int Hi, int Wi, const void* wpacked, const float* bias, float* out, long long out_ns,
                        int C8out, int Do, int Ho, int Wo, int mode, int K, int stride, int accumulate, int flags,
                        float* stats_ws, int stats_c8, const tta_norm_bwd_seg* segs, int nsegs, int* q_ksplit, int* q_grid,
                        int* q_nbuf, cudaStream_t stream) {
  const int split = in_dtype == TTA_F16_HI ? 0 : 1;
  const bool query = q_ksplit != nullptr;  // shape the launch only: report split-K factor and grid
  // flags bit 16: `out` is one fp16 plane in the chunk layout (out_n_stride in ELEMENTS as always): the scaled
  // input gradient of a dgrad goes to the norm backward in 2 instead of 4 bytes.  fp16 has no vector atomics:
  // no split-K; gradient fan-in (accumulate) is a 16-byte read-modify-write
  const bool out_f16 = (flags & 65536) != 0;
  if (out_f16) flags |= 2;
  TTA_REQUIRE(!out_f16 || (stats_ws == nullptr && segs == nullptr && !(flags & 16384) && ((flags >> 8) & 7) == 0),
              "tta_conv_tc: flags bit 16 (fp16 result) excludes fused statistics and the small-Cout transposed kernel");
  TTA_REQUIRE(query || (in_hi && (in_lo || !split) && wpacked && out), "tta_conv_tc: null pointer");
  int geom = geom_of(mode, K, stride);
  TTA_REQUIRE(geom != GEOM_NONE, "tta_conv_tc: unsupported geometry mode=%d K=%d stride=%d", mode, K, stride);
  TTA_REQUIRE(in_dtype >= 0 && in_dtype <= 2, "tta_conv_tc: bad dtype");
  // small planes (8^3 level): two d-planes per 128-row tile (see issue_group); flags bit5 keeps the
  // one-plane tiles (testing: both must agree)
  // flags bits 8..10: real output-channel count (1..4) of a transposed stride-2 conv whose weights were
  // packed for the dense-GEMM + col2im kernel (layout.pack_weights_tc(t2s=True)); 0 = generic path
  // flags bit 14 (with bits 8..10): that kernel writes the COMPACT fp32 layout [N][D][H][W][4] (out_n_stride in
  // floats) instead of the 8-channel chunk layout -- for the fused full-resolution head, whose three tensors
  // would otherwise carry five pad channels through every pass
  const int t2s_small_cout = (flags >> 8) & 7;
  TTA_REQUIRE(!(flags & 16384) || t2s_small_cout > 0, "tta_conv_tc: flags bit 14 needs the small-Cout transposed kernel");
  TTA_REQUIRE(t2s_small_cout == 0 || t2s_of(geom, C8in * 8, t2s_small_cout, in_dtype == TTA_F16_HI ? 0 : 1),
              "tta_conv_tc: flags ask for the small-Cout transposed kernel but the layer does not qualify");
  if (t2s_small_cout > 0 && t2s_of(geom, C8in * 8, t2s_small_cout, in_dtype == TTA_F16_HI ? 0 : 1)) {
    TTA_REQUIRE(C8out == 1 && segs == nullptr, "tta_conv_tc: small-Cout transposed conv writes one channel chunk");
    return conv_t2s_impl(in_hi, in_lo, in_ns, N, C8in, Di, Hi, Wi, wpacked, bias, out, out_ns, t2s_small_cout, Do, Ho,
                         Wo, accumulate, stats_ws, stats_c8, (flags & 16384) ? 4 : 8, q_ksplit, q_grid, q_nbuf, stream);
  }
  // flags bit 15: the input planes are COMPACT ([N][D][H][W][4], in_n_stride in 16-bit elements, C8in == 1): stride-2
  // convs only (the stem, the dgrad of the head convT); weights packed with layout.pack_weights_tc(s2c4=True)
  const bool c4in = (flags & 32768) != 0;
  TTA_REQUIRE(!c4in || (geom == GEOM_S2 && C8in == 1), "tta_conv_tc: flags bit 15 (compact input) needs a stride-2 conv over one chunk");
  if (c4in) geom = GEOM_S2C4;
  const bool stacked = stacked_of(geom, C8in * 8, C8out * 8, in_dtype == TTA_F16_HI ? 0 : 1);
  if (stacked) geom = geom == GEOM_S1 ? GEOM_S1K : GEOM_S1TK;
  const bool pl2 = (geom == GEOM_S1 || geom == GEOM_S1T) && Ho <= 8 && Do >= 2 && !(flags & 32);
  // OPT-IN (flags bit13): the same nine (kd, kh) pipeline groups for ordinary planes (one plane per
  // accumulator, 16-row boxes): smaller stages, more of them in flight.  Bit-identical; measured on the
  // whole step 2.346 ms vs 2.334 ms with the three kd groups (the halo rows are fetched three times)
  const bool g9w = (geom == GEOM_S1 || geom == GEOM_S1T) && !pl2 && (flags & 8192);
  const bool g9 = pl2 || g9w;
  if (g9) geom = geom == GEOM_S1 ? GEOM_S1P : GEOM_S1TP;
  // transposed stride-2 conv over small INPUT planes: same idea, rows 64..127 = input plane d0 + 1;
  // the (jh = 0, jh = 1) halo rows become two 8-row boxes so that the plane pitch stays 8 rows
  const bool pl2t = geom == GEOM_T2 && Hi <= 8 && Di >= 2 && !(flags & 32);
  const int ppa = (pl2 || pl2t) ? 2 : 1;  // d-planes per accumulator
  const long long Vi = (long long)Di * Hi * Wi;
  TTA_REQUIRE(c4in || in_ns % (Vi * 8) == 0, "tta_conv_tc: n_stride must be a whole number of channel chunks");
  const int c8_pitch = c4in ? 1 : (int)(in_ns / (Vi * 8));
  if (geom == GEOM_S2 || geom == GEOM_S2C4) {
    TTA_REQUIRE(Di % 2 == 0 && Hi % 2 == 0 && Wi % 2 == 0 && Do == Di / 2 && Ho == Hi / 2 && Wo == Wi / 2,
                "tta_conv_tc: stride-2 conv needs even input dims");
  } else if (geom == GEOM_T2) {
    TTA_REQUIRE(Do == 2 * Di && Ho == 2 * Hi && Wo == 2 * Wi, "tta_conv_tc: transposed s2 output dims");
  } else {
    TTA_REQUIRE(Do == Di && Ho == Hi && Wo == Wi, "tta_conv_tc: stride-1 dims must match");
  }
  EncodeTiledFn enc = query ? nullptr : get_encode();
  TTA_REQUIRE(query || enc != nullptr, "tta_conv_tc: cuTensorMapEncodeTiled entry point not found");

  TcParams P;
  memset(&P, 0, sizeof(P));
  const int cout_pad = C8out * 8;
  P.ntile = ntile_of(geom, cout_pad, split);
  P.n_ntiles = (cout_pad + P.ntile - 1) / P.ntile;
  P.ncblk = (C8in + 1) / 2;
  P.c8_view = c8_pitch;
  P.c8in = C8in;
  P.single_chunk = C8in == 1;
  P.out_mul = geom == GEOM_T2 ? 2 : 1;
  P.d_mul = (geom == GEOM_S2 || geom == GEOM_S2C4) ? 2 : 1;
  P.Do = Do; P.Ho = Ho; P.Wo = Wo; P.C8out = C8out; P.accumulate = accumulate;
  P.out_ns = out_ns; P.wpacked = (const uint8_t*)wpacked; P.bias = bias; P.out = out;
  P.out_f16 = out_f16 ? 1 : 0;
  const int fmt = in_dtype == TTA_BF16 ? 1 : 0;
  const int idesc0 = (1 << 4) | (fmt << 7) | (fmt << 10) | ((128 >> 4) << 24);
  P.idesc0 = idesc0;
  P.idesc_n = idesc0 | ((P.ntile >> 3) << 17);
  // flags bit 17 (EXPERIMENT, split-plane operands only): drop the A_lo * B_hi product -- activations rounded to fp16,
  // weights still hi + lo.  DESIGN.md 6 has the measured error / time.
  if ((flags & 131072) && split) P.idesc_n = 0;
  P.idesc_2n = idesc0 | (((2 * P.ntile) >> 3) << 17);

  int Td, Th, Tw;  // extents of the tile space
  if (geom == GEOM_T2) { Td = Di; Th = Hi; Tw = Wi; } else { Td = Do; Th = Ho; Tw = Wo; }
  int hx, wx;      // halo extents of the single-box geometries
  if (geom == GEOM_K1) { hx = 16; wx = 8; } else if (pl2t) { hx = 8; wx = 9; } else if (geom == GEOM_T2) { hx = 17; wx = 9; } else if (pl2) { hx = 8; wx = 10; } else if (g9w) { hx = 16; wx = 10; } else { hx = 18; wx = 10; }
  // small-plane tiles regroup the SAME packed weights: 9 (kd, kh) groups of 3 kw entries instead of 3 kd groups of 9
  const int gmax = (g9 || c4in) ? 3 : tta_conv_tc_gmax(mode, K, stride);
  P.ngroups = g9 ? 9 : tta_conv_tc_ngroups(mode, K, stride);
  P.pl2 = (pl2 || pl2t) ? 1 : 0;
  P.s2pair = s2pair_of(geom, C8in * 8) ? 1 : 0;   // (GEOM_S2 only: the compact geometry has its own pairing)
  P.t2_jh16 = pl2t ? 144 : 9;
  const int acc_cols = split ? 2 * P.ntile : P.ntile;
  P.b_entry_bytes = 2 * acc_cols * 16;  // [kchunk 2][hi NT (| lo NT) rows][16 B]
  P.b_blob_bytes = gmax * P.b_entry_bytes;
  if (stacked) {  // entry (kh, kw) = [kchunk 2][slot 3][hi NT (| lo NT)][16 B]; one blob per channel block
    P.b_entry_bytes = 3 * 2 * acc_cols * 16;
    P.b_blob_bytes = 9 * P.b_entry_bytes;
  }

  // ---- shape the work item: TD d-planes (B-operand reuse), TMEM double buffering, pipeline depth.
  // TMEM columns: nbuf * nacc * 2*NT <= 512.
  // split-K factor for `items` tile-level work items: the one that minimises (waves of CTAs) x (channel
  // blocks per split + ~2 blocks' worth of pipeline fill / epilogue per wave).  Rounding the factor UP
  // to fill the SMs (e.g. 64 items x 3 = 192 CTAs on 148 SMs) costs a second wave.
  auto choose_ks = [&](long long items) -> int {
    if ((flags & 2) || items >= num_sms() || P.ncblk < 4) return 1;
    if (flags & 64) {  // A/B switch: round the factor up to fill the SMs
      long long k = (num_sms() + items - 1) / items;
      if (k > P.ncblk / 2) k = P.ncblk / 2;
      const int cbs = (P.ncblk + (int)k - 1) / (int)k;
      return (P.ncblk + cbs - 1) / cbs;
    }
    int best = 1;
    long long best_cost = -1;
    for (int k = 1; k <= P.ncblk / 2; ++k) {
      const int cbs = (P.ncblk + k - 1) / k, real = (P.ncblk + cbs - 1) / cbs;
      const long long waves = (items * real + num_sms() - 1) / num_sms();
      const long long cost = waves * (cbs + 2);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = real; }
    }
    return best;
  };
  const bool conv_like = geom == GEOM_S1 || geom == GEOM_S1T || geom == GEOM_K1 || g9 || stacked;
  // td = accumulators per work item (each ppa d-planes)
  int td_max = 1;
  if (conv_like) {
    td_max = 512 / acc_cols;
    if (td_max > 4) td_max = 4;
    if (td_max > (Td + ppa - 1) / ppa) td_max = (Td + ppa - 1) / ppa;
    if (td_max < 1) td_max = 1;
    if (flags & 1) td_max = 1;
  }
  // kd-stacked convs: three (else two) halo planes per stage when the (td + 2)-plane halo splits evenly
  // -- fewer, longer stages: the per-stage barrier round trip was on the critical path of the
  // single-product dgrads.  Same-box A/B per step: 1 plane 2.285 ms, 2 planes 2.265 / 2.225, 3 planes 2.217.
  // flags bit7: one plane per stage, bit11: at most two (A/B switches)
  auto pps_of = [&](int td_) {
    if (!stacked || (flags & 128)) return 1;
    if (!(flags & 2048) && (td_ + 2) % 3 == 0) return 3;
    return (td_ + 2) % 2 == 0 ? 2 : 1;
  };
  auto a_plane_of = [&](int td_) {
    if (pl2t) return 2 * 4608;  // [kchunk 2][jh 2][2 planes][8 rows][9 w][16 B]
    if (stacked) return 2 * round128(hx * wx * pps_of(td_) * 16);  // a stage is pps halo planes
    if (geom == GEOM_S2C4) return round128(3 * 2304);   // three (kh) boxes [16 rows][144 B]; no second k-chunk copy
    return geom == GEOM_S2 ? 18176 : 2 * round128(hx * wx * td_ * ppa * 16);
  };
  const int a_planes = split ? 2 : 1;
  // small-channel layers (C <= 32: the full-resolution levels, where items are many): keep ALL
  // weights of the single n-tile resident -> the per-stage B re-fetch (up to 2/3 of the L2->SM fill
  // traffic of the stride-2 stem) disappears
  const int b_total = (stacked ? 1 : P.ngroups) * P.ncblk * P.b_blob_bytes;
  // (transposed stride-2 convs: the head convT 64->3 re-fetched 28 KB of weights per 10 KB of A and ran at
  // the L2->SM fill cap; its blobs reserve 18 entries per group, so the budget is what four A stages leave)
  const int a_only_stage = round128((split ? 2 : 1) * a_plane_of(1));
  const int b_res_cap = geom == GEOM_T2 ? 227 * 1024 - 14336 - 1024 - 4 * a_only_stage : 112 * 1024;
  const bool b_res = stacked || (P.n_ntiles == 1 && b_total <= b_res_cap && !(flags & 16));
  P.b_res = b_res ? 1 : 0;
  auto stage_bytes_of = [&](int td_) { return round128(a_planes * a_plane_of(td_) + (b_res ? 0 : P.b_blob_bytes)); };
  const int smem_budget = 227 * 1024 - 14336 - (b_res ? b_total : 0);  // static smem: barriers, bias, statistics slots
  int td = td_max;
  while (td > 1 && smem_budget / stage_bytes_of(td) < 2) --td;
  int nacc = geom == GEOM_T2 ? 8 : td;
  int nbuf = (2 * nacc * acc_cols <= 512) ? 2 : 1;
  {
    // many work items: trade TD for the second accumulator buffer (epilogue/main-loop overlap)
    const long long tiles_now = (long long)((Tw + 7) / 8) * ((Th + 15) / 16) * ((Td + td * ppa - 1) / (td * ppa));
    if (nbuf == 1 && conv_like && td > 1 && tiles_now * N * P.n_ntiles > 2LL * num_sms()) {
      const int t2 = td / 2;
      if (2 * t2 * acc_cols <= 512) { td = t2; nacc = td; nbuf = 2; }
    }
  }
  if (conv_like && !(flags & 2)) {
    // SM-starved layers (16^3 / 8^3 levels): even with the split-K factor capped at ncblk/2 the grid
    // may fill less than half of the SMs -> fewer d-planes per item (more, shorter items) until it does
    auto items_of = [&](int td_) {
      const long long it = (long long)((Tw + 7) / 8) * ((Th + 15) / 16) * ((Td + td_ * ppa - 1) / (td_ * ppa)) * P.n_ntiles * N;
      return it * choose_ks(it);
    };
    while (td > 1 && items_of(td) * 2 <= num_sms()) td = (td + 1) / 2;
    nacc = td;
    nbuf = (2 * nacc * acc_cols <= 512) ? 2 : 1;
  }
  P.td = td * ppa; P.nacc = nacc; P.nbuf = nbuf;   // P.td: d-planes per work item
  P.a_plane_bytes = a_plane_of(td);
  P.stage_bytes = stage_bytes_of(td);
  P.b_off = a_planes * P.a_plane_bytes;
  P.pps = pps_of(td);
  P.lbo16[0] = pl2t ? 4608 / 16 : round128(hx * wx * (stacked ? P.pps : td * ppa) * 16) / 16;
  if (stacked) P.ngroups = (td + 2) / P.pps;
  P.b_nblob = stacked ? P.ncblk : P.ncblk * P.ngroups;
  P.b_cb_bytes = stacked ? P.b_blob_bytes : P.ngroups * P.b_blob_bytes;
  P.b_g_bytes = stacked ? 0 : P.b_blob_bytes;
  // pipeline depth from what THIS instantiation leaves free (the statistics variants carry 8-10 KB of
  // static slots); td / split-K above were shaped with the most conservative budget so that
  // tta_conv_tc_query and the launch always agree on the grid
  const int static_smem = (stats_ws != nullptr && stats_c8 > 0) ? 11264 : ((segs != nullptr && nsegs > 0) ? 13312 : 3072);
  P.nstages = (227 * 1024 - static_smem - 1024 - (b_res ? b_total : 0)) / P.stage_bytes;
  TTA_REQUIRE(P.nstages >= 1, "tta_conv_tc: stage of %d bytes does not fit shared memory", P.stage_bytes);
  if (P.nstages > kMaxStages) P.nstages = kMaxStages;
  P.b_res_off = P.nstages * P.stage_bytes;
  int cols = 32;
  while (cols < nbuf * nacc * acc_cols) cols *= 2;
  TTA_REQUIRE(cols <= 512, "tta_conv_tc: %d accumulator columns exceed TMEM", nbuf * nacc * acc_cols);
  P.tmem_cols = cols;
  P.tiles_w = (Tw + 7) / 8; P.tiles_h = (Th + 15) / 16; P.tiles_d = (Td + P.td - 1) / P.td;

  // ---- split-K over channel blocks when the tile grid cannot fill the SMs
  {
    const long long items = (long long)P.tiles_w * P.tiles_h * P.tiles_d * P.n_ntiles * N;
    const int ks = choose_ks(items);
    P.cb_per_split = (P.ncblk + ks - 1) / ks;
    P.ksplit = (P.ncblk + P.cb_per_split - 1) / P.cb_per_split;
    P.work_items = (int)(items * P.ksplit);
  }
  {
    // OPT-IN (flags bit12): CTA pairs (cluster of 2) with multicast weight blobs, where weights are streamed
    // per stage and every tile has a partner.  Bit-identical results; measured on the whole step 2.540 ms
    // with pairs vs 2.531 ms without -- halving the weight fetch per CTA does not pay, the layers without
    // resident weights are not L2->SM bandwidth bound, and the lockstep of the pair costs a little.
    const long long tiles_all = (long long)P.tiles_w * P.tiles_h * P.tiles_d * N;
    P.cl2 = (!P.b_res && (flags & 4096) && !(flags & 4) && tiles_all % 2 == 0 && tiles_all >= 2 && num_sms() % 2 == 0) ? 2 : 0;
    P.pair_items = (int)(tiles_all / 2) * P.n_ntiles * P.ksplit;
  }
  {
    // magic(d) = floor(2^32 / d) + 1: __umulhi(x, magic) == x / d whenever x * d < 2^32; d == 1 -> 0 (identity)
    auto magic = [](unsigned d) -> unsigned { return d <= 1 ? 0u : (unsigned)((0x100000000ull / d) + 1ull); };
    const unsigned dmax = (unsigned)std::max(std::max(P.n_ntiles * P.ksplit, P.tiles_w * P.tiles_h * P.tiles_d),
                                             std::max(P.tiles_w, P.tiles_h));
    TTA_REQUIRE((unsigned long long)(P.work_items + 1) * dmax < 0x100000000ull,
                "tta_conv_tc: %d work items exceed the index arithmetic range", P.work_items);
    P.mg_per_tile = magic((unsigned)(P.n_ntiles * P.ksplit));
    P.mg_tiles = magic((unsigned)(P.tiles_w * P.tiles_h * P.tiles_d));
    P.mg_ksplit = magic((unsigned)P.ksplit);
    P.mg_tiles_w = magic((unsigned)P.tiles_w);
    P.mg_tiles_h = magic((unsigned)P.tiles_h);
  }
  if (getenv("TTA_TC_DEBUG"))
    fprintf(stderr, "tta_conv_tc: geom %d split %d N %d C8 %d->%d in %dx%dx%d | nt %d x%d ncblk %d | accs %d planes %d nbuf %d "
            "stages %d x %d B bres %d cl2 %d | tiles %dx%dx%d ksplit %d (cb %d) items %d waves %.2f\n",
            geom, split, N, C8in, C8out, Di, Hi, Wi, P.ntile, P.n_ntiles, P.ncblk, P.nacc, P.td, P.nbuf, P.nstages,
            P.stage_bytes, P.b_res, P.cl2, P.tiles_d, P.tiles_h, P.tiles_w, P.ksplit, P.cb_per_split, P.work_items,
            (double)P.work_items / num_sms());
  if (query) {
    *q_ksplit = P.ksplit;
    if (q_nbuf) *q_nbuf = P.nbuf;
    *q_grid = (flags & 4) ? P.work_items : (P.work_items < num_sms() ? P.work_items : num_sms());
    if (P.cl2) *q_grid = 2 * (P.pair_items < num_sms() / 2 ? P.pair_items : num_sms() / 2);
    return TTA_OK;
  }
  // ---- fused norm statistics (see the epilogue): needs the whole K sum in one CTA and plain stores
  P.stats_c8 = 0;
  if (stats_ws != nullptr && stats_c8 > 0) {
    TTA_REQUIRE(P.ksplit == 1 && !accumulate && split,
                "tta_conv_tc: fused statistics need split-plane operands, ksplit == 1 and no accumulate");
    TTA_REQUIRE(stats_c8 <= C8out, "tta_conv_tc: stats_c8 %d > C8out %d", stats_c8, C8out);
    P.stats_c8 = stats_c8;
    P.n_batch = N;
    P.stats = stats_ws + 1024;  // same workspace convention as tta_norm_stats: [1024 counters][partials]
  }
  // ---- fused norm-backward reductions (dgrad whose output is the complete gradient of <= 2 norms)
  P.nseg = 0;
  if (segs != nullptr && nsegs > 0) {
    TTA_REQUIRE(nsegs <= 2 && P.ksplit == 1 && !split && P.stats_c8 == 0,
                "tta_conv_tc: fused norm-backward sums need <= 2 segments, ksplit == 1, one fp16 plane");
    for (int i = 0; i < nsegs; ++i) {
      const tta_norm_bwd_seg& h = segs[i];
      TTA_REQUIRE(h.y && h.mean && h.rstd && h.gamma && h.beta && h.partial && h.c8_begin >= 0 && h.c8_count > 0 &&
                      h.c8_begin + h.c8_count <= C8out,
                  "tta_conv_tc: bad norm-backward segment %d", i);
      TcParams::BwdSeg& d = P.seg[i];
      d.c8_begin = h.c8_begin; d.c8_end = h.c8_begin + h.c8_count; d.relu = h.relu; d.pad = 0;
      d.y = h.y; d.y_ns = h.y_n_stride; d.mean = h.mean; d.rstd = h.rstd; d.gamma = h.gamma; d.beta = h.beta;
      d.partial = h.partial;
    }
    P.nseg = nsegs;
    P.n_batch = N;
  }

  // ---- tensor maps (one k-chunk = 8 channels per TMA box; zero padding = OOB fill)
  // merged (n, chunk) extent ends at the LAST chunk of this view: chunks past it (odd C8in, or the
  // neighbouring slice of a concat buffer) are out of bounds -> zero filled, never read.
  const cuuint64_t nc_extent = (cuuint64_t)((long long)(N - 1) * c8_pitch + C8in);
  const cuuint32_t es5[5] = {1, 1, 1, 1, 1};
  // stride-2 convs: 5-D map over one (h, w) parity class, 16 B inner box (8 channels of one voxel)
  auto encode_par = [&](CUtensorMap* m, const uint16_t* base, int par_h, int par_w, int bw, int bh) -> bool {
    const uint16_t* ptr = base + ((long long)par_h * Wi + par_w) * 8;
    cuuint64_t gdim[5] = {8, (cuuint64_t)((Wi - par_w + 1) / 2), (cuuint64_t)((Hi - par_h + 1) / 2), (cuuint64_t)Di,
                          nc_extent};
    cuuint64_t gstr[4] = {32, (cuuint64_t)32 * Wi, (cuuint64_t)16 * Wi * Hi, (cuuint64_t)16 * Vi};
    cuuint32_t box[5] = {8, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, (void*)ptr, gdim, gstr, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  // everything else: 4-D map whose inner dimension is a whole w-row (W*8 contiguous 16-bit values),
  // so a halo row of wx voxels is ONE wx*16-byte TMA row instead of wx 16-byte rows
  auto encode_row = [&](CUtensorMap* m, const uint16_t* base, int bw, int bh, int bd) -> bool {
    cuuint64_t gdim[4] = {(cuuint64_t)Wi * 8, (cuuint64_t)Hi, (cuuint64_t)Di, nc_extent};
    cuuint64_t gstr[3] = {(cuuint64_t)16 * Wi, (cuuint64_t)16 * Wi * Hi, (cuuint64_t)16 * Vi};
    cuuint32_t box[4] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, (cuuint32_t)bd, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)base, gdim, gstr, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  // stride-2 convs over a w-parity-split input ([..][H][2][W/2][8], flags bit3): every (h, w) parity
  // class is a dense run of W/2 voxels per row -> one bw*16-byte TMA row per halo row instead of bw
  // 16-byte requests (the 5-D path is TMA-issue bound: ~2 cycles per 16 B)
  auto encode_par_rows = [&](CUtensorMap* m, const uint16_t* base, int par_h, int par_w, int bw, int bh) -> bool {
    const uint16_t* ptr = base + ((long long)par_h * Wi + (long long)par_w * (Wi / 2)) * 8;
    cuuint64_t gdim[4] = {(cuuint64_t)(Wi / 2) * 8, (cuuint64_t)((Hi - par_h + 1) / 2), (cuuint64_t)Di, nc_extent};
    cuuint64_t gstr[3] = {(cuuint64_t)32 * Wi, (cuuint64_t)16 * Wi * Hi, (cuuint64_t)16 * Vi};
    cuuint32_t box[4] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)ptr, gdim, gstr, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  bool ok = true;
  P.s2_rows = (geom == GEOM_S2 && (flags & 8)) ? 1 : 0;
  if (geom == GEOM_S2C4) {
    // one 4-D map per input-row parity and operand plane: rows h = 2*i + par of the compact tensor, a row =
    // W*4 contiguous 16-bit values; box = 18 voxels x 16 rows (OOB -> zero = the conv's padding)
    for (int par = 0; par < 2; ++par)
      for (int pl = 0; pl < (split ? 2 : 1); ++pl) {
        const uint16_t* base = (pl ? in_lo : in_hi) + (long long)par * Wi * 4;
        cuuint64_t gdim[4] = {(cuuint64_t)Wi * 4, (cuuint64_t)((Hi - par + 1) / 2), (cuuint64_t)Di, (cuuint64_t)N};
        cuuint64_t gstr[3] = {(cuuint64_t)16 * Wi, (cuuint64_t)8 * Wi * Hi, (cuuint64_t)2 * in_ns};
        cuuint32_t box[4] = {72, 16, 1, 1};
        ok = ok && enc(&P.amap[par * 2 + pl], CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)base, gdim, gstr, box, es5,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }
  } else if (geom == GEOM_S2) {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const int m = ph * 2 + pw;
        if (P.s2_rows) {
          ok = ok && encode_par_rows(&P.amap[m * 2 + 0], in_hi, ph, pw, pw ? 9 : 8, ph ? 17 : 16);
          if (split) ok = ok && encode_par_rows(&P.amap[m * 2 + 1], in_lo, ph, pw, pw ? 9 : 8, ph ? 17 : 16);
        } else {
          ok = ok && encode_par(&P.amap[m * 2 + 0], in_hi, ph, pw, pw ? 9 : 8, ph ? 17 : 16);
          if (split) ok = ok && encode_par(&P.amap[m * 2 + 1], in_lo, ph, pw, pw ? 9 : 8, ph ? 17 : 16);
        }
      }
  } else {
    const int box_d = geom == GEOM_T2 ? ppa : (stacked ? P.pps : td * ppa);
    ok = ok && encode_row(&P.amap[0], in_hi, wx, hx, box_d);
    if (split) ok = ok && encode_row(&P.amap[1], in_lo, wx, hx, box_d);
  }
  TTA_REQUIRE(ok, "tta_conv_tc: cuTensorMapEncodeTiled failed (dims %d,%d,%d C8 pitch %d)", Di, Hi, Wi, c8_pitch);

  // ---- group tables (A loads; the MMA schedule is compile-time in issue_group<GEOM>)
  if (geom == GEOM_S1 || geom == GEOM_S1T) {
    for (int kd = 0; kd < 3; ++kd) {
      TcGroup& G = P.grp[kd];
      G.nloads = 1; G.nmma = 9;
      G.ld[0] = {0, -1, -1, geom == GEOM_S1 ? kd - 1 : 1 - kd, 0, 18 * 10 * td * 16, P.lbo16[0] * 16, 0};
      G.tx_bytes = 2 * G.ld[0].bytes;
    }
    for (int p = 0; p < td; ++p) P.acc_pd[p] = (signed char)p;
  } else if (stacked) {
    for (int j = 0; j < P.ngroups; ++j) {
      TcGroup& G = P.grp[j];
      G.nloads = 1; G.nmma = 9;
      G.ld[0] = {0, -1, -1, j * P.pps - 1, 0, 18 * 10 * P.pps * 16, P.lbo16[0] * 16, 0};
      G.tx_bytes = 2 * G.ld[0].bytes;
    }
    for (int p = 0; p < td; ++p) P.acc_pd[p] = (signed char)p;
  } else if (g9) {
    for (int g = 0; g < 9; ++g) {
      const int kd = g / 3, kh = g % 3;
      TcGroup& G = P.grp[g];
      G.nloads = 1; G.nmma = 3;
      G.ld[0] = {0, -1, geom == GEOM_S1P ? kh - 1 : 1 - kh, geom == GEOM_S1P ? kd - 1 : 1 - kd, 0,
                 hx * 10 * td * ppa * 16, P.lbo16[0] * 16, 0};
      G.tx_bytes = 2 * G.ld[0].bytes;
    }
    for (int p = 0; p < td; ++p) P.acc_pd[p] = (signed char)(ppa * p);
  } else if (geom == GEOM_K1) {
    TcGroup& G = P.grp[0];
    G.nloads = 1; G.nmma = 1;
    G.ld[0] = {0, 0, 0, 0, 0, 16 * 8 * td * 16, P.lbo16[0] * 16, 0};
    G.tx_bytes = 2 * G.ld[0].bytes;
    for (int p = 0; p < td; ++p) P.acc_pd[p] = (signed char)p;
  } else if (geom == GEOM_S2C4) {
    for (int kd = 0; kd < 3; ++kd) {
      TcGroup& G = P.grp[kd];
      G.nloads = 3; G.nmma = 3;
      for (int kh = 0; kh < 3; ++kh)   // input row 2h + kh - 1: kh = 1 -> parity 0, row h; kh = 0 / 2 -> parity 1, row h - 1 / h
        G.ld[kh] = {kh == 1 ? 0 : 1, 0, kh == 0 ? -1 : 0, kd - 1, kh * 2304, 2304, 0, 0};
      G.tx_bytes = 2 * 3 * 2304;
    }
    P.lbo16[0] = 1;
    P.acc_pd[0] = 0;
  } else if (geom == GEOM_S2) {
    int off[4], o = 0;
    const int hxs[2] = {16, 17}, wxs[2] = {8, 9};
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const int m = ph * 2 + pw;
        P.lbo16[m] = round128(hxs[ph] * wxs[pw] * 16) / 16;
        off[m] = o;
        o += 2 * P.lbo16[m] * 16;
      }
    for (int kd = 0; kd < 3; ++kd) {
      TcGroup& G = P.grp[kd];
      G.nloads = 4; G.nmma = P.s2pair ? 5 : 9;
      int abytes = 0;
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const int m = ph * 2 + pw;
          G.ld[m] = {m, pw ? -1 : 0, ph ? -1 : 0, kd - 1, off[m], hxs[ph] * wxs[pw] * 16, P.lbo16[m] * 16, 0};
          abytes += G.ld[m].bytes;
        }
      G.tx_bytes = 2 * abytes;
    }
    P.acc_pd[0] = 0;
  } else {  // GEOM_T2
    for (int jd = 0; jd < 2; ++jd) {
      TcGroup& G = P.grp[jd];
      G.nloads = 1;
      G.nmma = jd == 0 ? 18 : 9;
      if (pl2t) {
        G.nloads = 2;
        for (int jh = 0; jh < 2; ++jh) G.ld[jh] = {0, 0, jh, jd, jh * 2304, 2304, 4608, 0};
        G.tx_bytes = 2 * 2 * 2304;
      } else {
        G.ld[0] = {0, 0, 0, jd, 0, 17 * 9 * 16, P.lbo16[0] * 16, 0};
        G.tx_bytes = 2 * G.ld[0].bytes;
      }
    }
    for (int a = 0; a < 8; ++a) {
      P.acc_pd[a] = 0; P.acc_qd[a] = (signed char)(a >> 2); P.acc_qh[a] = (signed char)((a >> 1) & 1);
      P.acc_qw[a] = (signed char)(a & 1);
    }
  }

  const size_t smem = (size_t)P.nstages * P.stage_bytes + (P.b_res ? b_total : 0) + 1024;
  const bool pdl_ok = !(P.ksplit > 1 && !accumulate);  // a memset precedes the split-K launch
  if (P.ksplit > 1 && !accumulate) {
    // partial sums are combined with float4 atomics: start from zero (timing experiment, round 2: skipping all 14
    // memset nodes of a step -- wrong results -- gains 0.5 %, so a shared pre-zeroed arena is not worth building)
    const long long Vo = (long long)Do * Ho * Wo;
    const long long per_n = (long long)C8out * Vo * 8;
    static const bool per_sample_memset = getenv("TTA_MEMSET_PER_SAMPLE") != nullptr;   // A/B switch
    if (out_ns == per_n && !per_sample_memset) {   // the view is the whole buffer: one memset node instead of one per sample
      if (cudaMemsetAsync(out, 0, (size_t)N * per_n * sizeof(float), stream) != cudaSuccess)
        return tta_check_launch("tta_conv_tc(memset)");
    } else {
      for (int nn = 0; nn < N; ++nn)
        if (cudaMemsetAsync(out + nn * out_ns, 0, (size_t)per_n * sizeof(float), stream) != cudaSuccess)
          return tta_check_launch("tta_conv_tc(memset)");
    }
  }
  int grid_x = P.work_items < num_sms() ? P.work_items : num_sms();
  if (flags & 4) grid_x = P.work_items;
  if (P.cl2) grid_x = 2 * (P.pair_items < num_sms() / 2 ? P.pair_items : num_sms() / 2);
  const dim3 grid(grid_x, 1, 1);
#define TTA_TC_LAUNCH_ST(G, T, S, ST)                                                                         \
  do {                                                                                                     \
    static bool configured = false;                                                                        \
    if (!configured) {                                                                                     \
      cudaFuncAttributes fa;                                                                               \
      if (cudaFuncGetAttributes(&fa, conv_tc_kernel<G, T, S, ST>) != cudaSuccess)                             \
        return tta_check_launch("tta_conv_tc(attrs)");                                                     \
      const int max_dyn = 227 * 1024 - (int)fa.sharedSizeBytes;                                            \
      if (cudaFuncSetAttribute(conv_tc_kernel<G, T, S, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                               max_dyn) != cudaSuccess)                                                    \
        return tta_check_launch("tta_conv_tc(cudaFuncSetAttribute)");                                      \
      configured = true;                                                                                   \
    }                                                                                                      \
    if (P.cl2) tta_launch_cluster(conv_tc_kernel<G, T, S, ST>, grid, kTcThreads, smem, stream, 2, P);      \
    else tta_launch(conv_tc_kernel<G, T, S, ST>, grid, kTcThreads, smem, stream, pdl_ok && tta_pdl_family(8), P); \
  } while (0)
#define TTA_TC_LAUNCH(G, T)                                                                                \
  do {                                                                                                     \
    if (P.stats_c8 > 0) { /* fused statistics: forward (split-plane) convs only */                        \
      TTA_TC_LAUNCH_ST(G, T, 1, 1);                                                                        \
    } else if (P.nseg > 0) { /* fused norm-backward sums: single-plane dgrads only */                      \
      TTA_TC_LAUNCH_ST(G, T, 0, 2);                                                                        \
    } else if (split) {                                                                                    \
      TTA_TC_LAUNCH_ST(G, T, 1, 0);                                                                        \
    } else {                                                                                               \
      TTA_TC_LAUNCH_ST(G, T, 0, 0);                                                                        \
    }                                                                                                      \
  } while (0)
#define TTA_TC_LAUNCH_TD(G)                                                                                \
  switch (td) {                                                                                            \
    case 1: TTA_TC_LAUNCH(G, 1); break;                                                                    \
    case 2: TTA_TC_LAUNCH(G, 2); break;                                                                    \
    case 3: TTA_TC_LAUNCH(G, 3); break;                                                                    \
    default: TTA_TC_LAUNCH(G, 4); break;                                                                   \
  }
  switch (geom) {
    case GEOM_S1: TTA_TC_LAUNCH_TD(GEOM_S1); break;
    case GEOM_S1T: TTA_TC_LAUNCH_TD(GEOM_S1T); break;
    case GEOM_K1: TTA_TC_LAUNCH_TD(GEOM_K1); break;
    case GEOM_S1K: TTA_TC_LAUNCH_TD(GEOM_S1K); break;
    case GEOM_S1TK: TTA_TC_LAUNCH_TD(GEOM_S1TK); break;
    case GEOM_S1P: TTA_TC_LAUNCH_TD(GEOM_S1P); break;
    case GEOM_S1TP: TTA_TC_LAUNCH_TD(GEOM_S1TP); break;
    case GEOM_S2: TTA_TC_LAUNCH(GEOM_S2, 1); break;
    case GEOM_S2C4: TTA_TC_LAUNCH(GEOM_S2C4, 1); break;
    default: TTA_TC_LAUNCH(GEOM_T2, 1); break;
  }
#undef TTA_TC_LAUNCH_TD
#undef TTA_TC_LAUNCH_ST
#undef TTA_TC_LAUNCH
  return tta_check_launch("tta_conv_tc");
}

extern "C" {

int tta_conv_tc(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns, int in_dtype, int N, int C8in,
                int Di, int Hi, int Wi, const void* wpacked, const float* bias, float* out, long long out_ns,
                int C8out, int Do, int Ho, int Wo, int mode, int K, int stride, int accumulate, int flags,
                float* stats_ws, int stats_c8, cudaStream_t stream) {
  TTA_RECORDABLE(tta_conv_tc(in_hi, in_lo, in_ns, in_dtype, N, C8in, Di, Hi, Wi, wpacked, bias, out, out_ns, C8out, Do, Ho, Wo, mode, K, stride, accumulate, flags, stats_ws, stats_c8, s_));
  return conv_tc_impl(in_hi, in_lo, in_ns, in_dtype, N, C8in, Di, Hi, Wi, wpacked, bias, out, out_ns, C8out, Do, Ho,
                      Wo, mode, K, stride, accumulate, flags, stats_ws, stats_c8, nullptr, 0, nullptr, nullptr, nullptr,
                      stream);
}

// tta_conv_tc for an input-gradient conv whose result is the COMPLETE gradient w.r.t. the outputs of
// up to two norm layers (channel segments [c8_begin, c8_begin + c8_count) of the output view): the
// epilogue additionally leaves per-CTA partial sums of dz = g*[z > 0] and dz*xhat for each segment in
// seg.partial ([N][c8_count][grid][16] floats, grid from tta_conv_tc_query), which
// tta_norm_bwd_finalize(splits = grid) reduces.  segs is a HOST pointer.
int tta_conv_tc_bwd_norm(const uint16_t* in_hi, const uint16_t* in_lo, long long in_ns, int in_dtype, int N, int C8in,
                         int Di, int Hi, int Wi, const void* wpacked, float* out, long long out_ns, int C8out, int Do,
                         int Ho, int Wo, int mode, int K, int stride, int accumulate, int flags,
                         const tta_norm_bwd_seg* segs, int nsegs, cudaStream_t stream) {
  TTA_RECORDABLE(tta_conv_tc_bwd_norm(in_hi, in_lo, in_ns, in_dtype, N, C8in, Di, Hi, Wi, wpacked, out, out_ns, C8out, Do, Ho, Wo, mode, K, stride, accumulate, flags, segs, nsegs, s_));
  return conv_tc_impl(in_hi, in_lo, in_ns, in_dtype, N, C8in, Di, Hi, Wi, wpacked, nullptr, out, out_ns, C8out, Do, Ho,
                      Wo, mode, K, stride, accumulate, flags, nullptr, 0, segs, nsegs, nullptr, nullptr, nullptr, stream);
}

// Launch shape of tta_conv_tc for these arguments: *ksplit = split-K factor (fused statistics need
// 1), *grid = CTAs (= the `splits` of the statistics partials), *nbuf = TMEM accumulator buffers
// (2: the epilogue, and with it the statistics reduction, overlaps the next item's MMAs).
int tta_conv_tc_query(int in_dtype, int N, int C8in, int Di, int Hi, int Wi, int C8out, int Do, int Ho, int Wo,
                      int mode, int K, int stride, int accumulate, int flags, int* ksplit, int* grid, int* nbuf) {
  TTA_REQUIRE(ksplit && grid, "tta_conv_tc_query: null pointer");
  const long long Vi = (long long)Di * Hi * Wi;
  return conv_tc_impl(nullptr, nullptr, (long long)C8in * Vi * 8, in_dtype, N, C8in, Di, Hi, Wi, nullptr, nullptr,
                      nullptr, 0, C8out, Do, Ho, Wo, mode, K, stride, accumulate, flags, nullptr, 0, nullptr, 0, ksplit,
                      grid, nbuf, nullptr);
}

long long tta_conv_tc_packed_bytes(int mode, int K, int stride, int cin, int cout) {
  const int geom = geom_of(mode, K, stride);
  if (geom == GEOM_NONE) return 0;
  const int nt = ntile_of(geom, (cout + 7) / 8 * 8, 1);
  const int nnt = ((cout + 7) / 8 * 8 + nt - 1) / nt;
  const int ncb = (cin + 15) / 16;
  return (long long)nnt * ncb * tta_conv_tc_ngroups(mode, K, stride) * tta_conv_tc_gmax(mode, K, stride) * 2 * 2 *
         nt * 16;
}

}  // extern "C"
