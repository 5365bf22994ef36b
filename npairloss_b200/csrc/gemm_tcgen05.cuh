// gemm_tcgen05.cuh -- persistent, warp-specialised split-operand GEMM for sm_100a.
//
//   C[M x Nn] = sum over passes (sa,sb) of  A_sa[M x K] . B_sb[Nn x K]^T      (both operands K-major, 2-byte elements)
//
// "Split operand": an fp32 matrix is stored as NSPLIT 2-byte pieces whose sum reproduces it
//   NSPLIT=1 bf16            1 pass  (hi.hi)                                  -- throughput mode
//   NSPLIT=2 fp16 hi+lo      3 passes (hh, hl, lh)      ~2^-22 relative       -- fp32-faithful, pre-scaled inputs
//   NSPLIT=3 bf16 hi+mid+lo  6 passes (hh,hm,mh,mm,hl,lh) ~2^-24 relative     -- fp32-faithful, any dynamic range
// All passes accumulate into the same fp32 TMEM accumulator, so the tensor pipe sees one long K loop.
//
// Structure (persistent, one CTA per SM, 384 threads):
//   warp 0    TMA producer: cp.async.bulk.tensor 3-D boxes {BK, rows, 1 piece} -> 128B/64B-swizzled smem ring
//   warp 1    MMA issuer  : one thread issues tcgen05.mma.kind::f16 (128 x 256 x 16; pair mode 256 x 256 x 16), fp32 accum in TMEM
//   warp 2    TMEM allocator (512 columns = two 128x256 fp32 accumulators, double buffered)
//   warps 4-11 epilogue   : tcgen05.ld 32x32b -> registers -> fused epilogue (similarity store + row statistics,
//                           or scaled store of a gradient tile), overlapped with the next tile's MMAs; warp w drains TMEM
//                           lanes 32*(w%4) and the column half (w-4)/4 of the tile
// NCTA = 2 (pair mode, similarity epilogues): a cluster of two CTAs computes a 256 x 256 block with
// tcgen05.mma.cta_group::2 -- see GemmCfg.
// Replaces the reference's cublasSgemm calls: sim GEMM npair_multi_class_loss.cu:218 and the six backward
// GEMMs .cu:448-460; the EPI_SIM epilogue also replaces GetLabelDiffMtx (.cu:44-66) and the host statistics loop
// (.cu:225-265: min_within / max_between / max_all) -- they are computed while the tile is in registers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cfloat>
#include <stdint.h>

#include "kernels.cuh"
#include "ptx.cuh"
#include "thresholds.cuh"

namespace npair {

// EPI_SIM     : similarity tile -> S + fused row statistics (any world size)
// EPI_OUT     : alpha * acc (+ beta * out) gradient tile
// EPI_SIM_SYM : world == 1 only.  S = X X^T is symmetric, so only tiles that touch the upper triangle are computed
//               (tile list from the host, ~52 % of the tiles); every strictly-upper 128-column block is also written
//               MIRRORED (second TMA store) and contributes COLUMN statistics (warp redux) to the rows it mirrors into.
//               S comes out bitwise symmetric, which the backward weight builder relies on.
enum { EPI_SIM = 0, EPI_OUT = 1, EPI_SIM_SYM = 2 };

struct GemmParams {
  int M, Nn;           // logical output extent
  int num_kblocks;     // K_pad / BK
  int tiles_m, tiles_n;
  const int2* tile_list;      // optional explicit (m_blk, n_blk) list (EPI_SIM_SYM); tiles_m*tiles_n entries are then ignored
  int num_tiles_list;
  int splits, kb_per_split;   // split-K (EPI_OUT only): tile = (m_blk*tiles_n + n_blk)*splits + split, k-blocks [split*kb_per_split, ...)
  float* part;                // splits > 1: partial products [split][M][ldo]; a reduce kernel sums them in fixed order
  // ---- EPI_SIM ----
  float* S;            // [M x ldS] fp32 similarities
  long long ldS;       // multiple of 32
  const float* dev_scale;  // device scalar: inverse operand pre-scale (power of two) or NULL.  EPI_SIM multiplies the
                           // accumulators by its square (both operands were pre-scaled), EPI_OUT multiplies alpha by it
  const float* lab_rows;   // [M]  labels of this rank's rows
  const float* lab_cols;   // [Nn] labels of all columns
  int self_offset;         // rank * Q : column index of row 0's self pair
  uint32_t* st_minw;       // ordered-uint per-row statistics, pre-initialised
  uint32_t* st_maxw;
  uint32_t* st_maxb;
  uint32_t* st_maxall;
  int* cnt_same;           // per-row number of same-label non-self columns
  // fused threshold pick (.cu:275-337): the last CTA to finish runs thresholds_one_block (thresholds.cuh); fuse_thr = 0: separate kernel
  int fuse_thr;
  RowArrays ra;
  MiningParams mp;
  BlockScalars* bs;
  // ---- EPI_OUT ----
  float* out;          // [M x ldo]
  long long ldo;
  float alpha, beta;   // out = alpha*acc + beta*out
};

// BK_ = K-block in elements = one swizzle span per smem row (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B).  Measured on B200
// (gpurun_out/tune1.log): the short-K similarity GEMM (K = D) prefers 64, the long-K gradient GEMM (K = N) with two
// fp16 pieces prefers 32 (four 48 KB stages hide the TMA latency better than two 96 KB ones: 164 -> 141 us).
// NCTA = 2: CTA-pair mode (tcgen05.mma.cta_group::2).  A cluster of two CTAs on one TPC computes a 256 x 256 block: each CTA
// stages its own 128 rows of A and 128 of the 256 B rows, the leader issues one M = 256 MMA that reads both CTAs' shared
// memory, and each CTA's tensor memory receives its 128 x 256 half.  Per SM the operand traffic through shared memory drops
// from 12 KB to 8 KB per K = 16 step, which is what bounds these kernels.
template <int NSPLIT, int BK_, int NCTA = 1>
struct GemmCfg {
  static constexpr int BM = 128, BN = 256;
  static constexpr int B_ROWS = BN / NCTA;                    // B rows staged by one CTA
  static constexpr int BK = BK_;
  static constexpr int ROW_BYTES = BK * 2;                    // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  static constexpr uint32_t LAYOUT = (ROW_BYTES == 128) ? 2u : 4u;
  static constexpr uint32_t SBO = 8 * ROW_BYTES;              // byte distance between 8-row groups
  static constexpr int A_PIECE = BM * ROW_BYTES;
  static constexpr int B_PIECE = B_ROWS * ROW_BYTES;
  static constexpr int STAGE_BYTES = NSPLIT * (A_PIECE + B_PIECE);
  // pair mode: two TMA-store staging tiles per epilogue warp (direct + mirrored box in flight together), 160 KB of stages
  static constexpr int STORE_BUFS = (NCTA == 2) ? 2 : 1;
  static constexpr int STAGES = ((NCTA == 2 ? 160 : 192) * 1024) / STAGE_BYTES;   // fill 192 / 160 KB with operand stages
  static constexpr int NPASS = (NSPLIT == 1) ? 1 : (NSPLIT == 2 ? 3 : 6);
  static constexpr int STORE_STAGE_BYTES = 8 * 4096 * STORE_BUFS;   // 32x32 fp32 TMA-store staging tiles, per epilogue warp
  static constexpr int SMEM_AUX = 2048;                       // barriers + tmem ptr + column labels
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_STAGE_BYTES + SMEM_AUX + 1024 /*alignment slack*/;
  static constexpr int THREADS = 384;                         // 4 control warps + 8 epilogue warps
};

__device__ __forceinline__ void pass_pieces(int nsplit, int p, int& sa, int& sb) {
  // (A piece, B piece) of pass p; largest terms first
  if (nsplit == 1) { sa = 0; sb = 0; return; }
  if (nsplit == 2) { sa = (p == 2) ? 1 : 0; sb = (p == 1) ? 1 : 0; return; }
  // nsplit == 3: hh, hm, mh, mm, hl, lh
  const int A[6] = {0, 0, 1, 1, 0, 2};
  const int B[6] = {0, 1, 0, 1, 2, 0};
  sa = A[p]; sb = B[p];
}

// Per-thread statistics of 32 consecutive similarities against 32 labels staged in shared memory.
// fast: no bounds / self-pair checks, branch-free (predicated).  Otherwise entry c is valid iff idx0 + c < limit and
// idx0 + c != self_idx.
__device__ __forceinline__ void stats32(const float (&v)[32], const float* __restrict__ lab, float lab_i, bool fast, int idx0,
                                        int limit, int self_idx, float& minw, float& maxw, float& maxb, int& cnt) {
  if (fast) {
    // four independent accumulator sets (one per element of a float4): an epilogue warp is alone on its scheduler most of the
    // time, so instruction-level parallelism, not warp-level, has to hide the 4-cycle ALU latency of the min/max chains
    float mnw[4] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX}, mxw[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    float mxb[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    int cn[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 l4 = *reinterpret_cast<const float4*>(lab + 4 * q);
      const float ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float x = v[4 * q + e];
        if (ll[e] == lab_i) { mnw[e] = fminf(mnw[e], x); mxw[e] = fmaxf(mxw[e], x); ++cn[e]; }
        else mxb[e] = fmaxf(mxb[e], x);
      }
    }
    minw = fminf(minw, fminf(fminf(mnw[0], mnw[1]), fminf(mnw[2], mnw[3])));
    maxw = fmaxf(maxw, fmaxf(fmaxf(mxw[0], mxw[1]), fmaxf(mxw[2], mxw[3])));
    maxb = fmaxf(maxb, fmaxf(fmaxf(mxb[0], mxb[1]), fmaxf(mxb[2], mxb[3])));
    cnt += (cn[0] + cn[1]) + (cn[2] + cn[3]);
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int idx = idx0 + c;
      const bool valid = (idx < limit) && (idx != self_idx);
      const bool same = (lab[c] == lab_i);
      if (valid && same) { minw = fminf(minw, v[c]); maxw = fmaxf(maxw, v[c]); ++cnt; }
      if (valid && !same) maxb = fmaxf(maxb, v[c]);
    }
  }
}

// Statistics of 32 similarities none of which is a same-label pair (the caller has excluded it by the chunk's label range): only
// the hardest negative moves.  Eight independent chains, then a tree.
__device__ __forceinline__ float max32(const float (&v)[32]) {
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = fmaxf(fmaxf(v[e], v[8 + e]), fmaxf(v[16 + e], v[24 + e]));
  return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}

// NCTA = 2: launched with cluster dimension 2; p.tiles_m and the tile list count 256-row PAIR blocks, tmapB has 128-row boxes.
// 256-bit global store (sm_100: STG.E.256): one full 32-byte sector per lane
__device__ __forceinline__ void st_global_v8(float* dst, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// How the similarity epilogue writes S (measured at B = 8192, D = 512, pair kernel; MMA-only floor 74 us):
//   NPAIR_EPI_STG = 1 (default): DIRECT chunks leave the registers with 256-bit stores (lane = row, 128 contiguous bytes per
//     lane, no staging / proxy fence / bulk-store wait): 86 us with direct stores only, against 93 us for TMA-stored boxes.
//   NPAIR_EPI_MSTG = 0 (default): MIRRORED chunks are transposed through a 128B-swizzled per-warp staging tile and leave as one
//     32x32 fp32 TMA-store box; = 1 reads the transposed rows back and uses the 256-bit store (measured slower: 129 vs 110 us
//     with both on TMA).
#ifndef NPAIR_EPI_STG
#define NPAIR_EPI_STG 1
#endif
#ifndef NPAIR_EPI_MSTG
#define NPAIR_EPI_MSTG 0
#endif

template <int NSPLIT, bool BF16, int EPI, int BK_, int NCTA = 1>
__global__ void __launch_bounds__(384, 1)
split_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                  const __grid_constant__ CUtensorMap tmapS, const GemmParams p) {
  using Cfg = GemmCfg<NSPLIT, BK_, NCTA>;
  static_assert(NCTA == 1 || (NCTA == 2 && EPI != EPI_OUT), "pair mode is implemented for the similarity epilogues");
  const int cta_rank = (NCTA == 2) ? static_cast<int>(blockIdx.x & 1u) : 0;      // cluster = blocks {2c, 2c+1}
  const int worker = static_cast<int>(blockIdx.x) / NCTA, num_workers = static_cast<int>(gridDim.x) / NCTA;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer in the shared address space (offset arithmetic, no integer round trip): LDS/STS, not generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* store_stage = smem + STAGES * Cfg::STAGE_BYTES;          // [4 warps][32 rows][128 B], 128B-swizzled
  uint8_t* aux = store_stage + Cfg::STORE_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);            // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                           // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_lab = reinterpret_cast<float*>(aux + 256);                // [256] column labels of the current tile (EPI_SIM*)
  float* s_labr = reinterpret_cast<float*>(aux + 256 + 1024);        // [128] row labels of the current tile (EPI_SIM_SYM)
  float2* s_rng = reinterpret_cast<float2*>(aux + 256 + 1024 + 512); // [8] {min, max} label of each 32-column chunk, [8..12) of each 32-row group

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tile_list ? p.num_tiles_list : p.tiles_m * p.tiles_n * p.splits;
  const float inv_scale = p.dev_scale ? *p.dev_scale : 1.f;
  const float out_scale = (EPI != EPI_OUT) ? inv_scale * inv_scale : 1.f;
  const float alpha = p.alpha * inv_scale;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmapA);
    ptx::prefetch_tmap(&tmapB);
    if (EPI != EPI_OUT) ptx::prefetch_tmap(&tmapS);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 8 * NCTA); }
    ptx::fence_mbar_init();
  }
  if (NCTA == 2) ptx::cluster_sync_all();      // both CTAs of the pair are resident before the pair-wide tensor-memory allocation
  if (warp == 2) {
    ptx::tmem_alloc<512, NCTA>(tmem_ptr);
    ptx::tmem_relinquish<NCTA>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (NCTA == 2) ptx::cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const int mn = tile / p.splits, split = tile - mn * p.splits;
        int m_blk = mn / p.tiles_n, n_blk = mn % p.tiles_n;
        if (p.tile_list) { const int2 tl = p.tile_list[tile]; m_blk = tl.x; n_blk = tl.y; }
        m_blk = m_blk * NCTA + cta_rank;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
          if (NCTA == 1) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
            for (int s = 0; s < NSPLIT; ++s) {
              ptx::tma_load_3d(st + s * Cfg::A_PIECE, &tmapA, &full_bar[stage], kb * BK, m_blk * BM, s);
              ptx::tma_load_3d(st + NSPLIT * Cfg::A_PIECE + s * Cfg::B_PIECE, &tmapB, &full_bar[stage], kb * BK, n_blk * BN, s);
            }
          } else {
            // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of the whole pair stage (the peer's
            // bytes may land before the leader's expect_tx: the transaction count may go negative inside a phase)
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES * NCTA);
            const uint32_t lead_full = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
#pragma unroll
            for (int s = 0; s < NSPLIT; ++s) {
              ptx::tma_load_3d_pair(st + s * Cfg::A_PIECE, &tmapA, lead_full, kb * BK, m_blk * BM, s);
              ptx::tma_load_3d_pair(st + NSPLIT * Cfg::A_PIECE + s * Cfg::B_PIECE, &tmapB, lead_full, kb * BK,
                                    n_blk * BN + cta_rank * Cfg::B_ROWS, s);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BF16, BM * NCTA, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int split = tile % p.splits;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a0 = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b0 = a0 + NSPLIT * Cfg::A_PIECE;
#pragma unroll
          for (int ps = 0; ps < Cfg::NPASS; ++ps) {
            int sa, sb;
            pass_pieces(NSPLIT, ps, sa, sb);
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              const uint64_t ad = ptx::make_kmajor_desc(a0 + sa * Cfg::A_PIECE + k4 * 32, Cfg::SBO, Cfg::LAYOUT);
              const uint64_t bd = ptx::make_kmajor_desc(b0 + sb * Cfg::B_PIECE + k4 * 32, Cfg::SBO, Cfg::LAYOUT);
              if (NCTA == 1) ptx::mma_f16_ss(d_tmem, ad, bd, idesc, ((kb - kb0) | ps | k4) != 0 ? 1u : 0u);
              else ptx::mma_f16_ss_pair(d_tmem, ad, bd, idesc, ((kb - kb0) | ps | k4) != 0 ? 1u : 0u);
            }
          }
          // smem slot reusable once these MMAs retire (pair mode: in both CTAs)
          if (NCTA == 1) ptx::mma_commit(&empty_bar[stage]); else ptx::mma_commit_pair(&empty_bar[stage], 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (pair mode: each CTA drains its own 128 rows)
        if (NCTA == 1) ptx::mma_commit(&tfull_bar[acc]); else ptx::mma_commit_pair(&tfull_bar[acc], 3);
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    // 8 epilogue warps: warp w reads TMEM lanes 32*(w%4) (hardware rule) and the column half (w-4)/4 of the 256-wide tile
    const int ew = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;            // 0..255
    constexpr int SB = Cfg::STORE_BUFS;
    uint8_t* const stg0 = store_stage + (warp - 4) * 4096 * SB;
    uint32_t sgrp = 0;                            // bulk-store groups issued by this warp (SB == 2: alternate the two tiles)
    int it = 0;
    const uint32_t lead_tempty = (NCTA == 2) ? ptx::mapa_u32(ptx::smem_u32(&tempty_bar[0]), 0) : 0u;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
      const int mn = tile / p.splits, split = tile - mn * p.splits;
      int m_blk = mn / p.tiles_n, n_blk = mn % p.tiles_n;
      if (p.tile_list) { const int2 tl = p.tile_list[tile]; m_blk = tl.x; n_blk = tl.y; }
      m_blk = m_blk * NCTA + cta_rank;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_blk * BM + ew * 32 + lane;
      const int col_base = n_blk * BN;
      float lab_i = 0.f;
      if (EPI != EPI_OUT) {
        asm volatile("bar.sync 1, 256;" ::: "memory");   // previous tile's readers are done with s_lab
        s_lab[et] = (col_base + et < p.Nn) ? p.lab_cols[col_base + et] : 0.f;
        if (row < p.M) lab_i = p.lab_rows[row];
        if (EPI == EPI_SIM_SYM && half == 0) s_labr[ew * 32 + lane] = lab_i;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // label range of every 32-column chunk (epilogue warp w: chunk w) and of every 32-row group (warps 0..3): a row whose label
        // lies outside a chunk's range has no same-label pair in it, and its statistics reduce to one running maximum
        {
          const float lc = s_lab[(warp - 4) * 32 + lane];
          float mn = lc, mx = lc;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
          if (lane == 0) s_rng[warp - 4] = make_float2(mn, mx);
          if (EPI == EPI_SIM_SYM && warp - 4 < 4) {
            const float lr = s_labr[(warp - 4) * 32 + lane];
            float rn = lr, rx = lr;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { rn = fminf(rn, __shfl_xor_sync(0xffffffffu, rn, o)); rx = fmaxf(rx, __shfl_xor_sync(0xffffffffu, rx, o)); }
            if (lane == 0) s_rng[8 + warp - 4] = make_float2(rn, rx);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;

      float minw = FLT_MAX, maxw = -FLT_MAX, maxb = -FLT_MAX, maxall = -FLT_MAX;
      int cnt = 0;
      const int self_col = row + p.self_offset;
#pragma unroll 1
      for (int ch = half * 4; ch < half * 4 + 4; ++ch) {
        const int col0 = col_base + ch * 32;
        const int cb = col0 >> 7;                          // 128-wide column block (EPI_SIM_SYM bookkeeping)
        if (EPI == EPI_SIM_SYM && cb < m_blk) continue;    // lower-triangle half of a straddling tile: produced by mirroring
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + ch * 32, r);
        ptx::tmem_ld_wait();
#ifndef NPAIR_DBG_EPI_LEVEL
#define NPAIR_DBG_EPI_LEVEL 9
#endif
        if (EPI != EPI_OUT && NPAIR_DBG_EPI_LEVEL == 0) continue;
        if (EPI != EPI_OUT) {
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(r[c]) * out_scale;
          // registers -> 128B-swizzled staging tile -> one TMA store of a 32x32 fp32 box (full 128-byte lines;
          // rows >= M and columns >= Nn are clipped by the tensor map)
          if (NPAIR_EPI_STG) {
            // ldS is a multiple of 32, so a partial last chunk stores zeros (zero-filled operand rows) into the row padding
            if (NPAIR_DBG_EPI_LEVEL >= 1 && row < p.M && col0 < p.Nn) {
              float* dst = p.S + static_cast<long long>(row) * p.ldS + col0;
#pragma unroll
              for (int q = 0; q < 4; ++q) st_global_v8(dst + 8 * q, v + 8 * q);
            }
          } else if (NPAIR_DBG_EPI_LEVEL >= 1 && m_blk * BM + ew * 32 < p.M && col0 < p.Nn) {          // warp-uniform
            // this warp's previous box (SB == 2: the one before it, which used the same tile) has been read out of smem
            if (lane == 0) { if (SB == 2) ptx::tma_store_wait_read<1>(); else ptx::tma_store_wait_read<0>(); }
            __syncwarp();
            uint8_t* const stg = stg0 + (SB == 2 ? (sgrp & 1u) * 4096u : 0u);
            ++sgrp;
            uint8_t* srow = stg + lane * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(srow + ((q ^ (lane & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            ptx::fence_proxy_async_smem();
            __syncwarp();
#ifndef NPAIR_DBG_SKIP_DSTORE
            if (lane == 0) { ptx::tma_store_2d(&tmapS, stg, col0, m_blk * BM + ew * 32); ptx::tma_store_commit(); }
#endif
          }
          // statistics AFTER issuing the store: the ~200 ALU instructions hide the bulk store's shared-memory read
          if (NPAIR_DBG_EPI_LEVEL >= 2 && col0 < p.Nn) {
            const float2 rg = s_rng[ch];
            // warp-uniform: every row of this warp is outside the chunk's label range (and the chunk is whole: the self pair is a
            // same-label pair, so it cannot be in such a chunk)
            if (__all_sync(0xffffffffu, row >= p.M || lab_i < rg.x || lab_i > rg.y) && col0 + 32 <= p.Nn) {
              if (row < p.M) maxb = fmaxf(maxb, max32(v));
            } else if (row < p.M)
              stats32(v, s_lab + ch * 32, lab_i, col0 + 32 <= p.Nn && (self_col < col0 || self_col >= col0 + 32), col0, p.Nn, self_col,
                      minw, maxw, maxb, cnt);
          }
          if (NPAIR_DBG_EPI_LEVEL >= 3 && EPI == EPI_SIM_SYM && cb > m_blk && m_blk * BM + ew * 32 < p.M && col0 < p.Nn) {
            // ---- mirrored store: staging row c holds S[col0 + c][rows of this warp]; box lands at (x = row block, y = col0) ----
            if (!NPAIR_EPI_MSTG && lane == 0) { if (SB == 2) ptx::tma_store_wait_read<1>(); else ptx::tma_store_wait_read<0>(); }
            __syncwarp();                                  // (register-store variant: the previous read-back is complete)
            uint8_t* const stg = stg0 + ((!NPAIR_EPI_MSTG && SB == 2) ? (sgrp & 1u) * 4096u : 0u);
            ++sgrp;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              *reinterpret_cast<float*>(stg + c * 128 + ((((lane >> 2) ^ (c & 7))) << 4) + ((lane & 3) << 2)) = v[c];
            if (!NPAIR_EPI_MSTG) ptx::fence_proxy_async_smem();
            __syncwarp();
#ifndef NPAIR_DBG_SKIP_MSTORE
            if (!NPAIR_EPI_MSTG && lane == 0) { ptx::tma_store_2d(&tmapS, stg, m_blk * BM + ew * 32, col0); ptx::tma_store_commit(); }
#endif
            // ---- mirrored statistics: the staging tile is the transposed chunk, so lane L reads back ROW gc = col0 + L of the
            //      symmetric matrix (32 entries against this warp's 32 row labels) and reuses the per-thread statistics ----
            const int gc = col0 + lane;
            float vt[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t4 = *reinterpret_cast<const float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4));
              vt[4 * q] = t4.x; vt[4 * q + 1] = t4.y; vt[4 * q + 2] = t4.z; vt[4 * q + 3] = t4.w;
            }
            float t_minw = FLT_MAX, t_maxw = -FLT_MAX, t_maxb = -FLT_MAX;
            int t_cnt = 0;
            const int r0 = m_blk * BM + ew * 32;
            if (NPAIR_EPI_MSTG && gc < p.Nn) {             // row gc of the symmetric matrix, columns r0 .. r0 + 31
              float* dst = p.S + static_cast<long long>(gc) * p.ldS + r0;
#pragma unroll
              for (int q = 0; q < 4; ++q) st_global_v8(dst + 8 * q, vt + 8 * q);
            }
            if (NPAIR_DBG_EPI_LEVEL >= 4) {
              const float lab_c = s_lab[ch * 32 + lane];
              const float2 rr = s_rng[8 + ew];
              const bool plain = __all_sync(0xffffffffu, gc >= p.Nn || lab_c < rr.x || lab_c > rr.y) && r0 + 32 <= p.M;
              if (gc < p.Nn) {
              if (plain) t_maxb = max32(vt);
              else stats32(vt, s_labr + ew * 32, lab_c, r0 + 32 <= p.M, r0, p.M, -1, t_minw, t_maxw, t_maxb, t_cnt);
              if (t_cnt) {
                atomicMin(&p.st_minw[gc], f2ord(t_minw));
                atomicMax(&p.st_maxw[gc], f2ord(t_maxw));
                atomicAdd(&p.cnt_same[gc], t_cnt);
              }
              atomicMax(&p.st_maxb[gc], f2ord(t_maxb));
              atomicMax(&p.st_maxall[gc], f2ord(fmaxf(t_maxw, t_maxb)));
              }
            }
          }
        } else {
          if (row < p.M) {
            float* obase = p.splits > 1 ? p.part + static_cast<long long>(split) * p.M * p.ldo : p.out;
            const float beta = p.splits > 1 ? 0.f : p.beta;
            float* dst = obase + static_cast<long long>(row) * p.ldo + col0;
            if (col0 + 32 <= p.Nn && (p.ldo & 3) == 0) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 o = make_float4(alpha * __uint_as_float(r[4 * q]), alpha * __uint_as_float(r[4 * q + 1]),
                                       alpha * __uint_as_float(r[4 * q + 2]), alpha * __uint_as_float(r[4 * q + 3]));
                if (beta != 0.f) {
                  const float4 old = reinterpret_cast<float4*>(dst)[q];
                  o.x += beta * old.x; o.y += beta * old.y; o.z += beta * old.z; o.w += beta * old.w;
                }
                reinterpret_cast<float4*>(dst)[q] = o;
              }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (col0 + c < p.Nn) {
                  float o = alpha * __uint_as_float(r[c]);
                  if (beta != 0.f) o += beta * dst[c];
                  dst[c] = o;
                }
            }
          }
        }
      }
      // accumulator drained -> MMA warp may overwrite it
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (NCTA == 1) ptx::mbar_arrive(&tempty_bar[acc]); else ptx::mbar_arrive_cluster(lead_tempty + 8u * acc); }
      if (EPI != EPI_OUT && row < p.M) {
        maxall = fmaxf(maxw, maxb);                       // every valid column is either same- or diff-label
        if (cnt) {
          atomicMin(&p.st_minw[row], f2ord(minw));
          atomicMax(&p.st_maxw[row], f2ord(maxw));
          atomicAdd(&p.cnt_same[row], cnt);
        }
        atomicMax(&p.st_maxb[row], f2ord(maxb));
        atomicMax(&p.st_maxall[row], f2ord(maxall));
      }
    }
  }
  if (EPI != EPI_OUT && warp >= 4 && lane == 0) ptx::tma_store_wait<0>();   // bulk stores complete before exit
  ptx::tc_fence_before();
  __syncthreads();
  if (NCTA == 2) ptx::cluster_sync_all();      // no CTA of the pair exits while the other may still signal or read it
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512, NCTA>(tmem_base);
  }
  if (EPI != EPI_OUT && p.fuse_thr) {
    // every CTA's statistics atomics are out; the last CTA to get here picks the thresholds for the whole block of rows
    // (a static __shared__ flag would push static + dynamic shared memory past the 227 KB a CTA may ask for)
    volatile int& s_last_cta = *reinterpret_cast<volatile int*>(aux + 192);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last_cta = (atomicAdd(&p.bs->ticket2, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last_cta) {
      __threadfence();
      thresholds_one_block(p.ra, p.M, p.Nn, p.mp, p.bs, smem);     // the operand ring is idle: reuse its first bytes as scratch
      if (threadIdx.x == 0) p.bs->ticket2 = 0;
    }
  }
}

}  // namespace npair
