// grad_fused.cuh -- gradient GEMM with the weight builder fused in as its A-operand producer (sm_100a).
//
//   dX[j][:] = alpha * sum_m H[j][m] * X_total[m][:]        H[j][m] = g'(S[j][m]; row j) + (1/world) g'(S[j][m]; row m)
//
// Replaces Get_Query_Diff_Part x3 + six cublasSgemm + the N x D all-reduce of the reference (npair_multi_class_loss.cu:438-497).
// H (Q x N) is never written to HBM: per 32-column K block the 16 producer warps read a 128 x 32 fp32 tile of S (TMA,
// 128B-swizzled) and the 32 column records (bulk copy), evaluate the weights in registers (row record in registers,
// thread = row), split them into the 2-byte operand pieces and store those with tcgen05.st straight into TENSOR MEMORY,
// from where tcgen05.mma takes its A operand (the multi-pass GEMM is shared-memory-bandwidth bound in 1-CTA mode: every
// pass re-reads its operands from smem, so keeping A out of smem removes a third of that traffic).  Requires a bitwise
// symmetric S (EPI_SIM_SYM tiles at world == 1, K-concatenated operands across ranks) -- see gemm_tcgen05.cuh / kernels.cu.
//
// Chunked accumulation.  Every tcgen05.mma TRUNCATES the fp32 accumulator (measured: a bias of about -0.45 * 2^-24 * |acc| per
// accumulating instruction, tests/diag_gemm_error.py), and this GEMM's K is the database size: at N = 8192 the 1536 instructions
// per output element put the gradient 2.4e-5 (normwise) from the exact value -- over the 1e-5 parity bar, and growing with N.  The
// K range is therefore cut into chunks of `chunk_kb` K blocks: each chunk accumulates in tensor memory from zero, the producer
// warps drain it and ADD it to the output in fp32 round-to-nearest (first chunk: plain store, later chunks: red.global.add.v4.f32
// by the same thread to the same addresses, hence in program order and deterministic).  The error is then bounded by the chunk
// length (2.7e-6 measured at 32 K blocks) for any N.  The drain is done by four DEDICATED warps: when the weight producers did it
// themselves every drain cost ~6 us (they idled until the tensor pipe had caught up, then the pipe idled until they had refilled
// the ring; polling for the accumulator between K blocks was worse still: the producers are the critical path of this kernel).
//
// CTA = 704 threads: warp 0 TMA producer (B pieces of X^T, S tile, column records), warp 1 MMA issuer, warps 4-19 weight
// producers (warp w: TMEM lanes 32*(w%4), 8 of the 32 K columns), warps 2, 3, 20, 21 accumulator drain (one per TMEM lane
// quarter; warp 2 also allocates the tensor memory).  NCTA = 2: CTA-pair mode, see FusedCfg.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace npair {

#define NPAIR_LOG2E_F 1.4426950408889634f

struct FusedGradParams {
  int Q, N, D;
  int num_kblocks;              // ceil(N / 32)
  int tiles_m, tiles_n;         // 128-row blocks, 256-column tiles of D
  int splits, kb_per_split;     // split-K over the sample index (few row blocks when Q = B / world is small)
  int chunk_kb;                 // accumulation chunk in K blocks (0 = the whole K range in one accumulator), see below
  float* part;                  // split-K partials [split][Q][ldo]
  const float* S;               // only for address checks; tiles come through the tensor map
  const float* rowrec;          // [Q][8]  this rank's row records {m2c, thr_n, m2, label | thr_p, cA, cT, 0} (lse_rows_kernel)
  const float* colrec;          // [N][8]  records of every column's row (== rowrec when world == 1)
  int self_offset;              // global column of local row 0
  float inv_world, log2_world;
  float sgn_p, sgn_n;           // +-1: direction of the same-/diff-label selection compare
  float* out; long long ldo;
  float alpha, beta;
  const float* dev_scale;       // inverse operand pre-scale (power of two) or NULL
};

// NCTA = 2: CTA-pair mode (cluster of 2, tcgen05.mma.cta_group::2, M = 256).  Each CTA produces the weights of its own 128 rows
// into its own tensor memory and stages 128 of the 256 rows of the B tile; the leader's MMA reads both halves.  Halves the
// B traffic through each SM's shared memory (the bound of this kernel); p.tiles_m then counts 256-row pair blocks.
template <int NSPLIT, int NCTA = 1>
struct FusedCfg {
  static constexpr int BM = 128, BN = 256, BK = 32;
  static constexpr int B_ROWS = BN / NCTA;
  static constexpr int B_PIECE = B_ROWS * 64;             // 64-byte rows (32 x 2-byte), SWIZZLE_64B
  static constexpr int S_TILE = BM * 128;                 // 128-byte rows (32 x fp32), SWIZZLE_128B
  static constexpr int CREC = BK * 32;                    // 32 column records of 32 bytes
  static constexpr int STAGE_BYTES = NSPLIT * B_PIECE + S_TILE + CREC;
  static constexpr int STAGES = (NCTA == 1) ? ((NSPLIT == 1) ? 6 : (NSPLIT == 2 ? 4 : 3)) : ((NSPLIT == 1) ? 8 : (NSPLIT == 2 ? 6 : 5));
  // tensor memory: columns [0,256) = the 128 x 256 fp32 accumulator; A operand pieces behind it,
  // 16 packed columns (32 two-byte K elements) per piece and stage
  static constexpr int TMEM_A0 = 256;
  static constexpr int A_COLS = BK / 2;
  static constexpr int NPASS = (NSPLIT == 1) ? 1 : (NSPLIT == 2 ? 3 : 6);
  static constexpr int DRAIN_STAGE = 4 * 4096;            // one 32 x 32 fp32 transposition tile per drain warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + DRAIN_STAGE + 1024 /*barriers*/ + 1024 /*alignment*/;
  static constexpr int THREADS = 704;                         // 22 warps: TMA, MMA, 16 weight producers (4..19), 4 drain warps (2, 3, 20, 21)
};

// two fp32 weights -> packed 2-byte pieces (lo 16 bits = first value)
template <int NSPLIT, bool BF16>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t (&out)[3]) {
  if (NSPLIT == 1) {
    out[0] = static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(a))) | (static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(b))) << 16);
  } else if (NSPLIT == 2) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    out[0] = *reinterpret_cast<const uint32_t*>(&h);
    out[1] = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const float r1a = a - hf.x, r1b = b - hf.y;
    const __nv_bfloat162 m = __floats2bfloat162_rn(r1a, r1b);
    const float2 mf = __bfloat1622float2(m);
    const __nv_bfloat162 l = __floats2bfloat162_rn(r1a - mf.x, r1b - mf.y);
    out[0] = *reinterpret_cast<const uint32_t*>(&h);
    out[1] = *reinterpret_cast<const uint32_t*>(&m);
    out[2] = *reinterpret_cast<const uint32_t*>(&l);
  }
}

// fp32 round-to-nearest adds performed at the L2 (REDG.E.ADD.F32x4.RN): same thread, same address => program order
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* dst, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(a) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

// End of the accumulation chunk that starts at K block c0 of the range [kb0, kb1).  The FIRST chunk of a tile is shortened by an
// amount that depends on the tile (`key`: its 256-row block, column block and split -- NOT on which CTA computes it, so the
// summation order, hence every bit of the result, is independent of the grid and of the 1-CTA / CTA-pair variant): neighbouring
// tiles do not all drain at the same moment (bursts of 128 KB of L2 reductions per SM).
__device__ __forceinline__ int chunk_end(int c0, int kb0, int kb1, int ch, int key) {
  if (ch <= 0) return kb1;
  if (c0 == kb0) { const int first = max(1, (((key & 7) + 1) * ch) >> 3); return min(kb1, kb0 + first); }
  return min(kb1, c0 + ch);
}

// Weights of 8 consecutive diff-label pairs of one row (the common case): per pair TWO exponentials whose arguments already carry
// the factors 1/T (and 1/world for the transposed term) -- see lse_rows_kernel -- switched off by a -inf argument when the pair
// is not selected.  NEG: the diff-label rule compares -s (HARD / RELATIVE_HARD negatives), folded into the compare's operand sign.
template <bool NEG>
__device__ __forceinline__ bool produce8(const float (&sv)[8], const float4* __restrict__ crec, float r_m2r, float r_tn, float r_lab, float (&g)[8]) {
  bool any_same = false;
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    const float4 ca = crec[2 * cc];              // {m2c, thr_n, m2, label} of the column's row
    const float s = sv[cc];
    const float a1 = fmaf(s, NPAIR_LOG2E_F, -r_m2r);
    const float a2 = fmaf(s, NPAIR_LOG2E_F, -ca.x);
    const bool s1 = NEG ? (-s <= r_tn) : (s <= r_tn);
    const bool s2 = NEG ? (-s <= ca.y) : (s <= ca.y);
    float e1, e2;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1 ? a1 : -INFINITY));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(s2 ? a2 : -INFINITY));
    g[cc] = e1 + e2;
    any_same |= (ca.w == r_lab);
  }
  return any_same;
}

// Measured and dropped (round 2, profiles/r02_experiments.md): forming the transposed term's exponential from the row term's
// (one ex2 per pair plus per-row / per-column constants) was SLOWER (180 vs 166 us): the extra shuffles and selects cost more
// than the saved MUFU issue slots.
template <int NSPLIT, bool BF16, int NCTA = 1>
__global__ void __launch_bounds__(704, 1)     // 22 warps x 80 registers (88 would not fit: the register file is handed out in 512-register units per warp)
fused_grad_kernel(const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapS, const FusedGradParams p) {
  using Cfg = FusedCfg<NSPLIT, NCTA>;
  static_assert(Cfg::TMEM_A0 + Cfg::STAGES * NSPLIT * Cfg::A_COLS <= 512, "A pieces do not fit behind the accumulator");
  const int cta_rank = (NCTA == 2) ? static_cast<int>(blockIdx.x & 1u) : 0;      // cluster = blocks {2c, 2c+1}
  const int worker = static_cast<int>(blockIdx.x) / NCTA, num_workers = static_cast<int>(gridDim.x) / NCTA;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* drain_stage = smem + STAGES * Cfg::STAGE_BYTES;
  uint8_t* aux = drain_stage + Cfg::DRAIN_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);            // [STAGES] TMA bytes landed (B, S tile, column records)
  uint64_t* bfull_bar = full_bar + STAGES;                           // [STAGES] pair mode: both CTAs' B halves landed (leader's copy is used)
  uint64_t* aready_bar = bfull_bar + STAGES;                         // [STAGES] producers wrote the A pieces (pair mode: of both CTAs, leader's copy)
  uint64_t* empty_bar = aready_bar + STAGES;                         // [STAGES] MMAs of the stage retired
  uint64_t* tfull_bar = empty_bar + STAGES;                          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n * p.splits;
  const float inv_scale = p.dev_scale ? *p.dev_scale : 1.f;
  const float alpha = p.alpha * inv_scale;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmapB); ptx::prefetch_tmap(&tmapS); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&bfull_bar[s], 1); ptx::mbar_init(&aready_bar[s], 16 * NCTA); ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 4 * NCTA); }
    ptx::fence_mbar_init();
  }
  if (NCTA == 2) ptx::cluster_sync_all();      // both CTAs of the pair are resident before the pair-wide tensor-memory allocation
  if (warp == 2) { ptx::tmem_alloc<512, NCTA>(tmem_ptr); ptx::tmem_relinquish<NCTA>(); }
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
        const int m_blk = (mn / p.tiles_n) * NCTA + cta_rank, n_blk = mn % p.tiles_n;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const int m0 = kb * BK;
          const uint32_t crec_bytes = static_cast<uint32_t>(min(BK, p.N - m0)) * 32u;
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
          if (NCTA == 1) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], NSPLIT * Cfg::B_PIECE + Cfg::S_TILE + crec_bytes);
#pragma unroll
            for (int s = 0; s < NSPLIT; ++s)
              ptx::tma_load_3d(st + s * Cfg::B_PIECE, &tmapB, &full_bar[stage], m0, n_blk * BN, s);
          } else {
            // own S tile + column records on the local barrier (the local producers wait for them); the B halves of both CTAs
            // complete on the leader's bfull barrier (the MMA issuer waits for it)
            ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::S_TILE + crec_bytes);
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&bfull_bar[stage], NSPLIT * Cfg::B_PIECE * NCTA);
            const uint32_t lead_bfull = ptx::mapa_u32(ptx::smem_u32(&bfull_bar[stage]), 0);
#pragma unroll
            for (int s = 0; s < NSPLIT; ++s)
              ptx::tma_load_3d_pair(st + s * Cfg::B_PIECE, &tmapB, lead_bfull, m0, n_blk * BN + cta_rank * Cfg::B_ROWS, s);
          }
          ptx::tma_load_2d(st + NSPLIT * Cfg::B_PIECE, &tmapS, &full_bar[stage], m0, m_blk * BM);
          bulk_copy_g2s(st + NSPLIT * Cfg::B_PIECE + Cfg::S_TILE, p.colrec + 8ll * m0, crec_bytes, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BF16, BM * NCTA, BN);
      int stage = 0; uint32_t phase = 0;
      uint32_t gen = 0;                          // accumulator generations (one per chunk), drained by warps 2, 3, 20, 21
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const uint32_t d_tmem = tmem_base;
        const int mn = tile / p.splits, split = tile - mn * p.splits;
        const int ckey = ((mn / p.tiles_n) >> (NCTA == 1 ? 1 : 0)) + mn % p.tiles_n + split;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
        const int ch = p.chunk_kb;
        for (int c0 = kb0, c1; c0 < kb1; c0 = c1, ++gen) {
          c1 = chunk_end(c0, kb0, kb1, ch, ckey);
          ptx::mbar_wait(&tempty_bar[0], (gen & 1u) ^ 1u);
          ptx::tc_fence_after();
          for (int kb = c0; kb < c1; ++kb) {
            ptx::mbar_wait(NCTA == 1 ? &full_bar[stage] : &bfull_bar[stage], phase);
            ptx::mbar_wait(&aready_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t b0 = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t a_t = tmem_base + Cfg::TMEM_A0 + stage * NSPLIT * Cfg::A_COLS;
#pragma unroll
            for (int ps = 0; ps < Cfg::NPASS; ++ps) {
              int sa, sb;
              pass_pieces(NSPLIT, ps, sa, sb);
#pragma unroll
              for (int k2 = 0; k2 < BK / 16; ++k2) {
                const uint64_t bd = ptx::make_kmajor_desc(b0 + sb * Cfg::B_PIECE + k2 * 32, 512u, 4u);   // SWIZZLE_64B, 8 rows = 512 B
                if (NCTA == 1) ptx::mma_f16_ts(d_tmem, a_t + sa * Cfg::A_COLS + k2 * 8, bd, idesc, ((kb - c0) | ps | k2) != 0 ? 1u : 0u);
                else ptx::mma_f16_ts_pair(d_tmem, a_t + sa * Cfg::A_COLS + k2 * 8, bd, idesc, ((kb - c0) | ps | k2) != 0 ? 1u : 0u);
              }
            }
            if (NCTA == 1) ptx::mma_commit(&empty_bar[stage]); else ptx::mma_commit_pair(&empty_bar[stage], 3);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (NCTA == 1) ptx::mma_commit(&tfull_bar[0]); else ptx::mma_commit_pair(&tfull_bar[0], 3);
        }
      }
    }
  } else if (warp >= 4 && warp < 20) {
    // ===================================== weight producers =====================================
    const int ew = (warp - 4) & 3;               // TMEM lane group / 32-row group
    const int qt = (warp - 4) >> 2;              // quarter: 8 of the 32 K columns
    int stage = 0; uint32_t phase = 0;
    int prev_stage = -1;                         // stage whose tcgen05.st are issued but not yet published to the MMA issuer
    const uint32_t lead_aready = (NCTA == 2) ? ptx::mapa_u32(ptx::smem_u32(&aready_bar[0]), 0) : 0u;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      const int mn = tile / p.splits, split = tile - mn * p.splits;
      const int m_blk = (mn / p.tiles_n) * NCTA + cta_rank;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
      const int rl = ew * 32 + lane;             // row inside the tile
      const int row = m_blk * BM + rl;
      // row record (neutral when the row does not exist: thresholds -inf -> nothing selected)
      float r_m2 = 0.f, r_m2r = INFINITY, r_tp = -INFINITY, r_tn = -INFINITY, r_cA = 0.f, r_lab = 0.f;
      if (row < p.Q) {
        const float4 a = *reinterpret_cast<const float4*>(p.rowrec + 8ll * row);
        const float4 b = *reinterpret_cast<const float4*>(p.rowrec + 8ll * row + 4);
        r_m2r = a.x - p.log2_world; r_tn = a.y; r_m2 = a.z; r_lab = a.w; r_tp = b.x; r_cA = b.y;   // the row term carries no 1/world
      }
      const int self_col = row + p.self_offset;
      const bool neg_n = p.sgn_n < 0.f;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
        const uint8_t* s_tile = st + NSPLIT * Cfg::B_PIECE;
        const float4* crec = reinterpret_cast<const float4*>(s_tile + Cfg::S_TILE) + 2 * (8 * qt);
        // my 8 similarities: chunks 2*qt, 2*qt+1 of row rl (128B swizzle: chunk ^ (row & 7))
        float sv[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 t4 = *reinterpret_cast<const float4*>(s_tile + rl * 128 + (((2 * qt + q) ^ (rl & 7)) << 4));
          sv[4 * q] = t4.x; sv[4 * q + 1] = t4.y; sv[4 * q + 2] = t4.z; sv[4 * q + 3] = t4.w;
        }
        const int m0 = kb * BK + 8 * qt;         // global column of sv[0]
        float g[8];
        const bool any_same = neg_n ? produce8<true>(sv, crec, r_m2r, r_tn, r_lab, g) : produce8<false>(sv, crec, r_m2r, r_tn, r_lab, g);
        // rare fix-ups: same-label pairs (the other selection rule and weight), the self pair, columns beyond N
        if (any_same || (self_col >= m0 && self_col < m0 + 8) || m0 + 8 > p.N) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 ca = crec[2 * cc], cb = crec[2 * cc + 1];   // cb = {thr_p, cA, cT, -}
            if (ca.w == r_lab) {
              const float s = sv[cc];
              float e1, e2;                      // same exponential as the forward row pass (fast_exp_m2)
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(s, NPAIR_LOG2E_F, -r_m2)));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(s, NPAIR_LOG2E_F, -ca.z)));
              const float key = s * p.sgn_p;
              const float w1 = (key <= r_tp) ? e1 * r_cA : 0.f;
              const float w2 = (key <= cb.x) ? e2 * cb.y : 0.f;
              g[cc] = fmaf(w2, p.inv_world, w1);
            }
            if (m0 + cc == self_col || m0 + cc >= p.N) g[cc] = 0.f;
          }
        }
        // pieces -> tensor memory: lane = row, 4 packed columns (8 K elements) at column offset 4*qt of each piece
        uint32_t pk[NSPLIT][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t o[3];
          split_pair<NSPLIT, BF16>(g[2 * q], g[2 * q + 1], o);
#pragma unroll
          for (int s = 0; s < NSPLIT; ++s) pk[s][q] = o[s];
        }
        // the previous K block's tensor-memory stores have had this block's arithmetic to complete: publish them now
        if (prev_stage >= 0) {
          ptx::tmem_st_wait();
          ptx::tc_fence_before();                 // order the tcgen05.st before the arrive that releases the MMA issuer
          __syncwarp();
          if (lane == 0) { if (NCTA == 1) ptx::mbar_arrive(&aready_bar[prev_stage]); else ptx::mbar_arrive_cluster(lead_aready + 8u * prev_stage); }
        }
        const uint32_t a_t = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + Cfg::TMEM_A0 + stage * NSPLIT * Cfg::A_COLS + 4 * qt;
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s) ptx::tmem_st_32x32b_x4(a_t + s * Cfg::A_COLS, pk[s][0], pk[s][1], pk[s][2], pk[s][3]);
        prev_stage = stage;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (prev_stage >= 0) {
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (NCTA == 1) ptx::mbar_arrive(&aready_bar[prev_stage]); else ptx::mbar_arrive_cluster(lead_aready + 8u * prev_stage); }
    }
  }
  if (warp == 2 || warp == 3 || warp >= 20) {
    // ===================================== accumulator drain (4 warps, one per TMEM lane quarter) =====================================
    // Dedicated warps, so that the producers never wait for the tensor pipe: they hand every accumulator generation (one per
    // chunk) back to the MMA issuer as soon as its 128 x 256 values are in flight to the output.
    const int ew = warp & 3;                      // TMEM lane quarter this warp may access (warp id % 4)
    const uint32_t lead_tempty = (NCTA == 2) ? ptx::mapa_u32(ptx::smem_u32(&tempty_bar[0]), 0) : 0u;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    uint8_t* const stg = drain_stage + ew * 4096;       // this warp's 32 x 32 fp32 transposition tile
    uint32_t gen = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      const int mn = tile / p.splits, split = tile - mn * p.splits;
      const int m_blk = (mn / p.tiles_n) * NCTA + cta_rank, n_blk = mn % p.tiles_n;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
      const int row0 = m_blk * BM + ew * 32, row = row0 + lane;
      float* obase = p.splits > 1 ? p.part + static_cast<long long>(split) * p.Q * p.ldo : p.out;
      const float beta = p.splits > 1 ? 0.f : p.beta;
      for (int c0 = kb0, c1; c0 < kb1; c0 = c1, ++gen) {
        c1 = chunk_end(c0, kb0, kb1, p.chunk_kb, ((mn / p.tiles_n) >> (NCTA == 1 ? 1 : 0)) + n_blk + split);
        const bool first = (c0 == kb0);          // first chunk of the tile stores (+ beta * out), later chunks add (fp32 RN)
        ptx::mbar_wait(&tfull_bar[0], gen & 1u);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_row + ch * 32, r);
          ptx::tmem_ld_wait();
          if (ch == BN / 32 - 1) {                // everything is in registers / on its way out: the MMA issuer may start the next chunk
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (NCTA == 1) ptx::mbar_arrive(&tempty_bar[0]); else ptx::mbar_arrive_cluster(lead_tempty); }
          }
          const int col0 = n_blk * BN + ch * 32;
          if (col0 + 32 <= p.D && (p.ldo & 3) == 0) {
            // lane = row in registers -> 128B-swizzled 32 x 32 staging tile -> lane = (row / 4 groups, 16-byte column chunk): every
            // global request covers whole 128-byte lines (4 rows x 128 B per instruction instead of 32 half-used sectors)
            __syncwarp();                         // the previous tile's reads of the staging tile are complete
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                  make_float4(alpha * __uint_as_float(r[4 * q]), alpha * __uint_as_float(r[4 * q + 1]), alpha * __uint_as_float(r[4 * q + 2]),
                              alpha * __uint_as_float(r[4 * q + 3]));
            __syncwarp();
            const int cq = lane & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + (lane >> 3);
              const int grow = row0 + rr;
              float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + ((cq ^ (rr & 7)) << 4));
              if (grow < p.Q) {
                float* dst = obase + static_cast<long long>(grow) * p.ldo + col0 + 4 * cq;
                if (first) {
                  if (beta != 0.f) { const float4 old = *reinterpret_cast<const float4*>(dst); o.x += beta * old.x; o.y += beta * old.y; o.z += beta * old.z; o.w += beta * old.w; }
                  *reinterpret_cast<float4*>(dst) = o;
                } else red_add_v4(dst, o.x, o.y, o.z, o.w);
              }
            }
          } else if (row < p.Q) {                 // ragged D / unaligned rows: lane = row, element by element
            float* dst = obase + static_cast<long long>(row) * p.ldo + col0;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (col0 + c < p.D) {
                const float o = alpha * __uint_as_float(r[c]);
                if (first) dst[c] = (beta != 0.f) ? o + beta * dst[c] : o;
                else red_add_f32(dst + c, o);
              }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (NCTA == 2) ptx::cluster_sync_all();      // no CTA of the pair exits while the other may still signal or read it
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512, NCTA>(tmem_base);
  }
}

}  // namespace npair
