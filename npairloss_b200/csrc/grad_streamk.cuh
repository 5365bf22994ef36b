// grad_streamk.cuh -- stream-K variant of the fused gradient kernel (CTA-pair mode only).  OPT-IN (NPAIR_GRAD_STREAMK=1):
// written after round 1's GPU budget was spent, not run yet.
//
// fused_grad_kernel gives each cluster whole 256 x 256 output blocks; at D = 512 there are only Q/128 = 64 of them for 74
// clusters (20 SMs idle) and for Q = B/world small it needs split-K partials plus a reduce kernel.  Here the (block, K block)
// units are laid end to end and every cluster takes the same number of CONSECUTIVE units, i.e. at most a tail of one block,
// whole blocks, and a head of another.  A partial accumulator goes to a workspace slot; the cluster that holds a block's
// HEAD (its last segment, so the other parts were written long before) adds the parts in K order -- deterministic -- and
// stores the block.  Roles, pipeline, weight producer and tensor-memory operand path are those of grad_fused.cuh.
#pragma once
#include "grad_fused.cuh"

namespace npair {

struct StreamKParams {
  float* ws;            // [blocks][max_slots][2 CTAs][128 rows][256] raw fp32 partial accumulators
  uint32_t* flags;      // [blocks][max_slots][2][16]: epoch of the last completed partial, per CTA rank and epilogue warp
  uint32_t epoch;       // this launch's number (> 0, increasing): flags never need a reset
  int upc;              // K-block units per cluster
  int max_slots;        // non-head parts a block can have
};

// a cluster's contiguous range of (block, K block) units, cut into per-block segments
struct SkWalk {
  long long u, u1;
  int nkb;
  __device__ SkWalk(int cluster, int upc, long long total, int nkb_)
      : u(static_cast<long long>(cluster) * upc), u1(min(total, static_cast<long long>(cluster + 1) * upc)), nkb(nkb_) {}
  __device__ bool next(int& block, int& kb0, int& kb1) {
    if (u >= u1) return false;
    block = static_cast<int>(u / nkb);
    kb0 = static_cast<int>(u - static_cast<long long>(block) * nkb);
    kb1 = static_cast<int>(min(static_cast<long long>(nkb), kb0 + (u1 - u)));
    u += kb1 - kb0;
    return true;
  }
};

template <int NSPLIT, bool BF16, bool ONE_EX2 = false>
__global__ void __launch_bounds__(640, 1)
fused_grad_sk_kernel(const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapS, const FusedGradParams p,
                     const StreamKParams sk) {
  constexpr int NCTA = 2;
  using Cfg = FusedCfg<NSPLIT, NCTA>;
  static_assert(Cfg::TMEM_A0 + Cfg::STAGES * NSPLIT * Cfg::A_COLS <= 512, "A pieces do not fit behind the accumulator");
  const int cta_rank = (NCTA == 2) ? static_cast<int>(blockIdx.x & 1u) : 0;      // cluster = blocks {2c, 2c+1}
  const int worker = static_cast<int>(blockIdx.x) / NCTA, num_workers = static_cast<int>(gridDim.x) / NCTA;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* aux = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);            // [STAGES] TMA bytes landed (B, S tile, column records)
  uint64_t* bfull_bar = full_bar + STAGES;                           // [STAGES] pair mode: both CTAs' B halves landed (leader's copy is used)
  uint64_t* aready_bar = bfull_bar + STAGES;                         // [STAGES] producers wrote the A pieces (pair mode: of both CTAs, leader's copy)
  uint64_t* empty_bar = aready_bar + STAGES;                         // [STAGES] MMAs of the stage retired
  uint64_t* tfull_bar = empty_bar + STAGES;                          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total_units = static_cast<long long>(p.tiles_m) * p.tiles_n * p.num_kblocks;
  const float inv_scale = p.dev_scale ? *p.dev_scale : 1.f;
  const float alpha = p.alpha * inv_scale;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmapB); ptx::prefetch_tmap(&tmapS); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&bfull_bar[s], 1); ptx::mbar_init(&aready_bar[s], 16 * NCTA); ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 16 * NCTA); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) { ptx::tmem_alloc<512>(tmem_ptr); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  if (NCTA == 2) ptx::cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      SkWalk walk(worker, sk.upc, total_units, p.num_kblocks);
      int mn, kb0, kb1;
      while (walk.next(mn, kb0, kb1)) {
        const int m_blk = (mn / p.tiles_n) * NCTA + cta_rank, n_blk = mn % p.tiles_n;
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
      int it = 0;
      SkWalk walk(worker, sk.upc, total_units, p.num_kblocks);
      int mn, kb0, kb1;
      for (; walk.next(mn, kb0, kb1); ++it) {
        const int acc = 0;                       // one accumulator: the producers run the epilogue themselves
        const uint32_t acc_phase = it & 1;
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base;
        for (int kb = kb0; kb < kb1; ++kb) {
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
              if (NCTA == 1) ptx::mma_f16_ts(d_tmem, a_t + sa * Cfg::A_COLS + k2 * 8, bd, idesc, ((kb - kb0) | ps | k2) != 0 ? 1u : 0u);
              else ptx::mma_f16_ts_pair(d_tmem, a_t + sa * Cfg::A_COLS + k2 * 8, bd, idesc, ((kb - kb0) | ps | k2) != 0 ? 1u : 0u);
            }
          }
          if (NCTA == 1) ptx::mma_commit(&empty_bar[stage]); else ptx::mma_commit_pair(&empty_bar[stage], 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (NCTA == 1) ptx::mma_commit(&tfull_bar[acc]); else ptx::mma_commit_pair(&tfull_bar[acc], 3);
      }
    }
  } else if (warp >= 4) {
    // ===================================== weight producers + epilogue =====================================
    const int ew = (warp - 4) & 3;               // TMEM lane group / 32-row group
    const int qt = (warp - 4) >> 2;              // quarter: 8 of the 32 K columns while producing, 64 of the 256 D columns in the epilogue
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    const uint32_t lead_aready = (NCTA == 2) ? ptx::mapa_u32(ptx::smem_u32(&aready_bar[0]), 0) : 0u;
    const uint32_t lead_tempty = (NCTA == 2) ? ptx::mapa_u32(ptx::smem_u32(&tempty_bar[0]), 0) : 0u;
    SkWalk walk(worker, sk.upc, total_units, p.num_kblocks);
    int mn, kb0, kb1;
    for (; walk.next(mn, kb0, kb1); ++it) {
      const int m_blk = (mn / p.tiles_n) * NCTA + cta_rank, n_blk = mn % p.tiles_n;
      const int rl = ew * 32 + lane;             // row inside the tile
      const int row = m_blk * BM + rl;
      // row record (neutral when the row does not exist: thresholds -inf -> nothing selected)
      float r_m2 = 0.f, r_tp = -INFINITY, r_tn = -INFINITY, r_cA = 0.f, r_cT = 0.f, r_lab = 0.f;
      if (row < p.Q) {
        const float4 a = *reinterpret_cast<const float4*>(p.rowrec + 8ll * row);
        const float4 b = *reinterpret_cast<const float4*>(p.rowrec + 8ll * row + 4);
        r_m2 = a.x; r_tn = a.y; r_cT = a.z; r_lab = a.w; r_tp = b.x; r_cA = b.y;
      }
      const int self_col = row + p.self_offset;
      float r_R = 0.f;                             // ONE_EX2: 2^(m2_i)
      bool row_slow = false;
      if (ONE_EX2) {
        row_slow = !(fabsf(r_m2) <= 30.f);
        if (!row_slow) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r_R) : "f"(r_m2));
      }
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
        bool any_same = false;
        bool one_ex2 = false;
        float myC = 0.f;
        if (ONE_EX2) {
          const float4 cl = crec[2 * (lane & 7)];                    // lane l evaluates column l & 7 of this warp's eight
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(myC) : "f"(-cl.x));
          myC *= cl.z * p.inv_world;                                 // 2^(-m2_j) * cT_j / world
          // |m2| <= 30 on both sides: the row/column constants stay within 2^+-30, and an e1 that underflows (s*log2e - m2_i
          // < -126) belongs to a pair whose true e2 is below 2^-66; the 1e18 bound keeps r_R * cj finite when a row's T is tiny
          one_ex2 = !__any_sync(0xffffffffu, row_slow || !(fabsf(cl.x) <= 30.f) || !(fabsf(myC) <= 1e18f));
        }
        if (ONE_EX2 && one_ex2) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 ca = crec[2 * cc];
            const float s = sv[cc];
            float e1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(s, NPAIR_LOG2E_F, -r_m2)));
            const float key = s * p.sgn_n;
            const float cj = __shfl_sync(0xffffffffu, myC, cc);
            const float a = (key <= r_tn) ? r_cT : 0.f;
            const float b = (key <= ca.y) ? r_R * cj : 0.f;
            g[cc] = e1 * (a + b);
            any_same |= (ca.w == r_lab);
          }
        } else
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const float4 ca = crec[2 * cc];        // {m2, thr_n, cT, label}: everything a diff-label pair needs
          const float s = sv[cc];
          float e1, e2;                          // same formula as the forward row pass (fast_exp_m2)
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(s, NPAIR_LOG2E_F, -r_m2)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(s, NPAIR_LOG2E_F, -ca.x)));
          const float key = s * p.sgn_n;
          const float w1 = (key <= r_tn) ? e1 * r_cT : 0.f;
          const float w2 = (key <= ca.y) ? e2 * ca.z : 0.f;
          g[cc] = fmaf(w2, p.inv_world, w1);
          any_same |= (ca.w == r_lab);
        }
        // rare fix-ups: same-label pairs (the other selection rule and weight), the self pair, columns beyond N
        if (any_same || (self_col >= m0 && self_col < m0 + 8) || m0 + 8 > p.N) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 ca = crec[2 * cc], cb = crec[2 * cc + 1];   // cb = {thr_p, cA, -, -}
            if (ca.w == r_lab) {
              const float s = sv[cc];
              float e1, e2;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(s, NPAIR_LOG2E_F, -r_m2)));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(s, NPAIR_LOG2E_F, -ca.x)));
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
        const uint32_t a_t = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + Cfg::TMEM_A0 + stage * NSPLIT * Cfg::A_COLS + 4 * qt;
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s) ptx::tmem_st_32x32b_x4(a_t + s * Cfg::A_COLS, pk[s][0], pk[s][1], pk[s][2], pk[s][3]);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();                   // order the tcgen05.st before the arrive that releases the MMA issuer
        __syncwarp();
        if (lane == 0) { if (NCTA == 1) ptx::mbar_arrive(&aready_bar[stage]); else ptx::mbar_arrive_cluster(lead_aready + 8u * stage); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      // ---- epilogue of this segment ----
      //   whole tile          : alpha * acc -> out
      //   tail / middle part  : raw accumulator -> workspace slot, then publish this warp's flag (release)
      //   head part (kb0 == 0): the LAST segment of this cluster; waits for the later parts of the tile (written by clusters
      //                         worker+1.. as their first segments, i.e. long ago), sums head + slot 0 + slot 1 .. and stores
      const int acc = 0;
      const uint32_t acc_phase = it & 1;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      const bool whole = (kb0 == 0 && kb1 == p.num_kblocks);
      const bool head = (kb0 == 0 && !whole);
      const long long tile_u0 = static_cast<long long>(mn) * p.num_kblocks;
      const int first_cluster = static_cast<int>(tile_u0 / sk.upc);                                  // owner of the tile's first unit
      const int last_cluster = min(num_workers - 1, static_cast<int>((tile_u0 + p.num_kblocks - 1) / sk.upc));
      const int n_other = head ? last_cluster - worker : 0;
      const int my_slot = worker - first_cluster - 1;                                                 // valid when !whole && !head
      const int w16 = warp - 4;
      auto ws_ptr = [&](int slot) {
        return sk.ws + ((((static_cast<long long>(mn) * sk.max_slots + slot) * 2 + cta_rank) * BM + rl) * BN);
      };
      auto flag_ptr = [&](int slot) { return sk.flags + (((static_cast<long long>(mn) * sk.max_slots + slot) * 2 + cta_rank) * 16 + w16); };
      if (head) {
        for (int s = 0; s < n_other; ++s) {
          const uint32_t* f = flag_ptr(s);
          uint32_t v;
          do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory"); } while (v != sk.epoch);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int ch = qt * 2; ch < qt * 2 + 2; ++ch) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + ch * 32, r);
        ptx::tmem_ld_wait();
        if (!whole && !head) {
          float4* dst = reinterpret_cast<float4*>(ws_ptr(my_slot) + ch * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
          continue;
        }
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(r[c]);
        for (int s = 0; s < n_other; ++s) {                    // fixed order: head, then the parts by increasing K
          const float4* src = reinterpret_cast<const float4*>(ws_ptr(s) + ch * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t4 = __ldcg(src + q);
            v[4 * q] += t4.x; v[4 * q + 1] += t4.y; v[4 * q + 2] += t4.z; v[4 * q + 3] += t4.w;
          }
        }
        const int col0 = n_blk * BN + ch * 32;
        if (row < p.Q) {
          float* dst = p.out + static_cast<long long>(row) * p.ldo + col0;
          if (col0 + 32 <= p.D && (p.ldo & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 o = make_float4(alpha * v[4 * q], alpha * v[4 * q + 1], alpha * v[4 * q + 2], alpha * v[4 * q + 3]);
              if (p.beta != 0.f) {
                const float4 old = reinterpret_cast<float4*>(dst)[q];
                o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              }
              reinterpret_cast<float4*>(dst)[q] = o;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (col0 + c < p.D) {
                float o = alpha * v[c];
                if (p.beta != 0.f) o += p.beta * dst[c];
                dst[c] = o;
              }
          }
        }
      }
      if (!whole && !head) {
        __threadfence();                                        // this warp's partial is visible device-wide ...
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag_ptr(my_slot)), "r"(sk.epoch) : "memory");   // ... before its flag
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (NCTA == 1) ptx::mbar_arrive(&tempty_bar[acc]); else ptx::mbar_arrive_cluster(lead_tempty + 8u * acc); }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (NCTA == 2) ptx::cluster_sync_all();      // no CTA of the pair exits while the other may still signal or read it
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace npair
