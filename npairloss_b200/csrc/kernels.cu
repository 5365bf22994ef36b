// kernels.cu -- HBM-bound kernels of the N-pair hot path (everything except the two tensor-core contractions).
// Each kernel cites the reference code it replaces (paths relative to /root/reference).
#include <cstdlib>
#include "kernels.cuh"
#include <cuda.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cfloat>

#include "thresholds.cuh"   // f2ord / ord2f come with kernels.cuh

namespace npair {
unsigned long long g_kernel_launches = 0;


// --------------------------------------------------------------------------------------------
// small helpers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// split an fp32 value into 2-byte pieces (see gemm_tcgen05.cuh header)
template <int PREC>
__device__ __forceinline__ void split3(float v, uint16_t& p0, uint16_t& p1, uint16_t& p2) {
  if (PREC == PREC_BF16) {
    p0 = __bfloat16_as_ushort(__float2bfloat16_rn(v)); p1 = 0; p2 = 0;
  } else if (PREC == PREC_FP16X2) {
    const __half h = __float2half_rn(v);
    const float r = v - __half2float(h);
    p0 = __half_as_ushort(h); p1 = __half_as_ushort(__float2half_rn(r)); p2 = 0;
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    p0 = __bfloat16_as_ushort(h); p1 = __bfloat16_as_ushort(m); p2 = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
  }
}
__host__ __device__ inline int nsplit_of(int prec) { return prec == PREC_BF16 ? 1 : (prec == PREC_FP16X2 ? 2 : 3); }

// exp(x) for x <= 0 as one FMUL + MUFU.EX2 (relative error ~ (2 + 1.44|x|) ulp: 3e-7 for the |x| <= 2 of unit-norm
// embeddings, 1e-5 only beyond |x| ~ 80 where the terms are ~1e-35 anyway).  Cheap enough to evaluate for EVERY pair,
// which keeps the row pass and the weight builder branch-free (the reference's expf, .cu:131, under a selection branch
// costs ~20 instructions per divergent hit).  The same function is used forward and backward, so W = e / A stays consistent.
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
#define NPAIR_LOG2E 1.4426950408889634f
// exp(s - max) with the row constant pre-multiplied: m2 = max * log2(e)  ->  one FFMA + one MUFU.EX2
__device__ __forceinline__ float fast_exp_m2(float sv, float m2) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(sv, NPAIR_LOG2E, -m2)));
  return y;
}

// Every selection rule of .cu:79-120 is rewritten as ONE compare  sgn*s <= thr'  with a per-row transformed threshold:
//   s <  t  <=>   s <= nextbelow(t)         s <= t  <=>   s <= t
//   s >= t  <=>  -s <= -t                   s >  t  <=>  -s <= nextbelow(-t)          ALL  <=>  s <= +inf
// (exact for every float incl. +-0 and the -FLT_MAX / FLT_MAX sentinels; NaN compares false on both sides).
__host__ __device__ __forceinline__ float ap_sign(int m) { return (m == M_EASY || m == M_RELATIVE_EASY) ? -1.f : 1.f; }
__host__ __device__ __forceinline__ float an_sign(int m) { return (m == M_HARD || m == M_RELATIVE_HARD) ? -1.f : 1.f; }
__device__ __forceinline__ float ap_thr(float t, int m) {     // same-label rule on t = posi_thr + margin_ident
  switch (m) {
    case M_HARD: return nextafterf(t, -INFINITY);            // s <  t
    case M_EASY: return -t;                                  // s >= t
    case M_RAND: return INFINITY;
    case M_RELATIVE_HARD: return t;                          // s <= t
    default: return -t;                                      // RELATIVE_EASY: s >= t
  }
}
__device__ __forceinline__ float an_thr(float t, int m) {     // diff-label rule on t = nega_thr + margin_diff
  switch (m) {
    case M_HARD: return nextafterf(-t, -INFINITY);           // s >  t
    case M_EASY: return t;                                   // s <= t
    case M_RAND: return INFINITY;
    case M_RELATIVE_HARD: return -t;                         // s >= t
    default: return t;                                       // RELATIVE_EASY: s <= t
  }
}

__device__ __forceinline__ bool sel_ap(float s, float tp, int m) {   // .cu:79-98
  switch (m) {
    case M_HARD: return s < tp;
    case M_EASY: return s >= tp;
    case M_RAND: return true;
    case M_RELATIVE_HARD: return s <= tp;
    default: return s >= tp;
  }
}
__device__ __forceinline__ bool sel_an(float s, float tn, int m) {   // .cu:100-119
  switch (m) {
    case M_HARD: return s > tn;
    case M_EASY: return s <= tn;
    case M_RAND: return true;
    case M_RELATIVE_HARD: return s >= tn;
    default: return s <= tn;
  }
}

// --------------------------------------------------------------------------------------------
// absmax / asum  (caffe_gpu_asum .cu:400; operand pre-scale for PREC_FP16X2)
// --------------------------------------------------------------------------------------------
// One kernel: per-block partial |x| sums (local rows) and max|x| (all rows), the reset of the per-row statistics, and --
// in the last block to finish (ticket) -- the final asum, the power-of-two operand scale and the reset of the step state.
// Returns true in the block that finished last (it has written the step's scalars to *bs).
__device__ __forceinline__ bool prep_reduce_body(const float* __restrict__ xl, long long nl, const float* __restrict__ xt, long long ntot,
                                                 float* __restrict__ partial, int want_scale, RowArrays ra, int Q, BlockScalars* bs) {
  __shared__ float s_sum[8], s_max[8];
  __shared__ int s_last;
  float sum = 0.f, mx = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // 16-byte loads, four in flight per thread (cudaMalloc'd / framework blobs are 16-byte aligned; otherwise the scalar loop)
  const bool same_range = want_scale && xl == xt && nl == ntot;    // world == 1: one sweep gives both the sum and the maximum
  if ((reinterpret_cast<uintptr_t>(xl) & 15) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(xl);
    const long long n4 = nl >> 2;
    long long i = t0;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + stride), c = __ldg(x4 + i + 2 * stride), d = __ldg(x4 + i + 3 * stride);
      const float a0 = fabsf(a.x), a1 = fabsf(a.y), a2 = fabsf(a.z), a3 = fabsf(a.w), b0 = fabsf(b.x), b1 = fabsf(b.y), b2 = fabsf(b.z), b3 = fabsf(b.w);
      const float c0 = fabsf(c.x), c1 = fabsf(c.y), c2 = fabsf(c.z), c3 = fabsf(c.w), d0 = fabsf(d.x), d1 = fabsf(d.y), d2 = fabsf(d.z), d3 = fabsf(d.w);
      sum += (a0 + a1) + (a2 + a3) + (b0 + b1) + (b2 + b3) + (c0 + c1) + (c2 + c3) + (d0 + d1) + (d2 + d3);
      if (same_range) {
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)), fmaxf(fmaxf(b0, b1), fmaxf(b2, b3))));
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)), fmaxf(fmaxf(d0, d1), fmaxf(d2, d3))));
      }
    }
    for (; i < n4; i += stride) {
      const float4 a = __ldg(x4 + i);
      sum += (fabsf(a.x) + fabsf(a.y)) + (fabsf(a.z) + fabsf(a.w));
      if (same_range) mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
    for (long long j = (n4 << 2) + t0; j < nl; j += stride) { sum += fabsf(xl[j]); if (same_range) mx = fmaxf(mx, fabsf(xl[j])); }
  } else {
    for (long long i = t0; i < nl; i += stride) { sum += fabsf(xl[i]); if (same_range) mx = fmaxf(mx, fabsf(xl[i])); }
  }
  if (want_scale && !same_range) {
    if ((reinterpret_cast<uintptr_t>(xt) & 15) == 0) {
      const float4* x4 = reinterpret_cast<const float4*>(xt);
      const long long n4 = ntot >> 2;
      long long i = t0;
      for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + stride), c = __ldg(x4 + i + 2 * stride), d = __ldg(x4 + i + 3 * stride);
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))), fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))), fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)))));
      }
      for (; i < n4; i += stride) { const float4 a = __ldg(x4 + i); mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)))); }
      for (long long j = (n4 << 2) + t0; j < ntot; j += stride) mx = fmaxf(mx, fabsf(xt[j]));
    } else {
      for (long long i = t0; i < ntot; i += stride) mx = fmaxf(mx, fabsf(xt[i]));
    }
  }
  for (long long i = t0; i < Q; i += stride) {                  // caffe_set of the stat blobs, .cu:230-236
    ra.st_minw[i] = f2ord(FLT_MAX); ra.st_maxw[i] = f2ord(-FLT_MAX);
    ra.st_maxb[i] = f2ord(-FLT_MAX); ra.st_maxall[i] = f2ord(-FLT_MAX);
    ra.cnt_same[i] = 0;
  }
  sum = warp_sum(sum); mx = warp_max(mx);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_sum[w] = sum; s_max[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = 0.f; mx = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) { sum += s_sum[k]; mx = fmaxf(mx, s_max[k]); }
    partial[blockIdx.x] = sum; partial[1024 + blockIdx.x] = mx;
    __threadfence();
    s_last = (atomicAdd(&bs->ticket0, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  __shared__ double s_dsum[8];
  double dsum = 0.0; mx = 0.f;
  for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += blockDim.x) { dsum += __ldcg(&partial[b]); mx = fmaxf(mx, __ldcg(&partial[1024 + b])); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { dsum += __shfl_xor_sync(0xffffffffu, dsum, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if (l == 0) { s_dsum[w] = dsum; s_max[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    dsum = 0.0; mx = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) { dsum += s_dsum[k]; mx = fmaxf(mx, s_max[k]); }
    bs->asum = static_cast<float>(dsum);
    bs->x_absmax = mx;
    float sc = 1.f, inv = 1.f;
    if (want_scale && mx > 0.f && isfinite(mx)) {
      int e; frexpf(mx, &e);                 // mx = m * 2^e, m in [0.5,1)
      sc = ldexpf(1.f, -e); inv = ldexpf(1.f, e);
    }
    bs->x_scale = sc; bs->x_inv_scale = inv;
    bs->err = 0; bs->ticket = 0; bs->ticket2 = 0; bs->ticket0 = 0; bs->ticket3 = 0; bs->sel_active[0] = 0; bs->sel_active[1] = 0;
    bs->cand_n[0] = 0; bs->cand_n[1] = 0;
    bs->n_same = 0; bs->n_diff = 0;
  }
  return true;
}
__global__ void __launch_bounds__(256) prep_reduce_kernel(const float* __restrict__ xl, long long nl, const float* __restrict__ xt, long long ntot,
                                                          float* __restrict__ partial, int want_scale, RowArrays ra, int Q, BlockScalars* bs) {
  prep_reduce_body(xl, nl, xt, ntot, partial, want_scale, ra, Q, bs);
}
void launch_prep_reduce(const float* x_local, long long n_local, const float* x_total, long long n_total, float* partial,
                        int want_scale, RowArrays ra, int Q, BlockScalars* bs, cudaStream_t st) {
  long long nmax = n_local > n_total ? n_local : n_total;
  int nb = static_cast<int>((nmax + 256 * 16 - 1) / (256 * 16));
  if (nb < 1) nb = 1; if (nb > 592) nb = 592;
  prep_reduce_kernel<<<nb, 256, 0, st>>>(x_local, n_local, x_total, n_total, partial, want_scale, ra, Q, bs);
  count_launch();
}

__global__ void absmax_asum_partial_kernel(const float* __restrict__ xl, long long nl, const float* __restrict__ xt, long long ntot,
                                           float* __restrict__ partial, int want_scale) {
  __shared__ float s_sum[32], s_max[32];
  float sum = 0.f, mx = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nl; i += stride) sum += fabsf(xl[i]);
  if (want_scale)
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < ntot; i += stride) mx = fmaxf(mx, fabsf(xt[i]));
  sum = warp_sum(sum); mx = warp_max(mx);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_sum[w] = sum; s_max[w] = mx; }
  __syncthreads();
  if (w == 0) {
    sum = (l < (blockDim.x >> 5)) ? s_sum[l] : 0.f;
    mx = (l < (blockDim.x >> 5)) ? s_max[l] : 0.f;
    sum = warp_sum(sum); mx = warp_max(mx);
    if (l == 0) { partial[blockIdx.x] = sum; partial[1024 + blockIdx.x] = mx; }
  }
}
__global__ void absmax_asum_final_kernel(const float* __restrict__ partial, int nb, BlockScalars* bs, int want_scale) {
  __shared__ double s_sum[32];
  __shared__ float s_max[32];
  double sum = 0.0; float mx = 0.f;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) { sum += partial[b]; mx = fmaxf(mx, partial[1024 + b]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_sum[w] = sum; s_max[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = 0.0; mx = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) { sum += s_sum[k]; mx = fmaxf(mx, s_max[k]); }
    bs->asum = static_cast<float>(sum);
    bs->x_absmax = mx;
    float sc = 1.f, inv = 1.f;
    if (want_scale && mx > 0.f && isfinite(mx)) {
      int e; frexpf(mx, &e);                 // mx = m * 2^e, m in [0.5,1)
      sc = ldexpf(1.f, -e); inv = ldexpf(1.f, e);
    }
    bs->x_scale = sc; bs->x_inv_scale = inv;
  }
}
void launch_absmax_asum(const float* x_local, long long n_local, const float* x_total, long long n_total,
                        float* partial, BlockScalars* bs, int want_scale, cudaStream_t st) {
  long long nmax = n_local > n_total ? n_local : n_total;
  int nb = static_cast<int>((nmax + 256 * 8 - 1) / (256 * 8));
  if (nb < 1) nb = 1; if (nb > 1024) nb = 1024;
  absmax_asum_partial_kernel<<<nb, 256, 0, st>>>(x_local, n_local, x_total, n_total, partial, want_scale);
  count_launch();
  absmax_asum_final_kernel<<<1, 256, 0, st>>>(partial, nb, bs, want_scale);
  count_launch();
}

// --------------------------------------------------------------------------------------------
// operand split: x_total fp32 [N x D] -> Xs[s][N][ldXs] (K-major for the similarity GEMM) and the transposed
// XsT[s][D][ldXsT] (K-major for the gradient GEMM whose K is the sample index); XlT = local columns only.
// --------------------------------------------------------------------------------------------
// Block = 256 threads, tile = 32 rows (n) x 64 features (d).  Thread (nl = t/8, dg = t%8) converts 8 consecutive features of
// one row: two 16-byte loads, one 16-byte store per piece / section; the transposed pieces go through a shared tile so
// that they, too, are written as 16-byte row segments.
struct SplitArgs {
  const float* x; int N, D;
  uint16_t* Xs; long long ldXs; uint16_t* XsT; long long ldXsT; uint16_t* XlT; long long ldXlT; int row0, Q;
  uint16_t *XcatA, *XcatB; long long Dp;
};
// The 8 features thread t of a block converts in tile (tile_d, tile_n): two 16-byte loads (issued early by the fused kernel).
__device__ __forceinline__ void split_load(const SplitArgs& a, int tile_d, int tile_n, float (&v)[8]) {
  const float* __restrict__ x = a.x; const int N = a.N, D = a.D;
  const int t = threadIdx.x, nl = t >> 3, dg = t & 7;
  const int n = tile_n * 32 + nl, d = tile_d * 64 + 8 * dg;
  const bool rowok = n < N;
  if (rowok && d + 7 < D && (D & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {   // 16-byte loads need an aligned base
    const float4 a4 = *reinterpret_cast<const float4*>(x + static_cast<long long>(n) * D + d);
    const float4 b4 = *reinterpret_cast<const float4*>(x + static_cast<long long>(n) * D + d + 4);
    v[0] = a4.x; v[1] = a4.y; v[2] = a4.z; v[3] = a4.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (rowok && d + e < D) ? x[static_cast<long long>(n) * D + d + e] : 0.f;
  }
}
// One 32-row x 64-feature tile (tile_n, tile_d) by one block of 256 threads from the values split_load fetched.  `buf` alternates
// between consecutive tiles of a block: the transposition tile is double-buffered, so ONE barrier per tile is enough.
template <int PREC>
__device__ __forceinline__ void split_tile(const SplitArgs& a, float sc, int tile_d, int tile_n, const float (&v)[8], int buf) {
  const int N = a.N, D = a.D;
  uint16_t* __restrict__ Xs = a.Xs; const long long ldXs = a.ldXs; uint16_t* __restrict__ XsT = a.XsT; const long long ldXsT = a.ldXsT;
  uint16_t* __restrict__ XlT = a.XlT; const long long ldXlT = a.ldXlT; const int row0 = a.row0, Q = a.Q;
  uint16_t* __restrict__ XcatA = a.XcatA; uint16_t* __restrict__ XcatB = a.XcatB; const long long Dp = a.Dp;
  constexpr int NS = (PREC == PREC_BF16) ? 1 : (PREC == PREC_FP16X2 ? 2 : 3);
  __shared__ __align__(16) uint16_t tile2[2][NS][64][40];  // [buffer][piece][d][n], row padded to 80 bytes (16-byte aligned, spreads banks)
  uint16_t (*tile)[64][40] = tile2[buf];
  const int n0 = tile_n * 32, d0 = tile_d * 64;
  const int t = threadIdx.x, nl = t >> 3, dg = t & 7;
  const int n = n0 + nl, d = d0 + 8 * dg;
  uint16_t p[8][3];
  const bool rowok = n < N;
#pragma unroll
  for (int e = 0; e < 8; ++e) split3<PREC>(v[e] * sc, p[e][0], p[e][1], p[e][2]);
  uint4 pk[3];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    pk[s] = make_uint4(p[0][s] | (static_cast<uint32_t>(p[1][s]) << 16), p[2][s] | (static_cast<uint32_t>(p[3][s]) << 16),
                       p[4][s] | (static_cast<uint32_t>(p[5][s]) << 16), p[6][s] | (static_cast<uint32_t>(p[7][s]) << 16));
#pragma unroll
    for (int e = 0; e < 8; ++e) tile[s][8 * dg + e][nl] = p[e][s];
  }
  // every destination row is padded to a multiple of 64 elements (Dp), so whole 16-byte groups can be stored even when
  // D is ragged: the excess elements are zeros (v = 0 above) and lie beyond the TMA extent anyway
  if (rowok && d < Dp) {
    const long long ps = static_cast<long long>(N) * ldXs;
    if (Xs)      // NULL when the similarity GEMM reads the K-concatenated operands below
#pragma unroll
      for (int s = 0; s < NS; ++s) *reinterpret_cast<uint4*>(Xs + s * ps + static_cast<long long>(n) * ldXs + d) = pk[s];
    // K-concatenated operands of the bitwise-symmetric similarity GEMM (one MMA pass over K_cat):
    //   fp16x2 : A row = [ hi | hi(8) lo(8) ... ]                         B row = [ hi | lo(8) hi(8) ... ]                  K_cat = 3*Dp
    //   bf16x3 : A row = [ hi | mid | hi(8) mid(8) ... | hi(8) lo(8) ... ]   B row = [ hi | mid | mid(8) hi(8) ... | lo(8) hi(8) ... ]   K_cat = 6*Dp
    // ONE K=16 MMA then sums 8 products p_j*q_m and the 8 mirrored products q_j*p_m: swapping the operand roles only
    // permutes the products inside an instruction, whose sum is order-invariant (measured: tests/diag_mma_symmetry.py),
    // so S[j][m] == S[m][j] bit for bit, on one rank and across ranks.
    if (PREC != PREC_BF16 && XcatA) {
      const long long kcat = (PREC == PREC_FP16X2 ? 3 : 6) * Dp;
      uint16_t* ra = XcatA + static_cast<long long>(n) * kcat;
      uint16_t* rb = XcatB + static_cast<long long>(n) * kcat;
      const bool local = (n >= row0 && n < row0 + Q);        // only the rank's own rows are ever an A operand
      *reinterpret_cast<uint4*>(rb + d) = pk[0];
      if (local) *reinterpret_cast<uint4*>(ra + d) = pk[0];
      if (PREC == PREC_FP16X2) {
        *reinterpret_cast<uint4*>(rb + Dp + 2 * d) = pk[1]; *reinterpret_cast<uint4*>(rb + Dp + 2 * d + 8) = pk[0];
        if (local) { *reinterpret_cast<uint4*>(ra + Dp + 2 * d) = pk[0]; *reinterpret_cast<uint4*>(ra + Dp + 2 * d + 8) = pk[1]; }
      } else {
        *reinterpret_cast<uint4*>(rb + Dp + d) = pk[1];
        *reinterpret_cast<uint4*>(rb + 2 * Dp + 2 * d) = pk[1]; *reinterpret_cast<uint4*>(rb + 2 * Dp + 2 * d + 8) = pk[0];
        *reinterpret_cast<uint4*>(rb + 4 * Dp + 2 * d) = pk[2]; *reinterpret_cast<uint4*>(rb + 4 * Dp + 2 * d + 8) = pk[0];
        if (local) {
          *reinterpret_cast<uint4*>(ra + Dp + d) = pk[1];
          *reinterpret_cast<uint4*>(ra + 2 * Dp + 2 * d) = pk[0]; *reinterpret_cast<uint4*>(ra + 2 * Dp + 2 * d + 8) = pk[1];
          *reinterpret_cast<uint4*>(ra + 4 * Dp + 2 * d) = pk[0]; *reinterpret_cast<uint4*>(ra + 4 * Dp + 2 * d + 8) = pk[2];
        }
      }
    }
  }
  __syncthreads();
  // transposed pieces: thread (dl = t/4, nc = t%4) stores 8 consecutive rows n of feature d0 + dl
  const int dl = t >> 2, nc = t & 3;
  const int dd = d0 + dl, nn = n0 + 8 * nc;
  if (dd < D && nn < N) {
    const long long pt = static_cast<long long>(D) * ldXsT;
    const long long pl = static_cast<long long>(D) * ldXlT;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const uint4 q = *reinterpret_cast<const uint4*>(&tile[s][dl][8 * nc]);
      *reinterpret_cast<uint4*>(XsT + s * pt + static_cast<long long>(dd) * ldXsT + nn) = q;     // ldXsT, nn multiples of 8
      if (XlT && nn >= row0 && nn < row0 + Q) {
        if (((nn - row0) & 7) == 0 && nn + 8 <= row0 + Q) *reinterpret_cast<uint4*>(XlT + s * pl + static_cast<long long>(dd) * ldXlT + (nn - row0)) = q;
        else
          for (int e = 0; e < 8; ++e)
            if (nn + e < row0 + Q && nn + e < N) XlT[s * pl + static_cast<long long>(dd) * ldXlT + (nn + e - row0)] = tile[s][dl][8 * nc + e];
      } else if (XlT && nn < row0 && nn + 8 > row0) {
        for (int e = 0; e < 8; ++e)
          if (nn + e >= row0 && nn + e < row0 + Q && nn + e < N) XlT[s * pl + static_cast<long long>(dd) * ldXlT + (nn + e - row0)] = tile[s][dl][8 * nc + e];
      }
    }
  }
}
template <int PREC>
__global__ void __launch_bounds__(256) split_kernel(SplitArgs a, const BlockScalars* __restrict__ bs) {
  float v[8];
  split_load(a, blockIdx.x, blockIdx.y, v);
  split_tile<PREC>(a, (PREC == PREC_FP16X2) ? bs->x_scale : 1.f, blockIdx.x, blockIdx.y, v, 0);
}
// Measured and dropped (round 2): reduce + grid-wide release/acquire + split in ONE cooperative launch (prep_fused_kernel) took
// 31.4 us against 29.4 us for the two launches below -- the blocks spend ~6 us spinning for the last block's scalars, more than the
// launch gap they save (ncu: CCTL.IVALL 38 k times, barrier stall 12 per issue).
void launch_split(const float* x_total, int N, int D, int prec, const BlockScalars* bs, uint16_t* Xs, long long ldXs,
                  uint16_t* XsT, long long ldXsT, uint16_t* XlT, long long ldXlT, int row0_local, int Q,
                  uint16_t* XcatA, uint16_t* XcatB, long long Dp, cudaStream_t st) {
  dim3 grid((D + 63) / 64, (N + 31) / 32);
  const SplitArgs a{x_total, N, D, Xs, ldXs, XsT, ldXsT, XlT, ldXlT, row0_local, Q, XcatA, XcatB, Dp};
  if (prec == PREC_BF16) split_kernel<PREC_BF16><<<grid, 256, 0, st>>>(a, bs);
  else if (prec == PREC_FP16X2) split_kernel<PREC_FP16X2><<<grid, 256, 0, st>>>(a, bs);
  else split_kernel<PREC_BF16X3><<<grid, 256, 0, st>>>(a, bs);
  count_launch();
}

// --------------------------------------------------------------------------------------------
// statistics init / reference row statistics (caffe_set of the three stat blobs, .cu:230-236)
// --------------------------------------------------------------------------------------------
// One block per row; same outputs as the sim-GEMM epilogue.  Used by the SIMT cross-check backend and by tests.
__global__ void row_stats_ref_kernel(const float* __restrict__ S, long long ldS, int Q, int N, const float* __restrict__ lab_rows,
                                     const float* __restrict__ lab_cols, int self_offset, RowArrays ra) {
  const int i = blockIdx.x;
  const float li = lab_rows[i];
  const int self_col = i + self_offset;
  float minw = FLT_MAX, maxw = -FLT_MAX, maxb = -FLT_MAX, maxall = -FLT_MAX;
  int cnt = 0;
  const float* row = S + static_cast<long long>(i) * ldS;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    if (j == self_col) continue;
    const float v = row[j];
    maxall = fmaxf(maxall, v);
    if (lab_cols[j] == li) { minw = fminf(minw, v); maxw = fmaxf(maxw, v); ++cnt; } else maxb = fmaxf(maxb, v);
  }
  minw = warp_min(minw); maxw = warp_max(maxw); maxb = warp_max(maxb); maxall = warp_max(maxall); cnt = warp_sum_i(cnt);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&ra.st_minw[i], f2ord(minw)); atomicMax(&ra.st_maxw[i], f2ord(maxw));
    atomicMax(&ra.st_maxb[i], f2ord(maxb)); atomicMax(&ra.st_maxall[i], f2ord(maxall));
    if (cnt) atomicAdd(&ra.cnt_same[i], cnt);
  }
}
void launch_row_stats_ref(const float* S, long long ldS, int Q, int N, const float* lab_rows, const float* lab_cols,
                          int self_offset, RowArrays ra, cudaStream_t st) {
  row_stats_ref_kernel<<<Q, 256, 0, st>>>(S, ldS, Q, N, lab_rows, lab_cols, self_offset, ra);
  count_launch();
}

// --------------------------------------------------------------------------------------------
// thresholds (.cu:275-337).  One block.  Non-relative modes and the pos==size-1 relative shortcut are
// closed forms of the row statistics; general relative modes arm the radix selects below.
// --------------------------------------------------------------------------------------------
// Multi-block: every block reduces its slice of the row statistics and writes the LOCAL-region thresholds of its rows
// (they need no global value); the last block to finish (ticket) combines the per-block partials into the block-wide
// sizes / extrema, the GLOBAL-region thresholds (read directly by the row pass) and the radix-select arming.
struct ThrPartial { unsigned long long ns; float mn, mxw, mxb; int err; };
__global__ void __launch_bounds__(256) thresholds_kernel(RowArrays ra, int Q, int N, MiningParams mp, BlockScalars* bs, ThrPartial* part,
                                                         float* __restrict__ xout /*world scope: 6 floats of block statistics, else NULL*/) {
  __shared__ unsigned long long s_ns[8];
  __shared__ float s_mn[8], s_mxw[8], s_mxb[8];
  __shared__ int s_err, s_last;
  if (threadIdx.x == 0) s_err = 0;
  __syncthreads();
  unsigned long long ns = 0; float mn = FLT_MAX, mxw = -FLT_MAX, mxb = -FLT_MAX;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += gridDim.x * blockDim.x) {
    const int cs = ra.cnt_same[i];
    const float r_mn = ord2f(ra.st_minw[i]), r_mxw = ord2f(ra.st_maxw[i]), r_mxb = ord2f(ra.st_maxb[i]);
    ns += static_cast<unsigned long long>(cs);
    mn = fminf(mn, r_mn); mxw = fmaxf(mxw, r_mxw); mxb = fmaxf(mxb, r_mxb);
    if (mp.ap_region == REGION_LOCAL) {
      if (!is_rel(mp.ap_method)) ra.posi_thr[i] = r_mxb;                                                   // .cu:279
      else if (sn_is_max(mp.identsn)) { if (cs == 0) atomicOr(&s_err, DERR_EMPTY_LIST); ra.posi_thr[i] = clamp_thr(r_mxw); }
    }
    if (mp.an_region == REGION_LOCAL) {
      if (!is_rel(mp.an_method)) ra.nega_thr[i] = r_mn;                                                    // .cu:310
      else if (sn_is_max(mp.diffsn)) { if (N - 1 - cs == 0) atomicOr(&s_err, DERR_EMPTY_LIST); ra.nega_thr[i] = clamp_thr(r_mxb); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
  mn = warp_min(mn); mxw = warp_max(mxw); mxb = warp_max(mxb);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_ns[w] = ns; s_mn[w] = mn; s_mxw[w] = mxw; s_mxb[w] = mxb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ThrPartial t; t.ns = 0; t.mn = FLT_MAX; t.mxw = -FLT_MAX; t.mxb = -FLT_MAX; t.err = s_err;
    for (int k = 0; k < (blockDim.x >> 5); ++k) { t.ns += s_ns[k]; t.mn = fminf(t.mn, s_mn[k]); t.mxw = fmaxf(t.mxw, s_mxw[k]); t.mxb = fmaxf(t.mxb, s_mxb[k]); }
    part[blockIdx.x] = t;
    __threadfence();
    s_last = (atomicAdd(&bs->ticket2, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  unsigned long long n_same = 0; float gmin_w = FLT_MAX, gmax_w = -FLT_MAX, gmax_b = -FLT_MAX; int err = 0;
  for (int g = 0; g < static_cast<int>(gridDim.x); ++g) {
    const ThrPartial t = part[g];
    n_same += t.ns; gmin_w = fminf(gmin_w, t.mn); gmax_w = fmaxf(gmax_w, t.mxw); gmax_b = fmaxf(gmax_b, t.mxb); err |= t.err;
  }
  bs->ticket2 = 0;
  if (xout) {          // world scope: this rank's block statistics go to the exchange; thresholds_world_kernel finishes
    xout[0] = __uint_as_float(static_cast<uint32_t>(n_same)); xout[1] = __uint_as_float(static_cast<uint32_t>(n_same >> 32));
    xout[2] = gmin_w; xout[3] = gmax_w; xout[4] = gmax_b; xout[5] = __int_as_float(err);
    return;
  }
  finish_thresholds(n_same, static_cast<unsigned long long>(Q) * static_cast<unsigned long long>(N - 1) - n_same, gmin_w, gmax_w, gmax_b, err, mp, bs);
}
// world scope (npair_config.global_scope): every rank reduces the world's block statistics in the same order -> identical thresholds
__global__ void thresholds_world_kernel(const float* __restrict__ xall /*[world][xstride]*/, int xstride, int world, long long N, MiningParams mp,
                                        BlockScalars* bs) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long n_same = 0; float gmin_w = FLT_MAX, gmax_w = -FLT_MAX, gmax_b = -FLT_MAX; int err = 0;
  for (int r = 0; r < world; ++r) {
    const float* x = xall + static_cast<long long>(r) * xstride;
    n_same += static_cast<unsigned long long>(__float_as_uint(x[0])) | (static_cast<unsigned long long>(__float_as_uint(x[1])) << 32);
    gmin_w = fminf(gmin_w, x[2]); gmax_w = fmaxf(gmax_w, x[3]); gmax_b = fmaxf(gmax_b, x[4]); err |= __float_as_int(x[5]);
  }
  finish_thresholds(n_same, static_cast<unsigned long long>(N) * static_cast<unsigned long long>(N - 1) - n_same, gmin_w, gmax_w, gmax_b, err, mp, bs);
}
void launch_thresholds_world(const float* xall, int xstride, int world, long long N, MiningParams mp, BlockScalars* bs, cudaStream_t st) {
  thresholds_world_kernel<<<1, 32, 0, st>>>(xall, xstride, world, N, mp, bs);
  count_launch();
}
void launch_thresholds(RowArrays ra, int Q, int N, MiningParams mp, BlockScalars* bs, float* scratch, float* xout, cudaStream_t st) {
  int grid = (Q + 255) / 256; if (grid > 64) grid = 64; if (grid < 1) grid = 1;
  thresholds_kernel<<<grid, 256, 0, st>>>(ra, Q, N, mp, bs, reinterpret_cast<ThrPartial*>(scratch), xout);
  count_launch();
}

// --------------------------------------------------------------------------------------------
// Relative thresholds = order statistics of the masked similarities (replaces the unconditional std::sorts of .cu:266-273
// and the list indexing of .cu:282-290, :300-304, :313-321, :331-335).  MSB-first radix select on the order-preserving
// uint32 keys, digits of 11 / 11 / 10 bits.  Similarities of one row are clustered (a few binades), so the first digit
// already narrows the k-th element down to a few percent of the row: those CANDIDATES are compacted (21-bit remainders) and the
// last two digits are decided on the compact list -- S is read once from HBM (LOCAL: one more time from L1/L2; GLOBAL: twice).
// Both sides (same-label list for AP, diff-label list for AN) are handled in the same sweep when both are relative.
// The self pair is counted by the vectorised sweep and taken out again by one thread (it is always a same-label entry).
// --------------------------------------------------------------------------------------------
#define NPAIR_SEL_BINS 2048

// Histogram increment as ONE shared-memory reduction per lane.  A plain atomicAdd(&hist[d], 1) is rewritten by the compiler into a
// loop over the warp's distinct addresses (leader election + ATOMS.POPC.INC per address): ~20 instructions per distinct bin, the
// bulk of the select kernels' instruction count in the first round-2 version.  The hardware resolves same-address conflicts itself.
__device__ __forceinline__ void smem_inc(unsigned int* p) {
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))) : "memory");
}
__device__ __forceinline__ void smem_inc_off(unsigned int* base, uint32_t byte_off) {
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(base)) + byte_off) : "memory");
}
__device__ __forceinline__ void smem_dec(unsigned int* p) {
  asm volatile("red.shared.add.u32 [%0], -1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))) : "memory");
}

// Sweep of one row: calls f(key, j, side) for every column j < N with side = 0 (same label as the row) or 1 (different);
// the self pair is NOT excluded here.  16-byte loads of S (row stride is a multiple of 32 floats) and of the labels.
template <class F>
__device__ __forceinline__ void sweep_row(const float* __restrict__ row, int N, const float* __restrict__ lab_cols, float li, bool lab_aligned,
                                          bool want_same, bool want_diff, F f) {
  for (int j4 = threadIdx.x * 4; j4 < N; j4 += blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + j4));
    float ll[4];
    if (lab_aligned && j4 + 3 < N) {
      const float4 l = __ldg(reinterpret_cast<const float4*>(lab_cols + j4));
      ll[0] = l.x; ll[1] = l.y; ll[2] = l.z; ll[3] = l.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) ll[c] = (j4 + c < N) ? __ldg(lab_cols + j4 + c) : li;
    }
    const float vv[4] = {v.x, v.y, v.z, v.w};
    if (j4 + 3 < N && ll[0] != li && ll[1] != li && ll[2] != li && ll[3] != li) {
      if (want_diff) {
#pragma unroll
        for (int c = 0; c < 4; ++c) f(f2ord(vv[c]), j4 + c, 1);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (j4 + c >= N) continue;
        const int side = (ll[c] == li) ? 0 : 1;
        if (side == 0 ? want_same : want_diff) f(f2ord(vv[c]), j4 + c, side);
      }
    }
  }
}

// Block-parallel search of the bin that holds 0-based rank r in hist[0..nbins): returns the bin (or nbins when r is out of range),
// the rank inside it and its population.  nbins <= 2048, blockDim.x threads (a multiple of 32, <= 1024).  All threads get the result.
template <class CT>
__device__ __forceinline__ int find_bin(const CT* hist, int nbins, unsigned long long r, unsigned long long* r_in, unsigned long long* pop,
                                        unsigned long long* s_scan /*[33]*/, int* s_res /*[1]*/, unsigned long long* s_out /*[2]*/) {
  const int per = (nbins + blockDim.x - 1) / blockDim.x;
  const int b0 = threadIdx.x * per;
  unsigned long long mine = 0;
  for (int b = b0; b < b0 + per && b < nbins; ++b) mine += hist[b];
  unsigned long long incl = mine;                                 // inclusive scan over the block
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_scan[w] = incl;
  if (threadIdx.x == 0) *s_res = nbins;
  __syncthreads();
  if (w == 0) {
    unsigned long long x = (lane < static_cast<int>(blockDim.x >> 5)) ? s_scan[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
    s_scan[lane] = x;                                             // inclusive warp totals
  }
  __syncthreads();
  const unsigned long long before = (w ? s_scan[w - 1] : 0ull) + incl - mine;
  if (mine && r >= before && r < before + mine) {                 // exactly one thread
    unsigned long long cum = before;
    int b = b0;
    for (; b < b0 + per && b < nbins; ++b) { const unsigned long long h = hist[b]; if (cum + h > r) break; cum += h; }
    *s_res = b; s_out[0] = r - cum; s_out[1] = hist[b];
  }
  __syncthreads();
  *r_in = s_out[0]; *pop = s_out[1];
  return *s_res;
}

// ---- LOCAL: ONE WARP per row, warp-private histogram -- no block barriers, no block-wide scans ----
// sweep 1  digit 1 (top 10 bits of the RAW float bits: 3 instructions per element, no label branch) of every column into the warp's
//          histogram; the few same-label entries (and the self pair) are kept in a small list on the side
// pick     the excluded keys (same-label entries, self pair) are taken out of the histogram again; bins are walked in value order
//          (negative floats: descending raw digit) to find the bin of the wanted rank
// sweep 2  (L1 / L2) elements of that bin -> per-LANE private candidate lists in shared memory (two predicated instructions per
//          match: no ballots, no atomics); 22-bit remainders
// tail     three more digits (8 + 7 + 7 bits) over the candidate lists, excluded keys subtracted per digit
// Anything that does not fit the fast path (more than 128 same-label entries, a lane with more than 48 candidates) is redone
// by slow_select_row: plain sweeps of the row, one digit per sweep, label test per element.
#define NPAIR_LSEL_WARPS 8
#define NPAIR_LSEL_D1 1024                 // bins of the first digit
#define NPAIR_LSEL_LCAP 48                 // candidates per lane
#define NPAIR_LSEL_SCAP 128                // same-label entries kept per row
#define NPAIR_LSEL_U 4                     // 16-byte loads in flight per lane and array
struct LselWarp {
  unsigned int hist[NPAIR_LSEL_D1];
  uint32_t cand[32 * NPAIR_LSEL_LCAP];     // [slot][lane]: lane-private lists, bank = lane
  uint32_t same[NPAIR_LSEL_SCAP];          // raw bits of the same-label entries (self pair excluded)
  unsigned int n_same, pad_[3];            // keeps sizeof a multiple of 16 (16-byte stores into hist)
};
static_assert(sizeof(LselWarp) % 16 == 0, "LselWarp must keep 16-byte alignment in an array");
// value order <-> raw 10-bit digit (sign, 8 exponent bits, 1 mantissa bit): order o in [0,512) are the negative floats, descending raw
__device__ __forceinline__ uint32_t d1_raw_of_order(uint32_t o) { return o < 512u ? 1023u - o : o - 512u; }
__device__ __forceinline__ uint32_t d1_order_of_raw(uint32_t r) { return r >= 512u ? 1023u - r : r + 512u; }

// rank r (0-based) within bins[0..nb) taken in index order; nb a multiple of 32.  Returns the bin, *r_in, *pop (warp-uniform); nb if out of range.
__device__ __forceinline__ int warp_find_bin(const unsigned int* bins, int nb, unsigned int r, unsigned int* r_in, unsigned int* pop, int lane) {
  const int per = nb >> 5;
  unsigned int mine = 0;
  for (int b = 0; b < per; ++b) mine += bins[lane * per + b];
  unsigned int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const unsigned int before = incl - mine;
  const unsigned int hit = __ballot_sync(0xffffffffu, mine && r >= before && r < before + mine);
  if (!hit) return nb;
  const int src = __ffs(hit) - 1;
  int bin = 0; unsigned int ri = 0, pp = 0;
  if (lane == src) {
    unsigned int cum = before;
    int b = 0;
    for (; b < per; ++b) { const unsigned int h = bins[lane * per + b]; if (cum + h > r) { pp = h; break; } cum += h; }
    bin = lane * per + b; ri = r - cum;
  }
  *r_in = __shfl_sync(0xffffffffu, ri, src); *pop = __shfl_sync(0xffffffffu, pp, src);
  return __shfl_sync(0xffffffffu, bin, src);
}

// Generic (slow) select of one side of one row by a warp: 32-bit ordered keys, digits of 10/10/10/2 bits, one sweep of the row per digit.
__device__ __noinline__ uint32_t slow_select_row(const float* __restrict__ row, int N, const float* __restrict__ lab_cols, float li, int self_col,
                                                 int side, unsigned int rank, unsigned int* hist /*[1024]*/, int lane) {
  uint32_t prefix = 0, mask = 0;
  int shift = 22;
  for (int pass = 0; pass < 4; ++pass) {
    const int bits = pass < 3 ? 10 : 2;
    if (pass == 3) shift = 0;
    const int nb = 1 << bits;
    for (int b = lane; b < 1024; b += 32) hist[b] = 0;
    __syncwarp();
    for (int j = lane; j < N; j += 32) {
      if (j == self_col) continue;
      if ((lab_cols[j] == li) != (side == 0)) continue;
      const uint32_t key = f2ord(row[j]);
      if ((key & mask) == prefix) smem_inc(&hist[(key >> shift) & (nb - 1)]);
    }
    __syncwarp();
    unsigned int r2, pp;
    const int d = warp_find_bin(hist, nb < 32 ? 32 : nb, rank, &r2, &pp, lane);
    prefix |= static_cast<uint32_t>(d) << shift; mask |= static_cast<uint32_t>(nb - 1) << shift; rank = r2;
    shift -= 10;
    __syncwarp();
  }
  return prefix;
}

__global__ void __launch_bounds__(32 * NPAIR_LSEL_WARPS, 2) local_select_kernel(const float* __restrict__ S, long long ldS, int Q, int N,
                                                                              const float* __restrict__ lab_rows, const float* __restrict__ lab_cols,
                                                                              int self_offset, int side_mask /*1 AP, 2 AN*/, float sn_ap, float sn_an,
                                                                              RowArrays ra, BlockScalars* bs) {
  extern __shared__ __align__(16) unsigned char lsel_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  LselWarp& W = reinterpret_cast<LselWarp*>(lsel_smem)[w];
  const bool want_same = side_mask & 1, want_diff = side_mask & 2;
  const bool lab_aligned = (reinterpret_cast<uintptr_t>(lab_cols) & 15) == 0;
  const int nwarps = gridDim.x * NPAIR_LSEL_WARPS;
  for (int i = blockIdx.x * NPAIR_LSEL_WARPS + w; i < Q; i += nwarps) {
    const float li = lab_rows[i];
    const int self_col = i + self_offset;
    const float* row = S + static_cast<long long>(i) * ldS;
    const int cs = ra.cnt_same[i];
    // ---------------- sweep 1 ----------------
    for (int b = lane * 4; b < NPAIR_LSEL_D1; b += 128) *reinterpret_cast<uint4*>(&W.hist[b]) = make_uint4(0u, 0u, 0u, 0u);
    if (lane == 0) W.n_same = 0;
    __syncwarp();
    const int n_vec = lab_aligned ? (N & ~127) : 0;               // whole 128-column groups with aligned labels: 16-byte loads
    for (int j4 = lane * 4; j4 < n_vec; j4 += 128 * NPAIR_LSEL_U) {   // NPAIR_LSEL_U groups (16-byte loads of S and of the labels) in flight per lane
      uint4 v[NPAIR_LSEL_U]; float4 l[NPAIR_LSEL_U];
#pragma unroll
      for (int u = 0; u < NPAIR_LSEL_U; ++u) {
        const int jj = j4 + 128 * u;
        if (jj < n_vec) { v[u] = __ldg(reinterpret_cast<const uint4*>(row + jj)); l[u] = __ldg(reinterpret_cast<const float4*>(lab_cols + jj)); }
      }
#pragma unroll
      for (int u = 0; u < NPAIR_LSEL_U; ++u) {
        const int jj = j4 + 128 * u;
        if (jj >= n_vec) continue;
        const uint32_t vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const float ll[4] = {l[u].x, l[u].y, l[u].z, l[u].w};
        if (want_diff) {
#pragma unroll
          for (int c = 0; c < 4; ++c) smem_inc_off(W.hist, (vv[c] >> 20) & 0xFFCu);
        }
        if (ll[0] == li || ll[1] == li || ll[2] == li || ll[3] == li) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (ll[c] == li && jj + c != self_col) { const unsigned int k = atomicAdd(&W.n_same, 1u); if (k < NPAIR_LSEL_SCAP) W.same[k] = vv[c]; }
        }
      }
    }
    for (int j = n_vec + lane; j < N; j += 32) {                  // ragged tail / unaligned labels
      const uint32_t b = __float_as_uint(row[j]);
      if (want_diff) smem_inc(&W.hist[b >> 22]);
      if (lab_cols[j] == li && j != self_col) { const unsigned int k = atomicAdd(&W.n_same, 1u); if (k < NPAIR_LSEL_SCAP) W.same[k] = b; }
    }
    __syncwarp();
    const unsigned int ns = W.n_same;                             // == cs
    const uint32_t self_bits = __float_as_uint(row[self_col]);
    // ---------------- AP side: the same-label list is short ----------------
    if (want_same) {
      unsigned long long pos = 0;
      float thr = 0.f;
      if (cs == 0) { if (lane == 0) atomicOr(&bs->err, DERR_EMPTY_LIST); }
      else if (!pos_index(sn_ap, static_cast<unsigned long long>(cs), pos)) { if (lane == 0) atomicOr(&bs->err, DERR_POS_RANGE); }
      else if (ns <= 32) {                                        // rank by counting inside the warp
        const uint32_t key = lane < static_cast<int>(ns) ? f2ord(__uint_as_float(W.same[lane])) : 0xFFFFFFFFu;
        unsigned int rk = 0;
        for (unsigned int t = 0; t < ns; ++t) { const uint32_t kt = __shfl_sync(0xffffffffu, key, t); rk += (kt < key || (kt == key && static_cast<int>(t) < lane)) ? 1u : 0u; }
        const unsigned int hit = __ballot_sync(0xffffffffu, lane < static_cast<int>(ns) && rk == static_cast<unsigned int>(pos));
        thr = clamp_thr(ord2f(__shfl_sync(0xffffffffu, key, __ffs(hit) - 1)));
      } else {
        thr = clamp_thr(ord2f(slow_select_row(row, N, lab_cols, li, self_col, 0, static_cast<unsigned int>(pos), W.hist + 0, lane)));
        // the slow path used the histogram: rebuild digit 1 for the diff side below by falling into its slow path as well
        if (want_diff && lane == 0) W.n_same = NPAIR_LSEL_SCAP + 1;
      }
      if (lane == 0) ra.posi_thr[i] = thr;                        // .cu:288
      __syncwarp();
    }
    // ---------------- AN side ----------------
    if (want_diff) {
      const unsigned long long size = static_cast<unsigned long long>(N - 1 - cs);
      unsigned long long pos = 0;
      float thr = 0.f;
      if (size == 0) { if (lane == 0) atomicOr(&bs->err, DERR_EMPTY_LIST); }
      else if (!pos_index(sn_an, size, pos)) { if (lane == 0) atomicOr(&bs->err, DERR_POS_RANGE); }
      else if (W.n_same > NPAIR_LSEL_SCAP) {
        thr = clamp_thr(ord2f(slow_select_row(row, N, lab_cols, li, self_col, 1, static_cast<unsigned int>(pos), W.hist, lane)));
      } else {
        // excluded keys (same-label entries + the self pair) leave the histogram; then walk the bins in value order
        for (unsigned int e = lane; e <= ns; e += 32) smem_dec(&W.hist[(e < ns ? W.same[e] : self_bits) >> 22]);
        __syncwarp();
        // permute into value order in place is not needed: lanes own 32 consecutive ORDER positions and read the raw bins they map to
        unsigned int mine = 0;
        for (int b = 0; b < 32; ++b) mine += W.hist[d1_raw_of_order(lane * 32 + b)];
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const unsigned int before = incl - mine, r0 = static_cast<unsigned int>(pos);
        const unsigned int hit = __ballot_sync(0xffffffffu, mine && r0 >= before && r0 < before + mine);
        const int src = __ffs(hit) - 1;                           // exists: pos < size = sum of the bins
        // the winning lane's 32 bins, one per lane: a second warp scan instead of a serial walk
        const unsigned int base = __shfl_sync(0xffffffffu, before, src);
        const uint32_t my_raw = d1_raw_of_order(src * 32 + lane);
        const unsigned int h1 = W.hist[my_raw];
        unsigned int inc2 = h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += t; }
        const unsigned int bef2 = base + inc2 - h1;
        const int src2 = __ffs(__ballot_sync(0xffffffffu, h1 && r0 >= bef2 && r0 < bef2 + h1)) - 1;
        const uint32_t raw = __shfl_sync(0xffffffffu, my_raw, src2);
        unsigned int rank = r0 - __shfl_sync(0xffffffffu, bef2, src2);
        const bool negative = raw >= 512u;                        // remainders of negative floats sort descending
        // ---------------- sweep 2: that bin's elements -> lane-private candidate lists ----------------
        unsigned int cnt = 0;
        for (int j4 = lane * 4; j4 < n_vec; j4 += 128 * NPAIR_LSEL_U) {
          uint4 v[NPAIR_LSEL_U];
#pragma unroll
          for (int u = 0; u < NPAIR_LSEL_U; ++u) if (j4 + 128 * u < n_vec) v[u] = __ldg(reinterpret_cast<const uint4*>(row + j4 + 128 * u));
#pragma unroll
          for (int u = 0; u < NPAIR_LSEL_U; ++u) {
            if (j4 + 128 * u >= n_vec) continue;
            const uint32_t vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if ((vv[c] >> 22) == raw) { if (cnt < NPAIR_LSEL_LCAP) W.cand[cnt * 32 + lane] = vv[c] & 0x3FFFFFu; ++cnt; }
          }
        }
        for (int j = n_vec + lane; j < N; j += 32) {
          const uint32_t b = __float_as_uint(row[j]);
          if ((b >> 22) == raw) { if (cnt < NPAIR_LSEL_LCAP) W.cand[cnt * 32 + lane] = b & 0x3FFFFFu; ++cnt; }
        }
        if (__any_sync(0xffffffffu, cnt > NPAIR_LSEL_LCAP)) {
          thr = clamp_thr(ord2f(slow_select_row(row, N, lab_cols, li, self_col, 1, static_cast<unsigned int>(pos), W.hist, lane)));
        } else {
          // ---------------- tail: 8 + 7 + 7 bits over the candidates; excluded keys of this bin are subtracted per digit ----------------
          // in remainder space the order is ascending for positive floats and descending for negative ones: flip the remainders of negatives
          const uint32_t flip = negative ? 0x3FFFFFu : 0u;
          uint32_t pre = 0, msk = 0;
          const int shifts[3] = {14, 7, 0}, nbits[3] = {8, 7, 7};
#pragma unroll
          for (int ps = 0; ps < 3; ++ps) {
            const int nb = 1 << nbits[ps];
            for (int b = lane * 4; b < nb; b += 128) *reinterpret_cast<uint4*>(&W.hist[b]) = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            for (unsigned int e = 0; e < cnt; ++e) {
              const uint32_t k = W.cand[e * 32 + lane] ^ flip;
              if ((k & msk) == pre) smem_inc(&W.hist[(k >> shifts[ps]) & (nb - 1)]);
            }
            __syncwarp();
            for (unsigned int e = lane; e <= ns; e += 32) {
              const uint32_t b = e < ns ? W.same[e] : self_bits;
              const uint32_t k = (b & 0x3FFFFFu) ^ flip;
              if ((b >> 22) == raw && (k & msk) == pre) smem_dec(&W.hist[(k >> shifts[ps]) & (nb - 1)]);
            }
            __syncwarp();
            unsigned int r2, p2;
            const int d = warp_find_bin(W.hist, nb, rank, &r2, &p2, lane);
            pre |= static_cast<uint32_t>(d & (nb - 1)) << shifts[ps]; msk |= static_cast<uint32_t>(nb - 1) << shifts[ps]; rank = r2;
            __syncwarp();
          }
          thr = clamp_thr(__uint_as_float((raw << 22) | (pre ^ flip)));
        }
      }
      if (lane == 0) ra.nega_thr[i] = thr;                        // .cu:319
      __syncwarp();
    }
  }
}

// ---- LOCAL, rows of up to 8192 columns: ONE BLOCK per row, the row stays in registers ----
// The warp-per-row kernel above is bound by its instruction count and by its few warps in flight (ncu: 56 M warp instructions, 29 %
// issue slots with 16 warps per SM).  This kernel holds a row in the registers of 256 threads (8 independent 16-byte loads each, S is
// read ONCE) and bins by VALUE with three instructions per entry:
//   pass 0   label test: entries with the row's label (the self pair among them) go to the short same-label list and are replaced by
//            NaN in the registers -- fminf / fmaxf skip NaN, so the value range [lo, hi] of the entries that stay needs no branch
//   pass 1   bin*4 = mantissa of fmaf(s, 4*2048/(hi-lo), 2^23 + 4 - lo*that): one FFMA, one AND, one shared-memory reduction.  The map
//            is monotone in s, so the wanted rank lies in the bin where the running count crosses it; bins hold a few dozen entries and
//            lanes rarely collide.  NaN lands in bin 4095, which nobody reads.
//   pick     the bin's entries (same registers, same three instructions) -> ordered keys in shared memory, ranked by counting
//            (<= 256 of them).  A fuller bin (outliers stretching the range, masses of duplicates), or a range the float map cannot
//            resolve, is refined from the registers instead, 11 bits of the ORDERED KEY at a time.
// The next row's loads are issued as soon as the registers are free, before the pick.  (Measured: keeping TWO rows in registers, the
// next row's loads a whole iteration ahead at 2 blocks per SM, is slower: 153 against 140 us.)
#define NPAIR_LSB_THREADS 256
#define NPAIR_LSB_VPT 8                    // 16-byte groups per thread: 256 * 8 * 4 = 8192 columns
#define NPAIR_LSB_BINS 2048
#define NPAIR_LSB_HIST 2304                // bins the find walks: 1 + 2048 + slack (the value map is shifted up by one bin); multiple of 256
#define NPAIR_LSB_CCAP 256                 // bin population ranked by counting
#ifndef NPAIR_LSB_MINB
#define NPAIR_LSB_MINB 3                   // resident blocks per SM (80 registers)
#endif
struct LselBlock {
  unsigned int hist[4096];                 // [0, NPAIR_LSB_HIST) are cleared and read; 4095 collects the NaN (excluded) entries
  uint32_t cand[NPAIR_LSB_CCAP];
  uint32_t same[NPAIR_LSEL_SCAP];
  unsigned int n_same, n_cand;
  unsigned int warp_tot[NPAIR_LSB_THREADS / 32];
  unsigned int out[3];                     // find: {bin, rank inside the bin, population}
  float red_min[NPAIR_LSB_THREADS / 32], red_max[NPAIR_LSB_THREADS / 32];
  uint32_t red_klo[NPAIR_LSB_THREADS / 32], red_khi[NPAIR_LSB_THREADS / 32];
};
// rank r within hist[0 .. PER * 256) in index order; every thread sums PER consecutive bins.  Two barriers; the result is in B.out
// afterwards (bin == PER * 256: rank out of range).
template <int PER>
__device__ __forceinline__ void block_find_bin_u32(LselBlock& B, unsigned int r) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  unsigned int h[PER], mine = 0;
#pragma unroll
  for (int q = 0; q < PER; ++q) { h[q] = B.hist[tid * PER + q]; mine += h[q]; }
  unsigned int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) B.warp_tot[w] = incl;
  if (tid == 0) B.out[0] = static_cast<unsigned int>(PER * NPAIR_LSB_THREADS);
  __syncthreads();
  unsigned int before = incl - mine;
#pragma unroll
  for (int k = 0; k < NPAIR_LSB_THREADS / 32; ++k) before += (k < w) ? B.warp_tot[k] : 0u;
  if (mine && r >= before && r < before + mine) {                  // exactly one thread
    unsigned int cum = before;
    int b = 0;
#pragma unroll
    for (int q = 0; q < PER - 1; ++q) { if (b == q && cum + h[q] <= r) { cum += h[q]; b = q + 1; } }
    unsigned int hb = h[0];
#pragma unroll
    for (int q = 1; q < PER; ++q) hb = (b == q) ? h[q] : hb;
    B.out[0] = static_cast<unsigned int>(tid * PER + b); B.out[1] = r - cum; B.out[2] = hb;
  }
  __syncthreads();
}

__device__ __forceinline__ uint4 ldg_stream_u4(const float* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// byte offset (bin * 4) of an entry in the histogram: monotone in f; NaN -> 0x3FFC
__device__ __forceinline__ uint32_t lsb_off(uint32_t bits, float s4, float c0) { return __float_as_uint(__fmaf_rn(__uint_as_float(bits), s4, c0)) & 0x3FFCu; }

__global__ void __launch_bounds__(NPAIR_LSB_THREADS, NPAIR_LSB_MINB) local_select_block_kernel(const float* __restrict__ S, long long ldS, int Q, int N,
                                                                                   const float* __restrict__ lab_rows, const float* __restrict__ lab_cols,
                                                                                   int self_offset, int side_mask /*1 AP, 2 AN*/, float sn_ap, float sn_an,
                                                                                   RowArrays ra, BlockScalars* bs) {
  __shared__ __align__(16) LselBlock B;
  constexpr uint32_t kNaN = 0x7FFFFFFFu;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool want_same = side_mask & 1, want_diff = side_mask & 2;
  // whole 4-column groups come through 16-byte loads (rows start 128-byte aligned: ldS is a multiple of 32); the last N % 4 columns
  // sit in one extra register of threads 0..2
  const int n4 = N & ~3;
  const bool has_tail = tid < N - n4;
  uint32_t v[4 * NPAIR_LSB_VPT + 1];       // [32] = the tail column (NaN where there is none)
  int i = blockIdx.x;
  auto load_row = [&](int r) {
    const float* row = S + static_cast<long long>(r) * ldS;
#pragma unroll
    for (int u = 0; u < NPAIR_LSB_VPT; ++u) {
      const int jj = (u * NPAIR_LSB_THREADS + tid) * 4;
      uint4 t = make_uint4(kNaN, kNaN, kNaN, kNaN);
      if (jj < n4) t = ldg_stream_u4(row + jj);
      v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
    }
    v[4 * NPAIR_LSB_VPT] = has_tail ? __float_as_uint(row[n4 + tid]) : kNaN;
  };
  if (i < Q) load_row(i);
  for (; i < Q; i += gridDim.x) {
    const float li = lab_rows[i];
    const int self_col = i + self_offset;
    const int cs = ra.cnt_same[i];
    for (int b = tid * 4; b < NPAIR_LSB_HIST; b += NPAIR_LSB_THREADS * 4) *reinterpret_cast<uint4*>(&B.hist[b]) = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) { B.n_same = 0; B.n_cand = 0; }
    __syncthreads();          // publishes the reset (the previous row's readers are behind the loop-end barrier)
    // ---------------- pass 0: labels -> same-label list, NaN in the registers; value range of the entries that stay ----------------
    float mn = FLT_MAX, mx = -FLT_MAX;
#pragma unroll
    for (int u = 0; u < NPAIR_LSB_VPT; ++u) {
      const int jj = (u * NPAIR_LSB_THREADS + tid) * 4;
      if (jj < n4) {
        const float4 l = __ldg(reinterpret_cast<const float4*>(lab_cols + jj));
        const float ll[4] = {l.x, l.y, l.z, l.w};
        if (l.x == li || l.y == li || l.z == li || l.w == li) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (ll[c] == li) {
              if (jj + c != self_col) { const unsigned int k = atomicAdd(&B.n_same, 1u); if (k < NPAIR_LSEL_SCAP) B.same[k] = v[4 * u + c]; }
              v[4 * u + c] = kNaN;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { mn = fminf(mn, __uint_as_float(v[4 * u + c])); mx = fmaxf(mx, __uint_as_float(v[4 * u + c])); }
      }
    }
    if (has_tail) {
      const int j = n4 + tid;
      if (lab_cols[j] == li) {
        if (j != self_col) { const unsigned int k = atomicAdd(&B.n_same, 1u); if (k < NPAIR_LSEL_SCAP) B.same[k] = v[4 * NPAIR_LSB_VPT]; }
        v[4 * NPAIR_LSB_VPT] = kNaN;
      }
      mn = fminf(mn, __uint_as_float(v[4 * NPAIR_LSB_VPT])); mx = fmaxf(mx, __uint_as_float(v[4 * NPAIR_LSB_VPT]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { B.red_min[w] = mn; B.red_max[w] = mx; }
    __syncthreads();
    const unsigned int ns = B.n_same;                              // == cs
    float lo = B.red_min[0], hi = B.red_max[0];
#pragma unroll
    for (int k = 1; k < NPAIR_LSB_THREADS / 32; ++k) { lo = fminf(lo, B.red_min[k]); hi = fmaxf(hi, B.red_max[k]); }
    bool slow_ap = false;
    unsigned long long pos_ap = 0, pos_an = 0;
    // ---------------- AP side: the same-label list is short; warp 0 ranks it by counting ----------------
    if (want_same) {
      if (cs == 0) { if (tid == 0) { atomicOr(&bs->err, DERR_EMPTY_LIST); ra.posi_thr[i] = 0.f; } }
      else if (!pos_index(sn_ap, static_cast<unsigned long long>(cs), pos_ap)) { if (tid == 0) { atomicOr(&bs->err, DERR_POS_RANGE); ra.posi_thr[i] = 0.f; } }
      else if (ns > NPAIR_LSEL_SCAP) slow_ap = true;
      else if (w == 0) {
        for (unsigned int e = lane; e < ns; e += 32) {
          const uint32_t key = f2ord(__uint_as_float(B.same[e]));
          unsigned int rk = 0;
          for (unsigned int t = 0; t < ns; ++t) { const uint32_t kt = f2ord(__uint_as_float(B.same[t])); rk += (kt < key || (kt == key && t < e)) ? 1u : 0u; }
          if (rk == static_cast<unsigned int>(pos_ap)) ra.posi_thr[i] = clamp_thr(ord2f(key));               // .cu:288
        }
      }
    }
    // ---------------- AN side (every condition below is block-uniform) ----------------
    bool have_an = false, refine = false, by_bin = false;
    float s4 = 0.f, c0 = 0.f;
    unsigned int rank = 0;
    uint32_t boff = 0;
    if (want_diff) {
      const unsigned long long size = static_cast<unsigned long long>(N - 1 - cs);
      if (size == 0) { if (tid == 0) { atomicOr(&bs->err, DERR_EMPTY_LIST); ra.nega_thr[i] = 0.f; } }
      else if (!pos_index(sn_an, size, pos_an)) { if (tid == 0) { atomicOr(&bs->err, DERR_POS_RANGE); ra.nega_thr[i] = 0.f; } }
      else have_an = true;
    }
    if (have_an) {
      rank = static_cast<unsigned int>(pos_an);
      // the value map: usable when it sends lo to bin >= 1 and hi to a bin the find walks (always, unless the range is empty or outside
      // what fp32 can scale -- then the key digits do the whole job)
      s4 = __fdiv_rn(4.f * NPAIR_LSB_BINS, hi - lo);
      c0 = __fmaf_rn(-lo, s4, 8388612.f);                           // 2^23 + 4: one bin of head room below lo
      const uint32_t o_lo = __float_as_uint(__fmaf_rn(lo, s4, c0)), o_hi = __float_as_uint(__fmaf_rn(hi, s4, c0));
      const bool map_ok = hi > lo && o_lo >= 0x4B000000u && o_hi >= o_lo && o_hi < 0x4B000000u + 4u * (NPAIR_LSB_HIST - 1);
      if (!map_ok) refine = true;
      else {
#pragma unroll
        for (int e = 0; e < 4 * NPAIR_LSB_VPT + 1; ++e) smem_inc_off(B.hist, lsb_off(v[e], s4, c0));
        __syncthreads();
        block_find_bin_u32<NPAIR_LSB_HIST / NPAIR_LSB_THREADS>(B, rank);
        boff = B.out[0] << 2; rank = B.out[1];
        if (B.out[2] > NPAIR_LSB_CCAP) { refine = true; by_bin = true; }
        else {
#pragma unroll
          for (int e = 0; e < 4 * NPAIR_LSB_VPT + 1; ++e)
            if (lsb_off(v[e], s4, c0) == boff) B.cand[atomicAdd(&B.n_cand, 1u)] = f2ord(__uint_as_float(v[e]));
        }
      }
      if (refine) {
        // Rare: the entries still in play (all of them, or one crowded bin) are narrowed by 11 bits of their ORDERED KEY per round, from
        // the registers: [klo, khi] always contains the wanted entry and `rank` counts inside it.
        uint32_t klo = 0xFFFFFFFFu, khi = 0u;
        auto in_play = [&](uint32_t bits) { return bits != kNaN && (!by_bin || lsb_off(bits, s4, c0) == boff); };
#pragma unroll
        for (int e = 0; e < 4 * NPAIR_LSB_VPT + 1; ++e)
          if (in_play(v[e])) { const uint32_t k = f2ord(__uint_as_float(v[e])); klo = min(klo, k); khi = max(khi, k); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, o)); khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, o)); }
        __syncthreads();                                            // readers of red_* / out of the steps above are done
        if (lane == 0) { B.red_klo[w] = klo; B.red_khi[w] = khi; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NPAIR_LSB_THREADS / 32; ++k) { klo = min(klo, B.red_klo[k]); khi = max(khi, B.red_khi[k]); }
        for (int round = 0; round < 4 && khi != klo; ++round) {
          const uint32_t range = khi - klo;
          const int shift = max(0, 32 - __clz(range) - 11);       // (range >> shift) < 2048
          for (int b = tid * 4; b < NPAIR_LSB_HIST; b += NPAIR_LSB_THREADS * 4) *reinterpret_cast<uint4*>(&B.hist[b]) = make_uint4(0u, 0u, 0u, 0u);
          __syncthreads();
#pragma unroll
          for (int e = 0; e < 4 * NPAIR_LSB_VPT + 1; ++e)
            if (in_play(v[e])) { const uint32_t k = f2ord(__uint_as_float(v[e])); if (k >= klo && k <= khi) smem_inc(&B.hist[(k - klo) >> shift]); }
          __syncthreads();
          block_find_bin_u32<NPAIR_LSB_HIST / NPAIR_LSB_THREADS>(B, rank);
          rank = B.out[1];
          klo += B.out[0] << shift;
          khi = min(khi, klo + ((shift ? (1u << shift) : 1u) - 1u));
        }
        if (tid == 0) ra.nega_thr[i] = clamp_thr(ord2f(klo));                                               // .cu:319
      }
    }
    // ---------------- the registers are free: the next row streams in while this row's pick runs ----------------
    const float* row = S + static_cast<long long>(i) * ldS;
    if (i + static_cast<int>(gridDim.x) < Q) load_row(i + static_cast<int>(gridDim.x));
    if (have_an && !refine) {
      __syncthreads();
      const unsigned int nc = B.n_cand;                            // <= NPAIR_LSB_CCAP = blockDim
      if (tid < static_cast<int>(nc)) {
        const uint32_t key = B.cand[tid];
        unsigned int rk = 0;
        for (unsigned int t = 0; t < nc; ++t) { const uint32_t kt = B.cand[t]; rk += (kt < key || (kt == key && static_cast<int>(t) < tid)) ? 1u : 0u; }
        if (rk == rank) ra.nega_thr[i] = clamp_thr(ord2f(key));                                             // .cu:319
      }
    }
    if (slow_ap) {                                                 // more than 128 same-label entries: warp 0 redoes the side with plain sweeps
      __syncthreads();
      if (w == 0) { const float t = clamp_thr(ord2f(slow_select_row(row, N, lab_cols, li, self_col, 0, static_cast<unsigned int>(pos_ap), B.hist, lane))); if (lane == 0) ra.posi_thr[i] = t; }
    }
    __syncthreads();
  }
}
void launch_local_select(const float* S, long long ldS, int Q, int N, const float* lab_rows, const float* lab_cols,
                         int self_offset, int side_mask, float sn_ap, float sn_an, RowArrays ra, BlockScalars* bs, int sms, bool force_warp_kernel, cudaStream_t st) {
  if (!force_warp_kernel && N <= NPAIR_LSB_THREADS * NPAIR_LSB_VPT * 4 && (reinterpret_cast<uintptr_t>(lab_cols) & 15) == 0 && (ldS & 3) == 0) {
    int grid = sms * NPAIR_LSB_MINB; if (grid > Q) grid = Q;
    local_select_block_kernel<<<grid, NPAIR_LSB_THREADS, 0, st>>>(S, ldS, Q, N, lab_rows, lab_cols, self_offset, side_mask, sn_ap, sn_an, ra, bs);
    count_launch();
    return;
  }
  const int smem = static_cast<int>(sizeof(LselWarp)) * NPAIR_LSEL_WARPS;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(local_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set = true; }
  const int per_sm = (227 * 1024) / (smem + 1024);
  int grid = sms * (per_sm < 1 ? 1 : per_sm);
  const int need = (Q + NPAIR_LSEL_WARPS - 1) / NPAIR_LSEL_WARPS;
  if (grid > need) grid = need;
  local_select_kernel<<<grid, 32 * NPAIR_LSEL_WARPS, smem, st>>>(S, ldS, Q, N, lab_rows, lab_cols, self_offset, side_mask, sn_ap, sn_an, ra, bs);
  count_launch();
}

// ---- GLOBAL: the rank's whole Q x N block.  Three kernels, each finished by its last block (ticket):
//   A  digit 1 (top 11 bits of the RAW float bits; three instructions per element on the all-different-label fast path) histogram
//      over S, 64-bit global counts; the last block walks the bins in value order -> bin, rank inside, population
//   B  second sweep of S: elements of that bin only (a shift and a compare per element): digit 2 histogram of their 21-bit
//      remainders, and -- when the bin fits the candidate buffer -- the remainders are compacted (per-block staging, one global
//      atomic per flush)
//   C  digit 3 over the candidates (or, oversized bin, over S once more) -> threshold, written to all rows
// Remainders of negative floats sort descending, so they are stored complemented ("flipped"): ascending everywhere.
struct GlobalSelectBufs {
  unsigned long long* hist;   // [2][2048]
  uint32_t* cand;             // [2][cap]
  unsigned int cap;
  int world_scope;            // 1: the digit counts are exchanged between the ranks before the decision
};
#define NPAIR_GSEL_STAGE 2048
__device__ __forceinline__ uint32_t d11_raw_of_order(uint32_t o) { return o < 1024u ? 2047u - o : o - 1024u; }

// Decides one digit of the GLOBAL select from the 64-bit counts in gb.hist (one block; `ordered` = 2048 x 8 bytes of shared memory)
__device__ void global_decide(int pass, bool act0, bool act1, GlobalSelectBufs gb, RowArrays ra, int Q, BlockScalars* bs,
                              unsigned long long* ordered, unsigned long long* s_scan, int* s_res, unsigned long long* s_out) {
  const int shift = pass == 1 ? 10 : 0;
  const int nbits = pass == 2 ? 10 : 11;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    if (!(side == 0 ? act0 : act1)) continue;
    const unsigned long long* gh = gb.hist + side * NPAIR_SEL_BINS;
    const int nb = 1 << nbits;
    __syncthreads();
    // pass 0 counted RAW digits: walk them in value order (negative floats: descending raw digit)
    for (int o = threadIdx.x; o < nb; o += blockDim.x) ordered[o] = __ldcg(&gh[pass == 0 ? d11_raw_of_order(o) : static_cast<uint32_t>(o)]);
    __syncthreads();
    unsigned long long r2, pp;
    const int d = find_bin(ordered, nb, bs->sel_rank[side], &r2, &pp, s_scan, s_res, s_out);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (d >= nb) { bs->err |= DERR_POS_RANGE; bs->sel_active[side] = 0; }
      else {
        bs->sel_rank[side] = r2;
        if (pass == 0) { bs->sel_prefix[side] = d11_raw_of_order(static_cast<uint32_t>(d)) << 21; bs->sel_cnt[side] = pp; bs->cand_n[side] = 0; }
        else bs->sel_prefix[side] |= static_cast<uint32_t>(d) << shift;
        if (pass == 2) {
          const uint32_t p = bs->sel_prefix[side];
          const uint32_t bits = (p & 0xFFE00000u) | ((p & 0x1FFFFFu) ^ ((p >> 21) >= 1024u ? 0x1FFFFFu : 0u));     // un-flip the remainder
          const float thr = clamp_thr(__uint_as_float(bits));                    // .cu:303 / :334
          if (side == 0) bs->posi_global = thr; else bs->nega_global = thr;
        }
      }
    }
    __syncthreads();
  }
  for (int b = threadIdx.x; b < 2 * NPAIR_SEL_BINS; b += blockDim.x) gb.hist[b] = 0ull;
  if (pass == 2) {
    __syncthreads();
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      if (!(side == 0 ? act0 : act1) || !bs->sel_active[side]) continue;
      const float thr = side == 0 ? bs->posi_global : bs->nega_global;
      float* out = side == 0 ? ra.posi_thr : ra.nega_thr;
      for (int i = threadIdx.x; i < Q; i += blockDim.x) out[i] = thr;
    }
  }
}

__global__ void __launch_bounds__(512) global_select_kernel(const float* __restrict__ S, long long ldS, int Q, int N, const float* __restrict__ lab_rows,
                                                            const float* __restrict__ lab_cols, int self_offset, int side_mask, int pass /*0,1,2*/,
                                                            GlobalSelectBufs gb, RowArrays ra, BlockScalars* bs) {
  __shared__ unsigned int hist[2][NPAIR_SEL_BINS];
  __shared__ uint32_t stage[2][NPAIR_GSEL_STAGE];
  __shared__ unsigned int s_nst[2], s_base[2];
  __shared__ unsigned long long s_scan[33], s_out[2];
  __shared__ int s_res, s_last;
  const bool act0 = (side_mask & 1) && bs->sel_active[0], act1 = (side_mask & 2) && bs->sel_active[1];
  if (!act0 && !act1) return;
  const bool lab_aligned = (reinterpret_cast<uintptr_t>(lab_cols) & 15) == 0;
  const int shift = pass == 1 ? 10 : 0;
  const int nbits = pass == 2 ? 10 : 11;
  const uint32_t dm = (1u << nbits) - 1u;
  // sel_prefix after pass 0: raw digit << 21; after pass 1: | flipped-remainder digit << 10
  const uint32_t raw0 = bs->sel_prefix[0] >> 21, raw1 = bs->sel_prefix[1] >> 21;
  const uint32_t flip0 = raw0 >= 1024u ? 0x1FFFFFu : 0u, flip1 = raw1 >= 1024u ? 0x1FFFFFu : 0u;
  const uint32_t mid0 = (bs->sel_prefix[0] >> 10) & 0x7FFu, mid1 = (bs->sel_prefix[1] >> 10) & 0x7FFu;   // pass 2: decided second digit
  const bool comp0 = act0 && pass == 1 && bs->sel_cnt[0] <= gb.cap, comp1 = act1 && pass == 1 && bs->sel_cnt[1] <= gb.cap;   // compaction this pass
  const bool list0 = act0 && pass == 2 && bs->sel_cnt[0] <= gb.cap, list1 = act1 && pass == 2 && bs->sel_cnt[1] <= gb.cap;   // read the list this pass
  for (int b = threadIdx.x; b < 2 * NPAIR_SEL_BINS; b += blockDim.x) (&hist[0][0])[b] = 0;
  if (threadIdx.x < 2) s_nst[threadIdx.x] = 0;
  __syncthreads();
  const bool sweep0 = act0 && !list0, sweep1 = act1 && !list1;

  // one element of the bin of `side` (passes 1 and 2): its flipped remainder goes to the digit histogram / the staging list
  auto take = [&](uint32_t bits, int side) {
    const uint32_t k = (bits & 0x1FFFFFu) ^ (side == 0 ? flip0 : flip1);
    if (pass == 2 && ((k >> 10) != (side == 0 ? mid0 : mid1))) return;
    smem_inc(&hist[side][(k >> shift) & dm]);
    if (side == 0 ? comp0 : comp1) {
      const unsigned int slot = atomicAdd(&s_nst[side], 1u);
      if (slot < NPAIR_GSEL_STAGE) stage[side][slot] = k;
      else {                                                     // staging full (rare): straight to the global list
        const unsigned int g = atomicAdd(&bs->cand_n[side], 1u);
        if (g < gb.cap) gb.cand[static_cast<size_t>(side) * gb.cap + g] = k;
      }
    }
  };

  if (sweep0 || sweep1) {
    for (int i = blockIdx.x; i < Q; i += gridDim.x) {
      const float li = lab_rows[i];
      const int self_col = i + self_offset;
      const float* row = S + static_cast<long long>(i) * ldS;
      // two 16-byte groups per thread in flight (S and labels): the sweeps are latency-bound otherwise
      for (int j0 = threadIdx.x * 4; j0 < N; j0 += blockDim.x * 8) {
        uint4 vq[2]; float lq[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j4 = j0 + u * blockDim.x * 4;
          if (j4 >= N) continue;
          vq[u] = __ldg(reinterpret_cast<const uint4*>(row + j4));              // row stride is a multiple of 32 floats: in bounds
          if (lab_aligned && j4 + 3 < N) { const float4 l = __ldg(reinterpret_cast<const float4*>(lab_cols + j4)); lq[u][0] = l.x; lq[u][1] = l.y; lq[u][2] = l.z; lq[u][3] = l.w; }
          else {
#pragma unroll
            for (int c = 0; c < 4; ++c) lq[u][c] = (j4 + c < N) ? __ldg(lab_cols + j4 + c) : li;
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
        const int j4 = j0 + u * blockDim.x * 4;
        if (j4 >= N) continue;
        const uint32_t vv[4] = {vq[u].x, vq[u].y, vq[u].z, vq[u].w};
        const float* ll = lq[u];
        if (j4 + 3 < N && ll[0] != li && ll[1] != li && ll[2] != li && ll[3] != li) {     // four diff-label pairs: the common case
          if (sweep1) {
            if (pass == 0) {
#pragma unroll
              for (int c = 0; c < 4; ++c) smem_inc_off(hist[1], (vv[c] >> 19) & 0x1FFCu);
            } else if ((vv[0] >> 21) == raw1 || (vv[1] >> 21) == raw1 || (vv[2] >> 21) == raw1 || (vv[3] >> 21) == raw1) {
#pragma unroll
              for (int c = 0; c < 4; ++c) if ((vv[c] >> 21) == raw1) take(vv[c], 1);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (j4 + c >= N || j4 + c == self_col) continue;                    // the self pair is in neither list (.cu:54)
            const int side = (ll[c] == li) ? 0 : 1;
            if (!(side == 0 ? sweep0 : sweep1)) continue;
            if (pass == 0) smem_inc(&hist[side][vv[c] >> 21]);
            else if ((vv[c] >> 21) == (side == 0 ? raw0 : raw1)) take(vv[c], side);
          }
        }
        }
      }
      if (pass == 1 && (comp0 || comp1)) {                         // flush a staging area that is at least half full
        __syncthreads();
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const unsigned int n = min(s_nst[side], static_cast<unsigned int>(NPAIR_GSEL_STAGE));
          if (n >= NPAIR_GSEL_STAGE / 2) {
            if (threadIdx.x == 0) s_base[side] = atomicAdd(&bs->cand_n[side], n);
            __syncthreads();
            for (unsigned int e = threadIdx.x; e < n; e += blockDim.x)
              if (s_base[side] + e < gb.cap) gb.cand[static_cast<size_t>(side) * gb.cap + s_base[side] + e] = stage[side][e];
            __syncthreads();
            if (threadIdx.x == 0) s_nst[side] = 0;
          }
        }
        __syncthreads();
      }
    }
  }
  if (list0 || list1) {                                            // pass 2 over the compact candidate lists (flipped remainders)
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      if (!(side == 0 ? list0 : list1)) continue;
      const unsigned int n = bs->cand_n[side];
      const uint32_t mid = side == 0 ? mid0 : mid1;
      const uint32_t* cl = gb.cand + static_cast<size_t>(side) * gb.cap;
      for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint32_t k = cl[e];
        if ((k >> 10) == mid) smem_inc(&hist[side][k & dm]);
      }
    }
  }
  __syncthreads();
  if (pass == 1) {                                                 // remaining staged candidates
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const unsigned int n = min(s_nst[side], static_cast<unsigned int>(NPAIR_GSEL_STAGE));
      if (n) {
        if (threadIdx.x == 0) s_base[side] = atomicAdd(&bs->cand_n[side], n);
        __syncthreads();
        for (unsigned int e = threadIdx.x; e < n; e += blockDim.x)
          if (s_base[side] + e < gb.cap) gb.cand[static_cast<size_t>(side) * gb.cap + s_base[side] + e] = stage[side][e];
        __syncthreads();
      }
    }
  }
  for (int b = threadIdx.x; b < 2 * NPAIR_SEL_BINS; b += blockDim.x) {
    const unsigned int h = (&hist[0][0])[b];
    if (h) atomicAdd(&gb.hist[b], static_cast<unsigned long long>(h));
  }
  // ---- last block: decide this digit (world scope: the counts are exchanged first, global_decide_kernel decides) ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&bs->ticket3, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) bs->ticket3 = 0;
  if (gb.world_scope) return;
  global_decide(pass, act0, act1, gb, ra, Q, bs, reinterpret_cast<unsigned long long*>(&hist[0][0]), s_scan, &s_res, s_out);
}
// world scope: sum the ranks' digit counts (same order on every rank -> identical decisions), then decide like the last block does
__global__ void __launch_bounds__(512) global_decide_kernel(const float* __restrict__ xall, int xstride, int world, int side_mask, int pass,
                                                            GlobalSelectBufs gb, RowArrays ra, int Q, BlockScalars* bs) {
  __shared__ unsigned long long ordered[NPAIR_SEL_BINS];
  __shared__ unsigned long long s_scan[33], s_out[2];
  __shared__ int s_res;
  const bool act0 = (side_mask & 1) && bs->sel_active[0], act1 = (side_mask & 2) && bs->sel_active[1];
  if (!act0 && !act1) return;
  for (int b = threadIdx.x; b < 2 * NPAIR_SEL_BINS; b += blockDim.x) {
    unsigned long long sum = 0;
    for (int r = 0; r < world; ++r) {
      const float* x = xall + static_cast<long long>(r) * xstride + 2 * b;
      sum += static_cast<unsigned long long>(__float_as_uint(x[0])) | (static_cast<unsigned long long>(__float_as_uint(x[1])) << 32);
    }
    gb.hist[b] = sum;
  }
  __syncthreads();
  global_decide(pass, act0, act1, gb, ra, Q, bs, ordered, s_scan, &s_res, s_out);
}
void launch_global_select_pass(const float* S, long long ldS, int Q, int N, const float* lab_rows, const float* lab_cols,
                               int self_offset, int side_mask, int pass, RowArrays ra, unsigned long long* hist, uint32_t* cand,
                               unsigned int cand_cap, int world_scope, BlockScalars* bs, int sms, cudaStream_t st) {
  int grid = sms * 4; if (grid > Q) grid = Q;
  GlobalSelectBufs gb; gb.hist = hist; gb.cand = cand; gb.cap = cand_cap; gb.world_scope = world_scope;
  global_select_kernel<<<grid, 512, 0, st>>>(S, ldS, Q, N, lab_rows, lab_cols, self_offset, side_mask, pass, gb, ra, bs);
  count_launch();
}
void launch_global_decide(const float* xall, int xstride, int world, int side_mask, int pass, RowArrays ra, int Q, unsigned long long* hist,
                          uint32_t* cand, unsigned int cand_cap, BlockScalars* bs, cudaStream_t st) {
  GlobalSelectBufs gb; gb.hist = hist; gb.cand = cand; gb.cap = cand_cap; gb.world_scope = 1;
  global_decide_kernel<<<1, 512, 0, st>>>(xall, xstride, world, side_mask, pass, gb, ra, Q, bs);
  count_launch();
}

// --------------------------------------------------------------------------------------------
// The forward row pass: one streaming read of S per row computes everything the reference spreads over
// GetSampledPairMtx (.cu:69-122), the count gemvs (.cu:355-360), Minus_Querywise_Maxval (.cu:124-156), the
// masked sums (.cu:373-380), ManipulateDIVandLOG (.cu:158-171) and GetRetrivePerformance (.cu:173-206).
// Retrieval uses the sort-free equivalence of SURVEY.md 9.4 Q11: with p* = max E over same-label non-self
// columns and c = #{non-self j : E_j >= p*},  hit_k  <=>  c <= min(k, N-2).
// --------------------------------------------------------------------------------------------
// Smallest float s (as an ordered key) with expf(s - max_all) >= pstar, searched downward from the best positive.
// Lets the retrieval count compare similarities instead of exponentials, so expf is only evaluated for SELECTED pairs.
__device__ __forceinline__ float retrieval_cut(float maxw, float max_all, int lane) {
  const float pstar = expf(maxw - max_all);
  if (!(pstar > 0.f)) return -INFINITY;                         // underflow: every entry ties with the best positive
  uint32_t base = f2ord(maxw), last_ok = base;
  for (int it = 0; it < 8; ++it) {                             // 256 ulps cover the flat steps of expf for |s-max| < ~80
    const uint32_t kc = base - static_cast<uint32_t>(lane);
    const bool ok = (kc <= base) && (expf(ord2f(kc) - max_all) >= pstar);
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int T = (bal == 0xffffffffu) ? 32 : (__ffs(~bal) - 1);
    if (T > 0) last_ok = base - static_cast<uint32_t>(T - 1);
    if (T < 32) return ord2f(last_ok);
    if (base < 64u) return ord2f(last_ok);
    base -= 32u;
  }
  // long flat step (denormal exponentials): bisection on the ordered keys, monotone expf assumed
  uint32_t lo = f2ord(-FLT_MAX), hi = last_ok;                 // invariant: hi satisfies
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (expf(ord2f(mid) - max_all) >= pstar) hi = mid; else lo = mid + 1u;
  }
  return ord2f(hi);
}

// 16-byte streaming load: read-only path, no L1 allocation (S is read exactly once by this kernel)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// One warp per anchor row, branch-free hot loop (12 instructions per element): one retrieval-count compare, one label
// compare, ONE selection compare with sign/threshold picked by the label predicate, a 2-instruction exponential, and two
// accumulations (T = A + B for every selected pair, A under the same-label predicate).
__device__ __forceinline__ void lse_elem(float sv, float lab, float li, float scut, float m2, float sgn_p, float thr_p,
                                         float sgn_n, float thr_n, float& A, float& T, int& c) {
  c += (sv >= scut) ? 1 : 0;                                  // == (exp(sv-max_all) >= exp(maxw-max_all)), SURVEY Q11
  const float e = fast_exp_m2(sv, m2);                        // .cu:130-131
  const bool same = (lab == li);
  const float key = sv * (same ? sgn_p : sgn_n);              // .cu:79-120 as one compare  +-s <= thr'
  const float es = (key <= (same ? thr_p : thr_n)) ? e : 0.f;
  T += es;
  if (same) A += es;
}

// The last block to finish also performs the job of the old finalize kernel (loss, retrieval ratios, asum, error word).
#ifndef NPAIR_LSE_U
#define NPAIR_LSE_U 4            // 16-byte loads in flight per lane and array (S, labels)
#endif
#ifndef NPAIR_LSE_MINB
#define NPAIR_LSE_MINB 3
#endif
__global__ void __launch_bounds__(256, NPAIR_LSE_MINB) lse_rows_kernel(const float* __restrict__ S, long long ldS, int Q, int N,
                                                       const float* __restrict__ lab_rows, const float* __restrict__ lab_cols,
                                                       int self_offset, MiningParams mp, RowArrays ra, BlockScalars* bs,
                                                       int num_tops, float* __restrict__ tops, float log2_world,
                                                       float* __restrict__ xout /*world scope: this rank's partial tops, else NULL*/,
                                                       int wpr /*warps per row: 1, 2, 4 or 8 (few rows per rank: keep the SMs full)*/,
                                                       unsigned int seq /*written behind the tops: the host polls it*/) {
  const int lane = threadIdx.x & 31;
  __shared__ float s_pA[8], s_pT[8];
  __shared__ int s_pc[8];
  // NPAIR_LSE_REV: walk the rows from the last to the first.  The similarity GEMM produced the high row blocks last, so
  // their tiles are the ones still resident in the 126 MB L2 when this kernel starts.
#ifndef NPAIR_LSE_REV
#define NPAIR_LSE_REV 1
#endif
  const int blk = NPAIR_LSE_REV ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int wib = threadIdx.x >> 5;
  const int i = blk * ((blockDim.x >> 5) / wpr) + wib / wpr;
  const int part = wib % wpr;                                   // this warp's column segment of the row
  float A = 0.f, T = 0.f; int c = 0;
  float m2 = 0.f, thr_p = 0.f, thr_n = 0.f, li = 0.f; int cs = 0;
  if (i < Q) {
    // the first block of this warp's segment is requested before anything else: the per-row set-up below (dependent loads of the row
    // statistics, the retrieval cut's expf search) then runs under the DRAM latency instead of in front of it
    constexpr int U = NPAIR_LSE_U;
    const float* row = S + static_cast<long long>(i) * ldS;
    // this warp's segment [c_lo, c_hi) of the row: multiples of 512 columns
    const int seg = ((N + wpr - 1) / wpr + 128 * U - 1) / (128 * U) * (128 * U);
    const int c_lo = min(N, part * seg), c_hi = min(N, c_lo + seg);
    const int n_full = c_lo + (c_hi - c_lo) / (128 * U) * (128 * U);
    const float4* srow4 = reinterpret_cast<const float4*>(row + c_lo) + lane;
    const float4* lab4 = reinterpret_cast<const float4*>(lab_cols + c_lo) + lane;
    const bool lab_aligned = (reinterpret_cast<uintptr_t>(lab_cols) & 15) == 0;
    const bool fast = lab_aligned && c_lo < n_full;
    float4 v[U], vn[U];
    if (fast) {
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg_stream(srow4 + 32 * u);
    }
    li = lab_rows[i];
    const int self_col = i + self_offset;
    const float max_all = ord2f(ra.st_maxall[i]);
    m2 = max_all * NPAIR_LOG2E;
    // GLOBAL-region thresholds are block-wide scalars (thresholds_kernel / global_pick_kernel); LOCAL ones are per row
    const float posi = mp.ap_region == REGION_GLOBAL ? bs->posi_global : ra.posi_thr[i];
    const float nega = mp.an_region == REGION_GLOBAL ? bs->nega_global : ra.nega_thr[i];
    if (lane == 0 && part == 0) { ra.posi_thr[i] = posi; ra.nega_thr[i] = nega; }   // kept per row for inspection (npair_debug_read)
    const float tp = posi + mp.margin_ident;                    // fp32 add as in .cu:81
    const float tn = nega + mp.margin_diff;                     // .cu:102
    const float sgn_p = ap_sign(mp.ap_method), sgn_n = an_sign(mp.an_method);
    thr_p = ap_thr(tp, mp.ap_method); thr_n = an_thr(tn, mp.an_method);
    cs = ra.cnt_same[i];
    const float scut = cs > 0 ? retrieval_cut(ord2f(ra.st_maxw[i]), max_all, lane) : INFINITY;
    // ---- full 512-column blocks: unguarded 128-bit loads (4 in flight per lane for S, 4 for the labels) ----
    int base = c_lo;
    if (fast) {
      // software pipeline: the 16-byte loads of block k+1 (streamed past the L1: every byte of S is used once) are in flight while
      // block k is evaluated; the labels (32 KB shared by every row) come from the L1 when they are needed
      for (; base < n_full; base += 128 * U, srow4 += 32 * U, lab4 += 32 * U) {
        const bool more = base + 128 * U < n_full;
        if (more) {
#pragma unroll
          for (int u = 0; u < U; ++u) vn[u] = ldg_stream(srow4 + 32 * U + 32 * u);
        }
        float4 l[U];
#pragma unroll
        for (int u = 0; u < U; ++u) l[u] = __ldg(lab4 + 32 * u);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j4 = base + u * 128 + lane * 4;
          const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          const float ll[4] = {l[u].x, l[u].y, l[u].z, l[u].w};
          const bool no_self = (self_col < j4 || self_col > j4 + 3);
          if (no_self && ll[0] != li && ll[1] != li && ll[2] != li && ll[3] != li) {
            // four diff-label pairs (all but ~cnt_same/4 groups of a row): no label-dependent selects, 8 instructions per pair
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              c += (vv[q] >= scut) ? 1 : 0;
              const float e = fast_exp_m2(vv[q], m2);
              if (vv[q] * sgn_n <= thr_n) T += e;
            }
          } else if (no_self) {
#pragma unroll
            for (int q = 0; q < 4; ++q) lse_elem(vv[q], ll[q], li, scut, m2, sgn_p, thr_p, sgn_n, thr_n, A, T, c);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (j4 + q != self_col) lse_elem(vv[q], ll[q], li, scut, m2, sgn_p, thr_p, sgn_n, thr_n, A, T, c);
          }
        }
        if (more) {
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = vn[u];
        }
      }
    }
    // ---- ragged tail (and the whole row when the label pointer is not 16-byte aligned) ----
    for (; base < c_hi; base += 512) {
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        const int j4 = base + u * 128 + lane * 4;
        if (j4 >= c_hi) continue;
        const float4 v4 = *reinterpret_cast<const float4*>(row + j4);      // row stride ldS is a multiple of 32: in bounds
        const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (j4 + q < c_hi && j4 + q != self_col) lse_elem(vv[q], lab_cols[j4 + q], li, scut, m2, sgn_p, thr_p, sgn_n, thr_n, A, T, c);
      }
    }
    A = warp_sum(A); T = warp_sum(T); c = warp_sum_i(c);
  }
  if (wpr > 1) {                                                // the row's segments, added in segment order by its first warp
    if (lane == 0) { s_pA[wib] = A; s_pT[wib] = T; s_pc[wib] = c; }
    __syncthreads();
    if (part == 0) {
      A = 0.f; T = 0.f; c = 0;
      for (int q = 0; q < wpr; ++q) { A += s_pA[wib + q]; T += s_pT[wib + q]; c += s_pc[wib + q]; }
    }
  }
  if (i < Q && part == 0) {
    if (lane == 0) {
      ra.A[i] = A; ra.T[i] = T;                                 // T = A + B (.cu:380)
      ra.logv[i] = (A == 0.f || T == 0.f) ? 0.f : logf(A / T);  // .cu:162-169
      const int lim = N - 2;
      ra.hits[i] = (cs > 0 && c <= min(1, lim)) ? 1 : 0;
      ra.hits[Q + i] = (cs > 0 && c <= min(5, lim)) ? 1 : 0;
      ra.hits[2 * Q + i] = (cs > 0 && c <= min(10, lim)) ? 1 : 0;
      const float invA = A == 0.f ? 0.f : 1.f / A;              // Get_Query_Diff_Part zero rules (.cu:410-415)
      const float invT = T == 0.f ? 0.f : 1.f / T;
      // first 16 bytes = all a diff-label pair needs.  Its backward weight exp(s - max) / T / world is evaluated by the gradient
      // kernel as ONE exponential 2^(s*log2(e) - m2c) with m2c = max*log2(e) + log2(T) + log2(world) (T == 0: +inf, weight 0):
      // {m2c, an_thr-transformed threshold, max*log2(e), label}; second 16 bytes = the same-label rule and the plain factors:
      // {ap_thr threshold, weight 1/T - 1/A, 1/T, 0}
      float4* rec = reinterpret_cast<float4*>(ra.rowscal + 8ll * i);
      rec[0] = make_float4(T == 0.f ? INFINITY : m2 + log2f(T) + log2_world, thr_n, m2, li);
      rec[1] = make_float4(thr_p, invT - invA, invT, 0.f);
    }
  }
  // ---- grid-level completion: the last block reduces the row results (fixed order -> deterministic) ----
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&bs->ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double s_l[8];
  __shared__ int s_h[3][8];
  double ls = 0.0; int h[3] = {0, 0, 0};
  // FU rows per thread and trip, all 4*FU loads issued before the first use: with one row per trip this tail was a chain of 32 L2
  // latencies at Q = 8192 (~20 us of the row pass).  The per-thread summation order (r ascending) is unchanged.
  constexpr int FU = 8;
  for (int r0 = threadIdx.x; r0 < Q; r0 += FU * blockDim.x) {
    float lv[FU]; int h0[FU], h1[FU], h2[FU];
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const int r = r0 + u * blockDim.x;
      const bool ok = r < Q;
      lv[u] = ok ? __ldcg(&ra.logv[r]) : 0.f;
      h0[u] = ok ? __ldcg(&ra.hits[r]) : 0; h1[u] = ok ? __ldcg(&ra.hits[Q + r]) : 0; h2[u] = ok ? __ldcg(&ra.hits[2 * Q + r]) : 0;
    }
#pragma unroll
    for (int u = 0; u < FU; ++u) { ls += lv[u]; h[0] += h0[u]; h[1] += h1[u]; h[2] += h2[u]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ls += __shfl_xor_sync(0xffffffffu, ls, o);
    h[0] += __shfl_xor_sync(0xffffffffu, h[0], o); h[1] += __shfl_xor_sync(0xffffffffu, h[1], o); h[2] += __shfl_xor_sync(0xffffffffu, h[2], o);
  }
  const int w = threadIdx.x >> 5;
  if (lane == 0) { s_l[w] = ls; s_h[0][w] = h[0]; s_h[1][w] = h[1]; s_h[2][w] = h[2]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ls = 0.0; h[0] = h[1] = h[2] = 0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) { ls += s_l[k]; h[0] += s_h[0][k]; h[1] += s_h[1][k]; h[2] += s_h[2][k]; }
    if (xout) {        // world scope: sums only; tops_world_kernel divides by the world's N after the exchange
      const unsigned long long lb = static_cast<unsigned long long>(__double_as_longlong(ls));
      xout[0] = __uint_as_float(static_cast<uint32_t>(lb)); xout[1] = __uint_as_float(static_cast<uint32_t>(lb >> 32));
      xout[2] = __int_as_float(h[0]); xout[3] = __int_as_float(h[1]); xout[4] = __int_as_float(h[2]);
      xout[5] = bs->asum; xout[6] = __int_as_float(bs->err);
      bs->ticket = 0;
      return;
    }
    float out[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    out[0] = static_cast<float>(ls) / static_cast<float>(-Q);                       // .cu:384-385
    for (int t = 1; t <= num_tops - 2 && t <= 3; ++t) out[t] = static_cast<float>(h[t - 1]) / static_cast<float>(Q);   // .cu:205
    out[num_tops - 1] = bs->asum / static_cast<float>(Q);                           // .cu:400-401 (always the LAST top)
    for (int t = 0; t < 5; ++t) tops[t] = out[t];
    reinterpret_cast<int*>(tops)[5] = bs->err;
    bs->ticket = 0;
    __threadfence_system();
    reinterpret_cast<volatile unsigned int*>(tops)[6] = seq;     // tops are visible on the host before the sequence number
  }
}
void launch_lse_rows(const float* S, long long ldS, int Q, int N, const float* lab_rows, const float* lab_cols,
                     int self_offset, MiningParams mp, RowArrays ra, BlockScalars* bs, int num_tops, float* tops_dev, int world, float* xout,
                     unsigned int seq, cudaStream_t st) {
  // few rows per rank (anchor sharding over many GPUs): several warps share a row so that every SM still holds ~32 warps
  int wpr = 1;
  while (wpr < 8 && static_cast<long long>(Q) * wpr < 4096 && N / (2 * wpr) >= 512) wpr *= 2;
  // measured on one rank's strip of a sharded HL job (tests/tune_rank.py, N = 8192): Q = 1024: 2 warps per row 18.6 us, 4: 22.7, 8: 25.5;
  // Q = 2048: 2: 29.8, 8: 37.1 -- beyond two warps the per-warp set-up outweighs the fuller SMs
  if (Q >= 1024 && wpr > 2) wpr = 2;
  // (measured: splitting rows over two warps at Q = 8192 to even out the 2.31 waves of one-warp-per-row blocks is slower, 80.6 vs 77 us)
#ifdef NPAIR_LSE_WPR_FORCE      // tuning builds only
  wpr = NPAIR_LSE_WPR_FORCE;
#endif
  if (wpr > 1) {
    const int rows_per_blk = 8 / wpr;
    const int grid = (Q + rows_per_blk - 1) / rows_per_blk;
    lse_rows_kernel<<<grid, 256, 0, st>>>(S, ldS, Q, N, lab_rows, lab_cols, self_offset, mp, ra, bs, num_tops, tops_dev,
                                          xout ? 0.f : log2f(static_cast<float>(world)), xout, wpr, seq);
    count_launch();
    return;
  }
  // 8 warps per block, 4 blocks per SM.  Measured at Q = 8192 (1.73 waves): 7 warps (1.98 waves, less idle tail) is SLOWER
  // (81.3 vs 78.7 us; 6: 83.7, 5: 87.6) -- the pass is latency-bound, more resident warps win.  NPAIR_LSE_WPB overrides.
  int wpb = 8;
  while (wpb > 1 && (Q + wpb - 1) / wpb < 296) wpb >>= 1;     // keep >= 2 blocks per SM when the rank has few rows
  const int grid = (Q + wpb - 1) / wpb;
  lse_rows_kernel<<<grid, wpb * 32, 0, st>>>(S, ldS, Q, N, lab_rows, lab_cols, self_offset, mp, ra, bs, num_tops, tops_dev,
                                             xout ? 0.f : log2f(static_cast<float>(world)), xout, 1, seq);
  count_launch();
}

// Measured and dropped (round 2, profiles/r02_experiments.md): a TMA-fed tile pass that walked only the upper triangle of the
// symmetric S (half the HBM bytes) ran in 84 us against 80 us for lse_rows_kernel -- the pass is issue-bound, not byte-bound.
// --------------------------------------------------------------------------------------------
// Backward weight builder: replaces Get_Query_Diff_Part x3 (.cu:405-419, :438-446).  The reference
// materialises W1,W2,W3 in fp32 and runs six GEMMs; here one pass over S produces the unit gradient
// weight   g'(i,c) = sel(i,c) * expf(S[i,c]-max_all_i) * (same ? 1/T_i - 1/A_i : 1/T_i)     ( = -W1+W2+W3 )
// directly as split 2-byte operand tiles for the tensor-core GEMM:
//   world == 1 :  H[j][m]  = g'(j,m) + g'(m,j)            dX = (lw/Q)/2 * H . X        (.cu:448-497 folded)
//   world  > 1 :  H[j][m]  = g'(j,m)  and  HT[m][j] = g'(j,m)
//                 local = H . X_total ,  total = HT . X_local , then reduce-scatter and blend (.cu:462-497)
// --------------------------------------------------------------------------------------------
struct RowScal { float maxall, tp, tn, cA, cT, lab; };   // maxall: max_all*log2e; tp/tn: ap_thr / an_thr transformed thresholds

// r.maxall holds max_all * log2(e) (see lse_rows_kernel)
__device__ __forceinline__ float gprime(float sv, bool same, const RowScal& r, float sgn_p, float sgn_n) {
  const float e = fast_exp_m2(sv, r.maxall);
  const float key = sv * (same ? sgn_p : sgn_n);
  return (key <= (same ? r.tp : r.tn)) ? e * (same ? r.cA : r.cT) : 0.f;
}

// four consecutive weights -> NS pieces, one 8-byte store per piece
template <int PREC>
__device__ __forceinline__ void store_quad(uint16_t* __restrict__ base, long long piece_stride, long long off, const float g[4]) {
  constexpr int NS = (PREC == PREC_BF16) ? 1 : (PREC == PREC_FP16X2 ? 2 : 3);
  if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f && g[3] == 0.f) {          // the common case under margin mining
#pragma unroll
    for (int s = 0; s < NS; ++s) *reinterpret_cast<uint2*>(base + s * piece_stride + off) = make_uint2(0u, 0u);
    return;
  }
  uint16_t p[4][3];
#pragma unroll
  for (int e = 0; e < 4; ++e) split3<PREC>(g[e], p[e][0], p[e][1], p[e][2]);
#pragma unroll
  for (int s = 0; s < NS; ++s)
    *reinterpret_cast<uint2*>(base + s * piece_stride + off) =
        make_uint2(static_cast<uint32_t>(p[0][s]) | (static_cast<uint32_t>(p[1][s]) << 16), static_cast<uint32_t>(p[2][s]) | (static_cast<uint32_t>(p[3][s]) << 16));
}

// 64 x 64 tiles, 256 threads; thread (tr = t/16, tc = t%16) owns the 4 x 4 micro-tile rows 4tr.., columns 4tc.. :
// one 16-byte load per row of the micro-tile, one 8-byte store per row and operand piece -- every request covers whole
// sectors, nothing is transposed.
// SYM (world == 1): the similarity GEMM wrote a bitwise symmetric S (EPI_SIM_SYM), so
//     H[j][m] = g'(S[j][m]; row j) + g'(S[j][m]; row m)                       (= G + G^T, .cu:448-497 folded)
//   needs only the row scalars of BOTH indices, which are local when world == 1.
// !SYM (world > 1): H[j][m] = g'(j,m) and the transposed copy HT[m][j] (micro-tile transposed in registers).
// MODE BW_ROWSCAL (world > 1): the similarity GEMM is bitwise symmetric ACROSS ranks (K-concatenated operands), so the
//   transposed term G[m][j] of row m on another rank is evaluated here from S[j][m] and row m's all-gathered scalars:
//     H[j][m] = g'(S[j][m]; row j) + (1/world) g'(S[j][m]; row m)          -- no N x D reduce-scatter (.cu:455-497)
template <int PREC, int MODE>
__global__ void __launch_bounds__(256, 4) build_weights_kernel(const float* __restrict__ S, long long ldS, int Q, int N,
                                                            const float* __restrict__ lab_rows, const float* __restrict__ lab_cols,
                                                            int self_offset, float inv_world, const float* __restrict__ rs_total,
                                                            MiningParams mp, RowArrays ra,
                                                            uint16_t* __restrict__ H, long long ldH, uint16_t* __restrict__ HT, long long ldHT) {
  constexpr bool SYM = (MODE != BW_SPLIT);      // both symmetric modes add the row-m term
  constexpr int TS = 64;
  const int ta = blockIdx.y, tb = blockIdx.x;
  __shared__ RowScal sc_a[TS], sc_b[TS];
  const int a0 = ta * TS, b0 = tb * TS;
  const int t = threadIdx.x, tr = t >> 4, tc = t & 15;
  const float sgn_p = ap_sign(mp.ap_method), sgn_n = an_sign(mp.an_method);
  if (t < TS) {
    const int j = a0 + t;
    RowScal r = {0.f, -INFINITY, -INFINITY, 0.f, 0.f, 0.f};
    if (j < Q) { const float* b = ra.rowscal + 8ll * j; r.maxall = b[2]; r.tn = b[1]; r.cT = b[6]; r.lab = b[3]; r.tp = b[4]; r.cA = b[5]; }
    sc_a[t] = r;
  } else if (t < 2 * TS) {
    const int mm = t - TS, m = b0 + mm;
    RowScal r = {0.f, -INFINITY, -INFINITY, 0.f, 0.f, 0.f};
    if (MODE == BW_SYM) {   // world == 1: column m is also a local row
      if (m < Q) { const float* b = ra.rowscal + 8ll * m; r.maxall = b[2]; r.tn = b[1]; r.cT = b[6]; r.lab = b[3]; r.tp = b[4]; r.cA = b[5]; }
    } else if (MODE == BW_ROWSCAL) {   // all-gathered [N][8] row records; the 1/world of .cu:474 folded into the weights
      if (m < N) {
        const float* b = rs_total + 8ll * m;
        r.maxall = b[2]; r.tn = b[1]; r.cT = b[6] * inv_world; r.lab = b[3]; r.tp = b[4]; r.cA = b[5] * inv_world;
      }
    } else if (m < N) r.lab = lab_cols[m];
    sc_b[mm] = r;
  }
  __syncthreads();
  const int ja0 = a0 + 4 * tr, mb0 = b0 + 4 * tc;       // my rows of block a, my columns of block b
  RowScal rb4[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) rb4[e] = sc_b[4 * tc + e];
  // block-uniform fast path: tile fully inside the matrix and not touching the self-pair diagonal
  const bool interior = (a0 + TS <= Q) && (b0 + TS <= N) && (a0 + self_offset + TS <= b0 || b0 + TS <= a0 + self_offset);
  const long long psH = static_cast<long long>(Q) * ldH;
  float gT[4][4];                                        // BW_SPLIT: transposed copy for HT
  float4 v4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {                          // all four 16-byte loads in flight before the arithmetic
    v4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ja0 + i < Q && mb0 < N) v4[i] = *reinterpret_cast<const float4*>(S + static_cast<long long>(ja0 + i) * ldS + mb0);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const RowScal rsa = sc_a[4 * tr + i];
    const float sv[4] = {v4[i].x, v4[i].y, v4[i].z, v4[i].w};
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = ja0 + i, m = mb0 + e;
      float x = 0.f;
      if (interior || (j < Q && m < N && m != j + self_offset)) {
        const bool same = rsa.lab == rb4[e].lab;
        x = gprime(sv[e], same, rsa, sgn_p, sgn_n);
        if (SYM) x += gprime(sv[e], same, rb4[e], sgn_p, sgn_n);
      }
      g[e] = x;
      if (!SYM) gT[e][i] = x;
    }
    if (ja0 + i < Q && mb0 < ldH) store_quad<PREC>(H, psH, static_cast<long long>(ja0 + i) * ldH + mb0, g);
  }
  if (!SYM) {
    const long long psT = static_cast<long long>(N) * ldHT;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (mb0 + e < N && ja0 < ldHT) store_quad<PREC>(HT, psT, static_cast<long long>(mb0 + e) * ldHT + ja0, gT[e]);   // rows beyond Q are zero
  }
}
void launch_build_weights(const float* S, long long ldS, int Q, int N, const float* lab_rows, const float* lab_cols,
                          int self_offset, int world, int mode, const float* rs_total, MiningParams mp, RowArrays ra, int prec,
                          uint16_t* H, long long ldH, uint16_t* HT, long long ldHT, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (Q + 63) / 64);
  const float inv_world = 1.f / static_cast<float>(world);
#define NPAIR_BW_ARGS S, ldS, Q, N, lab_rows, lab_cols, self_offset, inv_world, rs_total, mp, ra, H, ldH, HT, ldHT
#define NPAIR_LAUNCH_BW(P)                                                                            \
  do {                                                                                                \
    if (mode == BW_SYM) build_weights_kernel<P, BW_SYM><<<grid, 256, 0, st>>>(NPAIR_BW_ARGS);         \
    else if (mode == BW_ROWSCAL) build_weights_kernel<P, BW_ROWSCAL><<<grid, 256, 0, st>>>(NPAIR_BW_ARGS); \
    else build_weights_kernel<P, BW_SPLIT><<<grid, 256, 0, st>>>(NPAIR_BW_ARGS);                      \
  } while (0)
  if (prec == PREC_BF16) NPAIR_LAUNCH_BW(PREC_BF16);
  else if (prec == PREC_FP16X2) NPAIR_LAUNCH_BW(PREC_FP16X2);
  else NPAIR_LAUNCH_BW(PREC_BF16X3);
  count_launch();
#undef NPAIR_LAUNCH_BW
#undef NPAIR_BW_ARGS
}

__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n, float a) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += a * src[i];
}
// --------------------------------------------------------------------------------------------
// L2Normalize producer (usage/def.prototxt:115-120; the layer's source is not in the reference tree, so the semantics are
// stated here): y = x / ||x||_2 per sample, a zero row stays zero; backward dx = (dy - y (y . dy)) / ||x||.
// One warp per row, 16-byte loads, fixed summation order (lane-strided partial sums, then the shuffle tree).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, int rows, int dim, float* __restrict__ y,
                                                         float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + static_cast<long long>(r) * dim;
  float* yr = y + static_cast<long long>(r) * dim;
  const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  float ss = 0.f;
  if (vec) for (int d = lane * 4; d < dim; d += 128) { const float4 v = *reinterpret_cast<const float4*>(xr + d); ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss); }
  else for (int d = lane; d < dim; d += 32) ss = fmaf(xr[d], xr[d], ss);
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
  if (lane == 0 && inv_norm) inv_norm[r] = inv;
  if (vec) for (int d = lane * 4; d < dim; d += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + d);
    *reinterpret_cast<float4*>(yr + d) = nrm > 0.f ? make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else for (int d = lane; d < dim; d += 32) yr[d] = nrm > 0.f ? xr[d] / nrm : 0.f;
}
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv_norm, const float* __restrict__ dy,
                                                         int rows, int dim, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* yr = y + static_cast<long long>(r) * dim;
  const float* gr = dy + static_cast<long long>(r) * dim;
  float* dr = dx + static_cast<long long>(r) * dim;
  const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  float dot = 0.f;
  if (vec) for (int d = lane * 4; d < dim; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(yr + d), g = *reinterpret_cast<const float4*>(gr + d);
    dot = fmaf(a.x, g.x, dot); dot = fmaf(a.y, g.y, dot); dot = fmaf(a.z, g.z, dot); dot = fmaf(a.w, g.w, dot);
  } else for (int d = lane; d < dim; d += 32) dot = fmaf(yr[d], gr[d], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[r];
  if (vec) for (int d = lane * 4; d < dim; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(yr + d), g = *reinterpret_cast<const float4*>(gr + d);
    *reinterpret_cast<float4*>(dr + d) = make_float4((g.x - a.x * dot) * inv, (g.y - a.y * dot) * inv, (g.z - a.z * dot) * inv, (g.w - a.w * dot) * inv);
  } else for (int d = lane; d < dim; d += 32) dr[d] = (gr[d] - yr[d] * dot) * inv;
}
// world scope: tops from the ranks' partial sums, normalised by the world's N (identical on every rank)
__global__ void tops_world_kernel(const float* __restrict__ xall, int xstride, int world, long long N, int num_tops, float* __restrict__ tops,
                                  unsigned int seq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ls = 0.0; long long h[3] = {0, 0, 0}; double asum = 0.0; int err = 0;
  for (int r = 0; r < world; ++r) {
    const float* x = xall + static_cast<long long>(r) * xstride;
    const unsigned long long lb = static_cast<unsigned long long>(__float_as_uint(x[0])) | (static_cast<unsigned long long>(__float_as_uint(x[1])) << 32);
    ls += __longlong_as_double(static_cast<long long>(lb));
    h[0] += __float_as_int(x[2]); h[1] += __float_as_int(x[3]); h[2] += __float_as_int(x[4]);
    asum += x[5]; err |= __float_as_int(x[6]);
  }
  float out[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  out[0] = static_cast<float>(ls) / static_cast<float>(-N);
  for (int t = 1; t <= num_tops - 2 && t <= 3; ++t) out[t] = static_cast<float>(h[t - 1]) / static_cast<float>(N);
  out[num_tops - 1] = static_cast<float>(asum) / static_cast<float>(N);
  for (int t = 0; t < 5; ++t) tops[t] = out[t];
  reinterpret_cast<int*>(tops)[5] = err;
  __threadfence_system();
  reinterpret_cast<volatile unsigned int*>(tops)[6] = seq;
}
void launch_tops_world(const float* xall, int xstride, int world, long long N, int num_tops, float* tops_dev, unsigned int seq, cudaStream_t st) {
  tops_world_kernel<<<1, 32, 0, st>>>(xall, xstride, world, N, num_tops, tops_dev, seq);
  count_launch();
}

void launch_l2norm_fwd(const float* x, int rows, int dim, float* y, float* inv_norm, cudaStream_t st) {
  l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows, dim, y, inv_norm);
  count_launch();
}
void launch_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, int rows, int dim, float* dx, cudaStream_t st) {
  l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(y, inv_norm, dy, rows, dim, dx);
  count_launch();
}

void launch_axpy_rows(float* dst, const float* src, long long n, float a, cudaStream_t st) {
  int nb = static_cast<int>((n + 255) / 256); if (nb > 148 * 8) nb = 148 * 8; if (nb < 1) nb = 1;
  axpy_kernel<<<nb, 256, 0, st>>>(dst, src, n, a);
  count_launch();
}

}  // namespace npair
