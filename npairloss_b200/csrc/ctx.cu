// ctx.cu -- host side of libnpair_b200.so: the context behind the C ABI of include/npair_b200.h.
// Owns all device scratch (the reference keeps ~31 member Blobs, npair_multi_class_loss.hpp:59-78 / .cpp:44-154),
// builds the TMA tensor maps, enqueues the kernels of Forward_gpu (.cu:207-402) and Backward_gpu (.cu:420-499)
// on the caller's stream, and talks to NCCL (dlopen'ed, so the library loads on machines without it).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <sched.h>
#include <cudaTypedefs.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/npair_b200.h"
#include "gemm_tcgen05.cuh"
#include "grad_fused.cuh"
#include "kernels.cuh"

namespace npair {

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_create_err;

static std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap; va_start(ap, f); vsnprintf(buf, sizeof(buf), f, ap); va_end(ap);
  return std::string(buf);
}

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
struct NcclId { char internal[128]; };
typedef int (*fn_ncclGetUniqueId)(NcclId*);
typedef int (*fn_ncclCommInitRank)(void**, int, NcclId, int);
typedef int (*fn_ncclCommDestroy)(void*);
typedef int (*fn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_ncclReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclGroup)(void);
typedef const char* (*fn_ncclGetErrorString)(int);
struct NcclApi {
  void* h = nullptr;
  fn_ncclGetUniqueId GetUniqueId = nullptr;
  fn_ncclCommInitRank CommInitRank = nullptr;
  fn_ncclCommDestroy CommDestroy = nullptr;
  fn_ncclAllGather AllGather = nullptr;
  fn_ncclReduceScatter ReduceScatter = nullptr;
  fn_ncclGroup GroupStart = nullptr, GroupEnd = nullptr;
  fn_ncclGetErrorString GetErrorString = nullptr;
  std::string err;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* env = getenv("NPAIR_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.h) break;
  }
  if (!api.h) { api.err = "libnccl.so.2 not found (set NPAIR_NCCL_LIB)"; return &api; }
#define NPAIR_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.h, "nccl" #name)); if (!api.name) api.err = "missing symbol nccl" #name;
  NPAIR_SYM(GetUniqueId) NPAIR_SYM(CommInitRank) NPAIR_SYM(CommDestroy) NPAIR_SYM(AllGather) NPAIR_SYM(ReduceScatter)
  NPAIR_SYM(GroupStart) NPAIR_SYM(GroupEnd) NPAIR_SYM(GetErrorString)
#undef NPAIR_SYM
  return &api;
}
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

// One communicator per (process, unique id): a second context created with the same 128-byte id -- a solver's TRAIN and TEST nets,
// a repeated LayerSetUp, a wrapper that rebuilds its context for a new batch size -- shares the communicator of the first
// (an ncclUniqueId can only be consumed once by ncclCommInitRank).  Reference-counted; destroyed with its last context.
struct SharedComm { void* comm; int refs; };
static std::mutex g_comm_mu;
static std::map<std::string, SharedComm> g_comms;
static int acquire_comm(const void* id128, int world, int rank, void** out, std::string* err) {
  NcclApi* api = nccl_api();
  const std::string key(static_cast<const char*>(id128), 128);
  std::lock_guard<std::mutex> lk(g_comm_mu);
  auto it = g_comms.find(key);
  if (it != g_comms.end()) { ++it->second.refs; *out = it->second.comm; return 0; }
  NcclId id; memcpy(&id, id128, 128);
  void* comm = nullptr;
  const int r = api->CommInitRank(&comm, world, id, rank);
  if (r != 0) { *err = fmt("ncclCommInitRank: %s", api->GetErrorString(r)); return r; }
  g_comms[key] = SharedComm{comm, 1};
  *out = comm;
  return 0;
}
static void release_comm(void* comm) {
  NcclApi* api = nccl_api();
  std::lock_guard<std::mutex> lk(g_comm_mu);
  for (auto it = g_comms.begin(); it != g_comms.end(); ++it)
    if (it->second.comm == comm) {
      if (--it->second.refs == 0) { if (api->CommDestroy) api->CommDestroy(comm); g_comms.erase(it); }
      return;
    }
}

// ------------------------------------------------------------------------------------------------ TMA maps
static PFN_cuTensorMapEncodeTiled_v12000 tmap_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 3-D map over NSPLIT stacked 2-byte matrices [piece][rows][ld]; inner extent `cols` (<= ld), box {bk, box_rows, 1}
static bool make_tmap_pieces(CUtensorMap* m, const void* base, int cols, int rows, int pieces, long long ld_elems,
                             long long piece_stride_elems, int bk, int box_rows, std::string* err) {
  auto fn = tmap_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(pieces)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld_elems) * 2ull, static_cast<cuuint64_t>(piece_stride_elems) * 2ull};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(box_rows), 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUtensorMapSwizzle sw = (bk * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = fmt("cuTensorMapEncodeTiled failed (%d): cols=%d rows=%d pieces=%d ld=%lld", (int)r, cols, rows, pieces, ld_elems); return false; }
  return true;
}

// 2-D fp32 map over the similarity matrix [rows x ld], inner extent `cols`, box {32, 32}, 128B swizzle (TMA stores)
static bool make_tmap_f32_store(CUtensorMap* m, const void* base, int cols, int rows, long long ld_elems, std::string* err, int box_rows = 32) {
  auto fn = tmap_encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld_elems) * 4ull};
  cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = fmt("cuTensorMapEncodeTiled(S) failed (%d): cols=%d rows=%d ld=%lld", (int)r, cols, rows, ld_elems); return false; }
  return true;
}

// ------------------------------------------------------------------------------------------------ GEMM launchers
static inline int nsplit_of_prec(int prec) { return prec == PREC_BF16 ? 1 : (prec == PREC_FP16X2 ? 2 : 3); }
// K-block per (operand format, GEMM role); see GemmCfg
static inline int bk_of(int prec, int epi) {
  if (epi != EPI_OUT) return 64;                 // similarity GEMM: single pass, 64-element K blocks
  return prec == PREC_BF16 ? 64 : 32;
}
static inline int kcat_mult(int prec) { return prec == PREC_FP16X2 ? 3 : (prec == PREC_BF16X3 ? 6 : 1); }

template <int NSPLIT, bool BF16, int EPI, int BK>
static cudaError_t launch_split_gemm_t(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& sm, const GemmParams& p, int sms, cudaStream_t st) {
  using Cfg = GemmCfg<NSPLIT, BK>;
  auto kern = split_gemm_kernel<NSPLIT, BF16, EPI, BK>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = p.tile_list ? p.num_tiles_list : p.tiles_m * p.tiles_n * p.splits;
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(a, b, sm, p);
  count_launch();
  return cudaGetLastError();
}
// CTA-pair launch (cluster dimension 2, tcgen05 cta_group::2); p counts 256-row pair blocks (tiles_m / tile list)
template <bool BF16, int EPI>
static cudaError_t launch_pair_gemm_t(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& sm, const GemmParams& p, int sms, cudaStream_t st) {
  using Cfg = GemmCfg<1, 64, 2>;
  auto kern = split_gemm_kernel<1, BF16, EPI, 64, 2>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = p.tile_list ? p.num_tiles_list : p.tiles_m * p.tiles_n * p.splits;
  const int pairs = sms / 2;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  lc.blockDim = dim3(Cfg::THREADS);
  lc.dynamicSmemBytes = Cfg::SMEM_BYTES;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  count_launch();
  return cudaLaunchKernelEx(&lc, kern, a, b, sm, p);
}
static cudaError_t launch_sim_gemm_pair(int prec, bool sym_tiles, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& sm, const GemmParams& p, int sms, cudaStream_t st) {
  if (prec == PREC_FP16X2) return sym_tiles ? launch_pair_gemm_t<false, EPI_SIM_SYM>(a, b, sm, p, sms, st) : launch_pair_gemm_t<false, EPI_SIM>(a, b, sm, p, sms, st);
  return sym_tiles ? launch_pair_gemm_t<true, EPI_SIM_SYM>(a, b, sm, p, sms, st) : launch_pair_gemm_t<true, EPI_SIM>(a, b, sm, p, sms, st);
}
// `sm`: fp32 tensor map of the similarity matrix for EPI_SIM's TMA stores (ignored by EPI_OUT: pass any valid map)
// Similarity GEMM: always ONE MMA pass over K-concatenated operands (see split_kernel), fp16 or bf16 elements.
static cudaError_t launch_sim_gemm(int prec, bool sym_tiles, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& sm, const GemmParams& p, int sms, cudaStream_t st) {
  if (prec == PREC_FP16X2) return sym_tiles ? launch_split_gemm_t<1, false, EPI_SIM_SYM, 64>(a, b, sm, p, sms, st) : launch_split_gemm_t<1, false, EPI_SIM, 64>(a, b, sm, p, sms, st);
  return sym_tiles ? launch_split_gemm_t<1, true, EPI_SIM_SYM, 64>(a, b, sm, p, sms, st) : launch_split_gemm_t<1, true, EPI_SIM, 64>(a, b, sm, p, sms, st);
}
// Gradient GEMM: A = split gradient weights, B = split transposed features (EPI_OUT)
static cudaError_t launch_split_gemm(int prec, int epi, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& sm, const GemmParams& p, int sms, cudaStream_t st) {
  (void)epi;
  if (prec == PREC_BF16) return launch_split_gemm_t<1, true, EPI_OUT, 64>(a, b, sm, p, sms, st);
  if (prec == PREC_FP16X2) return launch_split_gemm_t<2, false, EPI_OUT, 32>(a, b, sm, p, sms, st);
  return launch_split_gemm_t<3, true, EPI_OUT, 32>(a, b, sm, p, sms, st);
}

template <int NSPLIT, bool BF16>
static cudaError_t launch_fused_grad_t(const CUtensorMap& b, const CUtensorMap& sm, const FusedGradParams& p, int sms, cudaStream_t st) {
  using Cfg = FusedCfg<NSPLIT>;
  auto kern = fused_grad_kernel<NSPLIT, BF16>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n * p.splits;
  kern<<<tiles < sms ? tiles : sms, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(b, sm, p);
  count_launch();
  return cudaGetLastError();
}
// CTA-pair launch of the fused gradient kernel (cluster dimension 2); p.tiles_m counts 256-row pair blocks, `b` has 128-row boxes
template <int NSPLIT, bool BF16>
static cudaError_t launch_fused_grad_pair_t(const CUtensorMap& b, const CUtensorMap& sm, const FusedGradParams& p, int sms, cudaStream_t st) {
  using Cfg = FusedCfg<NSPLIT, 2>;
  auto kern = fused_grad_kernel<NSPLIT, BF16, 2>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n * p.splits;
  const int pairs = sms / 2;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  lc.blockDim = dim3(Cfg::THREADS);
  lc.dynamicSmemBytes = Cfg::SMEM_BYTES;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  count_launch();
  return cudaLaunchKernelEx(&lc, kern, b, sm, p);
}
static cudaError_t launch_fused_grad_pair(int prec, const CUtensorMap& b, const CUtensorMap& sm, const FusedGradParams& p, int sms, cudaStream_t st) {
  if (prec == PREC_BF16) return launch_fused_grad_pair_t<1, true>(b, sm, p, sms, st);
  if (prec == PREC_FP16X2) return launch_fused_grad_pair_t<2, false>(b, sm, p, sms, st);
  return launch_fused_grad_pair_t<3, true>(b, sm, p, sms, st);
}
static cudaError_t launch_fused_grad(int prec, const CUtensorMap& b, const CUtensorMap& sm, const FusedGradParams& p, int sms, cudaStream_t st) {
  if (prec == PREC_BF16) return launch_fused_grad_t<1, true>(b, sm, p, sms, st);
  if (prec == PREC_FP16X2) return launch_fused_grad_t<2, false>(b, sm, p, sms, st);
  return launch_fused_grad_t<3, true>(b, sm, p, sms, st);
}

// SIMT cross-check of the same contraction on the same split operands (tests only; NPAIR_GEMM_SIMT_CHECK).
template <int PREC>
__device__ __forceinline__ float piece_sum(const uint16_t* base, long long off, long long ps) {
  if (PREC == PREC_BF16) return __bfloat162float(__ushort_as_bfloat16(base[off]));
  if (PREC == PREC_FP16X2) return __half2float(__ushort_as_half(base[off])) + __half2float(__ushort_as_half(base[ps + off]));
  return __bfloat162float(__ushort_as_bfloat16(base[off])) + __bfloat162float(__ushort_as_bfloat16(base[ps + off])) +
         __bfloat162float(__ushort_as_bfloat16(base[2 * ps + off]));
}
template <int PREC, int EPI>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const uint16_t* __restrict__ A, long long lda, long long psA,
                                                        const uint16_t* __restrict__ B, long long ldb, long long psB, int K, GemmParams p) {
  __shared__ float As[16][65], Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, kk = e & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + kk;
      As[kk][r] = (gm < p.M && gk < K) ? piece_sum<PREC>(A, static_cast<long long>(gm) * lda + gk, psA) : 0.f;
      Bs[kk][r] = (gn < p.Nn && gk < K) ? piece_sum<PREC>(B, static_cast<long long>(gn) * ldb + gk, psB) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float inv = p.dev_scale ? *p.dev_scale : 1.f;
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= p.M) continue;
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.Nn) continue;
      if (EPI == EPI_SIM) p.S[static_cast<long long>(row) * p.ldS + col] = acc[i][j] * inv * inv;
      else {
        float* d = p.out + static_cast<long long>(row) * p.ldo + col;
        float o = p.alpha * inv * acc[i][j];
        if (p.beta != 0.f) o += p.beta * *d;
        *d = o;
      }
    }
  }
}
static cudaError_t launch_simt_gemm(int prec, int epi, const uint16_t* A, long long lda, long long psA, const uint16_t* B, long long ldb,
                                    long long psB, int K, const GemmParams& p, cudaStream_t st) {
  dim3 grid((p.Nn + 63) / 64, (p.M + 63) / 64);
#define NPAIR_SIMT(P)                                                                                   \
  do {                                                                                                  \
    if (epi == EPI_SIM) simt_gemm_kernel<P, EPI_SIM><<<grid, 256, 0, st>>>(A, lda, psA, B, ldb, psB, K, p); \
    else simt_gemm_kernel<P, EPI_OUT><<<grid, 256, 0, st>>>(A, lda, psA, B, ldb, psB, K, p);                \
  } while (0)
  if (prec == PREC_BF16) NPAIR_SIMT(PREC_BF16);
  else if (prec == PREC_FP16X2) NPAIR_SIMT(PREC_FP16X2);
  else NPAIR_SIMT(PREC_BF16X3);
  count_launch();
#undef NPAIR_SIMT
  return cudaGetLastError();
}

// out = sum_s part[s] + beta*out, fixed summation order (deterministic split-K)
// ---- peer-memory exchange (world > 1, one process per GPU, NVLink / NVSwitch) ----
// Every rank owns one exported region (cudaIpc-mapped into all ranks):
//     X[2][N][D] fp32 | LAB[2][N] | REC[2][N][8] | XCH[2][world][8192] (small world-scope reductions) | FLAGS[3 kinds][2][world] uint32
// and PUSHES its own rows into every rank's region with plain stores over NVLink, then raises one flag per peer (release.sys);
// consumers wait for the world's flags (acquire.sys).  Replaces GatherFeatureAndLabel's MPI_Allgather (reference .cu:17-43) and the
// backward's exchange (row records instead of the N x D all-reduce, .cu:462-489) without a collective rendezvous: nothing
// waits for a slower rank until its data is really needed.  Buffers are double-buffered by the parity of the step counter and
// the flags carry the step number (never reset), so a rank that runs one step ahead never overwrites data still in use;
// like any collective this requires all ranks to issue the same sequence of forward / backward calls.
__global__ void p2p_push_kernel(const float* __restrict__ srcA, long long nA, long long offA, const float* __restrict__ srcB, long long nB, long long offB,
                                float* const* __restrict__ peer_base, long long flags_off /*in floats*/, int flag_index, int world, uint32_t epoch,
                                unsigned int* ticket) {
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x, nth = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool vecA = (nA & 3) == 0 && (offA & 3) == 0 && (reinterpret_cast<uintptr_t>(srcA) & 15) == 0;
  if (vecA) {
    for (long long i = tid; i < (nA >> 2); i += nth) {
      const float4 v = reinterpret_cast<const float4*>(srcA)[i];
      for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peer_base[r] + offA)[i] = v;
    }
  } else {
    for (long long i = tid; i < nA; i += nth) { const float v = srcA[i]; for (int r = 0; r < world; ++r) peer_base[r][offA + i] = v; }
  }
  for (long long i = tid; i < nB; i += nth) { const float v = srcB[i]; for (int r = 0; r < world; ++r) peer_base[r][offB + i] = v; }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (static_cast<int>(threadIdx.x) < world) {
    uint32_t* f = reinterpret_cast<uint32_t*>(peer_base[threadIdx.x] + flags_off) + flag_index;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
  if (threadIdx.x == 0) *ticket = 0;
}
__global__ void p2p_wait_kernel(const uint32_t* __restrict__ flags /*[world]*/, int world, uint32_t epoch) {
  if (static_cast<int>(threadIdx.x) < world) {
    uint32_t v;
    do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory"); } while (static_cast<int32_t>(v - epoch) < 0);
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long long n, float* __restrict__ out, float beta) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 a = *reinterpret_cast<const float4*>(part + i);
      for (int s = 1; s < splits; ++s) { const float4 b = *reinterpret_cast<const float4*>(part + s * n + i); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
      if (beta != 0.f) { const float4 o = *reinterpret_cast<const float4*>(out + i); a.x += beta * o.x; a.y += beta * o.y; a.z += beta * o.z; a.w += beta * o.w; }
      *reinterpret_cast<float4*>(out + i) = a;
    } else {
      for (long long j = i; j < n; ++j) {
        float a = part[j];
        for (int s = 1; s < splits; ++s) a += part[s * n + j];
        if (beta != 0.f) a += beta * out[j];
        out[j] = a;
      }
    }
  }
}
// choose a split-K factor that fills the SMs when the output has few tiles (strong scaling: Q = B/world shrinks)
static void plan_splits(GemmParams* gp, int sms, long long max_part_floats) {
  const int tiles = gp->tiles_m * gp->tiles_n;
  int splits = sms / (tiles > 0 ? tiles : 1);
  if (splits > 16) splits = 16;
  if (splits > gp->num_kblocks / 4) splits = gp->num_kblocks / 4;       // keep >= 4 k-blocks per split
  while (splits > 1 && static_cast<long long>(splits) * gp->M * gp->ldo > max_part_floats) --splits;
  if (splits < 1) splits = 1;
  int kpb = (gp->num_kblocks + splits - 1) / splits;
  splits = (gp->num_kblocks + kpb - 1) / kpb;                            // no empty split
  gp->splits = splits; gp->kb_per_split = kpb;
}

static inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

}  // namespace npair

using namespace npair;

#define NPAIR_PROF_PHASES 9
#define NPAIR_INTERNAL_FULL_TILES (1 << 30)   // npair_config.flags, internal: world == 1 computes every tile (symmetry self-check)
#define NPAIR_XCH_FLOATS 8192          // largest small exchange: two sides x 2048 64-bit digit counts

// ------------------------------------------------------------------------------------------------ context
struct npair_ctx {
  npair_config cfg;
  int Q, N, D, world, rank, prec, nsplit, bk_sim, bk_grad, sms, device;
  long long Dp, Np, Qp, ldS;
  // device scratch
  float* Xtot_buf = nullptr;     // world > 1: all-gather target
  float* labtot_buf = nullptr;
  float* S = nullptr;
  uint16_t *Xs = nullptr, *XsT = nullptr, *XlT = nullptr, *H = nullptr, *HT = nullptr;
  float* OUT2 = nullptr;         // world > 1: N x D transposed-term product before the reduce-scatter
  int bwd_mode = 0;              // NPAIR_BWDMODE_*
  uint16_t *XcatA = nullptr, *XcatB = nullptr;   // row-scalar mode, fp16x2: K-concatenated operands [N][3*Dp]
  float* rs_total = nullptr;     // row-scalar mode: all-gathered [N][8] row records
  CUtensorMap tm_catA, tm_catB;
  CUtensorMap tm_simB2, tm_catB2;   // 128-row B boxes for the CTA-pair similarity GEMM
  bool sim_pair = false;            // similarity GEMM runs in CTA-pair mode (cta_group::2)
  int2* sym_tiles2 = nullptr;       // world == 1: (pair_m, n_blk) list of the pair kernel
  int n_sym_tiles2 = 0;
  CUtensorMap tm_fB2;            // 128-row boxes of X^T for the CTA-pair gradient kernel
  bool grad_pair = false;        // fused gradient kernel runs in CTA-pair mode
  float *Ynorm = nullptr, *dY = nullptr, *inv_norm = nullptr;   // normalize_input: x / ||x||, gradient w.r.t. it, 1 / ||x||
  int grad_chunk_kb = 64;        // accumulation chunk of the gradient GEMM in 32-column K blocks (grad_fused.cuh); 0 = unchunked
  CUtensorMap tm_fB, tm_fS;      // fused gradient kernel: X^T pieces with 32-wide K boxes, 128-row fp32 boxes of S
  bool fused_grad = false;
  bool rs_gathered = false;
  bool ext_gathered = false;      // the current forward came through npair_forward_gathered (external collectives)
  bool defer_sync = false;        // npair_forward_backward: the forward returns after enqueueing, the caller synchronises later
  // peer-memory exchange (world > 1 with a communicator; NPAIR_FLAG_NCCL_FEATURES / _RECORDS fall back to NCCL)
  bool p2p_feat = false, p2p_rec = false;
  float* p2p_region = nullptr;         // X[2][N][D] | LAB[2][N] | REC[2][N][8] | FLAGS
  long long p2p_offX = 0, p2p_offLab = 0, p2p_offRec = 0, p2p_offXch = 0, p2p_offFlags = 0;   // in floats
  uint32_t xch_epoch = 0;              // small exchanges of the world-scope mode (several per step)
  float* xch_src = nullptr;            // [8192] staging of this rank's contribution
  float* xch_all = nullptr;            // NCCL fallback: gathered [world][8192]
  float** p2p_peer_base = nullptr;     // device array [world] of the ranks' regions
  unsigned int* p2p_ticket = nullptr;
  std::vector<void*> p2p_opened;
  uint32_t p2p_fwd_epoch = 0, p2p_rec_epoch = 0;
  int2* sym_tiles = nullptr;     // world == 1: (m_blk, n_blk) of the similarity tiles touching the upper triangle
  int n_sym_tiles = 0;
  float* part = nullptr;         // split-K partial products of the gradient GEMM
  long long part_floats = 0;
  void* row_block = nullptr;     // backing store of RowArrays
  RowArrays ra;
  BlockScalars* bs = nullptr;
  float* partial = nullptr;
  unsigned long long* ghist = nullptr;   // [2][2048] 64-bit digit counts of the GLOBAL radix select
  uint32_t* gcand = nullptr; unsigned int gcand_cap = 0;   // [2][cap] compacted candidates of the GLOBAL radix select
  float* tops_pinned = nullptr;  // host-mapped: 5 tops + err(int) + sequence number of the forward that wrote them
  unsigned int tops_seq = 0;
  float* tops_dev = nullptr;
  CUtensorMap tm_simA, tm_simB, tm_S, tm_b1A, tm_b1B, tm_b2A, tm_b2B;
  // nccl
  void* comm = nullptr; bool own_comm = false;
  // per-step state
  const float* cur_feat = nullptr; const float* cur_label = nullptr;
  const float* y_local = nullptr;   // normalize_input: this rank's normalised rows
  const float *x_total = nullptr, *lab_total = nullptr;
  bool fwd_done = false;
  cudaStream_t last_stream = nullptr;
  size_t bytes = 0;
  std::string err;
  // optional per-phase CUDA-event timing (npair_profile_enable)
  bool prof = false;
  cudaEvent_t ev[NPAIR_PROF_PHASES + 1][2];
  bool ev_made = false;
  bool ev_used[NPAIR_PROF_PHASES] = {};
};

struct PhaseTimer {
  npair_ctx* c; int ph; cudaStream_t st;
  PhaseTimer(npair_ctx* c_, int ph_, cudaStream_t st_) : c(c_), ph(ph_), st(st_) {
    if (c->prof) { cudaEventRecord(c->ev[ph][0], st); }
  }
  ~PhaseTimer() {
    if (c->prof) { cudaEventRecord(c->ev[ph][1], st); c->ev_used[ph] = true; }
  }
};

#define CUDA_TRY(ctx, call)                                                                              \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      (ctx)->err = fmt("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);      \
      return NPAIR_E_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)

static inline bool is_rel_cfg(int m) { return m == NPAIR_RELATIVE_HARD || m == NPAIR_RELATIVE_EASY; }
static int validate(const npair_config* c, std::string* err) {
  if (!c) { *err = "null config"; return NPAIR_E_ARG; }
  if (c->Q < 1 || c->D < 1) { *err = "Q and D must be >= 1"; return NPAIR_E_ARG; }
  if (c->world < 1 || c->rank < 0 || c->rank >= c->world) { *err = "bad world/rank"; return NPAIR_E_ARG; }
  if (c->num_tops < 1 || c->num_tops > 5) { *err = "num_tops must be 1..5 (npair_multi_class_loss.hpp:32-34)"; return NPAIR_E_ARG; }
  if (c->ap_region < 0 || c->ap_region > 1 || c->an_region < 0 || c->an_region > 1) { *err = "bad mining region"; return NPAIR_E_ARG; }
  if (c->ap_method < 0 || c->ap_method > 4 || c->an_method < 0 || c->an_method > 4) { *err = "bad mining method"; return NPAIR_E_ARG; }
  if (c->sim_precision < 0 || c->sim_precision > 2) { *err = "bad sim_precision"; return NPAIR_E_ARG; }
  if (c->gemm_backend < 0 || c->gemm_backend > 1) { *err = "bad gemm_backend"; return NPAIR_E_ARG; }
  if (c->bwd_exchange < 0 || c->bwd_exchange > 1) { *err = "bad bwd_exchange"; return NPAIR_E_ARG; }
  if (static_cast<long long>(c->Q) * c->world > 0x7fffffffLL) { *err = "N = Q*world exceeds int32"; return NPAIR_E_ARG; }
  if (c->global_scope < 0 || c->global_scope > 1 || c->normalize_input < 0 || c->normalize_input > 1) { *err = "global_scope / normalize_input must be 0 or 1"; return NPAIR_E_ARG; }
  if (c->grad_chunk_cols > 0 && (c->grad_chunk_cols & 31)) { *err = "grad_chunk_cols must be a multiple of 32"; return NPAIR_E_ARG; }
  if (c->global_scope && c->world > 1 && (c->bwd_exchange != NPAIR_BWD_AUTO || c->gemm_backend != NPAIR_GEMM_TCGEN05 || (c->flags & NPAIR_FLAG_NO_FUSED_GRAD))) {
    *err = "global_scope needs the row-record backward (bwd_exchange AUTO, tcgen05 backend, fused gradient kernel)"; return NPAIR_E_ARG;
  }
  return NPAIR_OK;
}

static inline bool is_rel_cfg_early(int m) { return m == NPAIR_RELATIVE_HARD || m == NPAIR_RELATIVE_EASY; }
struct Sizes { long long N, Dp, Np, Qp, ldS; int ns; size_t total; };
static Sizes sizes_of(const npair_config* c) {
  Sizes s;
  s.N = static_cast<long long>(c->Q) * c->world;
  s.Dp = round_up(c->D, 64); s.Np = round_up(s.N, 64); s.Qp = round_up(c->Q, 64); s.ldS = round_up(s.N, 32);
  s.ns = nsplit_of_prec(c->sim_precision);
  const bool tc = c->gemm_backend == NPAIR_GEMM_TCGEN05;
  const bool rs = c->world > 1 && c->bwd_exchange != NPAIR_BWD_AUTO;
  const bool fused = tc && !rs;
  size_t t = 0;
  if (c->world > 1) t += sizeof(float) * (s.N * c->D + s.N);                       // all-gather targets
  t += sizeof(float) * c->Q * s.ldS;                                                // S
  t += 2ull * s.ns * s.N * s.Dp + 2ull * s.ns * c->D * s.Np;                        // operand pieces, transposed pieces
  if (tc && c->sim_precision != NPAIR_PREC_BF16) t += 2ull * 2 * s.N * kcat_mult(c->sim_precision) * s.Dp;   // K-concatenated operands
  if (!fused) t += 2ull * s.ns * c->Q * s.Np;                                       // materialised gradient weights
  if (rs) { t += 2ull * s.ns * c->D * s.Qp + 2ull * s.ns * s.N * s.Qp + sizeof(float) * s.N * c->D; }
  if (c->world > 1 && !rs) t += sizeof(float) * 8ull * s.N;                         // gathered row records
  if (c->normalize_input) t += sizeof(float) * (2ull * c->Q * c->D + c->Q) + (c->world > 1 ? 0 : 0);   // y, dy, 1/||x||
  if (tc) {                                                                         // split-K partials of the gradient GEMM
    const int tiles = ((c->Q + 127) / 128) * ((c->D + 255) / 256);
    int smax = 148 / (tiles > 0 ? tiles : 1); if (smax > 16) smax = 16;
    if (smax > 1) t += sizeof(float) * static_cast<size_t>(smax) * c->Q * c->D;
  }
  if (tc && (is_rel_cfg_early(c->ap_method) && c->ap_region == NPAIR_GLOBAL || is_rel_cfg_early(c->an_method) && c->an_region == NPAIR_GLOBAL)) {
    unsigned long long cap = static_cast<unsigned long long>(c->Q) * s.N / 8 + 4096;           // candidate lists of the GLOBAL radix select
    if (cap > (32ull << 20)) cap = 32ull << 20;
    t += 8ull * cap;
  }
  if (c->world > 1) t += sizeof(float) * (2ull * s.N * c->D + 2ull * s.N + 16ull * s.N) + (c->global_scope ? 8ull * c->world * 8192 : 0);   // exchange region
  t += 4ull * 21 * c->Q + 65536;                                                    // row arrays, records, scalars
  s.total = t;
  return s;
}

extern "C" {

const char* npair_version(void) { return "npairloss_b200 0.1 (abi 1; sm_100a tcgen05/TMA)"; }

void npair_config_default(npair_config* c, int32_t Q, int32_t D) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->Q = Q; c->D = D; c->world = 1; c->rank = 0; c->num_tops = 5;
  c->margin_ident = 0.f; c->margin_diff = 0.f; c->identsn = -1.f; c->diffsn = -1.f;       // caffe.proto:4-7
  c->ap_region = NPAIR_LOCAL; c->ap_method = NPAIR_RAND; c->an_region = NPAIR_LOCAL; c->an_method = NPAIR_RAND;   // :19-22
  c->sim_precision = NPAIR_PREC_FP32_FP16X2; c->gemm_backend = NPAIR_GEMM_TCGEN05; c->device = -1; c->bwd_exchange = NPAIR_BWD_AUTO;
  c->global_scope = 0; c->normalize_input = 0; c->grad_chunk_cols = 0; c->flags = 0;
}

size_t npair_workspace_bytes(const npair_config* cfg) {
  std::string e;
  if (validate(cfg, &e) != NPAIR_OK) return 0;
  return sizes_of(cfg).total;
}

const char* npair_last_error(const npair_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int npair_nccl_unique_id(void* out) {
  if (!out) return NPAIR_E_ARG;
  NcclApi* api = nccl_api();
  if (!api->h || !api->err.empty()) { g_create_err = api->err; return NPAIR_E_NCCL; }
  NcclId id;
  int r = api->GetUniqueId(&id);
  if (r != 0) { g_create_err = fmt("ncclGetUniqueId: %s", api->GetErrorString(r)); return NPAIR_E_NCCL; }
  memcpy(out, &id, 128);
  return NPAIR_OK;
}

void npair_destroy(npair_ctx* c) {
  if (!c) return;
  if (c->device >= 0) cudaSetDevice(c->device);
  for (void* q : c->p2p_opened) cudaIpcCloseMemHandle(q);
  cudaFree(c->p2p_region); cudaFree(c->p2p_peer_base); cudaFree(c->p2p_ticket);
  if (c->comm && c->own_comm) release_comm(c->comm);
  cudaFree(c->Xtot_buf); cudaFree(c->labtot_buf); cudaFree(c->S); cudaFree(c->Xs); cudaFree(c->XsT); cudaFree(c->XlT);
  cudaFree(c->H); cudaFree(c->HT); cudaFree(c->OUT2); cudaFree(c->part); cudaFree(c->sym_tiles); cudaFree(c->sym_tiles2); cudaFree(c->XcatA); cudaFree(c->XcatB); cudaFree(c->rs_total); cudaFree(c->row_block); cudaFree(c->bs); cudaFree(c->partial); cudaFree(c->ghist); cudaFree(c->gcand); cudaFree(c->Ynorm); cudaFree(c->dY); cudaFree(c->inv_norm); cudaFree(c->xch_src); cudaFree(c->xch_all);
  if (c->tops_pinned) cudaFreeHost(c->tops_pinned);
  if (c->ev_made) for (int i = 0; i < NPAIR_PROF_PHASES; ++i) { cudaEventDestroy(c->ev[i][0]); cudaEventDestroy(c->ev[i][1]); }
  delete c;
}

static int create_impl(const npair_config* cfg, const void* id128, void* ext_comm, npair_ctx** out);
// One-off device check (per process, device and operand format): a 192 x 192 similarity matrix computed with EVERY tile (no mirroring)
// must come out bitwise symmetric.
static bool mma_is_symmetric(int prec, int device) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, bool> cache;
  std::lock_guard<std::mutex> lk(mu);
  const std::pair<int, int> key(device, prec);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  bool ok = false;
  const int Q = 192, D = 96;
  npair_config cfg;
  npair_config_default(&cfg, Q, D);
  cfg.sim_precision = prec; cfg.device = device; cfg.flags = NPAIR_INTERNAL_FULL_TILES; cfg.num_tops = 2;
  npair_ctx* t = nullptr;
  if (create_impl(&cfg, nullptr, nullptr, &t) == NPAIR_OK) {
    std::vector<float> x(static_cast<size_t>(Q) * D), lab(Q), S(static_cast<size_t>(Q) * t->ldS);
    uint32_t rng = 12345u;
    for (int r = 0; r < Q; ++r) {
      double nrm = 0.0;
      for (int d = 0; d < D; ++d) { rng = rng * 1664525u + 1013904223u; const float v = static_cast<float>(static_cast<int32_t>(rng >> 8) % 2001 - 1000) * 1e-3f; x[static_cast<size_t>(r) * D + d] = v; nrm += static_cast<double>(v) * v; }
      const float inv = static_cast<float>(1.0 / sqrt(nrm > 0 ? nrm : 1.0));
      for (int d = 0; d < D; ++d) x[static_cast<size_t>(r) * D + d] *= inv;
      lab[r] = static_cast<float>(r / 2);
    }
    float *dx = nullptr, *dl = nullptr;
    float tops[5];
    if (cudaMalloc(&dx, sizeof(float) * x.size()) == cudaSuccess && cudaMalloc(&dl, sizeof(float) * Q) == cudaSuccess &&
        cudaMemcpy(dx, x.data(), sizeof(float) * x.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(dl, lab.data(), sizeof(float) * Q, cudaMemcpyHostToDevice) == cudaSuccess &&
        npair_forward(t, dx, dl, tops, nullptr) == NPAIR_OK && cudaDeviceSynchronize() == cudaSuccess &&
        cudaMemcpy(S.data(), t->S, sizeof(float) * S.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
      ok = true;
      for (int i = 0; i < Q && ok; ++i)
        for (int j = 0; j < i; ++j)
          if (memcmp(&S[static_cast<size_t>(i) * t->ldS + j], &S[static_cast<size_t>(j) * t->ldS + i], 4) != 0) { ok = false; break; }
    }
    cudaFree(dx); cudaFree(dl);
    npair_destroy(t);
  }
  cudaGetLastError();
  cache[key] = ok;
  return ok;
}

static int create_impl(const npair_config* cfg, const void* id128, void* ext_comm, npair_ctx** out) {
  if (!out) { g_create_err = "null out"; return NPAIR_E_ARG; }
  *out = nullptr;
  std::string e;
  int rc = validate(cfg, &e);
  if (rc != NPAIR_OK) { g_create_err = e; return rc; }
  npair_ctx* c = new npair_ctx();
  c->cfg = *cfg;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    g_create_err = "no CUDA device: libnpair_b200 has no CPU fallback (the oracle under oracle/ is test-only)";
    delete c; return NPAIR_E_CUDA;
  }
#define CREATE_TRY(call)                                                                                  \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess) {                                                                             \
      g_create_err = fmt("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);     \
      npair_destroy(c); return NPAIR_E_CUDA;                                                              \
    }                                                                                                     \
  } while (0)
  if (cfg->device >= 0) CREATE_TRY(cudaSetDevice(cfg->device));
  CREATE_TRY(cudaGetDevice(&c->device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major != 10) {
    g_create_err = fmt("device %d is sm_%d%d; this library contains sm_100a code only", c->device, prop.major, prop.minor);
    npair_destroy(c); return NPAIR_E_CUDA;
  }
  c->sms = prop.multiProcessorCount;
  const Sizes sz = sizes_of(cfg);
  c->Q = cfg->Q; c->D = cfg->D; c->world = cfg->world; c->rank = cfg->rank; c->N = static_cast<int>(sz.N);
  c->prec = cfg->sim_precision; c->nsplit = sz.ns; c->bk_sim = bk_of(c->prec, EPI_SIM); c->bk_grad = bk_of(c->prec, EPI_OUT);
  c->Dp = sz.Dp; c->Np = sz.Np; c->Qp = sz.Qp; c->ldS = sz.ldS; c->bytes = sz.total;
  const int Q = c->Q, D = c->D, N = c->N, ns = c->nsplit;
  if (c->world > 1) {
    CREATE_TRY(cudaMalloc(&c->Xtot_buf, sizeof(float) * static_cast<size_t>(N) * D));
    CREATE_TRY(cudaMalloc(&c->labtot_buf, sizeof(float) * N));
  }
  if (cfg->normalize_input) {
    CREATE_TRY(cudaMalloc(&c->Ynorm, sizeof(float) * static_cast<size_t>(Q) * D));
    CREATE_TRY(cudaMalloc(&c->dY, sizeof(float) * static_cast<size_t>(Q) * D));
    CREATE_TRY(cudaMalloc(&c->inv_norm, sizeof(float) * Q));
  }
  CREATE_TRY(cudaMalloc(&c->S, sizeof(float) * static_cast<size_t>(Q) * c->ldS));
  CREATE_TRY(cudaMemset(c->S, 0, sizeof(float) * static_cast<size_t>(Q) * c->ldS));
  CREATE_TRY(cudaMalloc(&c->Xs, 2ull * ns * N * c->Dp));
  CREATE_TRY(cudaMemset(c->Xs, 0, 2ull * ns * N * c->Dp));
  CREATE_TRY(cudaMalloc(&c->XsT, 2ull * ns * D * c->Np));
  CREATE_TRY(cudaMemset(c->XsT, 0, 2ull * ns * D * c->Np));
  // gradient weights are only materialised when the fused tensor-memory kernel is not used
  // (reduce-scatter exchange, SIMT cross-check backend, NPAIR_NO_FUSED_GRAD)
  {
    const bool multi_rs = c->world > 1 && (cfg->bwd_exchange != NPAIR_BWD_AUTO ||
                                           (cfg->gemm_backend == NPAIR_GEMM_TCGEN05 && !mma_is_symmetric(cfg->sim_precision, c->device)));
    c->fused_grad = cfg->gemm_backend == NPAIR_GEMM_TCGEN05 && !multi_rs && !(cfg->flags & NPAIR_FLAG_NO_FUSED_GRAD);
  }
  if (!c->fused_grad) {
    CREATE_TRY(cudaMalloc(&c->H, 2ull * ns * Q * c->Np));
    CREATE_TRY(cudaMemset(c->H, 0, 2ull * ns * Q * c->Np));
  }
  // The row-record exchange needs S[j][m] on rank r to equal S[m][j] on the rank that owns row m BIT FOR BIT, i.e. a tensor-core
  // MMA whose result does not change when the operand roles are swapped.  Measured true on B200 for every operand format; checked
  // once per process and format on this device -- if it ever fails, the reference's reduce-scatter form is used instead.
  const bool sym_ok = c->world == 1 || cfg->bwd_exchange != NPAIR_BWD_AUTO || cfg->gemm_backend != NPAIR_GEMM_TCGEN05 || mma_is_symmetric(cfg->sim_precision, c->device);
  c->bwd_mode = c->world == 1 ? NPAIR_BWDMODE_SINGLE
              : ((cfg->bwd_exchange == NPAIR_BWD_AUTO && sym_ok) ? NPAIR_BWDMODE_ROW_SCALARS : NPAIR_BWDMODE_REDUCE_SCATTER);
  if (c->bwd_mode == NPAIR_BWDMODE_ROW_SCALARS) CREATE_TRY(cudaMalloc(&c->rs_total, sizeof(float) * 8ull * N));
  if (c->prec != PREC_BF16 && cfg->gemm_backend == NPAIR_GEMM_TCGEN05) {
    const size_t cat_bytes = 2ull * N * kcat_mult(c->prec) * c->Dp;
    CREATE_TRY(cudaMalloc(&c->XcatA, cat_bytes));
    CREATE_TRY(cudaMemset(c->XcatA, 0, cat_bytes));
    CREATE_TRY(cudaMalloc(&c->XcatB, cat_bytes));
    CREATE_TRY(cudaMemset(c->XcatB, 0, cat_bytes));
  }
  if (c->bwd_mode == NPAIR_BWDMODE_REDUCE_SCATTER) {
    CREATE_TRY(cudaMalloc(&c->XlT, 2ull * ns * D * c->Qp));
    CREATE_TRY(cudaMemset(c->XlT, 0, 2ull * ns * D * c->Qp));
    CREATE_TRY(cudaMalloc(&c->HT, 2ull * ns * N * c->Qp));
    CREATE_TRY(cudaMemset(c->HT, 0, 2ull * ns * N * c->Qp));
    CREATE_TRY(cudaMalloc(&c->OUT2, sizeof(float) * static_cast<size_t>(N) * D));
  }
  {
    // split-K workspace for the gradient GEMM: at most (SMs / tiles) partial Q x D products, capped at 16
    const int tiles = ((Q + 127) / 128) * ((D + 255) / 256);
    int smax = c->sms / (tiles > 0 ? tiles : 1); if (smax > 16) smax = 16;
    if (smax > 1) {
      c->part_floats = static_cast<long long>(smax) * Q * D;
      CREATE_TRY(cudaMalloc(&c->part, sizeof(float) * c->part_floats));
    }
  }
  // row arrays: 5 uint32/int stats, 2 thr, 3 fwd, 3 hits = 13 arrays of Q 4-byte words + the [Q][8] row records
  CREATE_TRY(cudaMalloc(&c->row_block, 4ull * 21 * Q + 64));
  CREATE_TRY(cudaMemset(c->row_block, 0, 4ull * 21 * Q + 64));
  {
    uint32_t* w = static_cast<uint32_t*>(c->row_block);
    RowArrays& ra = c->ra;
    ra.st_minw = w; w += Q; ra.st_maxw = w; w += Q; ra.st_maxb = w; w += Q; ra.st_maxall = w; w += Q;
    ra.cnt_same = reinterpret_cast<int*>(w); w += Q;
    ra.posi_thr = reinterpret_cast<float*>(w); w += Q; ra.nega_thr = reinterpret_cast<float*>(w); w += Q;
    ra.A = reinterpret_cast<float*>(w); w += Q; ra.T = reinterpret_cast<float*>(w); w += Q; ra.logv = reinterpret_cast<float*>(w); w += Q;
    ra.hits = reinterpret_cast<int*>(w); w += 3 * Q;
    w = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(w) + 31) & ~static_cast<uintptr_t>(31));
    ra.rowscal = reinterpret_cast<float*>(w); w += 8ll * Q;
  }
  CREATE_TRY(cudaMalloc(&c->bs, sizeof(BlockScalars)));
  CREATE_TRY(cudaMemset(c->bs, 0, sizeof(BlockScalars)));
  CREATE_TRY(cudaMalloc(&c->partial, sizeof(float) * 2048));
  CREATE_TRY(cudaMalloc(&c->ghist, sizeof(unsigned long long) * 4096));
  CREATE_TRY(cudaMemset(c->ghist, 0, sizeof(unsigned long long) * 4096));
  {
    // GLOBAL relative select with a general SN: candidate lists of the chosen first-digit bucket (1/8 of the block, at most
    // 32 M entries per side; a bigger bucket -- heavily tied data -- takes the three-sweep path)
    const bool need = (is_rel_cfg(cfg->ap_method) && cfg->ap_region == NPAIR_GLOBAL) || (is_rel_cfg(cfg->an_method) && cfg->an_region == NPAIR_GLOBAL);
    if (need) {
      unsigned long long cap = static_cast<unsigned long long>(Q) * N / 8 + 4096;
      if (cap > (32ull << 20)) cap = 32ull << 20;
      c->gcand_cap = static_cast<unsigned int>(cap);
      CREATE_TRY(cudaMalloc(&c->gcand, sizeof(uint32_t) * 2ull * cap));
    }
  }
  CREATE_TRY(cudaHostAlloc(&c->tops_pinned, 64, cudaHostAllocMapped));
  memset(c->tops_pinned, 0, 64);
  CREATE_TRY(cudaHostGetDevicePointer(&c->tops_dev, c->tops_pinned, 0));
  // ---- TMA tensor maps (K-major boxes of one swizzle span) ----
  if (cfg->gemm_backend == NPAIR_GEMM_TCGEN05) {
    std::string te;
    const int bks = c->bk_sim, bkg = c->bk_grad;
    bool ok = true;
    // similarity: A = local rows of Xs, B = all rows of Xs; K = D
    ok = ok && make_tmap_pieces(&c->tm_simA, c->Xs + static_cast<long long>(c->rank) * Q * c->Dp, D, Q, ns, c->Dp, static_cast<long long>(N) * c->Dp, bks, 128, &te);
    ok = ok && make_tmap_pieces(&c->tm_simB, c->Xs, D, N, ns, c->Dp, static_cast<long long>(N) * c->Dp, bks, 256, &te);
    ok = ok && make_tmap_pieces(&c->tm_simB2, c->Xs, D, N, ns, c->Dp, static_cast<long long>(N) * c->Dp, bks, 128, &te);
    ok = ok && make_tmap_f32_store(&c->tm_S, c->S, N, Q, c->ldS, &te);
    // gradient 1: A = H [Q x N], B = XsT [D x N]; K = N
    if (c->H) ok = ok && make_tmap_pieces(&c->tm_b1A, c->H, N, Q, ns, c->Np, static_cast<long long>(Q) * c->Np, bkg, 128, &te);
    ok = ok && make_tmap_pieces(&c->tm_b1B, c->XsT, N, D, ns, c->Np, static_cast<long long>(D) * c->Np, bkg, 256, &te);
    if (c->fused_grad) {
      ok = ok && make_tmap_pieces(&c->tm_fB, c->XsT, N, D, ns, c->Np, static_cast<long long>(D) * c->Np, 32, 256, &te);
      ok = ok && make_tmap_pieces(&c->tm_fB2, c->XsT, N, D, ns, c->Np, static_cast<long long>(D) * c->Np, 32, 128, &te);
      ok = ok && make_tmap_f32_store(&c->tm_fS, c->S, N, Q, c->ldS, &te, 128);
    }
    if (c->XcatA) {       // bitwise-symmetric similarity: one pass over K_cat = 3*Dp (fp16x2) / 6*Dp (bf16x3)
      const long long kc = kcat_mult(c->prec) * c->Dp;
      ok = ok && make_tmap_pieces(&c->tm_catA, c->XcatA + static_cast<long long>(c->rank) * Q * kc, static_cast<int>(kc), Q, 1, kc, static_cast<long long>(N) * kc, 64, 128, &te);
      ok = ok && make_tmap_pieces(&c->tm_catB, c->XcatB, static_cast<int>(kc), N, 1, kc, static_cast<long long>(N) * kc, 64, 256, &te);
      ok = ok && make_tmap_pieces(&c->tm_catB2, c->XcatB, static_cast<int>(kc), N, 1, kc, static_cast<long long>(N) * kc, 64, 128, &te);
    }
    if (c->bwd_mode == NPAIR_BWDMODE_REDUCE_SCATTER) {   // gradient 2: A = HT [N x Q], B = XlT [D x Q]; K = Q
      ok = ok && make_tmap_pieces(&c->tm_b2A, c->HT, Q, N, ns, c->Qp, static_cast<long long>(N) * c->Qp, bkg, 128, &te);
      ok = ok && make_tmap_pieces(&c->tm_b2B, c->XlT, Q, D, ns, c->Qp, static_cast<long long>(D) * c->Qp, bkg, 256, &te);
    }
    if (!ok) { g_create_err = te; npair_destroy(c); return NPAIR_E_CUDA; }
  }
  if (c->world == 1 && cfg->gemm_backend == NPAIR_GEMM_TCGEN05 && !(cfg->flags & NPAIR_INTERNAL_FULL_TILES)) {
    // S = X X^T is symmetric: only tiles (m_blk, n_blk) whose 256 columns reach the 128-row block's diagonal or beyond
    std::vector<int2> tl;
    const int tm = (Q + 127) / 128, tn = (N + 255) / 256;
    for (int mb = 0; mb < tm; ++mb)
      for (int nb = mb / 2; nb < tn; ++nb) tl.push_back(make_int2(mb, nb));
    c->n_sym_tiles = static_cast<int>(tl.size());
    CREATE_TRY(cudaMalloc(&c->sym_tiles, sizeof(int2) * tl.size()));
    CREATE_TRY(cudaMemcpy(c->sym_tiles, tl.data(), sizeof(int2) * tl.size(), cudaMemcpyHostToDevice));
    // pair kernel: 256-row block pm covers m_blk 2pm, 2pm+1 -> the same set of 128 x 256 tiles (n_blk >= pm)
    std::vector<int2> tl2;
    for (int pm = 0; pm < (tm + 1) / 2; ++pm)
      for (int nb = pm; nb < tn; ++nb) tl2.push_back(make_int2(pm, nb));
    c->n_sym_tiles2 = static_cast<int>(tl2.size());
    CREATE_TRY(cudaMalloc(&c->sym_tiles2, sizeof(int2) * tl2.size()));
    CREATE_TRY(cudaMemcpy(c->sym_tiles2, tl2.data(), sizeof(int2) * tl2.size(), cudaMemcpyHostToDevice));
  }
  {
    // CTA-pair similarity GEMM whenever there are at least two 128-row blocks (NPAIR_SIM_1CTA=1 keeps the single-CTA kernel)
    c->sim_pair = cfg->gemm_backend == NPAIR_GEMM_TCGEN05 && Q > 128 && !(cfg->flags & NPAIR_FLAG_SIM_1CTA);
    c->grad_pair = c->fused_grad && Q > 128 && !(cfg->flags & NPAIR_FLAG_GRAD_1CTA);
    c->grad_chunk_kb = cfg->grad_chunk_cols > 0 ? (cfg->grad_chunk_cols + 31) / 32 : 64;     // default: 2048 database columns
    if (cfg->grad_chunk_cols < 0) c->grad_chunk_kb = 0;                                        // negative: one accumulator for the whole K range (diagnostic)
  }
  // ---- NCCL ----
  if (c->world > 1 && (id128 || ext_comm)) {
    NcclApi* api = nccl_api();
    if (!api->h || !api->err.empty()) { g_create_err = api->err; npair_destroy(c); return NPAIR_E_NCCL; }
    if (ext_comm) { c->comm = ext_comm; c->own_comm = false; }
    else {
      std::string ce;
      if (acquire_comm(id128, c->world, c->rank, &c->comm, &ce) != 0) { g_create_err = ce; c->comm = nullptr; npair_destroy(c); return NPAIR_E_NCCL; }
      c->own_comm = true;
    }
  }
  if (c->comm && c->world > 1 && c->world <= 32 && !getenv("NPAIR_NO_P2P")) {
    const bool want_feat = !(cfg->flags & NPAIR_FLAG_NCCL_FEATURES);
    const bool want_rec = !(cfg->flags & NPAIR_FLAG_NCCL_RECORDS) && c->bwd_mode == NPAIR_BWDMODE_ROW_SCALARS;
    if (want_feat || want_rec) {
      // one exported region per rank; handles travel over the NCCL communicator once (a 64-byte all-gather)
      NcclApi* api = nccl_api();
      const int W = c->world;
      const long long nX = static_cast<long long>(N) * D, nL = round_up(N, 4), nR = 8ll * N;
      c->p2p_offX = 0; c->p2p_offLab = 2 * nX; c->p2p_offRec = c->p2p_offLab + 2 * nL; c->p2p_offXch = c->p2p_offRec + 2 * nR;
      c->p2p_offFlags = c->p2p_offXch + (cfg->global_scope ? 2ll * W * NPAIR_XCH_FLOATS : 0);
      const long long total = c->p2p_offFlags + round_up(6ll * W, 4);
      CREATE_TRY(cudaMalloc(&c->p2p_region, sizeof(float) * static_cast<size_t>(total)));
      CREATE_TRY(cudaMemset(c->p2p_region, 0, sizeof(float) * static_cast<size_t>(total)));
      CREATE_TRY(cudaMalloc(&c->p2p_ticket, sizeof(unsigned int)));
      CREATE_TRY(cudaMemset(c->p2p_ticket, 0, sizeof(unsigned int)));
      cudaIpcMemHandle_t mine;
      static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
      CREATE_TRY(cudaIpcGetMemHandle(&mine, c->p2p_region));
      float *d_mine = nullptr, *d_all = nullptr;
      CREATE_TRY(cudaMalloc(&d_mine, 64));
      CREATE_TRY(cudaMalloc(&d_all, 64ull * W));
      CREATE_TRY(cudaMemcpy(d_mine, &mine, 64, cudaMemcpyHostToDevice));
      CREATE_TRY(cudaDeviceSynchronize());                       // the memset above has landed before any peer can write into the region
      int r = api->AllGather(d_mine, d_all, 16, NCCL_FLOAT32, c->comm, nullptr);
      if (r != 0) { g_create_err = fmt("ncclAllGather(ipc handles): %s", api->GetErrorString(r)); cudaFree(d_mine); cudaFree(d_all); npair_destroy(c); return NPAIR_E_NCCL; }
      CREATE_TRY(cudaStreamSynchronize(nullptr));
      std::vector<cudaIpcMemHandle_t> all(W);
      CREATE_TRY(cudaMemcpy(all.data(), d_all, 64ull * W, cudaMemcpyDeviceToHost));
      cudaFree(d_mine); cudaFree(d_all);
      std::vector<float*> pb(W);
      bool mapped = true;
      for (int q = 0; q < W && mapped; ++q) {
        if (q == c->rank) { pb[q] = c->p2p_region; continue; }
        void* a = nullptr;
        if (cudaIpcOpenMemHandle(&a, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = false; break; }
        c->p2p_opened.push_back(a);
        pb[q] = static_cast<float*>(a);
      }
      if (mapped) {
        CREATE_TRY(cudaMalloc(&c->p2p_peer_base, sizeof(float*) * W));
        CREATE_TRY(cudaMemcpy(c->p2p_peer_base, pb.data(), sizeof(float*) * W, cudaMemcpyHostToDevice));
        c->p2p_feat = want_feat; c->p2p_rec = want_rec;
      }
      // (no peer access between some pair of GPUs: the NCCL paths are used; every rank takes the same decision only if the
      // topology is symmetric, which holds on an NVSwitch box -- a mixed outcome is reported by the first exchange's timeout)
    }
  }
  if (cfg->global_scope && c->world > 1) {
    if (!c->comm) { g_create_err = "global_scope with world > 1 needs a communicator (the world-scope reductions are internal)"; npair_destroy(c); return NPAIR_E_ARG; }
    CREATE_TRY(cudaMalloc(&c->xch_src, sizeof(float) * NPAIR_XCH_FLOATS));
    if (!c->p2p_region) CREATE_TRY(cudaMalloc(&c->xch_all, sizeof(float) * static_cast<size_t>(NPAIR_XCH_FLOATS) * c->world));
  }
#undef CREATE_TRY
  *out = c;
  return NPAIR_OK;
}

int npair_create(const npair_config* cfg, const void* id128, npair_ctx** out) { return create_impl(cfg, id128, nullptr, out); }
int npair_create_with_comm(const npair_config* cfg, void* comm, npair_ctx** out) {
  if (cfg && cfg->world > 1 && !comm) { g_create_err = "null communicator"; return NPAIR_E_ARG; }
  return create_impl(cfg, nullptr, comm, out);
}

// World-scope mode: every rank contributes `n` floats (in c->xch_src, or `src` copied there) and gets the world's contributions as
// [world][NPAIR_XCH_FLOATS]; all ranks then reduce them in rank order, so decisions are identical everywhere.
static int xchg_small(npair_ctx* c, const float* src, int n, const float** all, cudaStream_t st) {
  if (src != c->xch_src) CUDA_TRY(c, cudaMemcpyAsync(c->xch_src, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (c->p2p_region) {
    const uint32_t ep = ++c->xch_epoch;
    const long long par = ep & 1u;
    const long long off = c->p2p_offXch + (par * c->world + c->rank) * NPAIR_XCH_FLOATS;
    int nb = (n / 4 + 255) / 256; if (nb > 8) nb = 8; if (nb < 1) nb = 1;
    p2p_push_kernel<<<nb, 256, 0, st>>>(c->xch_src, n, off, nullptr, 0, 0, c->p2p_peer_base, c->p2p_offFlags,
                                        4 * c->world + static_cast<int>(par) * c->world + c->rank, c->world, ep, c->p2p_ticket);
    count_launch();
    p2p_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(c->p2p_region + c->p2p_offFlags) + 4 * c->world + par * c->world, c->world, ep);
    count_launch();
    *all = c->p2p_region + c->p2p_offXch + par * c->world * NPAIR_XCH_FLOATS;
  } else {
    NcclApi* api = nccl_api();
    int r = api->AllGather(c->xch_src, c->xch_all, NPAIR_XCH_FLOATS, NCCL_FLOAT32, c->comm, st);
    if (r != 0) { c->err = fmt("ncclAllGather(world-scope reduction): %s", api->GetErrorString(r)); return NPAIR_E_NCCL; }
    *all = c->xch_all;
  }
  return NPAIR_OK;
}

static MiningParams mining_of(const npair_config& c) {
  MiningParams mp;
  mp.ap_region = c.ap_region; mp.ap_method = c.ap_method; mp.an_region = c.an_region; mp.an_method = c.an_method;
  mp.margin_ident = c.margin_ident; mp.margin_diff = c.margin_diff; mp.identsn = c.identsn; mp.diffsn = c.diffsn;
  return mp;
}
static inline bool is_rel_m(int m) { return m == NPAIR_RELATIVE_HARD || m == NPAIR_RELATIVE_EASY; }
static inline bool sn_max(float sn) { return sn >= 0.f && static_cast<int>(sn) == 0; }

// The reference blocks after its forward (host reads of loss / asum, .cu:384,400).  The five tops land in mapped pinned memory followed by
// this forward's sequence number: polling that word returns a few microseconds earlier than a stream synchronisation and does not
// wait for anything enqueued behind the row pass (the row-record push, a backward).  A fault in a kernel never writes the number:
// after ~2 s fall back to the synchronisation, which reports the error.
static int wait_tops(npair_ctx* c, cudaStream_t st) {
  volatile unsigned int* seqp = reinterpret_cast<volatile unsigned int*>(c->tops_pinned) + 6;
  unsigned long long spins = 0;
  while (*seqp != c->tops_seq) {
    if (++spins > (1ull << 28)) { CUDA_TRY(c, cudaStreamSynchronize(st)); if (*seqp != c->tops_seq) { c->err = "the forward kernels finished without publishing their results"; return NPAIR_E_CUDA; } break; }
    __builtin_ia32_pause();
    if ((spins & 0x3FFull) == 0) sched_yield();   // ranks that share a core (fewer cores than ranks, an inherited binding) take turns quickly
  }
  __sync_synchronize();
  return NPAIR_OK;
}

static int forward_impl(npair_ctx* c, const float* d_feat, const float* d_label, float tops_host[5], cudaStream_t st);

int npair_forward(npair_ctx* c, const float* d_feat, const float* d_label, float tops_host[5], void* stream) {
  if (!c) return NPAIR_E_ARG;
  if (!d_feat || !d_label || !tops_host) { c->err = "null pointer argument"; return NPAIR_E_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->fwd_done = false; c->last_stream = st; c->ext_gathered = false;
  const int Q = c->Q, D = c->D;
  if (c->cfg.normalize_input) {               // fused L2Normalize producer (usage/def.prototxt:115-120): the layer works on x / ||x||
    PhaseTimer pt(c, 1, st);
    launch_l2norm_fwd(d_feat, Q, D, c->Ynorm, c->inv_norm, st);
    d_feat = c->Ynorm; c->y_local = c->Ynorm;
  }
  // ---- GatherFeatureAndLabel (.cu:17-43): one NCCL group, device to device over NVLink ----
  if (c->world > 1 && c->p2p_feat) {
    PhaseTimer pt(c, 0, st);
    const uint32_t ep = ++c->p2p_fwd_epoch;
    const long long par = ep & 1u, N = c->N;
    const long long offX = c->p2p_offX + par * N * D + static_cast<long long>(c->rank) * Q * D;
    const long long offL = c->p2p_offLab + par * round_up(N, 4) + static_cast<long long>(c->rank) * Q;
    int nb = static_cast<int>((static_cast<long long>(Q) * D / 4 + 255) / 256); if (nb > 2 * c->sms) nb = 2 * c->sms; if (nb < 1) nb = 1;
    p2p_push_kernel<<<nb, 256, 0, st>>>(d_feat, static_cast<long long>(Q) * D, offX, d_label, Q, offL, c->p2p_peer_base, c->p2p_offFlags,
                                        static_cast<int>(par) * c->world + c->rank, c->world, ep, c->p2p_ticket);
    count_launch();
    p2p_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(c->p2p_region + c->p2p_offFlags) + par * c->world, c->world, ep);
    count_launch();
    c->x_total = c->p2p_region + c->p2p_offX + par * N * D;
    c->lab_total = c->p2p_region + c->p2p_offLab + par * round_up(N, 4);
  } else if (c->world > 1) {
    if (!c->comm) { c->err = "context was created without a communicator: use npair_forward_gathered"; return NPAIR_E_STATE; }
    PhaseTimer pt(c, 0, st);
    NcclApi* api = nccl_api();
    int r = api->GroupStart();
    if (r == 0) r = api->AllGather(d_feat, c->Xtot_buf, static_cast<size_t>(Q) * D, NCCL_FLOAT32, c->comm, st);
    if (r == 0) r = api->AllGather(d_label, c->labtot_buf, static_cast<size_t>(Q), NCCL_FLOAT32, c->comm, st);
    int r2 = api->GroupEnd();
    if (r == 0) r = r2;
    if (r != 0) { c->err = fmt("ncclAllGather: %s", api->GetErrorString(r)); return NPAIR_E_NCCL; }
    c->x_total = c->Xtot_buf; c->lab_total = c->labtot_buf;
  } else { c->x_total = d_feat; c->lab_total = d_label; }
  return forward_impl(c, d_feat, d_label, tops_host, st);
}

/* External-collectives variant: the caller already holds the all-gathered N x D features and N labels (rank r's rows are
 * [r*Q,(r+1)*Q)).  Lets a host framework keep its own communication layer, and lets tests emulate every rank on one GPU. */
int npair_forward_gathered(npair_ctx* c, const float* d_feat_total, const float* d_label_total, float tops_host[5], void* stream) {
  if (!c) return NPAIR_E_ARG;
  if (!d_feat_total || !d_label_total || !tops_host) { c->err = "null pointer argument"; return NPAIR_E_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->fwd_done = false; c->last_stream = st; c->ext_gathered = true;
  if (c->cfg.normalize_input) {               // the gathered bottoms are raw embeddings: normalise all N rows (1 / ||x|| kept for the local ones)
    PhaseTimer pt(c, 1, st);
    float* dst = c->world > 1 ? c->Xtot_buf : c->Ynorm;
    launch_l2norm_fwd(d_feat_total, c->N, c->D, dst, nullptr, st);
    launch_l2norm_fwd(d_feat_total + static_cast<long long>(c->rank) * c->Q * c->D, c->Q, c->D, c->Ynorm, c->inv_norm, st);
    c->y_local = c->Ynorm;
    d_feat_total = dst;
  }
  c->x_total = d_feat_total; c->lab_total = d_label_total;
  return forward_impl(c, d_feat_total + static_cast<long long>(c->rank) * c->Q * c->D, d_label_total + static_cast<long long>(c->rank) * c->Q, tops_host, st);
}

static int forward_impl(npair_ctx* c, const float* d_feat, const float* d_label, float tops_host[5], cudaStream_t st) {
  const int Q = c->Q, N = c->N, D = c->D;
  const MiningParams mp = mining_of(c->cfg);
  c->cur_feat = d_feat; c->cur_label = d_label;
  const int self_off = c->rank * Q;
  // ---- operand preparation: |x| sum (top asum, .cu:400), power-of-two pre-scale, split to tensor-core pieces ----
  {
    PhaseTimer pt(c, 1, st);
    // Xs (the un-concatenated K-major pieces) is only read by the single-pass bf16 similarity GEMM and the SIMT backend
    uint16_t* xs_dst = (c->XcatA && c->cfg.gemm_backend == NPAIR_GEMM_TCGEN05) ? nullptr : c->Xs;
    launch_prep_reduce(d_feat, static_cast<long long>(Q) * D, c->x_total, static_cast<long long>(N) * D, c->partial,
                       c->prec == PREC_FP16X2 ? 1 : 0, c->ra, Q, c->bs, st);
    launch_split(c->x_total, N, D, c->prec, c->bs, xs_dst, c->Dp, c->XsT, c->Np, c->XlT, c->Qp, self_off, Q, c->XcatA, c->XcatB, c->Dp, st);
  }
  // ---- S = X_local . X_total^T (.cu:218) with fused masks + row statistics (.cu:44-66, :225-265) ----
  GemmParams gp; memset(&gp, 0, sizeof(gp));
  gp.M = Q; gp.Nn = N; gp.num_kblocks = static_cast<int>((D + c->bk_sim - 1) / c->bk_sim);
  gp.tiles_m = (Q + 127) / 128; gp.tiles_n = (N + 255) / 256; gp.splits = 1; gp.kb_per_split = gp.num_kblocks;
  gp.S = c->S; gp.ldS = c->ldS; gp.dev_scale = &c->bs->x_inv_scale;
  gp.lab_rows = d_label; gp.lab_cols = c->lab_total; gp.self_offset = self_off;
  gp.st_minw = c->ra.st_minw; gp.st_maxw = c->ra.st_maxw; gp.st_maxb = c->ra.st_maxb; gp.st_maxall = c->ra.st_maxall; gp.cnt_same = c->ra.cnt_same;
  // the threshold pick rides in the similarity kernel's last CTA unless its result has to be exchanged first (world scope)
  const bool fuse_thr = c->cfg.gemm_backend == NPAIR_GEMM_TCGEN05 && !(c->cfg.global_scope && c->world > 1);
  gp.fuse_thr = fuse_thr ? 1 : 0; gp.ra = c->ra; gp.mp = mp; gp.bs = c->bs;
  if (c->cfg.gemm_backend == NPAIR_GEMM_TCGEN05) {
    PhaseTimer pt(c, 2, st);
    if (c->sym_tiles) { gp.tile_list = c->sym_tiles; gp.num_tiles_list = c->n_sym_tiles; }
    if (c->XcatA) { gp.num_kblocks = static_cast<int>(kcat_mult(c->prec) * c->Dp / 64); gp.kb_per_split = gp.num_kblocks; }
    if (c->sim_pair) {
      gp.tiles_m = (gp.tiles_m + 1) / 2;
      if (c->sym_tiles2) { gp.tile_list = c->sym_tiles2; gp.num_tiles_list = c->n_sym_tiles2; }
      CUDA_TRY(c, launch_sim_gemm_pair(c->prec, c->sym_tiles != nullptr, c->XcatA ? c->tm_catA : c->tm_simA, c->XcatA ? c->tm_catB2 : c->tm_simB2,
                                       c->tm_S, gp, c->sms, st));
    } else if (c->XcatA) {
      CUDA_TRY(c, launch_sim_gemm(c->prec, c->sym_tiles != nullptr, c->tm_catA, c->tm_catB, c->tm_S, gp, c->sms, st));
    } else
      CUDA_TRY(c, launch_sim_gemm(c->prec, c->sym_tiles != nullptr, c->tm_simA, c->tm_simB, c->tm_S, gp, c->sms, st));
  } else {
    CUDA_TRY(c, launch_simt_gemm(c->prec, EPI_SIM, c->Xs + static_cast<long long>(self_off) * c->Dp, c->Dp, static_cast<long long>(N) * c->Dp,
                                 c->Xs, c->Dp, static_cast<long long>(N) * c->Dp, D, gp, st));
    launch_row_stats_ref(c->S, c->ldS, Q, N, d_label, c->lab_total, self_off, c->ra, st);
  }
  // ---- thresholds (.cu:275-337) ----
  {
  PhaseTimer pt(c, 3, st);
  const bool wscope = c->cfg.global_scope && c->world > 1;      // world == 1: the block IS the world
  if (!fuse_thr) launch_thresholds(c->ra, Q, N, mp, c->bs, c->partial, wscope ? c->xch_src : nullptr, st);
  if (wscope) {
    const float* all = nullptr;
    const int rc = xchg_small(c, c->xch_src, 8, &all, st);
    if (rc != NPAIR_OK) return rc;
    launch_thresholds_world(all, NPAIR_XCH_FLOATS, c->world, N, mp, c->bs, st);
  }
  {
    // general relative SN: radix selects; both sides of a region share one sweep of S
    int local_mask = 0, global_mask = 0;
    if (is_rel_m(mp.ap_method) && !sn_max(mp.identsn)) (mp.ap_region == NPAIR_LOCAL ? local_mask : global_mask) |= 1;
    if (is_rel_m(mp.an_method) && !sn_max(mp.diffsn)) (mp.an_region == NPAIR_LOCAL ? local_mask : global_mask) |= 2;
    if (global_mask) {
      for (int pass = 0; pass < 3; ++pass) {
        launch_global_select_pass(c->S, c->ldS, Q, N, d_label, c->lab_total, self_off, global_mask, pass, c->ra, c->ghist, c->gcand, c->gcand_cap,
                                  wscope ? 1 : 0, c->bs, c->sms, st);
        if (wscope) {
          const float* all = nullptr;
          const int rc = xchg_small(c, reinterpret_cast<const float*>(c->ghist), NPAIR_XCH_FLOATS, &all, st);
          if (rc != NPAIR_OK) return rc;
          launch_global_decide(all, NPAIR_XCH_FLOATS, c->world, global_mask, pass, c->ra, Q, c->ghist, c->gcand, c->gcand_cap, c->bs, st);
        }
      }
    }
    if (local_mask) launch_local_select(c->S, c->ldS, Q, N, d_label, c->lab_total, self_off, local_mask, mp.identsn, mp.diffsn, c->ra, c->bs, c->sms, (c->cfg.flags & NPAIR_FLAG_LSEL_WARP) != 0, st);
  }
  }
  // ---- selection + counts + exp + masked sums + log + retrieval in one pass (.cu:343-398) ----
  {
    PhaseTimer pt(c, 4, st);
    const bool wscope = c->cfg.global_scope && c->world > 1;
    ++c->tops_seq;
    launch_lse_rows(c->S, c->ldS, Q, N, d_label, c->lab_total, self_off, mp, c->ra, c->bs, c->cfg.num_tops, c->tops_dev, c->world,
                    wscope ? c->xch_src : nullptr, c->tops_seq, st);
    if (wscope) {       // loss / retrieval / asum over the world's N rows, identical on every rank (the reference's are per rank, .cu:385)
      const float* all = nullptr;
      const int rc = xchg_small(c, c->xch_src, 8, &all, st);
      if (rc != NPAIR_OK) return rc;
      launch_tops_world(all, NPAIR_XCH_FLOATS, c->world, N, c->cfg.num_tops, c->tops_dev, c->tops_seq, st);
    }
  }
  c->rs_gathered = false;
  // NPAIR_RS_GATHER_FWD=1 enqueues the row-record exchange here instead of at the start of npair_backward.  Measured on
  // 8 x B200 (profiles/r01_bench_v6_n8.json): with the collective in front of the forward's host synchronisation the step is
  // SLOWER (0.293 ms against 0.218 ms with it in the backward, where its rendezvous overlaps the host's return), so it is opt-in.
  if (c->p2p_rec && c->comm && c->x_total != nullptr && !c->ext_gathered) {
    // peer-memory exchange: push this rank's 32-byte row records to every rank now; the backward only waits for the flags
    PhaseTimer pt(c, 8, st);
    const uint32_t ep = ++c->p2p_rec_epoch;
    const long long par = ep & 1u;
    const long long offR = c->p2p_offRec + par * 8ll * N + 8ll * c->rank * Q;
    int nb = (2 * Q + 255) / 256; if (nb > 64) nb = 64; if (nb < 1) nb = 1;
    p2p_push_kernel<<<nb, 256, 0, st>>>(c->ra.rowscal, 8ll * Q, offR, nullptr, 0, 0, c->p2p_peer_base, c->p2p_offFlags,
                                        2 * c->world + static_cast<int>(par) * c->world + c->rank, c->world, ep, c->p2p_ticket);
    count_launch();
  }
  const bool gather_in_fwd = false;      // measured slower on 8 GPUs (profiles/r01_bench_v6_n8.json): the gather stays in the backward
  if (gather_in_fwd && !c->p2p_rec && c->bwd_mode == NPAIR_BWDMODE_ROW_SCALARS && c->comm) {
    // The only backward exchange (8*Q floats per rank, replaces the N x D MPI_Allreduce of .cu:462-489) does not depend on
    // the loss weight, so it is enqueued here: it runs while the host wakes up from the synchronisation below.
    PhaseTimer pt(c, 8, st);
    NcclApi* api = nccl_api();
    int r = api->AllGather(c->ra.rowscal, c->rs_total, 8ull * Q, NCCL_FLOAT32, c->comm, st);
    if (r != 0) { c->err = fmt("ncclAllGather(row records): %s", api->GetErrorString(r)); return NPAIR_E_NCCL; }
    c->rs_gathered = true;
  }
  CUDA_TRY(c, cudaGetLastError());
  if (c->defer_sync) return NPAIR_OK;              // npair_forward_backward enqueues the backward first, then waits once
  { const int wrc = wait_tops(c, st); if (wrc != NPAIR_OK) return wrc; }
  const int derr = reinterpret_cast<int*>(c->tops_pinned)[5];
  if (derr & DERR_EMPTY_LIST) { c->err = "an empty same/diff list was indexed (undefined behaviour in the reference, .cu:296/:327/:288)"; return NPAIR_E_EMPTY_LIST; }
  if (derr & DERR_POS_RANGE) { c->err = "identsn/diffsn select a position outside the list (undefined behaviour in the reference, .cu:285-288)"; return NPAIR_E_POS_RANGE; }
  for (int t = 0; t < 5; ++t) tops_host[t] = t < c->cfg.num_tops ? c->tops_pinned[t] : 0.f;
  c->fwd_done = true;
  return NPAIR_OK;
}

static int backward_core(npair_ctx* c, float loss_weight, float* d_diff, float* d_total_ext, const float* d_rs_ext, cudaStream_t st);
// Backward_gpu (+ the projection of the fused L2Normalize producer: the kernels produce d loss / d y, the caller gets d loss / d x)
static int backward_impl(npair_ctx* c, float loss_weight, float* d_diff, float* d_total_ext, const float* d_rs_ext, cudaStream_t st) {
  if (!c->cfg.normalize_input) return backward_core(c, loss_weight, d_diff, d_total_ext, d_rs_ext, st);
  if (d_total_ext) { c->err = "normalize_input: the partial (pre-all-reduce) backward is not available, its sum over ranks would have to be projected"; return NPAIR_E_STATE; }
  const int rc = backward_core(c, loss_weight, c->dY, nullptr, d_rs_ext, st);
  if (rc != NPAIR_OK) return rc;
  PhaseTimer pt(c, 5, st);
  launch_l2norm_bwd(c->y_local, c->inv_norm, c->dY, c->Q, c->D, d_diff, st);
  return NPAIR_OK;
}

int npair_bwd_exchange_mode(const npair_ctx* c) { return c ? c->bwd_mode : NPAIR_E_ARG; }

int npair_backward(npair_ctx* c, float loss_weight, float* d_diff, void* stream) {
  if (!c) return NPAIR_E_ARG;
  if (!d_diff) { c->err = "null gradient pointer"; return NPAIR_E_ARG; }
  if (!c->fwd_done) { c->err = "npair_backward called without a successful npair_forward"; return NPAIR_E_STATE; }
  if (c->world > 1 && !c->comm) { c->err = "context was created without a communicator: use npair_backward_partial / npair_backward_gathered"; return NPAIR_E_STATE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->last_stream = st;
  return backward_impl(c, loss_weight, d_diff, nullptr, nullptr, st);
}

/* Forward + backward with ONE host synchronisation: the backward (whose loss weight is a constant of the net, top[0]'s diff)
 * is enqueued right behind the forward's kernels, then the call waits for the five tops.  Saves the host round trip between
 * the two calls (the GPU idles for it: ~20 us of a 0.4 ms step at B = 8192).  Same results as npair_forward + npair_backward;
 * when the forward reports an error the gradient buffer holds garbage and the context needs a new forward. */
int npair_forward_backward(npair_ctx* c, const float* d_feat, const float* d_label, float loss_weight, float* d_diff, float tops_host[5],
                           void* stream) {
  if (!c) return NPAIR_E_ARG;
  if (!d_feat || !d_label || !d_diff || !tops_host) { c->err = "null pointer argument"; return NPAIR_E_ARG; }
  if (c->world > 1 && !c->comm) { c->err = "context was created without a communicator"; return NPAIR_E_STATE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c->defer_sync = true;
  int rc = npair_forward(c, d_feat, d_label, tops_host, stream);
  c->defer_sync = false;
  if (rc != NPAIR_OK) return rc;
  c->fwd_done = true;                                  // enqueued; confirmed (or revoked) after the synchronisation below
  rc = backward_impl(c, loss_weight, d_diff, nullptr, nullptr, st);
  if (rc != NPAIR_OK) { c->fwd_done = false; return rc; }
  // wait for the forward's tops only: the gradient kernels keep running while the caller prepares (and enqueues) its next step
  { const int wrc = wait_tops(c, st); if (wrc != NPAIR_OK) { c->fwd_done = false; return wrc; } }
  const int derr = reinterpret_cast<int*>(c->tops_pinned)[5];
  if (derr) c->fwd_done = false;
  if (derr & DERR_EMPTY_LIST) { c->err = "an empty same/diff list was indexed (undefined behaviour in the reference, .cu:296/:327/:288)"; return NPAIR_E_EMPTY_LIST; }
  if (derr & DERR_POS_RANGE) { c->err = "identsn/diffsn select a position outside the list (undefined behaviour in the reference, .cu:285-288)"; return NPAIR_E_POS_RANGE; }
  for (int t = 0; t < 5; ++t) tops_host[t] = t < c->cfg.num_tops ? c->tops_pinned[t] : 0.f;
  return NPAIR_OK;
}

/* External-collectives variant of Backward_gpu up to the all-reduce (.cu:420-460):
 *   d_local_half : Q x D  = (1/2)(lw/Q) G . X_total
 *   d_total_half : N x D  = (1/2)(1/world)(lw/Q) G^T . X_local        (this rank's addend of the all-reduce)
 * so that bottom.diff of rank r = d_local_half + sum over ranks of d_total_half[rows of r]  (.cu:462-497).
 * world == 1: d_total_half may be NULL and d_local_half receives the complete gradient. */
int npair_backward_partial(npair_ctx* c, float loss_weight, float* d_local_half, float* d_total_half, void* stream) {
  if (!c) return NPAIR_E_ARG;
  if (!d_local_half || (c->world > 1 && !d_total_half)) { c->err = "null gradient pointer"; return NPAIR_E_ARG; }
  if (!c->fwd_done) { c->err = "npair_backward_partial called without a successful forward"; return NPAIR_E_STATE; }
  if (c->bwd_mode == NPAIR_BWDMODE_ROW_SCALARS) { c->err = "this context exchanges row scalars: use npair_row_scalars + npair_backward_gathered"; return NPAIR_E_STATE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->last_stream = st;
  return backward_impl(c, loss_weight, d_local_half, c->world > 1 ? d_total_half : nullptr, nullptr, st);
}

int npair_row_scalars(npair_ctx* c, float* d_out, void* stream) {
  if (!c || !d_out) return NPAIR_E_ARG;
  if (!c->fwd_done) { c->err = "npair_row_scalars called without a successful forward"; return NPAIR_E_STATE; }
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaMemcpyAsync(d_out, c->ra.rowscal, sizeof(float) * 8ull * c->Q, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return NPAIR_OK;
}

int npair_backward_gathered(npair_ctx* c, float loss_weight, const float* d_rs_total, float* d_diff, void* stream) {
  if (!c) return NPAIR_E_ARG;
  if (!d_rs_total || !d_diff) { c->err = "null pointer argument"; return NPAIR_E_ARG; }
  if (!c->fwd_done) { c->err = "npair_backward_gathered called without a successful forward"; return NPAIR_E_STATE; }
  if (c->bwd_mode != NPAIR_BWDMODE_ROW_SCALARS) { c->err = "this context does not exchange row scalars (see npair_bwd_exchange_mode)"; return NPAIR_E_STATE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->last_stream = st;
  return backward_impl(c, loss_weight, d_diff, nullptr, d_rs_total, st);
}

static int backward_core(npair_ctx* c, float loss_weight, float* d_diff, float* d_total_ext, const float* d_rs_ext, cudaStream_t st) {
  const int Q = c->Q, N = c->N, D = c->D;
  const MiningParams mp = mining_of(c->cfg);
  const int self_off = c->rank * Q;
  const bool wscope = c->cfg.global_scope && c->world > 1;
  // loss_weight / dot_normalizer (.cu:427,448); world scope: the normaliser is the world's batch and the transposed term is not
  // divided by the world size, i.e. exactly what a single rank holding the whole batch computes
  const float lw_over_q = loss_weight / static_cast<float>(wscope ? N : Q);
  const bool tc = c->cfg.gemm_backend == NPAIR_GEMM_TCGEN05;
  const float* rs_total = nullptr;
  int bw_mode = BW_SYM;
  if (c->bwd_mode == NPAIR_BWDMODE_ROW_SCALARS) {
    bw_mode = BW_ROWSCAL;
    if (d_rs_ext) rs_total = d_rs_ext;
    else {
      if (c->p2p_rec && !c->ext_gathered) {
        PhaseTimer pt(c, 8, st);
        const uint32_t ep = c->p2p_rec_epoch;
        const long long par = ep & 1u;
        p2p_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t*>(c->p2p_region + c->p2p_offFlags) + 2 * c->world + par * c->world, c->world, ep);
        count_launch();
        rs_total = c->p2p_region + c->p2p_offRec + par * 8ll * N;
      } else if (!c->rs_gathered) {
        // the only backward exchange: 8*Q floats per rank (replaces the N x D MPI_Allreduce of .cu:462-489)
        if (!c->comm) { c->err = "no communicator: use npair_backward_gathered with externally gathered row records"; return NPAIR_E_STATE; }
        PhaseTimer pt(c, 8, st);
        NcclApi* api = nccl_api();
        int r = api->AllGather(c->ra.rowscal, c->rs_total, 8ull * Q, NCCL_FLOAT32, c->comm, st);
        if (r != 0) { c->err = fmt("ncclAllGather(row records): %s", api->GetErrorString(r)); return NPAIR_E_NCCL; }
        c->rs_gathered = true;
        rs_total = c->rs_total;
      } else rs_total = c->rs_total;
    }
  } else if (c->bwd_mode == NPAIR_BWDMODE_REDUCE_SCATTER) bw_mode = BW_SPLIT;
  if (tc && c->fused_grad) {
    // weights are produced inside the gradient GEMM: no H in HBM
    FusedGradParams fp; memset(&fp, 0, sizeof(fp));
    fp.Q = Q; fp.N = N; fp.D = D; fp.num_kblocks = (N + 31) / 32;
    fp.tiles_m = (Q + 127) / 128; fp.tiles_n = (D + 255) / 256;
    fp.rowrec = c->ra.rowscal; fp.colrec = rs_total ? rs_total : c->ra.rowscal;
    fp.self_offset = self_off; fp.inv_world = wscope ? 1.f : 1.f / static_cast<float>(c->world);
    fp.log2_world = wscope ? 0.f : log2f(static_cast<float>(c->world));
    fp.sgn_p = (mp.ap_method == M_EASY || mp.ap_method == M_RELATIVE_EASY) ? -1.f : 1.f;
    fp.sgn_n = (mp.an_method == M_HARD || mp.an_method == M_RELATIVE_HARD) ? -1.f : 1.f;
    fp.out = d_diff; fp.ldo = D; fp.alpha = 0.5f * lw_over_q; fp.beta = 0.f; fp.dev_scale = &c->bs->x_inv_scale;
    fp.part = c->part; fp.splits = 1; fp.kb_per_split = fp.num_kblocks;
    fp.chunk_kb = c->grad_chunk_kb;
    if (c->grad_pair) fp.tiles_m = (fp.tiles_m + 1) / 2;       // 256-row pair blocks, one cluster of two CTAs each
    if (c->part) {
      const int tiles = fp.tiles_m * fp.tiles_n;
      int splits = (c->grad_pair ? c->sms / 2 : c->sms) / (tiles > 0 ? tiles : 1);
      if (splits > 16) splits = 16;
      if (splits > fp.num_kblocks / 8) splits = fp.num_kblocks / 8;        // keep >= 8 K blocks (256 columns) per split
      while (splits > 1 && static_cast<long long>(splits) * Q * D > c->part_floats) --splits;
      if (splits < 1) splits = 1;
      const int kpb = (fp.num_kblocks + splits - 1) / splits;
      fp.splits = (fp.num_kblocks + kpb - 1) / kpb; fp.kb_per_split = kpb;
    }
    {
      PhaseTimer pt(c, 6, st);
      if (c->grad_pair) CUDA_TRY(c, launch_fused_grad_pair(c->prec, c->tm_fB2, c->tm_fS, fp, c->sms, st));
      else CUDA_TRY(c, launch_fused_grad(c->prec, c->tm_fB, c->tm_fS, fp, c->sms, st));
      if (fp.splits > 1) {
        const long long n = static_cast<long long>(Q) * D;
        int nb = static_cast<int>((n / 4 + 255) / 256); if (nb > c->sms * 8) nb = c->sms * 8; if (nb < 1) nb = 1;
        splitk_reduce_kernel<<<nb, 256, 0, st>>>(c->part, fp.splits, n, d_diff, 0.f);
        count_launch();
      }
    }
    CUDA_TRY(c, cudaGetLastError());
    return NPAIR_OK;
  }
  {
    PhaseTimer pt(c, 5, st);
    launch_build_weights(c->S, c->ldS, Q, N, c->cur_label, c->lab_total, self_off, c->world, bw_mode, rs_total, mp, c->ra, c->prec, c->H, c->Np, c->HT, c->Qp, st);
  }
  GemmParams gp; memset(&gp, 0, sizeof(gp));
  gp.dev_scale = &c->bs->x_inv_scale;
  const bool rs_path = c->bwd_mode == NPAIR_BWDMODE_REDUCE_SCATTER;
  if (rs_path) {
    // total = (1/2)(1/k)(lw/Q) * G^T . X_local  (N x D)  -> reduce-scatter (== all-reduce + own slice, .cu:462-497)
    gp.M = N; gp.Nn = D; gp.num_kblocks = static_cast<int>((Q + c->bk_grad - 1) / c->bk_grad);
    gp.tiles_m = (N + 127) / 128; gp.tiles_n = (D + 255) / 256; gp.splits = 1; gp.kb_per_split = gp.num_kblocks;
    gp.out = d_total_ext ? d_total_ext : c->OUT2; gp.ldo = D; gp.alpha = 0.5f * (1.f / static_cast<float>(c->world)) * lw_over_q; gp.beta = 0.f;
    {
      PhaseTimer pt(c, 7, st);
      if (tc) CUDA_TRY(c, launch_split_gemm(c->prec, EPI_OUT, c->tm_b2A, c->tm_b2B, c->tm_S, gp, c->sms, st));
      else CUDA_TRY(c, launch_simt_gemm(c->prec, EPI_OUT, c->HT, c->Qp, static_cast<long long>(N) * c->Qp, c->XlT, c->Qp, static_cast<long long>(D) * c->Qp, Q, gp, st));
    }
    if (!d_total_ext) {
      PhaseTimer pt(c, 8, st);
      NcclApi* api = nccl_api();
      int r = api->ReduceScatter(c->OUT2, d_diff, static_cast<size_t>(Q) * D, NCCL_FLOAT32, NCCL_SUM, c->comm, st);
      if (r != 0) { c->err = fmt("ncclReduceScatter: %s", api->GetErrorString(r)); return NPAIR_E_NCCL; }
    }
  }
  // d_diff = (1/2)(lw/Q) * H . X_total  (H = G + G^T/world in the symmetric modes; accumulated onto the scattered term otherwise)
  gp.M = Q; gp.Nn = D; gp.num_kblocks = static_cast<int>((N + c->bk_grad - 1) / c->bk_grad);
  gp.tiles_m = (Q + 127) / 128; gp.tiles_n = (D + 255) / 256;
  gp.out = d_diff; gp.ldo = D; gp.alpha = 0.5f * lw_over_q; gp.beta = (rs_path && !d_total_ext) ? 1.f : 0.f;
  gp.splits = 1; gp.kb_per_split = gp.num_kblocks; gp.part = c->part;
  if (tc && c->part) plan_splits(&gp, c->sms, c->part_floats);
  {
    PhaseTimer pt(c, 6, st);
    if (tc) CUDA_TRY(c, launch_split_gemm(c->prec, EPI_OUT, c->tm_b1A, c->tm_b1B, c->tm_S, gp, c->sms, st));
    else CUDA_TRY(c, launch_simt_gemm(c->prec, EPI_OUT, c->H, c->Np, static_cast<long long>(Q) * c->Np, c->XsT, c->Np, static_cast<long long>(D) * c->Np, N, gp, st));
    if (tc && gp.splits > 1) {
      const long long n = static_cast<long long>(Q) * D;
      int nb = static_cast<int>((n / 4 + 255) / 256); if (nb > c->sms * 8) nb = c->sms * 8; if (nb < 1) nb = 1;
      splitk_reduce_kernel<<<nb, 256, 0, st>>>(c->part, gp.splits, n, d_diff, gp.beta);
      count_launch();
    }
  }
  CUDA_TRY(c, cudaGetLastError());
  return NPAIR_OK;
}

/* Per-phase CUDA-event timing on the caller's stream (bench.py's roofline leg).  Phases:
 * 0 forward all-gather   8 backward exchange (row-scalar all-gather or reduce-scatter)   1 operand prep (asum/absmax, split, stat init)
 * 2 similarity GEMM + fused statistics   3 thresholds + radix selects   4 forward row pass + finalize
 * 5 backward weight builder   6 gradient GEMM (G . X_total)   7 transposed gradient GEMM (G^T . X_local, world > 1) */
int npair_profile_enable(npair_ctx* c, int on) {
  if (!c) return NPAIR_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (on && !c->ev_made) {
    for (int i = 0; i < NPAIR_PROF_PHASES; ++i) { CUDA_TRY(c, cudaEventCreate(&c->ev[i][0])); CUDA_TRY(c, cudaEventCreate(&c->ev[i][1])); }
    c->ev_made = true;
  }
  c->prof = on != 0;
  for (int i = 0; i < NPAIR_PROF_PHASES; ++i) c->ev_used[i] = false;
  return NPAIR_OK;
}
/* milliseconds of each phase of the most recent forward+backward; synchronises the stream.  ms_out[9]. */
unsigned long long npair_kernel_launches(void) { return npair::g_kernel_launches; }

int npair_profile_read(npair_ctx* c, float* ms_out) {
  if (!c || !ms_out) return NPAIR_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaStreamSynchronize(c->last_stream));
  for (int i = 0; i < NPAIR_PROF_PHASES; ++i) {
    ms_out[i] = 0.f;
    if (c->ev_made && c->ev_used[i]) { float ms = 0.f; CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev[i][0], c->ev[i][1])); ms_out[i] = ms; }
    c->ev_used[i] = false;
  }
  return NPAIR_OK;
}

__global__ void decode_ord_kernel(const uint32_t* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ord2f(in[i]);
}
__global__ void int_to_float_kernel(const int* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}

int npair_debug_read(npair_ctx* c, int which, float* dst, size_t n) {
  if (!c || !dst) return NPAIR_E_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaStreamSynchronize(c->last_stream));
  const int Q = c->Q, N = c->N;
  if (which == 0) {
    if (n < static_cast<size_t>(Q) * N) { c->err = "buffer too small"; return NPAIR_E_ARG; }
    CUDA_TRY(c, cudaMemcpy2D(dst, sizeof(float) * N, c->S, sizeof(float) * c->ldS, sizeof(float) * N, Q, cudaMemcpyDeviceToHost));
    return NPAIR_OK;
  }
  if (which == 10) {
    if (n < 1) return NPAIR_E_ARG;
    CUDA_TRY(c, cudaMemcpy(dst, &c->bs->x_scale, sizeof(float), cudaMemcpyDeviceToHost));
    return NPAIR_OK;
  }
  if (n < static_cast<size_t>(Q)) { c->err = "buffer too small"; return NPAIR_E_ARG; }
  const float* src = nullptr; const uint32_t* osrc = nullptr; const int* isrc = nullptr;
  switch (which) {
    case 1: src = c->ra.posi_thr; break;
    case 2: src = c->ra.nega_thr; break;
    case 3: osrc = c->ra.st_minw; break;
    case 4: osrc = c->ra.st_maxb; break;
    case 5: osrc = c->ra.st_maxall; break;
    case 6: src = c->ra.A; break;
    case 7: src = c->ra.T; break;
    case 8: isrc = c->ra.cnt_same; break;
    case 9: osrc = c->ra.st_maxw; break;
    default: c->err = "unknown debug selector"; return NPAIR_E_ARG;
  }
  if (src) { CUDA_TRY(c, cudaMemcpy(dst, src, sizeof(float) * Q, cudaMemcpyDeviceToHost)); return NPAIR_OK; }
  float* tmp = nullptr;
  CUDA_TRY(c, cudaMalloc(&tmp, sizeof(float) * Q));
  if (osrc) decode_ord_kernel<<<(Q + 255) / 256, 256>>>(osrc, tmp, Q);
  else int_to_float_kernel<<<(Q + 255) / 256, 256>>>(isrc, tmp, Q);
  cudaError_t e = cudaMemcpy(dst, tmp, sizeof(float) * Q, cudaMemcpyDeviceToHost);
  cudaFree(tmp);
  CUDA_TRY(c, e);
  return NPAIR_OK;
}

__global__ void cvt_d2f_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = static_cast<float>(in[i]);
}
__global__ void cvt_f2d_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = static_cast<double>(in[i]);
}
/* Device-side dtype bridges for the Dtype=double instantiation of the Caffe layer (INSTANTIATE_CLASS, reference .cpp:190):
 * the reference's arithmetic is fp32 there too (expf/logf/FLT_MAX). */
int npair_util_f64_to_f32(const double* d_src, float* d_dst, size_t n, void* stream) {
  if (!d_src || !d_dst) return NPAIR_E_ARG;
  if (n == 0) return NPAIR_OK;
  int nb = static_cast<int>((n + 255) / 256); if (nb > 148 * 16) nb = 148 * 16;
  cvt_d2f_kernel<<<nb, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_dst, n);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? NPAIR_OK : NPAIR_E_CUDA;
}
int npair_util_f32_to_f64(const float* d_src, double* d_dst, size_t n, void* stream) {
  if (!d_src || !d_dst) return NPAIR_E_ARG;
  if (n == 0) return NPAIR_OK;
  int nb = static_cast<int>((n + 255) / 256); if (nb > 148 * 16) nb = 148 * 16;
  cvt_f2d_kernel<<<nb, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_dst, n);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? NPAIR_OK : NPAIR_E_CUDA;
}

int npair_l2normalize_forward(const float* d_x, int rows, int dim, float* d_y, float* d_inv_norm, void* stream) {
  if (!d_x || !d_y || rows < 1 || dim < 1) { g_create_err = "bad argument"; return NPAIR_E_ARG; }
  launch_l2norm_fwd(d_x, rows, dim, d_y, d_inv_norm, static_cast<cudaStream_t>(stream));
  return cudaGetLastError() == cudaSuccess ? NPAIR_OK : NPAIR_E_CUDA;
}
int npair_l2normalize_backward(const float* d_y, const float* d_inv_norm, const float* d_dy, int rows, int dim, float* d_dx, void* stream) {
  if (!d_y || !d_inv_norm || !d_dy || !d_dx || rows < 1 || dim < 1) { g_create_err = "bad argument"; return NPAIR_E_ARG; }
  launch_l2norm_bwd(d_y, d_inv_norm, d_dy, rows, dim, d_dx, static_cast<cudaStream_t>(stream));
  return cudaGetLastError() == cudaSuccess ? NPAIR_OK : NPAIR_E_CUDA;
}

int npair_debug_mma_symmetric(int precision) {
  if (precision < 0 || precision > 2) return NPAIR_E_ARG;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return NPAIR_E_CUDA;
  return mma_is_symmetric(precision, dev) ? 1 : 0;
}

int npair_debug_gemm(int precision, int backend, int M, int Nn, int K, const float* dA, const float* dB, float* dC, void* stream) {
  if (M < 1 || Nn < 1 || K < 1 || !dA || !dB || !dC || precision < 0 || precision > 2) { g_create_err = "bad argument"; return NPAIR_E_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ns = nsplit_of_prec(precision), bk = bk_of(precision, EPI_OUT);
  const long long Kp = round_up(K, 64);
  uint16_t *As = nullptr, *Bs = nullptr, *dummyT = nullptr;
  BlockScalars* bs = nullptr; float* partial = nullptr;
  int rc = NPAIR_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define DG_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { g_create_err = fmt("%s: %s", #call, cudaGetErrorString(e__)); rc = NPAIR_E_CUDA; goto done; } } while (0)
  {
    const long long Mt = round_up(M, 64), Nt = round_up(Nn, 64);
    const long long tmax = Mt > Nt ? Mt : Nt;
    DG_TRY(cudaMalloc(&As, 2ull * ns * M * Kp)); DG_TRY(cudaMalloc(&Bs, 2ull * ns * Nn * Kp));
    DG_TRY(cudaMalloc(&dummyT, 2ull * ns * K * tmax));
    DG_TRY(cudaMalloc(&bs, sizeof(BlockScalars))); DG_TRY(cudaMemset(bs, 0, sizeof(BlockScalars)));
    DG_TRY(cudaMalloc(&partial, sizeof(float) * 2048));
    // one common power-of-two scale over both operands (the layer multiplies X by X^T, i.e. a single matrix)
    launch_absmax_asum(dA, static_cast<long long>(M) * K, dA, static_cast<long long>(M) * K, partial, bs, precision == PREC_FP16X2, st);
    float sA = 1.f, sB = 1.f;
    if (precision == PREC_FP16X2) {
      DG_TRY(cudaStreamSynchronize(st));
      float mA = 0.f, mB = 0.f;
      DG_TRY(cudaMemcpy(&mA, &bs->x_absmax, 4, cudaMemcpyDeviceToHost));
      launch_absmax_asum(dB, static_cast<long long>(Nn) * K, dB, static_cast<long long>(Nn) * K, partial, bs, 1, st);
      DG_TRY(cudaStreamSynchronize(st));
      DG_TRY(cudaMemcpy(&mB, &bs->x_absmax, 4, cudaMemcpyDeviceToHost));
      const float mx = mA > mB ? mA : mB;
      int e = 0; if (mx > 0.f) frexpf(mx, &e);
      sA = ldexpf(1.f, -e); sB = ldexpf(1.f, e);
      float sc[2] = {sA, sB};
      DG_TRY(cudaMemcpy(&bs->x_scale, sc, 8, cudaMemcpyHostToDevice));
    } else {
      float sc[2] = {1.f, 1.f};
      DG_TRY(cudaMemcpy(&bs->x_scale, sc, 8, cudaMemcpyHostToDevice));
    }
    launch_split(dA, M, K, precision, bs, As, Kp, dummyT, tmax, nullptr, 0, 0, 0, nullptr, nullptr, Kp, st);
    launch_split(dB, Nn, K, precision, bs, Bs, Kp, dummyT, tmax, nullptr, 0, 0, 0, nullptr, nullptr, Kp, st);
    GemmParams gp; memset(&gp, 0, sizeof(gp));
    gp.M = M; gp.Nn = Nn; gp.num_kblocks = (K + bk - 1) / bk; gp.tiles_m = (M + 127) / 128; gp.tiles_n = (Nn + 255) / 256;
    gp.out = dC; gp.ldo = Nn; gp.alpha = 1.f; gp.beta = 0.f; gp.splits = 1; gp.kb_per_split = gp.num_kblocks;
    // EPI_OUT applies the inverse scale once; both operands were scaled -> fold the second factor into alpha
    gp.alpha = sB; gp.dev_scale = &bs->x_inv_scale;
    if (backend == NPAIR_GEMM_TCGEN05) {
      CUtensorMap ta, tb; std::string te;
      if (!make_tmap_pieces(&ta, As, K, M, ns, Kp, static_cast<long long>(M) * Kp, bk, 128, &te) ||
          !make_tmap_pieces(&tb, Bs, K, Nn, ns, Kp, static_cast<long long>(Nn) * Kp, bk, 256, &te)) { g_create_err = te; rc = NPAIR_E_CUDA; goto done; }
      DG_TRY(launch_split_gemm(precision, EPI_OUT, ta, tb, ta, gp, sms, st));
    } else {
      DG_TRY(launch_simt_gemm(precision, EPI_OUT, As, Kp, static_cast<long long>(M) * Kp, Bs, Kp, static_cast<long long>(Nn) * Kp, K, gp, st));
    }
    DG_TRY(cudaStreamSynchronize(st));
  }
done:
  cudaFree(As); cudaFree(Bs); cudaFree(dummyT); cudaFree(bs); cudaFree(partial);
#undef DG_TRY
  return rc;
}

}  // extern "C"
