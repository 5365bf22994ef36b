/*
 * npair_b200.h -- C ABI of libnpair_b200.so: the B200-native NPairMultiClassLoss hot path.
 *
 * This is the drop-in boundary for the reference layer's GPU methods.  Each entry point cites the
 * reference interface it replaces (paths relative to quziyan/NPairLoss):
 *
 *   npair_create / npair_destroy   <- NPairMultiClassLossLayer::LayerSetUp        npair_multi_class_loss.cpp:19-155
 *                                     (parameter read :32-42, scratch allocation :44-154; the reference leaks
 *                                      total_feature_/total_label_, .hpp:35 -- here the context owns and frees all scratch)
 *   npair_forward                  <- NPairMultiClassLossLayer::Forward_gpu       npair_multi_class_loss.cu:207-402
 *                                     incl. GatherFeatureAndLabel (.cu:17-43, MPI_Allgather -> ncclAllGather)
 *   npair_backward                 <- NPairMultiClassLossLayer::Backward_gpu      npair_multi_class_loss.cu:420-499
 *                                     incl. MPI_Allreduce + slice (.cu:462-497 -> ncclReduceScatter)
 *   npair_config                   <- message NPairLossParameter                  caffe.proto:3-23
 *                                     + fork statics Caffe::NUM_GPU / Caffe::RANK (.cu:214,220)
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns 0 on success or a
 * negative NPAIR_E_* code, with a human-readable message from npair_last_error().  No exceptions cross the ABI.
 * Device pointers are on the context's device; `stream` is a cudaStream_t passed as void* (NULL = legacy default).
 * A context is not re-entrant; use one per rank (one process per GPU).
 */
#ifndef NPAIR_B200_H_
#define NPAIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPAIR_ABI_VERSION 2

/* caffe.proto:8-11 */
enum { NPAIR_GLOBAL = 0, NPAIR_LOCAL = 1 };
/* caffe.proto:12-18 (RAND selects ALL pairs: there is no RNG in the reference, .cu:88-89) */
enum { NPAIR_HARD = 0, NPAIR_EASY = 1, NPAIR_RAND = 2, NPAIR_RELATIVE_HARD = 3, NPAIR_RELATIVE_EASY = 4 };

/* how the fp32 operands are fed to the tcgen05 tensor cores (fp32 accumulation in TMEM in every mode) */
enum {
  NPAIR_PREC_FP32_BF16X3 = 0, /* 3-way bf16 split, 6 MMA passes: ~2^-24 relative, any dynamic range (fp32-faithful) */
  NPAIR_PREC_BF16 = 1,        /* single bf16 pass: ~2^-9 relative (throughput mode, BASELINE config 3)            */
  NPAIR_PREC_FP32_FP16X2 = 2  /* 2-way fp16 split of power-of-two pre-scaled operands, 3 passes: ~2^-22 relative to
                                 max|x| (fp32-faithful for embeddings of bounded dynamic range, e.g. L2-normalised) */
};
enum { NPAIR_GEMM_TCGEN05 = 0, NPAIR_GEMM_SIMT_CHECK = 1 /* slow fp32 CUDA-core cross-check, tests only */ };

enum {
  NPAIR_OK = 0,
  NPAIR_E_ARG = -1,        /* bad argument / unsupported configuration                                   */
  NPAIR_E_CUDA = -2,       /* CUDA runtime / driver error                                                */
  NPAIR_E_NCCL = -3,       /* NCCL missing or failed                                                     */
  NPAIR_E_EMPTY_LIST = -4, /* reference would index an empty list (UB upstream: .cu:296, :327) */
  NPAIR_E_POS_RANGE = -5,  /* pos(SN,size) outside [0,size) (UB upstream: .cu:288, :303, :319, :334)     */
  NPAIR_E_STATE = -6       /* backward without forward etc.                                              */
};

typedef struct npair_ctx npair_ctx;

typedef struct {
  int32_t Q;            /* per-rank batch = bottom[0]->num()           (.cpp:24)                */
  int32_t D;            /* feature dim = channels*height*width          (.cu:215)                */
  int32_t world;        /* Caffe::NUM_GPU                               (.cu:214)                */
  int32_t rank;         /* Caffe::RANK                                  (.cu:220)                */
  int32_t num_tops;     /* 1..5 top blobs                               (.hpp:32-34)             */
  float margin_ident;   /* caffe.proto:4  default 0  */
  float margin_diff;    /* caffe.proto:5  default 0  */
  float identsn;        /* caffe.proto:6  default -1 */
  float diffsn;         /* caffe.proto:7  default -1 */
  int32_t ap_region;    /* caffe.proto:19 default LOCAL */
  int32_t ap_method;    /* caffe.proto:20 default RAND  */
  int32_t an_region;    /* caffe.proto:21 default LOCAL */
  int32_t an_method;    /* caffe.proto:22 default RAND  */
  int32_t sim_precision; /* NPAIR_PREC_*  */
  int32_t gemm_backend;  /* NPAIR_GEMM_*  */
  int32_t device;        /* CUDA device ordinal; -1 = current device */
  int32_t bwd_exchange;  /* world > 1 only.  NPAIR_BWD_AUTO: row-record exchange (every operand format is laid out so that the
                            similarity GEMM is bitwise symmetric; npair_create verifies that on the device and falls back to the
                            reduce-scatter form if the check fails).  NPAIR_BWD_REDUCE_SCATTER forces the reference's form
                            (all-reduce of the N x D transposed product, .cu:455-497). */
  /* ---- ABI 2: extensions beyond the reference layer (all 0 = reference behaviour) ---- */
  int32_t global_scope;    /* 1: GLOBAL-region mining lists and the loss / gradient normaliser span the WORLD's N x N pairs, so the
                              result does not depend on how the batch is sharded.  0: per rank, as the reference
                              (.cu:225-268 builds the lists from the rank's own Q x N block, :385/:427 divide by Q). */
  int32_t normalize_input; /* 1: the L2Normalize producer layer (usage/def.prototxt:115-120) is fused in: bottom[0] holds raw
                              embeddings, the layer works on x / ||x||_2 and returns the gradient w.r.t. the raw embeddings. */
  int32_t grad_chunk_cols; /* accumulation chunk of the gradient GEMM in database columns (multiple of 32); 0 = default (2048: gradient 5e-6 from exact at any N; 1024 halves that for +8 % kernel time), < 0 = one accumulator.
                              The tensor core truncates its fp32 accumulator on every MMA; chunks bound that error, see DESIGN 5 */
  int32_t flags;           /* NPAIR_FLAG_* */
} npair_config;

/* tuning / diagnostic switches (were environment variables in ABI 1) */
enum {
  NPAIR_FLAG_NO_FUSED_GRAD = 1,   /* materialise the gradient weights and run the plain split GEMM (cross-check path)        */
  NPAIR_FLAG_SIM_1CTA = 2,        /* similarity GEMM without CTA pairs                                                       */
  NPAIR_FLAG_GRAD_1CTA = 4,       /* gradient GEMM without CTA pairs                                                         */
  NPAIR_FLAG_NCCL_RECORDS = 8,    /* world > 1: exchange the row records with ncclAllGather instead of NVLink peer stores    */
  NPAIR_FLAG_NCCL_FEATURES = 16,  /* world > 1: gather the features with ncclAllGather instead of NVLink peer loads          */
  NPAIR_FLAG_LSEL_WARP = 32       /* LOCAL RELATIVE_* select: warp-per-row kernel also for rows that fit the block-per-row one */
};

enum { NPAIR_BWD_AUTO = 0, NPAIR_BWD_REDUCE_SCATTER = 1 };
/* what a context actually uses: 0 = single rank (symmetric tiles), 1 = reduce-scatter, 2 = row-scalar exchange */
enum { NPAIR_BWDMODE_SINGLE = 0, NPAIR_BWDMODE_REDUCE_SCATTER = 1, NPAIR_BWDMODE_ROW_SCALARS = 2 };
int npair_bwd_exchange_mode(const npair_ctx* ctx);

/* fills proto defaults (caffe.proto:4-7,19-22), world=1, rank=0, num_tops=5, fp32-faithful fp16x2, tcgen05, extensions off */
void npair_config_default(npair_config* cfg, int32_t Q, int32_t D);

/* Workspace the context will allocate on the device for this configuration (bytes). */
size_t npair_workspace_bytes(const npair_config* cfg);

/* 128-byte NCCL unique id (rank 0 calls this and ships the bytes to the other ranks out of band). */
int npair_nccl_unique_id(void* out_id_128B);

/* world == 1: nccl_unique_id_128B may be NULL.  world > 1: collective call, every rank passes the same id
 * (NULL id with world > 1 creates an external-collectives context, see npair_forward_gathered). */
int npair_create(const npair_config* cfg, const void* nccl_unique_id_128B, npair_ctx** out);
/* world > 1 with a communicator the host framework already owns (ncclComm_t passed as void*; not destroyed). */
int npair_create_with_comm(const npair_config* cfg, void* nccl_comm, npair_ctx** out);
void npair_destroy(npair_ctx* ctx);

/* Forward_gpu.  d_feat: Q x D fp32 row-major (bottom[0]->gpu_data()); d_label: Q fp32 (bottom[1]->gpu_data()).
 * tops_host[0..num_tops) are written exactly as the reference writes top[i]->mutable_cpu_data()[0]
 * (.cu:388-401): [loss, top1, top5, top10, feature_asum], the LAST top always being the asum.  The call
 * returns after the scalars are valid on the host (one stream synchronisation).  The feature / label buffers
 * must stay unchanged until npair_backward has been enqueued (the reference caches the blobs too, .cpp:29-30). */
int npair_forward(npair_ctx* ctx, const float* d_feat, const float* d_label, float tops_host[5], void* stream);

/* Backward_gpu.  loss_weight = top[0]->cpu_diff()[0] (.cu:435).  d_feat_diff: Q x D fp32, OVERWRITTEN
 * (bottom[0]->mutable_gpu_diff(); beta = 0 at .cu:448, propagate_down ignored).  Asynchronous on `stream`. */
int npair_backward(npair_ctx* ctx, float loss_weight, float* d_feat_diff, void* stream);

/* npair_forward + npair_backward with a single host synchronisation: the backward is enqueued behind the forward's kernels (the
 * loss weight -- top[0]'s diff, reference .cu:435 -- is a constant of the net), then the call waits for the five tops ONLY: like
 * npair_backward, the gradient is complete in stream order, not when the call returns.  Same results; saves the host round trip
 * during which the GPU idles.  On a forward error the gradient buffer is unspecified. */
int npair_forward_backward(npair_ctx* ctx, const float* d_feat, const float* d_label, float loss_weight, float* d_feat_diff,
                           float tops_host[5], void* stream);

/* External-collectives variants for host frameworks that keep their own communication layer (and for emulating all
 * ranks on one GPU in tests).  A context for world > 1 created with npair_create(cfg, NULL, ..) has no communicator
 * and only accepts these two calls.
 *   npair_forward_gathered : d_feat_total N x D and d_label_total N are the already all-gathered bottoms
 *                            (what GatherFeatureAndLabel produces, .cu:17-43); rank r's rows are [r*Q,(r+1)*Q).
 *   npair_backward_partial : Backward_gpu up to the all-reduce (.cu:420-460):
 *        d_local_half  Q x D = (1/2)(lw/Q) G . X_total
 *        d_total_half  N x D = (1/2)(1/world)(lw/Q) G^T . X_local    (this rank's addend of the all-reduce)
 *     so bottom.diff of rank r = d_local_half + sum_over_ranks d_total_half[rows of r]  (.cu:462-497).
 *     world == 1: d_total_half may be NULL; d_local_half then receives the complete gradient. */
int npair_forward_gathered(npair_ctx* ctx, const float* d_feat_total, const float* d_label_total, float tops_host[5], void* stream);
int npair_backward_partial(npair_ctx* ctx, float loss_weight, float* d_local_half, float* d_total_half, void* stream);
/* Row-scalar exchange form (contexts whose npair_bwd_exchange_mode is NPAIR_BWDMODE_ROW_SCALARS).  Because the similarity
 * GEMM is bitwise symmetric across ranks, rank r can evaluate the transposed gradient weights G[m][j] of every other rank
 * from its own S[j][m] and a 32-byte record of row m, so the backward exchange is an all-gather of 8*Q floats per rank instead
 * of the reference's N x D all-reduce:
 *   npair_row_scalars       : copies this rank's [Q][8] records (after a forward) to d_out_8Q
 *   npair_backward_gathered : d_rs_total = [N][8] records of all ranks in global row order; writes the complete bottom.diff */
int npair_row_scalars(npair_ctx* ctx, float* d_out_8Q, void* stream);
int npair_backward_gathered(npair_ctx* ctx, float loss_weight, const float* d_rs_total, float* d_feat_diff, void* stream);

/* The L2Normalize producer layer of the reference net (usage/def.prototxt:115-120; its source is not part of the reference tree):
 * y[r][:] = x[r][:] / ||x[r][:]||_2 (a zero row stays zero), and its backward dx = (dy - y (y . dy)) / ||x||.  Stand-alone entry
 * points for a host framework's own L2Normalize layer; npair_config.normalize_input = 1 runs the same kernels inside
 * npair_forward / npair_backward.  d_inv_norm: rows floats (1 / ||x||, 0 for a zero row). */
int npair_l2normalize_forward(const float* d_x, int rows, int dim, float* d_y, float* d_inv_norm, void* stream);
int npair_l2normalize_backward(const float* d_y, const float* d_inv_norm, const float* d_dy, int rows, int dim, float* d_dx, void* stream);

const char* npair_last_error(const npair_ctx* ctx);   /* ctx may be NULL: last create() error of this thread */
const char* npair_version(void);

/* Per-phase CUDA-event timing on the caller's stream (used by bench.py for the roofline of the dominant kernel).
 * ms_out[9]: 0 forward all-gather  1 operand prep  2 similarity GEMM (+fused statistics)  3 thresholds / radix selects
 *            4 forward row pass + finalize  5 backward weight builder  6 gradient GEMM  7 transposed gradient GEMM
 *            8 backward exchange (row-scalar all-gather or reduce-scatter) */
int npair_profile_enable(npair_ctx* ctx, int on);
int npair_profile_read(npair_ctx* ctx, float ms_out[9]);
/* Cumulative number of CUDA kernels this library has launched in the calling process (all contexts).  bench.py reports
 * the difference across its timed region as "gpu_launches". */
unsigned long long npair_kernel_launches(void);

/* Device-side dtype bridges for the Dtype=double instantiation of the Caffe layer (INSTANTIATE_CLASS, reference
 * npair_multi_class_loss.cpp:190); the reference's arithmetic is fp32 there too (expf/logf/FLT_MAX). */
int npair_util_f64_to_f32(const double* d_src, float* d_dst, size_t n, void* stream);
int npair_util_f32_to_f64(const float* d_src, double* d_dst, size_t n, void* stream);

/* Introspection for parity tests (copies device scratch to host; synchronises the context's last stream).
 * which: 0 = S (Q x N similarities, row-major, ld = N)      1 = posi_thr[Q]   2 = nega_thr[Q]
 *        3 = min_within[Q]  4 = max_between[Q]  5 = max_all[Q]  6 = A[Q]  7 = T[Q]  8 = same-label count[Q]
 *        9 = max_within[Q]  10 = operand pre-scale (1 float) */
int npair_debug_read(npair_ctx* ctx, int which, float* host_dst, size_t n_floats);

/* 1 if a similarity matrix computed on this device with every tile (no mirroring) in operand format `precision` comes out bitwise
 * symmetric, 0 if not (the row-record backward exchange then falls back to the reduce-scatter form), negative on error.  npair_create
 * runs and caches this check itself for world > 1. */
int npair_debug_mma_symmetric(int precision);

/* Stand-alone run of the split-operand GEMM  C[M x Nn] = A[M x K] . B[Nn x K]^T  on device fp32 inputs
 * (unit test of the tcgen05 path; not used by the layer). */
int npair_debug_gemm(int precision, int backend, int M, int Nn, int K, const float* d_A, const float* d_B, float* d_C,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NPAIR_B200_H_ */
