// Host-side declarations of the kernel launchers (one .cu per kernel family).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/orx.h"

namespace orx {

// Rigorous bounds on |fast score - canonical cosine| per scan path ("Exactness").
constexpr double EPS_GEMV_F32 = 3.0e-6;   // fp32 FMA chain of depth 32 + 5 shuffle adds
constexpr double EPS_GEMV_BF16 = 3.0e-6;  // same arithmetic on exactly-converted bf16 rows
constexpr double EPS_UMMA_TF32 = 2.5e-3;  // both operands cut to 10 mantissa bits: (1+2^-10)^2-1 = 1.96e-3, + accumulation
constexpr double EPS_UMMA_BF16 = 2.5e-3;  // RNE bf16 query: 2^-9 = 1.96e-3 (rows are exact bf16), + accumulation

// candidate list = 32 * slots keys per query: k <= 16 -> 32 candidates, k <= 32 -> 64
// candidate list = 32 * slots keys per query: k <= 16 -> 32, k <= 32 -> 64, k <= 128 -> 32 more than k at least
inline int slots_for_k(int k) { return k <= 16 ? 1 : k <= 32 ? 2 : k <= 96 ? 4 : 5; }
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

struct QueryPrep {            // per query, written by prep_queries
    double n2q;               // canonical |q|^2
    int nonfinite;            // 1 if any element is NaN/Inf
    int zero;                 // 1 if |q| == 0 (every distance is NaN)
};

// ---- where a search's results go and how its completion is signalled (finalize.cu, exchange.cu)
// Result arrays of the WHOLE search, indexed by the query's position in the search (a launch that covers queries
// [q_base, q_base + nq) of it is told q_base).  Device memory, mapped host memory or peer memory alike.
struct ResultOut {
    orx_id *ids;        // [nq_total, k]
    double *dist;       // [nq_total, k]
    int *counts;        // [nq_total]
    int *flags;         // [nq_total]  bit 0 unproven, bit 1 non-finite query, bit 2 zero query, bit 3 rank error
};
// Row-sharded search: finalize pushes this rank's block straight into the gather buffers of its targets (every
// rank, or only the root of a one-process group) with peer stores, and the LAST finalize CTA of the search raises
// this rank's arrival word on each target.  n_targets == 0: not sharded, results go to `ResultOut`.
struct PublishArgs {
    char *const *slot;          // device array [n_targets]: MY slot inside target t's gather buffer (current set)
    uint32_t *const *flag;      // device array [n_targets]: MY arrival word on target t (current set)
    int n_targets;
    uint32_t seq;               // value the arrival words take
    uint64_t dist_off, counts_off, flags_off;      // layout of a slot for (nq_total, k)
};
// Completion of a search made of several launches: every CTA (= query) counts itself on `counter`; the one that
// brings it to `total` resets it and signals -- arrival words (sharded finalize) or `done_host` (a mapped host word the
// host polls instead of synchronising the stream).
struct DoneArgs {
    unsigned int *counter;      // device word, 0 between searches
    unsigned int total;         // queries of the whole search
    uint32_t *done_host;        // mapped host word <- token (nullptr: nothing to signal here)
    uint32_t token;
};

// ---- finalize.cu
void launch_prep_queries(const float *q, int nq, float *q_copy, float *qhat, void *qhat_bf16,
                         QueryPrep *prep, cudaStream_t st);
// q / prep / partial / floor are launch-relative (query 0 of this launch); results are written at q_base + query
void launch_finalize(int dtype, const void *table, const double *n2, const orx_id *row_ids,
                     const float *q, const QueryPrep *prep, const uint64_t *partial, int nparts,
                     int slots, int nq, int k, uint32_t n_rows, double eps,
                     const ResultOut &out, int q_base, const PublishArgs &pub, const DoneArgs &done,
                     cudaStream_t st, const float *floor = nullptr);
// list_stride_bytes == 0: three dense [n_lists][...] arrays; otherwise list l of each array is at +l*stride
void launch_merge_topk(int n_lists, int nq, int k, const orx_id *ids, const double *dist,
                       const int *counts, size_t list_stride_bytes, orx_id *out_ids, double *out_dist,
                       int *out_counts, cudaStream_t st);
// exhaustive fallback: collect rows whose fast score may reach `cos_floor`, rescore, select
void launch_collect(int dtype, const void *table, const float *scale, uint32_t n_rows,
                    const float *qhat, float fast_floor, int collect_all,
                    uint32_t *list, uint32_t *count, cudaStream_t st);
void launch_rescore_list(int dtype, const void *table, const double *n2, const orx_id *row_ids,
                         const float *q, const QueryPrep *prep, const uint32_t *list,
                         const uint32_t *count, double *dist_out, cudaStream_t st);
void launch_select_list(const orx_id *row_ids, const uint32_t *list, const uint32_t *count,
                        const double *dist, int k, orx_id *out_ids, double *out_dist,
                        int *out_count, cudaStream_t st);

void launch_flags_from_prep(const QueryPrep *prep, int nq, int *flags, cudaStream_t st);
void launch_fill_flags(int *flags, int nq, int value, cudaStream_t st);

// ---- exchange.cu (row-sharded search over NVLink peer memory)
void launch_publish(const void *my_slot, void *const *peer_slot, uint32_t *const *peer_flag, int world,
                    size_t bytes, uint32_t seq, cudaStream_t st);
// err_host: mapped host word <- 1 when a rank never arrived within the bounded wait (the kernel then gives up
// WITHOUT trapping: the CUDA context and the resident table survive, the host reports ORX_ERR_CUDA)
void launch_merge_wait(int world, int rank, int nq, int k, const void *set_base, size_t slot_stride,
                       size_t dist_off, size_t counts_off, size_t flags_off, const uint32_t *arrival,
                       int arrival_stride_words, uint32_t seq, orx_id *out_ids, double *out_dist,
                       int *out_counts, int *flags_any, int *flags_mine, int *redo, uint32_t *err_host,
                       const DoneArgs &done, cudaStream_t st);
void launch_signal_done(const DoneArgs &done, cudaStream_t st);      // 1 thread: *done_host = token (paths without a last CTA)

// ---- scan_gemv.cu
int scan_gemv_grid(int device, uint32_t n_rows);
void launch_scan_gemv(int dtype, const void *table, const float *scale, uint32_t n_rows,
                      const float *qhat, int nq, int slots, uint64_t *partial, int grid,
                      cudaStream_t st);

// the same scan over the rows whose bit is set in allow_bits [(n_rows+31)/32 words; bits >= n_rows are 0]
int scan_gemv_batch_parts(int device, uint32_t n_rows, int nq);
void launch_scan_gemv_batch(int dtype, const void *table, const float *scale, uint32_t n_rows, const float *qhat, int nq,
                            int slots, uint64_t *partial, int parts, cudaStream_t st);
void launch_scan_gemv_filtered(int dtype, const void *table, const float *scale, uint32_t n_rows,
                               const uint32_t *allow_bits, const float *qhat, int nq, int slots,
                               uint64_t *partial, int grid, cudaStream_t st);

// ---- table_ops.cu
void launch_validate_rows(const float *src, uint64_t n, int *flag, cudaStream_t st);
void launch_commit_rows(int dtype, const float *src, const uint32_t *src_idx, const uint32_t *dst_row,
                        const orx_id *ids, uint32_t n, void *table, float *scale, double *n2,
                        orx_id *row_ids, const int *abort_flag, cudaStream_t st);
void launch_adopt_rows(int dtype, const void *table, uint32_t row0, uint32_t n, const orx_id *ids, float *scale,
                       double *n2, orx_id *row_ids, int *flag, cudaStream_t st);
void launch_move_rows(int dtype, const uint32_t *src_row, const uint32_t *dst_row, uint32_t n,
                      void *table, float *scale, double *n2, orx_id *row_ids, cudaStream_t st);
void launch_gather_rows(int dtype, const void *table, const uint32_t *rows, uint32_t n, float *out,
                        cudaStream_t st);

// ---- pgwire.cu (pgvector wire formats: COPY BINARY bulk load, text input)
void launch_mask_scale(const float *scale, const uint32_t *allow_bits, uint32_t n_rows, float *out, cudaStream_t st);
void launch_decode_pgvector(const uint8_t *raw, const uint64_t *payload_off, uint32_t n, float *out,
                            cudaStream_t st);

// launch `kernel` on `st` with the programmatic-stream-serialization attribute (common.cuh: PDL)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// SM count of the CURRENT device (cached per device): launch geometry is sized from it, never from a literal
int device_sms();
template <typename T> inline T cap_grid(T blocks, int ctas_per_sm) {     // persistent-style grids: <= ctas_per_sm x SMs
    const T cap = (T)device_sms() * (T)ctas_per_sm;
    return blocks > cap ? cap : blocks;
}

// ---- orx_api.cu: what the other translation units need from an index
int set_error(int code, const char *fmt, ...);     // records the thread-local message, returns code
int index_device(const orx_index *ix);
void index_count_launches(orx_index *ix, uint64_t n);

}  // namespace orx
