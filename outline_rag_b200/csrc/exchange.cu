// Row-sharded search: the exchange of per-GPU top-k lists FUSED with the kernels on either side of it, over
// NVLink peer memory.  No collective library call and no host round trip between the scan and
// the merged answer:
//
//   finalize_kernel  (csrc/finalize.cu) writes this rank's result block STRAIGHT INTO slot[rank] of every
//                    target GPU's gather buffer with P2P stores (NVLink5 / NVSwitch); the search's last finalize
//                    CTA fences at system scope and stores the sequence number into flag[rank] on each target;
//   merge_wait_kernel one CTA per query spins (bounded) until every rank's flag carries the sequence number,
//                    merges the G sorted lists with the ordering contract (distance ASC, NaN last, id ASC) --
//                    a candidate's rank is its position in its own list plus one binary search per other list --
//                    and writes the answer, straight into mapped host memory when the caller wants it there; the
//                    last CTA raises the host's completion word, which the host polls (no stream synchronise).
//   publish_kernel   the stand-alone push (CTA g copies a block that already sits in the local slot to target g),
//                    for the paths that do not end in finalize: an empty shard, the exact second round.
//
// Payload per rank: nq * (k*24 + 8) bytes (296 B at k=12, nq=1): latency-bound.  Chain per search on one stream:
// prep -> scan -> finalize(+publish) -> merge_wait, no host synchronisation inside.
// Two buffer sets alternate with the sequence number: a rank can run at most one search ahead of
// a peer (its merge needs the peer's publish), so set (seq & 1) is never overwritten while read.
#include "common.cuh"
#include "internal.h"

namespace orx {

__global__ void __launch_bounds__(256)
publish_kernel(const uint4 *__restrict__ my_slot, uint4 *const *__restrict__ peer_slot,
               uint32_t *const *__restrict__ peer_flag, int n_vec, uint32_t seq) {
    const int g = blockIdx.x;
    uint4 *dst = peer_slot[g];
    if (dst != my_slot) {
        for (int i = threadIdx.x; i < n_vec; i += blockDim.x) dst[i] = my_slot[i];
    }
    __threadfence_system();                 // my stores are visible system-wide before the flag is
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t *>(peer_flag[g]) = seq;
}

void launch_publish(const void *my_slot, void *const *peer_slot, uint32_t *const *peer_flag, int world,
                    size_t bytes, uint32_t seq, cudaStream_t st) {
    publish_kernel<<<world, 256, 0, st>>>(static_cast<const uint4 *>(my_slot),
                                          reinterpret_cast<uint4 *const *>(peer_slot), peer_flag,
                                          (int)(bytes / 16), seq);
}

constexpr int XMERGE_MAX = 1024;      // world * k candidates per query (8 GPUs x k = 128)
constexpr int XWORLD_MAX = 64;

__global__ void __launch_bounds__(256)
merge_wait_kernel(int world, int rank, int nq, int k, const char *__restrict__ set_base, size_t slot_stride,
                  size_t dist_off, size_t counts_off, size_t flags_off,
                  const uint32_t *__restrict__ arrival, int arrival_stride_words, uint32_t seq,
                  orx_id *__restrict__ out_ids, double *__restrict__ out_dist, int *__restrict__ out_counts,
                  int *__restrict__ flags_any, int *__restrict__ flags_mine, int *__restrict__ redo,
                  uint32_t *__restrict__ err_host, const DoneArgs done) {
    __shared__ double s_d[XMERGE_MAX];
    __shared__ uint64_t s_hi[XMERGE_MAX], s_lo[XMERGE_MAX];
    __shared__ int s_cnt[XWORLD_MAX];
    __shared__ int s_fail;
    const int qi = blockIdx.x;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    // ---- wait until every rank has published this search (peer GPUs run concurrently)
    if ((int)threadIdx.x < world) {
        const volatile uint32_t *f = arrival + (size_t)threadIdx.x * arrival_stride_words;
        uint64_t t0 = 0;
        for (uint32_t it = 0;; ++it) {
            if ((int32_t)(*f - seq) >= 0) break;
            if ((it & 255u) == 255u) {
                uint64_t now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 10000000000ull) {      // a rank never arrived: give up, report, keep the context
                    s_fail = 1;
                    break;
                }
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const bool failed = s_fail != 0;
    if (!failed) {
        const int total = world * k;
        if ((int)threadIdx.x < world)
            s_cnt[threadIdx.x] = min(k, max(0, __ldcg(reinterpret_cast<const int *>(set_base + (size_t)threadIdx.x * slot_stride + counts_off) + qi)));
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int l = e / k, r = e % k;
            const char *slot = set_base + (size_t)l * slot_stride;
            const size_t src = (size_t)qi * k + r;
            const ulonglong2 id = __ldcg(reinterpret_cast<const ulonglong2 *>(slot) + src);
            s_d[e] = __ldcg(reinterpret_cast<const double *>(slot + dist_off) + src);
            s_hi[e] = id.x;
            s_lo[e] = id.y;
        }
        __syncthreads();
        int valid = 0;
        for (int l = 0; l < world; ++l) valid += s_cnt[l];
        const int count = min(k, valid);
        // every list is sorted by the ordering contract: rank(e) = position in its own list + for every other list
        // the number of its entries that sort before e (an identical (distance, id) in an earlier list counts as
        // before -- cannot happen with disjoint shards, keeps the order total)
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int l = e / k, r = e % k;
            if (r >= s_cnt[l]) continue;
            const double d = s_d[e];
            const uint64_t hi = s_hi[e], lo = s_lo[e];
            int rank_e = r;
            for (int o = 0; o < world && rank_e < k; ++o) {
                if (o == l) continue;
                int lo_i = 0, hi_i = s_cnt[o];
                const int base = o * k;
                while (lo_i < hi_i) {
                    const int mid = (lo_i + hi_i) >> 1;
                    const int c = base + mid;
                    const bool before = o < l ? !sorts_before(d, hi, lo, s_d[c], s_hi[c], s_lo[c])
                                              : sorts_before(s_d[c], s_hi[c], s_lo[c], d, hi, lo);
                    if (before) lo_i = mid + 1;
                    else hi_i = mid;
                }
                rank_e += lo_i;
            }
            if (rank_e < k) {
                out_ids[(size_t)qi * k + rank_e].hi = hi;
                out_ids[(size_t)qi * k + rank_e].lo = lo;
                out_dist[(size_t)qi * k + rank_e] = d;
            }
        }
        for (int r = count + threadIdx.x; r < k; r += blockDim.x) {
            out_ids[(size_t)qi * k + r].hi = 0ull;
            out_ids[(size_t)qi * k + r].lo = 0ull;
            out_dist[(size_t)qi * k + r] = __longlong_as_double(0x7ff8000000000000ll);
        }
        if (threadIdx.x == 0) {
            out_counts[qi] = count;
            int any = 0;
            for (int l = 0; l < world; ++l)
                any |= __ldcg(reinterpret_cast<const int *>(set_base + (size_t)l * slot_stride + flags_off) + qi);
            flags_any[qi] = any;
            flags_mine[qi] = __ldcg(reinterpret_cast<const int *>(set_base + (size_t)rank * slot_stride + flags_off) + qi);
            if (any & 1) *redo = 1;
        }
    } else if (threadIdx.x == 0) {
        *reinterpret_cast<volatile uint32_t *>(err_host) = 1u;
    }
    // completion word for the host (which polls it instead of synchronising the stream)
    if (done.counter != nullptr) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            bool last = done.total == 1u;
            if (!last) {
                const unsigned int old = atomicAdd(done.counter, 1u);
                if (old + 1u == done.total) {
                    *done.counter = 0u;
                    __threadfence_system();
                    last = true;
                }
            }
            if (last && done.done_host != nullptr) *reinterpret_cast<volatile uint32_t *>(done.done_host) = done.token;
        }
    }
}

void launch_merge_wait(int world, int rank, int nq, int k, const void *set_base, size_t slot_stride,
                       size_t dist_off, size_t counts_off, size_t flags_off, const uint32_t *arrival,
                       int arrival_stride_words, uint32_t seq, orx_id *out_ids, double *out_dist,
                       int *out_counts, int *flags_any, int *flags_mine, int *redo, uint32_t *err_host,
                       const DoneArgs &done, cudaStream_t st) {
    if (nq <= 0) return;
    // Small batches: launched with the programmatic-serialization attribute, so the merge CTAs are resident and polling
    // the arrival words while this rank's finalize still runs (they synchronise through the words and system fences, not
    // through grid completion).  Safe only while the polling CTAs cannot crowd finalize's CTAs out of the SMs: nq <= 128.
    if (nq <= 128)
        launch_pdl(merge_wait_kernel, dim3(nq), dim3(256), 0, st, world, rank, nq, k, static_cast<const char *>(set_base),
                   slot_stride, dist_off, counts_off, flags_off, arrival, arrival_stride_words, seq, out_ids, out_dist,
                   out_counts, flags_any, flags_mine, redo, err_host, done);
    else
        merge_wait_kernel<<<nq, 256, 0, st>>>(world, rank, nq, k, static_cast<const char *>(set_base), slot_stride,
                                              dist_off, counts_off, flags_off, arrival, arrival_stride_words, seq,
                                              out_ids, out_dist, out_counts, flags_any, flags_mine, redo, err_host, done);
}

}  // namespace orx
