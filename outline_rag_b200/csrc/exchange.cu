// Row-sharded search: the exchange of per-GPU top-k lists FUSED with the merge, over NVLink peer
// memory.  No collective library call and no host round trip between the scan and
// the merged answer:
//
//   finalize_kernel  (csrc/finalize.cu)  writes this rank's result block into ITS OWN slot of its
//                    gather buffer;
//   publish_kernel   CTA g pushes that block with 128-bit peer stores into slot[rank] of GPU g's
//                    gather buffer (P2P over NVLink5 / NVSwitch), fences at system scope and then
//                    stores the search's sequence number into flag[rank] on GPU g;
//   merge_wait_kernel one CTA per query spins (bounded) until every rank's flag carries the
//                    sequence number, then merges the G lists with the ordering contract
//                    (distance ASC, NaN last, id ASC) and writes the answer -- straight into mapped
//                    host memory when the caller wants it there.
//
// Payload per rank: nq * (k*24 + 8) bytes (296 B at k=12, nq=1): latency-bound, so what matters is
// that the whole chain is 5 back-to-back launches on one stream with ONE host synchronisation.
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

__global__ void __launch_bounds__(256)
merge_wait_kernel(int world, int rank, int nq, int k, const char *__restrict__ set_base, size_t slot_stride,
                  size_t dist_off, size_t counts_off, size_t flags_off,
                  const uint32_t *__restrict__ arrival, int arrival_stride_words, uint32_t seq,
                  orx_id *__restrict__ out_ids, double *__restrict__ out_dist, int *__restrict__ out_counts,
                  int *__restrict__ flags_any, int *__restrict__ flags_mine, int *__restrict__ redo) {
    __shared__ double s_d[XMERGE_MAX];
    __shared__ uint64_t s_hi[XMERGE_MAX], s_lo[XMERGE_MAX];
    __shared__ unsigned char s_ok[XMERGE_MAX];
    __shared__ int s_valid;
    const int qi = blockIdx.x;
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
                else if (now - t0 > 10000000000ull) __trap();       // a rank never arrived: fail loudly
            }
        }
        __threadfence_system();
    }
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    const int total = world * k;
    int mine = 0;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int l = e / k, r = e % k;
        const char *slot = set_base + (size_t)l * slot_stride;
        const size_t src = (size_t)qi * k + r;
        const int cnt = __ldcg(reinterpret_cast<const int *>(slot + counts_off) + qi);
        const bool ok = r < cnt;
        const ulonglong2 id = __ldcg(reinterpret_cast<const ulonglong2 *>(slot) + src);
        s_ok[e] = ok;
        s_d[e] = __ldcg(reinterpret_cast<const double *>(slot + dist_off) + src);
        s_hi[e] = id.x;
        s_lo[e] = id.y;
        mine += ok;
    }
    atomicAdd(&s_valid, mine);
    __syncthreads();
    const int count = min(k, s_valid);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        if (!s_ok[e]) continue;
        int rank_e = 0;
        for (int c = 0; c < total; ++c)
            rank_e += (c != e && s_ok[c] &&
                       (sorts_before(s_d[c], s_hi[c], s_lo[c], s_d[e], s_hi[e], s_lo[e]) ||
                        (c < e && s_d[c] == s_d[e] && s_hi[c] == s_hi[e] && s_lo[c] == s_lo[e])));
        if (rank_e < k) {
            out_ids[(size_t)qi * k + rank_e].hi = s_hi[e];
            out_ids[(size_t)qi * k + rank_e].lo = s_lo[e];
            out_dist[(size_t)qi * k + rank_e] = s_d[e];
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
}

void launch_merge_wait(int world, int rank, int nq, int k, const void *set_base, size_t slot_stride,
                       size_t dist_off, size_t counts_off, size_t flags_off, const uint32_t *arrival,
                       int arrival_stride_words, uint32_t seq, orx_id *out_ids, double *out_dist,
                       int *out_counts, int *flags_any, int *flags_mine, int *redo, cudaStream_t st) {
    if (nq <= 0) return;
    merge_wait_kernel<<<nq, 256, 0, st>>>(world, rank, nq, k, static_cast<const char *>(set_base), slot_stride,
                                          dist_off, counts_off, flags_off, arrival, arrival_stride_words, seq,
                                          out_ids, out_dist, out_counts, flags_any, flags_mine, redo);
}

}  // namespace orx
