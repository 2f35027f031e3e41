// Shared device helpers: orderable score keys, the warp-distributed top-K list,
// the canonical binary64 halving-tree sum, streaming loads.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define ORX_DIM 1024
#define FULL_MASK 0xffffffffu

namespace orx {

// ---------------------------------------------------------------- score keys
// A candidate is one 64-bit key: high word = order-preserving image of the fp32
// fast score, low word = ~row.  A LARGER key is a BETTER candidate
// (score descending, then row ascending).  Key 0 is "empty".
//   ord 0xFFFFFFFF : "irregular" row (norm outside [2^-40, 2^40]) whose fast score
//                    is not trusted -> always a candidate, settled by the rescore.
//   ord 0x00000000 : zero-norm row, cosine distance NaN -> sorts after everything.
constexpr uint32_t ORD_ALWAYS = 0xFFFFFFFFu;
constexpr uint32_t ORD_NAN = 0u;
constexpr uint32_t ROW_INVALID = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t float_to_ord(float s) {
    uint32_t u = __float_as_uint(s);
    uint32_t o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    // keep the two sentinels free for regular scores (|s| <= ~1 for regular rows)
    o = min(max(o, 1u), 0xFFFFFFFEu);
    return o;
}
__device__ __forceinline__ float ord_to_float(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(u);
}
// score key of (dot * scale): scale is 1/|x| for a regular row, +inf for an
// irregular one, NaN for a zero-norm row (set by the upsert kernel).
__device__ __forceinline__ uint32_t score_ord(float dot, float scale) {
    float s = dot * scale;
    uint32_t o = float_to_ord(s);
    if (scale != scale) o = ORD_NAN;                       // zero-norm row
    else if (scale == __int_as_float(0x7f800000) || s != s) o = ORD_ALWAYS;
    return o;
}
__device__ __forceinline__ uint64_t make_key(uint32_t ord, uint32_t row) {
    return ((uint64_t)ord << 32) | (uint64_t)(~row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~(uint32_t)k; }
__device__ __forceinline__ uint32_t key_ord(uint64_t k) { return (uint32_t)(k >> 32); }

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL_MASK, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(FULL_MASK, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
    uint32_t lo = __shfl_up_sync(FULL_MASK, (uint32_t)v, d);
    uint32_t hi = __shfl_up_sync(FULL_MASK, (uint32_t)(v >> 32), d);
    return ((uint64_t)hi << 32) | lo;
}

// ------------------------------------------------ warp-distributed sorted top-K
// K = 32*S keys, sorted descending; position p lives in slot p/32 of lane p%32.
// All calls are warp-uniform (every lane passes the same key to insert()).
template <int S>
struct WarpTopK {
    uint64_t k[S];
    uint64_t thr;     // current K-th best (smallest kept key); warp-uniform

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < S; ++s) k[s] = 0ull;
        thr = 0ull;
    }
    // precondition: nk > thr, nk identical in all lanes
    __device__ __forceinline__ void insert(uint64_t nk, int lane) {
        bool placed = false;
        uint64_t carry = 0ull;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            uint64_t last = shfl_u64(k[s], 31);
            uint64_t up = shfl_up_u64(k[s], 1);
            if (!placed) {
                int pos = __popc(__ballot_sync(FULL_MASK, k[s] > nk));
                if (pos < 32) {
                    if (lane > pos) k[s] = up;
                    else if (lane == pos) k[s] = nk;
                    carry = last;
                    placed = true;
                }
            } else {
                k[s] = (lane == 0) ? carry : up;
                carry = last;
            }
        }
        thr = shfl_u64(k[S - 1], 31);
    }
    __device__ __forceinline__ void offer(uint64_t nk, int lane) {   // warp-uniform nk
        if (nk > thr) insert(nk, lane);
    }
    // every lane offers its own key (0 = nothing)
    __device__ __forceinline__ void offer_lanes(uint64_t mine, int lane) {
        unsigned m = __ballot_sync(FULL_MASK, mine > thr);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            uint64_t c = shfl_u64(mine, src);
            if (c > thr) insert(c, lane);
        }
    }
    // write the sorted list: dst[p], p = s*32 + lane
    __device__ __forceinline__ void store(uint64_t *dst, int lane) const {
#pragma unroll
        for (int s = 0; s < S; ++s) dst[s * 32 + lane] = k[s];
    }
};

// --------------------------------------------------- canonical binary64 tree sum
// The halving tree of oracle/cosine_topk.py:canon_sum over 1024 terms, for a warp
// in which lane l holds terms l + 32*j in p[j] (j = 0..31).  Result valid in lane 0
// (broadcast by the caller).  __dadd_rn forbids any re-association / contraction.
__device__ __forceinline__ double canon_tree_1024(double (&p)[32]) {
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
        for (int j = 0; j < h; ++j) p[j] = __dadd_rn(p[j], p[j + h]);
    }
    double v = p[0];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        double o = __shfl_down_sync(FULL_MASK, v, d);
        v = __dadd_rn(v, o);
    }
    return v;
}
__device__ __forceinline__ double bcast_lane0(double v) { return __shfl_sync(FULL_MASK, v, 0); }

// canonical cosine distance from the three canonical sums (pgvector cosine_distance
// [UPSTREAM vector.c] with binary64 accumulators): 1 - clamp(dot / sqrt(n2x*n2q))
__device__ __forceinline__ double canon_dist(double dot, double n2x, double n2q) {
    double sim = __ddiv_rn(dot, __dsqrt_rn(__dmul_rn(n2x, n2q)));
    if (sim > 1.0) sim = 1.0;
    else if (sim < -1.0) sim = -1.0;
    return __dsub_rn(1.0, sim);
}

// "a sorts before b" under the ordering contract (distance ASC, NaN last, id ASC)
__device__ __forceinline__ bool sorts_before(double da, uint64_t ahi, uint64_t alo,
                                             double db, uint64_t bhi, uint64_t blo) {
    bool na = da != da, nb = db != db;
    if (na != nb) return nb;
    if (!na && da != db) return da < db;
    if (ahi != bhi) return ahi < bhi;
    return alo < blo;
}

// ---------------------------------------------- programmatic dependent launch (PDL)
// The kernels of one search chain (prep -> scan -> finalize) are launched with the programmatic-stream-serialization
// attribute: a dependent grid may be scheduled while its predecessor still runs and parks at pdl_wait() until the
// predecessor has completed and its writes are visible, so the launch latency of each kernel hides behind the one
// before it.  Both instructions are no-ops for a kernel launched the plain way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------ streaming loads
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// element (l + 32*j) of a stored row as fp32, for either table dtype
template <typename T> __device__ __forceinline__ float row_elem(const T *row, int e);
template <> __device__ __forceinline__ float row_elem<float>(const float *row, int e) { return row[e]; }
template <> __device__ __forceinline__ float row_elem<__nv_bfloat16>(const __nv_bfloat16 *row, int e) {
    return __bfloat162float(row[e]);
}

}  // namespace orx
