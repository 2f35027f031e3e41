// Row-dot helpers shared by the GEMV scan and the threshold-collect fallback, so that a
// row's fast score is produced by the SAME instruction sequence in both (identical rows
// get identical fast scores; the collect pass sees exactly what the scan saw).
#pragma once
#include "common.cuh"

namespace orx {

template <typename T> struct RowVec;
template <> struct RowVec<float> {          // 4096 B row = 256 x 16 B
    static constexpr int NV = 8;            // 16-byte vectors per lane
    static constexpr int VEC_PER_ROW = 256;
};
template <> struct RowVec<__nv_bfloat16> {  // 2048 B row = 128 x 16 B
    static constexpr int NV = 4;
    static constexpr int VEC_PER_ROW = 128;
};

// q slice of one lane as 8 float4: for fp32 rows qv[j] pairs with vector (lane + 32 j);
// for bf16 rows qv[2j], qv[2j+1] pair with the 8 elements of vector (lane + 32 j).
template <typename T>
__device__ __forceinline__ void load_q_slice(const float *qhat, int lane, float4 (&qv)[8]) {
    const float4 *q4 = reinterpret_cast<const float4 *>(qhat);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[j] = q4[lane + 32 * j];
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            qv[2 * j] = q4[2 * (lane + 32 * j)];
            qv[2 * j + 1] = q4[2 * (lane + 32 * j) + 1];
        }
    }
}

template <typename T>
__device__ __forceinline__ void load_row_vecs(const uint4 *tab, uint32_t row, int lane,
                                              uint4 (&v)[RowVec<T>::NV]) {
    const uint4 *p = tab + (size_t)row * RowVec<T>::VEC_PER_ROW + lane;
#pragma unroll
    for (int j = 0; j < RowVec<T>::NV; ++j) v[j] = ldg_stream_u4(p + 32 * j);
}

// per-lane fp32 FMA chain (depth 32), then the xor butterfly (depth 5): every lane ends
// with the same dot.  |dot - exact| <= gamma_37 * sum|x_i q_i|  (DESIGN.md "Exactness").
template <typename T>
__device__ __forceinline__ float warp_row_dot(const uint4 (&v)[RowVec<T>::NV], const float4 (&qv)[8]) {
    float acc = 0.f;
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc = fmaf(__uint_as_float(v[j].x), qv[j].x, acc);
            acc = fmaf(__uint_as_float(v[j].y), qv[j].y, acc);
            acc = fmaf(__uint_as_float(v[j].z), qv[j].z, acc);
            acc = fmaf(__uint_as_float(v[j].w), qv[j].w, acc);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc = fmaf(bf16lo_to_f32(v[j].x), qv[2 * j].x, acc);
            acc = fmaf(bf16hi_to_f32(v[j].x), qv[2 * j].y, acc);
            acc = fmaf(bf16lo_to_f32(v[j].y), qv[2 * j].z, acc);
            acc = fmaf(bf16hi_to_f32(v[j].y), qv[2 * j].w, acc);
            acc = fmaf(bf16lo_to_f32(v[j].z), qv[2 * j + 1].x, acc);
            acc = fmaf(bf16hi_to_f32(v[j].z), qv[2 * j + 1].y, acc);
            acc = fmaf(bf16lo_to_f32(v[j].w), qv[2 * j + 1].z, acc);
            acc = fmaf(bf16hi_to_f32(v[j].w), qv[2 * j + 1].w, acc);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, d);
    return acc;
}

// canonical binary64 dot of a stored row with the ORIGINAL fp32 query (warp-wide; lane l
// owns elements l + 32 j).  Result broadcast to all lanes.
template <typename T>
__device__ __forceinline__ double warp_canon_dot(const T *row, const float *q, int lane) {
    double p[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int e = lane + 32 * j;
        p[j] = __dmul_rn((double)row_elem<T>(row, e), (double)q[e]);   // exact product
    }
    return bcast_lane0(canon_tree_1024(p));
}

}  // namespace orx
