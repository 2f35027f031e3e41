// TEST / BENCH TOOLING, not part of the product library (liborx.so does not contain it).
// Device side of the synthetic "bge-m3-shaped" generator -- bit-identical to
// orx_testkit/synth.py (same integer hash, same exactly-rounded fp32/fp64 steps in the
// same order; __f*_rn / __d*_rn intrinsics forbid FMA contraction).  Used by bench.py and
// the -m gpu tests to build 1M..100M-row tables in HBM without a 41 GB host upload
// (SURVEY.md 8d).  Stand-in for the remote bge-m3 service (reference app/llm_services.py:218-222).
// Built into orx_testkit/liborx_synth.so (orx_testkit/Makefile); C entry point: orxtk_synth_rows.
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "../../outline_rag_b200/csrc/common.cuh"

namespace orx {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
constexpr uint64_t GOLD = 0x9E3779B97F4A7C15ull;

// Irwin-Hall(4) of the four 16-bit fields of one hash -> ~N(0,1) fp32
__device__ __forceinline__ float gauss_elem(uint64_t key, uint64_t vec_index, uint32_t d) {
    const uint64_t h = mix64(key + (vec_index * (uint64_t)ORX_DIM + d) * GOLD);
    const int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)(h >> 48);
    return __fmul_rn((float)(s - 131070), __uint_as_float(0x37ddb3d7u) /* fp32(sqrt(3)/65536) */);
}

// normalise the 32 values a lane holds (elements lane + 32 j) with the canonical binary64 norm
__device__ __forceinline__ void warp_normalize(float (&y)[32]) {
    double p[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) p[j] = __dmul_rn((double)y[j], (double)y[j]);
    const double n2 = bcast_lane0(canon_tree_1024(p));
    const float inv = __double2float_rn(__ddiv_rn(1.0, __dsqrt_rn(n2)));
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = __fmul_rn(y[j], inv);
}

__global__ void __launch_bounds__(256)
synth_unit_kernel(uint64_t key, uint32_t n_vec, float *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t n_gw = gridDim.x * (blockDim.x >> 5);
    for (uint32_t v = gw; v < n_vec; v += n_gw) {
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = gauss_elem(key, v, lane + 32 * j);
        warp_normalize(y);
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[(size_t)v * ORX_DIM + lane + 32 * j] = y[j];
    }
}

__global__ void __launch_bounds__(256)
synth_rows_kernel(uint64_t key_noise, uint64_t key_cid, const float *__restrict__ mean,
                  const float *__restrict__ centres, uint32_t n_centres, uint64_t row_start,
                  uint64_t n_rows, float *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const uint64_t gw = blockIdx.x * (uint64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t n_gw = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = gw; i < n_rows; i += n_gw) {
        const uint64_t row = row_start + i;
        const uint32_t cid = (uint32_t)(mix64(key_cid + row * GOLD) % n_centres);
        const float *c = centres + (size_t)cid * ORX_DIM;
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int e = lane + 32 * j;
            const float g = gauss_elem(key_noise, row, e);
            const float a = __fadd_rn(__fmul_rn(0.5f, mean[e]), __fmul_rn(0.6f, c[e]));
            y[j] = __fadd_rn(a, __fmul_rn(0.019375f /* fp32(0.62/32) */, g));
        }
        warp_normalize(y);
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[i * ORX_DIM + lane + 32 * j] = y[j];
    }
}

static uint32_t grid_cap() {          // 8 CTAs per SM of the current device
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 1;
    return (uint32_t)sms * 8u;
}

void launch_synth_unit(uint64_t key, uint32_t n_vec, float *dst, cudaStream_t st) {
    if (n_vec == 0) return;
    uint32_t blocks = (n_vec + 7) / 8;
    if (blocks > grid_cap()) blocks = grid_cap();
    synth_unit_kernel<<<blocks, 256, 0, st>>>(key, n_vec, dst);
}

void launch_synth_rows(uint64_t key_noise, uint64_t key_cid, const float *mean, const float *centres,
                       uint32_t n_centres, uint64_t row_start, uint64_t n_rows, float *dst,
                       cudaStream_t st) {
    if (n_rows == 0) return;
    uint64_t blocks = (n_rows + 7) / 8;
    if (blocks > grid_cap()) blocks = grid_cap();
    synth_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(key_noise, key_cid, mean, centres, n_centres,
                                                        row_start, n_rows, dst);
}

uint64_t synth_stream_key(uint64_t seed, uint64_t tag) { return mix64(seed ^ tag); }

}  // namespace orx

// ------------------------------------------------------------------------------- C entry point
namespace {
struct SynthState {
    float *mean = nullptr, *centres = nullptr;
};
std::mutex g_mu;
std::map<std::tuple<int, uint64_t, uint32_t>, SynthState> g_state;
thread_local std::string g_err;
int fail(const char *what, cudaError_t e) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return -4;
}
}  // namespace

extern "C" const char *orxtk_last_error(void) { return g_err.c_str(); }

// Fills dst_device [n_rows, 1024] fp32 with rows row_start .. row_start+n_rows-1 of the synthetic
// table (bit-identical to orx_testkit/synth.py: Synth(n_centres, seed).rows(...)).  The mean and the
// centre table are built on first use per (device, seed, n_centres).  0 = OK.
extern "C" int orxtk_synth_rows(int device, void *cuda_stream, uint64_t seed, uint32_t n_centres, uint64_t row_start,
                                uint64_t n_rows, float *dst_device) {
    if (n_centres == 0) { g_err = "n_centres must be > 0"; return -1; }
    if (n_rows == 0) return 0;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail("cudaSetDevice", e);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    // stream tags: orx_testkit/synth.py TAG_*
    const uint64_t k_mean = orx::synth_stream_key(seed, 0x6D65616Eull);
    const uint64_t k_centre = orx::synth_stream_key(seed, 0x63656E74ull);
    const uint64_t k_noise = orx::synth_stream_key(seed, 0x6E6F6973ull);
    const uint64_t k_cid = orx::synth_stream_key(seed, 0x63696421ull);
    SynthState ss;
    int rc = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto key = std::make_tuple(device, seed, n_centres);
        auto it = g_state.find(key);
        if (it == g_state.end()) {
            e = cudaMalloc(&ss.mean, ORX_DIM * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc(&ss.centres, (size_t)n_centres * ORX_DIM * sizeof(float));
            if (e == cudaSuccess) {
                orx::launch_synth_unit(k_mean, 1, ss.mean, st);
                orx::launch_synth_unit(k_centre, n_centres, ss.centres, st);
                e = cudaStreamSynchronize(st);
            }
            if (e != cudaSuccess) rc = fail("synth state", e);
            else g_state.emplace(key, ss);
        } else ss = it->second;
    }
    if (rc == 0) {
        orx::launch_synth_rows(k_noise, k_cid, ss.mean, ss.centres, n_centres, row_start, n_rows, dst_device, st);
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail("synth_rows launch", e);
    }
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}
