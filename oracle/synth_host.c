/* Host (plain C, pthreads) generator of the synthetic "bge-m3-shaped" table -- TEST / BASELINE
 * INFRASTRUCTURE ONLY (data for the CPU arm of bench.py and for oracle checks at sizes where the NumPy
 * generator is too slow: it makes ~13 k rows/s, this one a few M rows/s).
 *
 * Bit-identical to orx_testkit/synth.py (Synth.rows) and orx_testkit/csrc/synth.cu: the same integer hash
 * (splitmix64 finaliser), Irwin-Hall(4) of the four 16-bit fields, and only exactly-rounded IEEE operations in
 * the same order -- fp32 mul/add, binary64 halving-tree norm (element i pairs with i + n/2), one fp32 scale.
 * Compiled with -ffp-contract=off and without any fast-math flag (oracle/Makefile) so that no FMA contraction
 * or re-association can change a bit.  tests/test_synth.py compares it with the NumPy generator.
 * SURVEY.md 8(d) defines the recipe; the stand-in for the remote bge-m3 service (reference
 * app/llm_services.py:218-222).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DIM 1024
static const uint64_t GOLD = 0x9E3779B97F4A7C15ull;

static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline float gauss_elem(uint64_t key, uint64_t vec, uint32_t d)
{
    const uint64_t h = mix64(key + (vec * (uint64_t) DIM + d) * GOLD);
    const int s = (int) (h & 0xFFFF) + (int) ((h >> 16) & 0xFFFF) + (int) ((h >> 32) & 0xFFFF) + (int) (h >> 48);
    union { uint32_t u; float f; } g = { 0x37ddb3d7u };          /* fp32(sqrt(3)/65536) */
    return (float) (s - 131070) * g.f;
}

/* y <- y * fp32(1/sqrt(canonical |y|^2)) */
static void normalize(float *y)
{
    double p[DIM];
    for (int i = 0; i < DIM; i++) p[i] = (double) y[i] * (double) y[i];
    for (int n = DIM / 2; n >= 1; n /= 2)
        for (int i = 0; i < n; i++) p[i] = p[i] + p[i + n];
    const float inv = (float) (1.0 / sqrt(p[0]));
    for (int i = 0; i < DIM; i++) y[i] = y[i] * inv;
}

/* unit vectors of stream `key`: vec_index0 .. vec_index0 + n - 1 (the mean and the centre table) */
void synth_host_unit(uint64_t key, uint64_t vec_index0, uint64_t n, float *dst)
{
    for (uint64_t v = 0; v < n; v++) {
        float *y = dst + v * DIM;
        for (uint32_t d = 0; d < DIM; d++) y[d] = gauss_elem(key, vec_index0 + v, d);
        normalize(y);
    }
}

struct job {
    uint64_t key_noise, key_cid;
    const float *mean, *centres;
    uint32_t n_centres;
    const uint64_t *index;          /* explicit row indices, or NULL: row_start + i */
    uint64_t row_start, i0, i1;
    float *dst;
};

static void *rows_worker(void *arg)
{
    const struct job *j = (const struct job *) arg;
    for (uint64_t i = j->i0; i < j->i1; i++) {
        const uint64_t row = j->index ? j->index[i] : j->row_start + i;
        const uint32_t cid = (uint32_t) (mix64(j->key_cid + row * GOLD) % j->n_centres);
        const float *c = j->centres + (size_t) cid * DIM;
        float *y = j->dst + i * DIM;
        for (uint32_t d = 0; d < DIM; d++) {
            const float g = gauss_elem(j->key_noise, row, d);
            const float a = 0.5f * j->mean[d] + 0.6f * c[d];       /* two products, one add: no contraction */
            y[d] = a + 0.019375f * g;                              /* fp32(0.62/32) */
        }
        normalize(y);
    }
    return NULL;
}

/* rows index[i] (or row_start + i when index is NULL), i < n, into dst [n, 1024]; `threads` workers */
int synth_host_rows(uint64_t key_noise, uint64_t key_cid, const float *mean, const float *centres, uint32_t n_centres,
                    const uint64_t *index, uint64_t row_start, uint64_t n, float *dst, int threads)
{
    if (n_centres == 0) return -1;
    if (threads < 1) threads = 1;
    if ((uint64_t) threads > n) threads = (int) (n ? n : 1);
    pthread_t *tid = (pthread_t *) malloc(sizeof(pthread_t) * threads);
    struct job *jobs = (struct job *) malloc(sizeof(struct job) * threads);
    if (!tid || !jobs) { free(tid); free(jobs); return -2; }
    const uint64_t per = (n + threads - 1) / threads;
    int started = 0;
    for (int t = 0; t < threads; t++) {
        struct job *j = &jobs[t];
        j->key_noise = key_noise; j->key_cid = key_cid; j->mean = mean; j->centres = centres; j->n_centres = n_centres;
        j->index = index; j->row_start = row_start; j->dst = dst;
        j->i0 = (uint64_t) t * per; j->i1 = j->i0 + per > n ? n : j->i0 + per;
        if (j->i0 >= j->i1) break;
        if (t == threads - 1 || pthread_create(&tid[t], NULL, rows_worker, j) != 0) { rows_worker(j); tid[t] = 0; }
        started = t + 1;
    }
    for (int t = 0; t < started; t++) if (tid[t]) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
    return 0;
}
