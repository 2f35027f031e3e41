/* C restatement of the pgvector sequential-scan arithmetic -- TEST / BASELINE
 * INFRASTRUCTURE ONLY (see oracle/cosine_topk.py header; parity unpinned).
 *
 * Follows the published algorithm of pgvector `src/vector.c` [UPSTREAM, not in
 * /root/reference; Docker tag pgvector/pgvector:pg16, reference README.md:139]:
 *   cosine_distance(a, b): float accumulators `similarity += a[i]*b[i]`,
 *   `norma += a[i]*a[i]`, `normb += b[i]*b[i]`; then in double
 *   `similarity / sqrt(norma * normb)`, clamp to [-1, 1], return 1 - similarity.
 * and of the executor's bounded top-N heapsort for `ORDER BY ... LIMIT k`
 * (SQL emitted by langchain-postgres 0.0.16 for reference app/rag.py:85-87).
 * Built with pgvector's own optimisation flags (oracle/Makefile).  Ties are
 * broken by row index ascending (the build's ordering contract); NaN sorts last.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

double pgv_cosine_distance(int dim, const float *a, const float *b)
{
    float similarity = 0.0f, norma = 0.0f, normb = 0.0f;
    for (int i = 0; i < dim; i++) {
        similarity += a[i] * b[i];
        norma += a[i] * a[i];
        normb += b[i] * b[i];
    }
    double s = (double) similarity / sqrt((double) norma * (double) normb);
    if (s > 1.0) s = 1.0;
    else if (s < -1.0) s = -1.0;
    return 1.0 - s;
}

/* "a sorts before b": distance ascending, NaN last, row ascending */
static inline int before(double da, int64_t ia, double db, int64_t ib)
{
    int na = isnan(da), nb = isnan(db);
    if (na != nb) return nb;
    if (!na && da != db) return da < db;
    return ia < ib;
}

/* Sequential scan of rows [row0, row1) keeping the k best in a bounded max-heap
 * (worst on top), then heap-sorted ascending.  Returns the number of results. */
int pgv_scan_topk(const float *X, int64_t row0, int64_t row1, int dim, const float *q,
                  int k, int64_t *out_row, double *out_dist)
{
    int m = 0;
    if (k <= 0) return 0;
    for (int64_t r = row0; r < row1; r++) {
        double d = pgv_cosine_distance(dim, X + (size_t) r * dim, q);
        if (m < k) {
            int i = m++;
            out_row[i] = r; out_dist[i] = d;
            while (i > 0) {                       /* sift up: parent must sort after child */
                int p = (i - 1) / 2;
                if (before(out_dist[p], out_row[p], out_dist[i], out_row[i])) {
                    double td = out_dist[p]; out_dist[p] = out_dist[i]; out_dist[i] = td;
                    int64_t tr = out_row[p]; out_row[p] = out_row[i]; out_row[i] = tr;
                    i = p;
                } else break;
            }
        } else if (before(d, r, out_dist[0], out_row[0])) {
            int i = 0;
            out_row[0] = r; out_dist[0] = d;
            for (;;) {                            /* sift down */
                int l = 2 * i + 1, rr = l + 1, w = i;
                if (l < m && before(out_dist[w], out_row[w], out_dist[l], out_row[l])) w = l;
                if (rr < m && before(out_dist[w], out_row[w], out_dist[rr], out_row[rr])) w = rr;
                if (w == i) break;
                double td = out_dist[w]; out_dist[w] = out_dist[i]; out_dist[i] = td;
                int64_t tr = out_row[w]; out_row[w] = out_row[i]; out_row[i] = tr;
                i = w;
            }
        }
    }
    /* insertion sort of the <= k survivors */
    for (int i = 1; i < m; i++) {
        double d = out_dist[i]; int64_t r = out_row[i]; int j = i - 1;
        while (j >= 0 && before(d, r, out_dist[j], out_row[j])) {
            out_dist[j + 1] = out_dist[j]; out_row[j + 1] = out_row[j]; j--;
        }
        out_dist[j + 1] = d; out_row[j + 1] = r;
    }
    return m;
}

/* All host threads: each scans a contiguous slice (like parallel seq scan
 * workers under a Gather Merge), slices merged with the same ordering. */
struct slice { const float *X, *q; int64_t a, b; int dim, k, cnt; int64_t *rows; double *dist; };

static void *slice_main(void *p)
{
    struct slice *s = (struct slice *) p;
    s->cnt = pgv_scan_topk(s->X, s->a, s->b, s->dim, s->q, s->k, s->rows, s->dist);
    return 0;
}

int pgv_scan_topk_mt(const float *X, int64_t n, int dim, const float *q, int k,
                     int n_threads, int64_t *out_row, double *out_dist)
{
    if (n_threads < 1) n_threads = 1;
    if (k <= 0) return 0;
    int64_t *rows = (int64_t *) malloc(sizeof(int64_t) * (size_t) k * n_threads);
    double *dist = (double *) malloc(sizeof(double) * (size_t) k * n_threads);
    struct slice *sl = (struct slice *) calloc(n_threads, sizeof(struct slice));
    pthread_t *th = (pthread_t *) calloc(n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        struct slice s = { X, q, n * t / n_threads, n * (t + 1) / n_threads, dim, k, 0,
                           rows + (size_t) t * k, dist + (size_t) t * k };
        sl[t] = s;
        if (t > 0) pthread_create(&th[t], 0, slice_main, &sl[t]);
    }
    slice_main(&sl[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], 0);
    int m = 0;
    for (int t = 0; t < n_threads; t++)
        for (int i = 0; i < sl[t].cnt; i++) {
            double d = sl[t].dist[i]; int64_t r = sl[t].rows[i];
            if (m < k) { out_dist[m] = d; out_row[m] = r; m++; }
            else if (before(d, r, out_dist[m - 1], out_row[m - 1])) { out_dist[m - 1] = d; out_row[m - 1] = r; }
            else continue;
            for (int j = m - 1; j > 0 && before(out_dist[j], out_row[j], out_dist[j - 1], out_row[j - 1]); j--) {
                double td = out_dist[j]; out_dist[j] = out_dist[j - 1]; out_dist[j - 1] = td;
                int64_t tr = out_row[j]; out_row[j] = out_row[j - 1]; out_row[j - 1] = tr;
            }
        }
    free(rows); free(dist); free(sl); free(th);
    return m;
}
