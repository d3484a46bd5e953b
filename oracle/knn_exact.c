/* TEST INFRASTRUCTURE ONLY - CPU oracle, never linked into the product library.
 *
 * Brute-force exact k-nearest-neighbours in float64, restating the arithmetic of the
 * third-party routine the reference calls (scikit-learn `kneighbors_graph`,
 * /root/reference/graph.py:709, 867, 887):
 *
 *   - data are float64 (sklearn upcasts float32 input; padded periodic clouds are
 *     float64 already, graph.py:827-855);
 *   - reduced distance  d = 0; for j in 0..2: tmp = q[j] - p[j]; d += tmp*tmp;
 *     (sklearn/metrics/_dist_metrics.pxd.tp, EuclideanDistance.rdist) - separate
 *     multiply and add, NO fused multiply-add (compile with -ffp-contract=off);
 *   - result rows ascending by distance.
 *
 * sklearn's order among exactly-equal distances is KD-tree traversal order and is not
 * reproducible; this oracle (and the CUDA kernel) use the documented total order
 * (d2, index).  Parity is asserted on tie-free inputs; `n_ties_out` reports, per call,
 * how many query rows have an exact tie at or across the k-th boundary or inside the
 * row, so tests can tell "tie-equivalent" rows from real mismatches.
 */
#include <stdint.h>
#include <stdlib.h>

static inline double rdist3(const double *q, const double *p) {
    double d = 0.0, t;
    t = q[0] - p[0]; d += t * t;
    t = q[1] - p[1]; d += t * t;
    t = q[2] - p[2]; d += t * t;
    return d;
}

/* x: (P,3) row-major float64.  Queries are rows [q_begin, q_end) of x.  k <= 64.
 * out_idx: (nq,k) int64, out_d2: (nq,k) float64 (may be NULL).
 * row_has_tie: (nq,) uint8 (may be NULL): 1 if any two of the k+1 best candidates of the
 * row have exactly equal d2 (so a different tie-break could change the row). */
int knn_exact(const double *x, int64_t P, int64_t q_begin, int64_t q_end, int k, int include_self,
              int64_t *out_idx, double *out_d2, uint8_t *row_has_tie) {
    if (k < 1 || k > 64 || q_end > P || q_begin < 0) return -1;
    if ((int64_t)k > P - (include_self ? 0 : 1)) return -2;
    const int kk = (k + 1 <= P - (include_self ? 0 : 1)) ? k + 1 : k; /* one extra to detect boundary ties */
    /* no OpenMP in this image: the Python wrapper splits [0,nq) over threads (ctypes drops the GIL) */
    for (int64_t i = q_begin; i < q_end; ++i) {
        double bd[65];
        int64_t bi[65];
        int n = 0;
        const double *q = x + 3 * i;
        for (int64_t j = 0; j < P; ++j) {
            if (!include_self && j == i) continue;
            double d = rdist3(q, x + 3 * j);
            if (n == kk && !(d < bd[n - 1])) continue; /* j ascending: equal d2 keeps lower index */
            int pos = (n < kk) ? n : kk - 1;
            while (pos > 0 && d < bd[pos - 1]) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = d;
            bi[pos] = j;
            if (n < kk) ++n;
        }
        for (int m = 0; m < k; ++m) {
            out_idx[i * k + m] = bi[m];
            if (out_d2) out_d2[i * k + m] = bd[m];
        }
        if (row_has_tie) {
            uint8_t t = 0;
            for (int m = 1; m < n; ++m)
                if (bd[m] == bd[m - 1]) t = 1;
            row_has_tie[i] = t;
        }
    }
    return 0;
}
