// knn.cu - exact k-nearest-neighbour graph on a (periodic) particle box, bit-compatible with the
// float64 KD-tree the reference calls (scikit-learn kneighbors_graph; /root/reference/graph.py:709,
// 887) and with the reference's padded-cube periodic images (graph.py:801-917).
//
// Algorithm (cell list):
//   1. bounding box per sample (non-periodic) or the unit box (periodic); G^3 uniform cells
//   2. counting sort of the particles by cell (count -> exclusive scan -> scatter); sorted records
//      are float4 {x, y, z, index | image-flags << 24}, x-fastest cell order so that a row of
//      neighbouring cells is ONE contiguous range of records
//   3. one thread per query, queries taken in sorted order (a warp = spatially adjacent queries,
//      so candidate loads hit the same lines): visit cell shells R = 0,1,2,... around the query's
//      cell; keep the k best in a register-resident sorted list; stop once the k-th best distance
//      is strictly inside the region already covered.
//
// Bit-exactness: distances are float64, tmp = q - p; d = ((tx*tx) + (ty*ty)) + (tz*tz) with
// separately rounded multiply/add (__dmul_rn/__dadd_rn: no FMA contraction), images are formed
// as (double)x + shift before the subtraction.  Total order on candidates = (d2, index); the
// result is therefore independent of the visiting order.
#include <stdlib.h>

#include "nbpc_common.cuh"
#include "scan.cuh"

#define KNN_IDX_MASK 0x00FFFFFF
#define KNN_FLAG_SHIFT 24
#define KNN_THREADS 128
#define KNN_PEND 4          // per-lane queue of accepted-but-not-yet-inserted candidates

struct KnnGridInfo {  // per sample, device
    float lo[3];
    float inv_h;
    double h;
};

__device__ __forceinline__ unsigned knn_enc_float(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float knn_dec_float(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__device__ __forceinline__ int knn_cell_coord(float x, float lo, float inv_h, int G) {
    int c = (int)floorf((x - lo) * inv_h);
    return nbpc_min(nbpc_max(c, 0), G - 1);
}

// ------------------------------------------------------------------ 1. bounding box
__global__ void knn_bbox_init(unsigned *mm, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 6) mm[i] = ((i % 6) < 3) ? 0xFFFFFFFFu : 0u;  // [min xyz | max xyz]
}

__global__ void knn_bbox_reduce(const float *__restrict__ xyz, int64_t sb, int64_t sn, int N, unsigned *mm) {
    const int b = blockIdx.y;
    const float *p = xyz + (int64_t)b * sb;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = p[(int64_t)i * sn + d];
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
#ifndef NBPC_HOST_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    }
    if ((threadIdx.x & 31) != 0) return;
#endif
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (mn[d] <= mx[d]) {
            atomicMin(&mm[b * 6 + d], knn_enc_float(mn[d]));
            atomicMax(&mm[b * 6 + 3 + d], knn_enc_float(mx[d]));
        }
    }
}

__global__ void knn_grid_setup(const unsigned *mm, int B, int G, int periodic, KnnGridInfo *info) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    KnnGridInfo gi;
    if (periodic) {
        gi.lo[0] = gi.lo[1] = gi.lo[2] = 0.f;
        gi.inv_h = (float)G;
        gi.h = 1.0 / (double)G;
    } else {
        float ext = 0.f;
        for (int d = 0; d < 3; ++d) {
            float lo = knn_dec_float(mm[b * 6 + d]), hi = knn_dec_float(mm[b * 6 + 3 + d]);
            gi.lo[d] = lo;
            ext = fmaxf(ext, hi - lo);
        }
        if (!(ext > 0.f)) ext = 1.f;
        double h = (double)ext * (1.0 + 1e-6) / (double)G;
        gi.h = h;
        gi.inv_h = (float)(1.0 / h);
    }
    info[b] = gi;
}

// ------------------------------------------------------------------ 2. counting sort by cell
__global__ void knn_count(const float *__restrict__ xyz, int64_t sb, int64_t sn, int B, int N, int G,
                          const KnnGridInfo *__restrict__ info, int32_t *__restrict__ cell_of_point,
                          int32_t *__restrict__ cell_count, int periodic, int32_t *__restrict__ status) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * N) return;
    const int b = (int)(t / N), i = (int)(t % N);
    const float *p = xyz + (int64_t)b * sb + (int64_t)i * sn;
    const KnnGridInfo gi = info[b];
    // periodic mode assumes the unit box (the shell lower bounds treat edge cells as ending at 0 / 1): count offenders
    if (periodic && status && !(p[0] >= 0.f && p[0] <= 1.f && p[1] >= 0.f && p[1] <= 1.f && p[2] >= 0.f && p[2] <= 1.f))
        atomicAdd(&status[0], 1);
    const int cx = knn_cell_coord(p[0], gi.lo[0], gi.inv_h, G);
    const int cy = knn_cell_coord(p[1], gi.lo[1], gi.inv_h, G);
    const int cz = knn_cell_coord(p[2], gi.lo[2], gi.inv_h, G);
    const int cell = b * G * G * G + (cz * G + cy) * G + cx;
    cell_of_point[t] = cell;
    atomicAdd(&cell_count[cell], 1);
}

__global__ void knn_scatter(const float *__restrict__ xyz, int64_t sb, int64_t sn, int B, int N,
                            int periodic, float lower_f, float upper_f,
                            const int32_t *__restrict__ cell_of_point, const int32_t *__restrict__ cell_start,
                            int32_t *__restrict__ cell_fill, float4 *__restrict__ sorted) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * N) return;
    const int b = (int)(t / N), i = (int)(t % N);
    const float *p = xyz + (int64_t)b * sb + (int64_t)i * sn;
    const float x = p[0], y = p[1], z = p[2];
    int flags = 0;
    if (periodic) {
        // graph.py:842  bound = where(x >= upper, -1, where(x <= lower, +1, 0)), float32 compare
        const int fx = (x >= upper_f) ? 2 : ((x <= lower_f) ? 1 : 0);
        const int fy = (y >= upper_f) ? 2 : ((y <= lower_f) ? 1 : 0);
        const int fz = (z >= upper_f) ? 2 : ((z <= lower_f) ? 1 : 0);
        flags = fx | (fy << 2) | (fz << 4);
    }
    const int cell = cell_of_point[t];
    const int pos = cell_start[cell] + atomicAdd(&cell_fill[cell], 1);
    sorted[pos] = make_float4(x, y, z, __int_as_float(i | (flags << KNN_FLAG_SHIFT)));
}

// ------------------------------------------------------------------ 3. query
template <int K>
struct KnnTopK {
    double d[K];
    int id[K];
    // slots [0, K-k) hold -inf sentinels so that d[K-1] is always the k-th best real entry
    __device__ __forceinline__ void init(int k) {
#pragma unroll
        for (int m = 0; m < K; ++m) {
            const bool sentinel = m < K - k;
            d[m] = sentinel ? -INFINITY : INFINITY;
            id[m] = sentinel ? -1 : 0x7FFFFFFF;
        }
    }
    __device__ __forceinline__ bool accepts(double dd, int ii) const {
        return dd < d[K - 1] || (dd == d[K - 1] && ii < id[K - 1]);
    }
    // precondition: accepts(dd, ii).  Fully predicated (no early exit): an early-exit loop makes the
    // compiler index the arrays dynamically, which moves them from registers to local memory.
    __device__ __forceinline__ void insert(double dd, int ii) {
        // the list is sorted, so "new sorts before slot m-1" implies "new sorts before slot m": with the precondition
        // (new sorts before slot K-1) slot m becomes  below[m-1] ? old[m-1] : (below[m] ? new : old[m]).
        // Bitwise | and & (no short-circuit) keep the comparison straight-line: 2 DSETP + 1 ISETP + 1 LOP per slot.
        bool below[K];
#pragma unroll
        for (int m = 0; m < K - 1; ++m) below[m] = (dd < d[m]) | ((dd == d[m]) & (ii < id[m]));
        below[K - 1] = true;
#pragma unroll
        for (int m = K - 1; m > 0; --m) {
            const double dn = below[m] ? dd : d[m];
            const int in = below[m] ? ii : id[m];
            d[m] = below[m - 1] ? d[m - 1] : dn;
            id[m] = below[m - 1] ? id[m - 1] : in;
        }
        d[0] = below[0] ? dd : d[0];
        id[0] = below[0] ? ii : id[0];
    }
    __device__ __forceinline__ void sort_ids_ascending() {  // bitonic network on the ids padded to a power of two
        constexpr int KP = K <= 8 ? 8 : (K <= 16 ? 16 : (K <= 32 ? 32 : 64));
        int t[KP];
#pragma unroll
        for (int m = 0; m < KP; ++m) t[m] = m < K ? id[m] : 0x7FFFFFFF;   // padding sorts last
        // slots [0, K-k) hold -1 sentinels, which sort first: the real ids stay in slots [K-k, K)
#pragma unroll
        for (int size = 2; size <= KP; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
                for (int m = 0; m < KP; ++m) {
                    const int p = m ^ stride;
                    if (p > m) {
                        const bool asc = (m & size) == 0;
                        const int a = t[m], b = t[p];
                        const bool sw = asc ? (a > b) : (a < b);
                        t[m] = sw ? b : a;
                        t[p] = sw ? a : b;
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < K; ++m) id[m] = t[m];
    }
};

struct KnnQueryParams {
    const float4 *sorted;
    const int32_t *cell_start;
    const KnnGridInfo *info;
    int32_t *idx_out;
    double *d2_out;
    int B, N, G, k, include_self, order;
    // thread-per-query kernels only: when set, thread t answers query todo[t] for t < *todo_count (the queries the
    // warp-cooperative kernel handed over) instead of query t
    const int32_t *todo, *todo_count;
};

template <int K, bool PERIODIC>
__global__ void __launch_bounds__(KNN_THREADS) knn_query(KnnQueryParams P) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (int64_t)P.B * P.N) return;
    const int b = (int)(s / P.N);
    const int G = P.G;
    const KnnGridInfo gi = P.info[b];
    const float4 me = P.sorted[s];
    const int my_id = __float_as_int(me.w) & KNN_IDX_MASK;
    const double px = (double)me.x, py = (double)me.y, pz = (double)me.z;
    const int cx = knn_cell_coord(me.x, gi.lo[0], gi.inv_h, G);
    const int cy = knn_cell_coord(me.y, gi.lo[1], gi.inv_h, G);
    const int cz = knn_cell_coord(me.z, gi.lo[2], gi.inv_h, G);
    // position inside the own cell in cell units (clamped to [0,1]: the cell index itself was clamped)
    const float frx = fminf(fmaxf((me.x - gi.lo[0]) * gi.inv_h - (float)cx, 0.f), 1.f);
    const float fry = fminf(fmaxf((me.y - gi.lo[1]) * gi.inv_h - (float)cy, 0.f), 1.f);
    const float frz = fminf(fmaxf((me.z - gi.lo[2]) * gi.inv_h - (float)cz, 0.f), 1.f);
    // distance (in cells) from the query to the nearest face of its own cell, in [0, 0.5]
    const float margin = fminf(fminf(fminf(frx, 1.f - frx), fminf(fry, 1.f - fry)), fminf(frz, 1.f - frz));
    const float inv_h2 = gi.inv_h * gi.inv_h * 1.0001f;
    // lower bound (cell units, minus a slack covering float cell-assignment fuzz) on the distance along one
    // axis between the query and any point of the cell at offset o
#define KNN_GAP(o, fr) fmaxf(((o) > 0 ? (float)(o) - (fr) : ((o) < 0 ? (float)(-(o)-1) + (fr) : 0.f)) - 1e-3f, 0.f)
    const int tmin = PERIODIC ? -G : 0, tmax = PERIODIC ? 2 * G - 1 : G - 1;  // extended cell coords
    int Rmax = nbpc_max(nbpc_max(cx - tmin, tmax - cx), nbpc_max(nbpc_max(cy - tmin, tmax - cy), nbpc_max(cz - tmin, tmax - cz)));
    const int64_t cellbase = (int64_t)b * G * G * G;
    const bool skip_self = !P.include_self;

    KnnTopK<K> top;
    top.init(P.k);

    for (int R = 0; R <= Rmax; ++R) {
        for (int dz = -R; dz <= R; ++dz) {
            int tz = cz + dz;
            if (tz < tmin || tz > tmax) continue;
            int sz = 0;
            if (PERIODIC) {
                if (tz < 0) { tz += G; sz = -1; } else if (tz >= G) { tz -= G; sz = 1; }
            }
            const bool zface = (dz == -R || dz == R);
            const float gz = KNN_GAP(dz, frz);
            if (gz * gz > (float)top.d[K - 1] * inv_h2) continue;   // whole plane is beyond the current k-th distance
            for (int dy = -R; dy <= R; ++dy) {
                int ty = cy + dy;
                if (ty < tmin || ty > tmax) continue;
                // prune with the CURRENT k-th distance (cell units): skip rows / row ends that cannot hold a closer point
                const float gy = KNN_GAP(dy, fry);
                const float rem = (float)top.d[K - 1] * inv_h2 - (gz * gz + gy * gy);
                if (rem < 0.f) continue;
                const float xr = fminf(sqrtf(rem) + 1e-3f, 8192.f);
                const int x_lo = (int)floorf((float)cx + frx - xr), x_hi = (int)floorf((float)cx + frx + xr);
                int sy = 0;
                if (PERIODIC) {
                    if (ty < 0) { ty += G; sy = -1; } else if (ty >= G) { ty -= G; sy = 1; }
                }
                const bool full_row = zface || dy == -R || dy == R;
                const int64_t rowbase = cellbase + ((int64_t)tz * G + ty) * G;
                // x ranges (extended coords): the whole run [cx-R, cx+R] on a shell face, otherwise its two ends
                const int nparts = full_row ? 1 : 2;
                for (int part = 0; part < nparts; ++part) {
                    int x0 = full_row ? cx - R : (part == 0 ? cx - R : cx + R);
                    int x1 = full_row ? cx + R : x0;
                    if (full_row) {
                        x0 = nbpc_max(x0, nbpc_max(tmin, x_lo));
                        x1 = nbpc_min(x1, nbpc_min(tmax, x_hi));
                    } else if (x0 < tmin || x0 > tmax || x0 < x_lo || x0 > x_hi) {
                        continue;  // a single end cell outside the (extended) grid or out of reach
                    }
                    // split into the shift -1 / 0 / +1 copies of the row
                    const int nseg = PERIODIC ? 3 : 1;
                    for (int seg = 0; seg < nseg; ++seg) {
                        const int sx = PERIODIC ? seg - 1 : 0;
                        const int lo = nbpc_max(x0, sx * G), hi = nbpc_min(x1, sx * G + G - 1);
                        if (lo > hi) continue;
                        const int jb = __ldg(&P.cell_start[rowbase + (lo - sx * G)]);
                        const int je = __ldg(&P.cell_start[rowbase + (hi - sx * G) + 1]);
                        // an image shifted by +1 exists only for particles flagged 1 (x <= lower),
                        // by -1 only for particles flagged 2 (x >= upper)
                        int req = 0, reqmask = 0;
                        if (PERIODIC) {
                            if (sx) { reqmask |= 3; req |= (sx > 0 ? 1 : 2); }
                            if (sy) { reqmask |= 3 << 2; req |= (sy > 0 ? 1 : 2) << 2; }
                            if (sz) { reqmask |= 3 << 4; req |= (sz > 0 ? 1 : 2) << 4; }
                        }
                        const bool unshifted = (sx | sy | sz) == 0;
                        const double ox = (double)sx, oy = (double)sy, oz = (double)sz;
                        for (int j = jb; j < je; ++j) {
                            const float4 c = __ldg(&P.sorted[j]);
                            const int w = __float_as_int(c.w);
                            if (PERIODIC && (((w >> KNN_FLAG_SHIFT) & reqmask) != req)) continue;
                            const int cid = w & KNN_IDX_MASK;
                            if (skip_self && unshifted && cid == my_id) continue;
                            const double tx = __dsub_rn(px, __dadd_rn((double)c.x, ox));
                            const double ty2 = __dsub_rn(py, __dadd_rn((double)c.y, oy));
                            const double tz2 = __dsub_rn(pz, __dadd_rn((double)c.z, oz));
                            const double dd = __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty2, ty2)), __dmul_rn(tz2, tz2));
                            if (top.accepts(dd, cid)) top.insert(dd, cid);
                        }
                    }
                }
            }
        }
        // everything not yet visited is at least (R + margin) cells away along some axis
        const double g = ((double)R + (double)margin - 1e-3) * gi.h;
        if (g > 0.0 && top.d[K - 1] < g * g) break;
    }

    const int64_t out = ((int64_t)b * P.N + my_id) * P.k;
    if (P.d2_out) {
#pragma unroll
        for (int m = 0; m < K; ++m)
            if (m >= K - P.k) P.d2_out[out + (m - (K - P.k))] = top.d[m];
    }
    if (P.order == NBPC_ORDER_INDEX) top.sort_ids_ascending();
#pragma unroll
    for (int m = 0; m < K; ++m)
        if (m >= K - P.k) P.idx_out[out + (m - (K - P.k))] = top.id[m];
}

// ------------------------------------------------------------------ 3b. query, warp-uniform control flow
// Same traversal, same pruning and the same total order as knn_query, but every loop has a warp-uniform trip count (the
// shell radius runs until the LAST lane of the warp is done, a row segment is walked for the longest lane's length,
// shorter or pruned lanes are predicated off), so the warp stays converged and the expensive sorted insertion can be
// DEFERRED: a candidate that beats the lane's (possibly stale) k-th best is parked in a 4-deep per-lane queue in
// shared memory, and the queues are drained by all lanes together when some lane's queue is full or a shell ends.
// In knn_query the insertion (~120 predicated instructions) ran whenever ANY lane accepted a candidate - on nearly
// every candidate with one or two lanes active (ncu: 10 of 32 lanes active per instruction on average).
#ifndef NBPC_HOST_EMU
template <int K, bool PERIODIC>
__global__ void __launch_bounds__(KNN_THREADS) knn_query_uniform(KnnQueryParams P) {
    constexpr unsigned FULLM = 0xffffffffu;
    __shared__ double pend_d_s[KNN_PEND * KNN_THREADS];
    __shared__ int pend_i_s[KNN_PEND * KNN_THREADS];
    double *pend_d = pend_d_s + threadIdx.x;
    int *pend_i = pend_i_s + threadIdx.x;
    int npend = 0;

    const int64_t total = P.todo ? (int64_t)*P.todo_count : (int64_t)P.B * P.N;
    if ((int64_t)blockIdx.x * blockDim.x >= total) return;
    const int64_t s_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s_raw < total;
    const int64_t s_lin = valid ? s_raw : total - 1;   // tail lanes shadow the last query and never write
    const int64_t s = P.todo ? (int64_t)P.todo[s_lin] : s_lin;
    const int b = (int)(s / P.N);
    const int G = P.G;
    const KnnGridInfo gi = P.info[b];
    const float4 me = P.sorted[s];
    const int my_id = __float_as_int(me.w) & KNN_IDX_MASK;
    const double px = (double)me.x, py = (double)me.y, pz = (double)me.z;
    const int cx = knn_cell_coord(me.x, gi.lo[0], gi.inv_h, G);
    const int cy = knn_cell_coord(me.y, gi.lo[1], gi.inv_h, G);
    const int cz = knn_cell_coord(me.z, gi.lo[2], gi.inv_h, G);
    const float frx = fminf(fmaxf((me.x - gi.lo[0]) * gi.inv_h - (float)cx, 0.f), 1.f);
    const float fry = fminf(fmaxf((me.y - gi.lo[1]) * gi.inv_h - (float)cy, 0.f), 1.f);
    const float frz = fminf(fmaxf((me.z - gi.lo[2]) * gi.inv_h - (float)cz, 0.f), 1.f);
    const float margin = fminf(fminf(fminf(frx, 1.f - frx), fminf(fry, 1.f - fry)), fminf(frz, 1.f - frz));
    const float inv_h2 = gi.inv_h * gi.inv_h * 1.0001f;
    const int tmin = PERIODIC ? -G : 0, tmax = PERIODIC ? 2 * G - 1 : G - 1;
    const int Rmax = nbpc_max(nbpc_max(cx - tmin, tmax - cx), nbpc_max(nbpc_max(cy - tmin, tmax - cy), nbpc_max(cz - tmin, tmax - cz)));
    const int64_t cellbase = (int64_t)b * G * G * G;
    const bool skip_self = !P.include_self;

    KnnTopK<K> top;
    top.init(P.k);
    auto flush = [&]() {
#pragma unroll
        for (int f = 0; f < KNN_PEND; ++f) {
            if (f < npend) {
                const double fd = pend_d[f * KNN_THREADS];
                const int fi = pend_i[f * KNN_THREADS];
                if (top.accepts(fd, fi)) top.insert(fd, fi);
            }
        }
        npend = 0;
    };

    bool done = !valid;
    for (int R = 0;; ++R) {
        const bool act = !done && R <= Rmax;
        if (!__any_sync(FULLM, act)) break;
        for (int dz = -R; dz <= R; ++dz) {
            int tz = cz + dz;
            bool plane_ok = act && tz >= tmin && tz <= tmax;
            int sz = 0;
            if (PERIODIC) {
                if (tz < 0) { tz += G; sz = -1; } else if (tz >= G) { tz -= G; sz = 1; }
            }
            const bool zface = (dz == -R || dz == R);
            const float gz = KNN_GAP(dz, frz);
            plane_ok = plane_ok && !(gz * gz > (float)top.d[K - 1] * inv_h2);
            if (!__any_sync(FULLM, plane_ok)) continue;
            for (int dy = -R; dy <= R; ++dy) {
                int ty = cy + dy;
                bool row_ok = plane_ok && ty >= tmin && ty <= tmax;
                const float gy = KNN_GAP(dy, fry);
                const float rem = (float)top.d[K - 1] * inv_h2 - (gz * gz + gy * gy);
                row_ok = row_ok && !(rem < 0.f);
                if (!__any_sync(FULLM, row_ok)) continue;
                const float xr = fminf(sqrtf(fmaxf(rem, 0.f)) + 1e-3f, 8192.f);
                const int x_lo = (int)floorf((float)cx + frx - xr), x_hi = (int)floorf((float)cx + frx + xr);
                int sy = 0;
                if (PERIODIC) {
                    if (ty < 0) { ty += G; sy = -1; } else if (ty >= G) { ty -= G; sy = 1; }
                }
                const bool full_row = zface || dy == -R || dy == R;   // warp-uniform
                const int64_t rowbase = cellbase + ((int64_t)tz * G + ty) * G;
                const int nparts = full_row ? 1 : 2;
                for (int part = 0; part < nparts; ++part) {
                    int x0 = full_row ? cx - R : (part == 0 ? cx - R : cx + R);
                    int x1 = full_row ? cx + R : x0;
                    bool part_ok = row_ok;
                    if (full_row) {
                        x0 = nbpc_max(x0, nbpc_max(tmin, x_lo));
                        x1 = nbpc_min(x1, nbpc_min(tmax, x_hi));
                    } else if (x0 < tmin || x0 > tmax || x0 < x_lo || x0 > x_hi) {
                        part_ok = false;
                    }
                    const int nseg = PERIODIC ? 3 : 1;
                    for (int seg = 0; seg < nseg; ++seg) {
                        const int sx = PERIODIC ? seg - 1 : 0;
                        const int lo = nbpc_max(x0, sx * G), hi = nbpc_min(x1, sx * G + G - 1);
                        int jb = 0, je = 0;
                        if (part_ok && lo <= hi) {
                            jb = __ldg(&P.cell_start[rowbase + (lo - sx * G)]);
                            je = __ldg(&P.cell_start[rowbase + (hi - sx * G) + 1]);
                        }
                        const int len = je - jb;
                        const int maxlen = __reduce_max_sync(FULLM, len);
                        if (maxlen <= 0) continue;
                        int req = 0, reqmask = 0;
                        if (PERIODIC) {
                            if (sx) { reqmask |= 3; req |= (sx > 0 ? 1 : 2); }
                            if (sy) { reqmask |= 3 << 2; req |= (sy > 0 ? 1 : 2) << 2; }
                            if (sz) { reqmask |= 3 << 4; req |= (sz > 0 ? 1 : 2) << 4; }
                        }
                        const bool unshifted = (sx | sy | sz) == 0;
                        const double ox = (double)sx, oy = (double)sy, oz = (double)sz;
                        float4 c_pre = make_float4(0.f, 0.f, 0.f, 0.f);   // candidate i + 1 is loaded while candidate i is evaluated
                        if (len > 0) c_pre = __ldg(&P.sorted[jb]);
                        for (int i = 0; i < maxlen; ++i) {
                            const float4 c = c_pre;
                            if (i + 1 < len) c_pre = __ldg(&P.sorted[jb + i + 1]);
                            if (i < len) {
                                const int w = __float_as_int(c.w);
                                const int cid = w & KNN_IDX_MASK;
                                bool take = !(PERIODIC && (((w >> KNN_FLAG_SHIFT) & reqmask) != req));
                                take = take && !(skip_self && unshifted && cid == my_id);
                                const double tx = __dsub_rn(px, __dadd_rn((double)c.x, ox));
                                const double ty2 = __dsub_rn(py, __dadd_rn((double)c.y, oy));
                                const double tz2 = __dsub_rn(pz, __dadd_rn((double)c.z, oz));
                                const double dd = __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty2, ty2)), __dmul_rn(tz2, tz2));
                                if (take && top.accepts(dd, cid)) {
                                    pend_d[npend * KNN_THREADS] = dd;
                                    pend_i[npend * KNN_THREADS] = cid;
                                    ++npend;
                                }
                            }
                            if (__any_sync(FULLM, npend == KNN_PEND)) flush();
                        }
                    }
                }
            }
        }
        if (__any_sync(FULLM, npend > 0)) flush();   // the termination test needs the true k-th best
        const double g = ((double)R + (double)margin - 1e-3) * gi.h;
        if (act && g > 0.0 && top.d[K - 1] < g * g) done = true;
        if (R >= Rmax) done = true;
    }

    if (!valid) return;
    const int64_t out = ((int64_t)b * P.N + my_id) * P.k;
    if (P.d2_out) {
#pragma unroll
        for (int m = 0; m < K; ++m)
            if (m >= K - P.k) P.d2_out[out + (m - (K - P.k))] = top.d[m];
    }
    if (P.order == NBPC_ORDER_INDEX) top.sort_ids_ascending();
#pragma unroll
    for (int m = 0; m < K; ++m)
        if (m >= K - P.k) P.idx_out[out + (m - (K - P.k))] = top.id[m];
}

// ------------------------------------------------------------------ 3c. query, one warp per query (lanes = candidates)
#include "knn_warp.cuh"
#endif

// ------------------------------------------------------------------ host side
static int knn_grid_cells(int N) {
    // ~1 particle per cell by default (NBPC_KNN_RHO overrides): with the per-row pruning above a query
    // only touches the cells that intersect its current k-th-neighbour ball
    static double rho = 0.0;
    if (rho == 0.0) {
        const char *e = getenv("NBPC_KNN_RHO");
        rho = e ? atof(e) : 1.0;
        if (!(rho > 0.01 && rho < 1000.0)) rho = 1.0;
    }
    int G = (int)floor(cbrt((double)N / rho));
    if (G < 1) G = 1;
    if (G > 256) G = 256;
    return G;
}

struct KnnWorkspace {
    unsigned *mm;
    KnnGridInfo *info;
    int32_t *cell_start;  // B*G^3 + 1
    int32_t *cell_fill;   // B*G^3
    int32_t *cell_of_point;
    float4 *sorted;
    int32_t *partials;
    int32_t *todo_count;  // [1] queries handed from the warp-cooperative kernel to the thread-per-query kernel
    size_t bytes;
};

static KnnWorkspace knn_carve(void *ws, size_t ws_bytes, int B, int N, int G) {
    NbpcArena a(ws, ws_bytes);
    KnnWorkspace w;
    const int64_t ncell = (int64_t)B * G * G * G;
    w.mm = a.take<unsigned>((size_t)B * 6);
    w.info = a.take<KnnGridInfo>((size_t)B);
    w.cell_start = a.take<int32_t>((size_t)ncell + 1);
    w.cell_fill = a.take<int32_t>((size_t)ncell);
    w.cell_of_point = a.take<int32_t>((size_t)B * N);
    w.sorted = a.take<float4>((size_t)B * N);
    w.partials = a.take<int32_t>(nbpc_scan_partials_count(ncell + 1));
    w.todo_count = a.take<int32_t>(4);
    w.bytes = a.off;
    return w;
}

// NBPC_KNN_V: 1 = per-thread traversal, 2 = warp-uniform thread-per-query kernel with deferred insertion,
// 3 = warp-per-query kernel (k <= 32) with kernel 2 answering the queries it hands over, 0 / unset = by k
static int knn_version_from_env() {
    const char *e = getenv("NBPC_KNN_V");
    const int v = e ? atoi(e) : NBPC_KNN_AUTO;
    return (v >= 0 && v <= 3) ? v : NBPC_KNN_AUTO;
}
static int g_knn_version = knn_version_from_env();

template <int K>
static int knn_launch_query(KnnQueryParams P, int periodic, int32_t *todo, int32_t *todo_count, cudaStream_t stream) {
    const int64_t nq = (int64_t)P.B * P.N;
    const int grid = nbpc_cdiv(nq, KNN_THREADS);
    P.todo = nullptr;
    P.todo_count = nullptr;
#ifndef NBPC_HOST_EMU
    // measured crossover: the warp-per-query kernel wins from the 32-wide list on (k > 16), see DESIGN.md 4.1
    const int version = g_knn_version != NBPC_KNN_AUTO ? g_knn_version : (P.k > 16 ? NBPC_KNN_WARP : NBPC_KNN_THREAD);
    void (*kern2)(KnnQueryParams) = periodic ? knn_query_uniform<K, true> : knn_query_uniform<K, false>;
    if (version == 3 && P.k <= 32) {
        if (nbpc_memset_async(todo_count, 0, sizeof(int32_t), stream)) return 1;
        void (*kern3)(KnnQueryParams, int32_t *, int32_t *) = periodic ? knn_query_warp<true> : knn_query_warp<false>;
        NBPC_LAUNCH_N("knn_query_warp", kern3, dim3(nbpc_cdiv(P.N, KNW_WARPS), P.B), KNW_THREADS, 0, stream, P, todo, todo_count);
        P.todo = todo;
        P.todo_count = todo_count;
        NBPC_LAUNCH_N("knn_query_todo", kern2, grid, KNN_THREADS, 0, stream, P);
        return 0;
    }
    if (version >= 2) {
        NBPC_LAUNCH_N("knn_query", kern2, grid, KNN_THREADS, 0, stream, P);
        return 0;
    }
#endif
    void (*kern)(KnnQueryParams) = periodic ? knn_query<K, true> : knn_query<K, false>;
    NBPC_LAUNCH_N("knn_query", kern, grid, KNN_THREADS, 0, stream, P);
    return 0;
}

extern "C" {

int nbpc_set_knn_kernel(int kernel) {
    if (kernel != NBPC_KNN_AUTO && kernel != 1 && kernel != NBPC_KNN_THREAD && kernel != NBPC_KNN_WARP) {
        nbpc_set_error("nbpc_set_knn_kernel: unknown kernel");
        return NBPC_EINVAL;
    }
    g_knn_version = kernel;
    return NBPC_OK;
}

int nbpc_get_knn_kernel(void) { return g_knn_version; }

size_t nbpc_knn_workspace_bytes(int B, int N, int k, int periodic) {
    (void)k; (void)periodic;
    if (B < 1 || N < 1) return 0;
    return knn_carve(nullptr, 0, B, N, knn_grid_cells(N)).bytes;
}

int nbpc_knn(const float *xyz, int64_t stride_b, int64_t stride_n, int B, int N, int k, int periodic,
             double boundary_threshold, int include_self, int order, int32_t *idx_out, double *d2_out,
             int32_t *status, void *workspace, size_t ws_bytes, void *stream_) {
    NBPC_TRY(nbpc_require_sm100());
    cudaStream_t stream = (cudaStream_t)stream_;
    NBPC_ARG(xyz && idx_out && workspace, "null pointer");
    NBPC_ARG(B >= 1 && N >= 1, "B and N must be positive");
    NBPC_ARG(N <= KNN_IDX_MASK, "N must be < 2^24");
    NBPC_ARG((int64_t)B * N < (int64_t)1 << 31, "B*N must fit int32");
    NBPC_ARG(k >= 1 && k <= NBPC_KNN_MAX_K, "k must be in [1, 64]");
    NBPC_ARG(k <= N - (include_self ? 0 : 1), "k exceeds the number of available neighbours");
    NBPC_ARG(stride_n >= 3, "stride_n must be >= 3");
    NBPC_ARG(order == NBPC_ORDER_DISTANCE || order == NBPC_ORDER_INDEX, "bad order");
    NBPC_ARG(!periodic || (boundary_threshold >= 0.0 && boundary_threshold <= 1.0), "boundary_threshold must be in [0,1]");
    const int G = knn_grid_cells(N);
    KnnWorkspace w = knn_carve(workspace, ws_bytes, B, N, G);
    if (w.bytes > ws_bytes) {
        nbpc_set_error("nbpc_knn: workspace too small");
        return NBPC_EWORKSPACE;
    }
    const int64_t ncell = (int64_t)B * G * G * G;
    const int64_t P_ = (int64_t)B * N;

    if (!periodic) {
        NBPC_LAUNCH(knn_bbox_init, nbpc_cdiv(B * 6, 64), 64, 0, stream, w.mm, B);
        const int bx = nbpc_min(nbpc_cdiv(N, 256 * 8), 256);
        NBPC_LAUNCH(knn_bbox_reduce, dim3(bx, B), 256, 0, stream, xyz, stride_b, stride_n, N, w.mm);
    }
    NBPC_LAUNCH(knn_grid_setup, nbpc_cdiv(B, 64), 64, 0, stream, w.mm, B, G, periodic, w.info);
    if (nbpc_memset_async(w.cell_start, 0, sizeof(int32_t) * (size_t)(ncell + 1), stream) ||
        nbpc_memset_async(w.cell_fill, 0, sizeof(int32_t) * (size_t)ncell, stream)) {
        nbpc_set_error("nbpc_knn: memset failed");
        return NBPC_ELAUNCH;
    }
    if (status && nbpc_memset_async(status, 0, sizeof(int32_t), stream)) {
        nbpc_set_error("nbpc_knn: memset failed");
        return NBPC_ELAUNCH;
    }
    NBPC_LAUNCH(knn_count, nbpc_cdiv(P_, 256), 256, 0, stream, xyz, stride_b, stride_n, B, N, G, w.info,
                w.cell_of_point, w.cell_start, periodic, status);
    NBPC_TRY(nbpc_exclusive_scan_i32(w.cell_start, ncell + 1, w.partials, stream));
    const float lower_f = (float)boundary_threshold, upper_f = (float)(1.0 - boundary_threshold);
    NBPC_LAUNCH(knn_scatter, nbpc_cdiv(P_, 256), 256, 0, stream, xyz, stride_b, stride_n, B, N, periodic, lower_f,
                upper_f, w.cell_of_point, w.cell_start, w.cell_fill, w.sorted);

    KnnQueryParams P;
    P.sorted = w.sorted; P.cell_start = w.cell_start; P.info = w.info;
    P.idx_out = idx_out; P.d2_out = d2_out;
    P.B = B; P.N = N; P.G = G; P.k = k; P.include_self = include_self; P.order = order;
    P.todo = nullptr; P.todo_count = nullptr;
    // cell_of_point is dead after the scatter: it becomes the todo list of the warp-per-query kernel
    int rc;
    if (k <= 8) rc = knn_launch_query<8>(P, periodic, w.cell_of_point, w.todo_count, stream);
    // the reference's default k = 14 gets its own list length (13 % faster than the padded 16-wide list at 8 x 32^3)
    else if (k <= 14) rc = knn_launch_query<14>(P, periodic, w.cell_of_point, w.todo_count, stream);
    else if (k <= 16) rc = knn_launch_query<16>(P, periodic, w.cell_of_point, w.todo_count, stream);
    else if (k <= 32) rc = knn_launch_query<32>(P, periodic, w.cell_of_point, w.todo_count, stream);
    else rc = knn_launch_query<64>(P, periodic, w.cell_of_point, w.todo_count, stream);
    if (rc) {
        nbpc_set_error("nbpc_knn: memset failed");
        return NBPC_ELAUNCH;
    }
    return nbpc_check_launch("nbpc_knn");
}

}  // extern "C"
