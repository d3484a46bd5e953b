// knn_warp.cuh - warp-cooperative exact kNN query: ONE WARP PER QUERY, lanes = candidates.
//
// Same cell list, same shells, same row pruning and the same result (bit for bit) as knn_query_uniform; what changes is
// who does the work.  The thread-per-query kernels spend ~60 % of their instructions in the sorted insertion (~140
// predicated instructions, executed by the whole warp whenever any lane has something to insert).  Here
//   * the rows of a shell are spread over the lanes (one row each); their record ranges are concatenated into a staging
//     list in shared memory (warp scan), so that every batch of 32 candidates keeps 32 lanes busy;
//   * a candidate costs ONE FP32 distance (the "key").  The keys of a batch are sorted across the lanes with a bitonic
//     network of SHFL + FMNMX pairs (no payload) and merged into the lane-distributed list of the 32 smallest keys seen
//     so far; T = the k-th entry of that list is the pruning / termination bound;
//   * candidates whose key is within the FP32 error margin of T (knw_thr) are remembered; when the search ends, those
//     still within the margin of the final T (k of them unless there are near-ties) are re-evaluated in FP64 with the
//     reference's operation order (one per lane, two per lane when there are 33..64), and the exact (d2, index) order
//     decides.  More than 64 such candidates (massive ties: exact lattice shells, duplicated points) or a keep-list
//     overflow send the query to the thread-per-query kernel through a todo list - the result never depends on which
//     kernel produced it.
//
// Error margin.  u = 2^-24.  Open box: t_f = fl(p - c) has relative error u, key = fma chain of three non-negative
// terms: |key - d2| <= 6u d2.  The k smallest keys are <= T, so the exact k-th distance d2* <= T / (1 - 6u), and any
// candidate with d2 <= d2* has key <= T (1 + 6u) / (1 - 6u) = T (1 + 7.2e-7); the margin used is T (1 + 2e-6).  Periodic
// images: t_f = fl(fl(p - c) - s), |error| <= u (1 + 2|t|) per axis for ANY coordinates, i.e. sqrt(key) is within
// sqrt(T) (1 + 14u) + 3.7u = sqrt(T) (1 + 8.3e-7) + 2.2e-7; the margin used is (sqrt(T) (1 + 2e-6) + 5e-7)^2.  A wider
// margin only costs extra FP64 evaluations (and, for k = 32, more queries with more than 32 survivors).
#pragma once

#define KNW_WARPS 8
#define KNW_THREADS (KNW_WARPS * 32)
#define KNW_STAGE 128      // staged candidate references per warp
#define KNW_KEEP 160       // remembered (key, reference) pairs per warp
#define KNW_FULL 0xffffffffu
#define KNW_DIRECT 24     // runs of at least this many records are evaluated straight from the record array

template <bool PERIODIC>
__device__ __forceinline__ float knw_thr(float T) {
    if (PERIODIC) {
        const float a = sqrtf(T) * 1.000002f + 5e-7f;
        return a * a * 1.000001f;
    }
    return T * 1.000002f + 1e-37f;
}

// one compare-exchange of the bitonic networks below: the lane whose bit `bit` is clear keeps the smaller value
__device__ __forceinline__ float knw_cx(float v, int xor_mask, int bit, int lane) {
    const float o = __shfl_xor_sync(KNW_FULL, v, xor_mask);
    return (lane & bit) == 0 ? fminf(v, o) : fmaxf(v, o);
}
__device__ __forceinline__ int knw_cx(int v, int xor_mask, int bit, int lane) {
    const int o = __shfl_xor_sync(KNW_FULL, v, xor_mask);
    return (lane & bit) == 0 ? min(v, o) : max(v, o);
}
// ascending sort of one value per lane: every merge level starts with a "flip" (i <-> size-1-i) so that all
// compare-exchanges point the same way and the direction depends on one lane bit only
template <class T>
__device__ __forceinline__ T knw_sort32(T v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
        v = knw_cx(v, size - 1, size >> 1, lane);
#pragma unroll
        for (int stride = size >> 2; stride > 0; stride >>= 1) v = knw_cx(v, stride, stride, lane);
    }
    return v;
}
// v is bitonic (first half ascending, second half descending or the elementwise min of an ascending and a descending run)
template <class T>
__device__ __forceinline__ T knw_merge32(T v, int lane) {
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) v = knw_cx(v, stride, stride, lane);
    return v;
}

__device__ __forceinline__ void knw_cx_pair(double &d, int &id, int xor_mask, int bit, int lane) {
    const double od = __shfl_xor_sync(KNW_FULL, d, xor_mask);
    const int oi = __shfl_xor_sync(KNW_FULL, id, xor_mask);
    const bool o_less = (od < d) | ((od == d) & (oi < id));
    const bool take = ((lane & bit) == 0) ? o_less : !o_less;
    d = take ? od : d;
    id = take ? oi : id;
}
__device__ __forceinline__ void knw_sort32_pair(double &d, int &id, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
        knw_cx_pair(d, id, size - 1, size >> 1, lane);
#pragma unroll
        for (int stride = size >> 2; stride > 0; stride >>= 1) knw_cx_pair(d, id, stride, stride, lane);
    }
}

__device__ __forceinline__ float knw_sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Candidate references: record index relative to the sample (24 bits) | shift code << 24.  Shift code, 2 bits per axis:
// 0 = the particle itself, 1 = its image shifted by +1 (exists only for particles flagged 1, x <= lower), 2 = shifted by
// -1 (flag 2, x >= upper): the code IS the required flag pattern.
__device__ __forceinline__ int knw_axis_code(int s) { return s > 0 ? 1 : (s < 0 ? 2 : 0); }
__device__ __forceinline__ float knw_axis_shift(int c2) { return (float)(c2 & 1) - (float)(c2 >> 1); }

__device__ __forceinline__ void knw_merge32_pair(double &d, int &id, int lane) {
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) knw_cx_pair(d, id, stride, stride, lane);
}

template <bool PERIODIC>
__global__ void __launch_bounds__(KNW_THREADS, 4) knn_query_warp(KnnQueryParams P, int32_t *__restrict__ todo,
                                                                  int32_t *__restrict__ todo_count) {
    __shared__ int st_s[KNW_WARPS][KNW_STAGE];
    __shared__ int kj_s[KNW_WARPS][KNW_KEEP];
    __shared__ float kk_s[KNW_WARPS][KNW_KEEP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int *st = st_s[wib], *kj = kj_s[wib];
    float *kk = kk_s[wib];

    // grid = (warps per sample / KNW_WARPS, B): no division, and everything derived from b lives in uniform registers
    const int b = blockIdx.y;
    const int s_local = blockIdx.x * KNW_WARPS + wib;
    if (s_local >= P.N) return;                                      // warp-uniform
    const int s = b * P.N + s_local;                                 // B*N < 2^31
    const int G = P.G, k = P.k;
    const KnnGridInfo gi = P.info[b];
    const float4 me = P.sorted[s];
    const int my_id = __float_as_int(me.w) & KNN_IDX_MASK;
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
    const int32_t *__restrict__ cells = P.cell_start + (int64_t)b * G * G * G;   // the sample's cells
    const float4 *__restrict__ recs = P.sorted + (int64_t)b * P.N;               // ... and records
    const int rec0 = b * P.N;
    const bool skip_self = !P.include_self;

    float Rl = INFINITY;          // lane i: the (i+1)-th smallest key seen so far
    float Tcur = INFINITY, thrT = INFINITY;
    int n_stage = 0;              // staged references
    int n_keep = 0, n_merged = 0; // remembered candidates; the first n_merged of them are already in Rl
    bool bail = false;

    // sort 32 keys and merge them into the list of the 32 smallest
    auto merge_keys = [&](float key) {
        const float sorted_keys = knw_sort32(key, lane);
        const float rev = __shfl_xor_sync(KNW_FULL, sorted_keys, 31);
        Rl = knw_merge32(fminf(Rl, rev), lane);
        Tcur = __shfl_sync(KNW_FULL, Rl, k - 1);
        thrT = knw_thr<PERIODIC>(Tcur);
    };
    auto merge_pending = [&](bool all) {
        while (n_keep - n_merged >= (all ? 1 : 32)) {
            const int i = n_merged + lane;
            const float key = i < n_keep ? kk[i] : INFINITY;
            n_merged = nbpc_min(n_merged + 32, n_keep);
            if (__any_sync(KNW_FULL, key < Tcur)) merge_keys(key);
        }
    };
    // keep only the remembered candidates that are still within the margin of the current bound (all of them merged)
    auto compact_keep = [&]() {
        int m = 0;
        for (int i0 = 0; i0 < n_keep; i0 += 32) {
            const int i = i0 + lane;
            const bool v = i < n_keep;
            const float key = v ? kk[i] : INFINITY;
            const int e = v ? kj[i] : 0;
            const bool keepit = v && key <= thrT;
            const unsigned bm = __ballot_sync(KNW_FULL, keepit);
            if (keepit) {                   // pos <= i: every entry of this round was read before the ballot
                const int pos = m + __popc(bm & lt_mask);
                kk[pos] = key;
                kj[pos] = e;
            }
            m += __popc(bm);
            __syncwarp();
        }
        n_keep = n_merged = m;
    };

    // evaluate one reference per lane (e < 0: none): FP32 key, remember the ones within the margin
    auto eval_refs = [&](int e) {
        float key = INFINITY;
        bool take = false;
        if (e >= 0) {
            const float4 c = __ldg(&recs[e & KNN_IDX_MASK]);
            const int w = __float_as_int(c.w);
            float tx = me.x - c.x, ty = me.y - c.y, tz = me.z - c.z;
            take = true;
            if (PERIODIC) {
                const int code = e >> KNN_FLAG_SHIFT;
                const int reqmask = ((code | (code >> 1)) & 0x15) * 3;
                take = ((((w >> KNN_FLAG_SHIFT) ^ code) & reqmask) == 0);
                tx -= knw_axis_shift(code & 3); ty -= knw_axis_shift((code >> 2) & 3); tz -= knw_axis_shift((code >> 4) & 3);
            }
            // (an unshifted reference has code 0: e >> 24 == 0)
            take = take && !(skip_self && (e >> KNN_FLAG_SHIFT) == 0 && (w & KNN_IDX_MASK) == my_id);
            key = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
        }
        // while the k-list is not full nothing is pending (n_merged == n_keep): merge this batch right away, so that the
        // keep filter below already sees a finite bound
        const bool filling = !(Tcur < INFINITY);
        if (filling && __any_sync(KNW_FULL, take && key < INFINITY)) merge_keys(take ? key : INFINITY);
        const bool keepit = take && key <= thrT;
        const unsigned bm = __ballot_sync(KNW_FULL, keepit);
        if (keepit) {
            const int pos = n_keep + __popc(bm & lt_mask);
            kk[pos] = key;
            kj[pos] = e;
        }
        n_keep += __popc(bm);
        if (filling) n_merged = n_keep;
        __syncwarp();
        merge_pending(false);
        if (n_keep > KNW_KEEP - 32) {
            merge_pending(true);
            compact_keep();
            if (n_keep > KNW_KEEP - 32) bail = true;
        }
    };
    auto eval_full_batches = [&]() {
        int base = 0;
        for (; n_stage - base >= 32 && !bail; base += 32) eval_refs(st[base + lane]);
        if (base) {             // move the partial batch to the front
            const int left = n_stage - base;
            const int tmp = lane < left ? st[base + lane] : 0;
            __syncwarp();
            if (lane < left) st[lane] = tmp;
            n_stage = left;
            __syncwarp();
        }
    };

    // stage, from every lane, the references ref_a .. ref_a + len_a - 1 followed by ref_b .. ref_b + len_b - 1
    auto append = [&](int ref_a, int len_a, int ref_b, int len_b) {
        // long runs (dense cells) skip the staging list: the warp walks them 32 consecutive records at a time
        const unsigned long_m = __ballot_sync(KNW_FULL, len_a >= KNW_DIRECT || len_b >= KNW_DIRECT);
        if (long_m) {
            for (unsigned m = long_m; m && !bail; m &= m - 1) {
                const int src = __ffs(m) - 1;
#pragma unroll
                for (int run = 0; run < 2; ++run) {
                    const int r = __shfl_sync(KNW_FULL, run ? ref_b : ref_a, src);
                    const int l = __shfl_sync(KNW_FULL, run ? len_b : len_a, src);
                    if (l < KNW_DIRECT) continue;
                    for (int i0 = 0; i0 < l && !bail; i0 += 32) eval_refs(i0 + lane < l ? r + i0 + lane : -1);
                }
            }
            if (len_a >= KNW_DIRECT) len_a = 0;
            if (len_b >= KNW_DIRECT) len_b = 0;
        }
        const int len = len_a + len_b;
        if (!__any_sync(KNW_FULL, len > 0)) return;
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(KNW_FULL, incl, o);
            incl += lane >= o ? t : 0;
        }
        const int tot = __shfl_sync(KNW_FULL, incl, 31);
        if (tot == 0) return;
        if (tot <= KNW_STAGE - n_stage) {                  // the usual case: everything fits
            int *dst = st + n_stage + (incl - len);
            const int n_max = __reduce_max_sync(KNW_FULL, len);
            const int fix = ref_b - len_a - ref_a;          // i >= len_a: ref_b + (i - len_a)
#pragma unroll 4
            for (int i = 0; i < n_max; ++i)
                if (i < len) dst[i] = ref_a + i + (i >= len_a ? fix : 0);
            n_stage += tot;
            __syncwarp();
            eval_full_batches();
            return;
        }
        // dense cells: stage window after window of the concatenated reference list
        const int ord0 = incl - len;                        // ordinal of this lane's first reference
        for (int off = 0; off < tot && !bail;) {
            const int chunk = nbpc_min(KNW_STAGE - n_stage, tot - off);
            const int first = nbpc_max(ord0, off), last = nbpc_min(ord0 + len, off + chunk);
            const int n_max = __reduce_max_sync(KNW_FULL, nbpc_max(last - first, 0));
            for (int i = 0; i < n_max; ++i) {
                const int o = first + i;                    // ordinal; o - ord0 = index in this lane's list
                if (o < last) st[n_stage + (o - off)] = (o - ord0 < len_a) ? ref_a + (o - ord0) : ref_b + (o - ord0 - len_a);
            }
            n_stage += chunk;
            off += chunk;
            __syncwarp();
            eval_full_batches();
        }
    };

    for (int R = 1, Rprev = -1; !bail; Rprev = R, ++R) {
        const int W = 2 * R + 1, nrows = W * W;
        const float inv_w = __fdividef(1.f, (float)W);
        const bool prune = Tcur < INFINITY;                 // warp-uniform
        for (int r0 = 0; r0 < nrows && !bail; r0 += 32) {
            const int r = r0 + lane;
            const int qz = (int)(((float)r + 0.5f) * inv_w);   // r / W (exact: r, W small)
            const int dz = qz - R, dy = r - qz * W - R;
            int tz = cz + dz, ty = cy + dy;
            bool ok = r < nrows && (unsigned)(tz - tmin) <= (unsigned)(tmax - tmin) && (unsigned)(ty - tmin) <= (unsigned)(tmax - tmin);
            int code_yz = 0;
            if (PERIODIC) {
                int sz = 0, sy = 0;
                if (tz < 0) { tz += G; sz = -1; } else if (tz >= G) { tz -= G; sz = 1; }
                if (ty < 0) { ty += G; sy = -1; } else if (ty >= G) { ty -= G; sy = 1; }
                code_yz = (knw_axis_code(sy) << 2) | (knw_axis_code(sz) << 4);
            }
            int x_lo = tmin, x_hi = tmax;
            if (prune) {                                    // cut the row to the cells that intersect the current k-th ball
                const float gz = KNN_GAP(dz, frz), gy = KNN_GAP(dy, fry);
                const float rem = thrT * inv_h2 - (gz * gz + gy * gy);
                ok = ok && !(rem < 0.f);
                const float xr = fminf(knw_sqrt_approx(fmaxf(rem, 0.f)) + 1.1e-3f, 8192.f);
                x_lo = nbpc_max(x_lo, (int)floorf((float)cx + frx - xr));
                x_hi = nbpc_min(x_hi, (int)floorf((float)cx + frx + xr));
            }
            if (!__any_sync(KNW_FULL, ok)) continue;
            const bool full_row = nbpc_max(abs(dz), abs(dy)) > Rprev;
            // left part (or the whole row) and right part of the new cells of this row, extended coordinates
            const int a0 = nbpc_max(cx - R, x_lo), a1 = nbpc_min(full_row ? cx + R : cx - Rprev - 1, x_hi);
            const int b0 = nbpc_max(cx + Rprev + 1, x_lo), b1 = nbpc_min(cx + R, x_hi);
            const int32_t *__restrict__ row = cells + (tz * G + ty) * G;
            // copies of the row (shift -1 / 0 / +1) that some lane's range reaches
            const int seg_first = (PERIODIC && !__any_sync(KNW_FULL, ok && a0 < 0)) ? 1 : 0;
            const int seg_last = PERIODIC ? (__any_sync(KNW_FULL, ok && nbpc_max(a1, b1) >= G) ? 2 : 1) : 0;
#pragma unroll 1
            for (int seg = seg_first; seg <= seg_last; ++seg) {
                const int sx = PERIODIC ? seg - 1 : 0;
                const int z0 = sx * G, z1 = sx * G + G - 1;          // the cells of this copy of the row
                int ref_a = 0, len_a = 0, ref_b = 0, len_b = 0;
                const int code = PERIODIC ? ((knw_axis_code(sx) | code_yz) << KNN_FLAG_SHIFT) : 0;
                if (ok) {
                    const int la = nbpc_max(a0, z0), ha = nbpc_min(a1, z1);
                    if (la <= ha) {
                        const int jb = __ldg(&row[la - z0]);
                        len_a = __ldg(&row[ha - z0 + 1]) - jb;
                        ref_a = (jb - rec0) | code;
                    }
                    const int lb = nbpc_max(b0, z0), hb = nbpc_min(b1, z1);
                    if (!full_row && lb <= hb) {
                        const int jb = __ldg(&row[lb - z0]);
                        len_b = __ldg(&row[hb - z0 + 1]) - jb;
                        ref_b = (jb - rec0) | code;
                    }
                }
                append(ref_a, len_a, ref_b, len_b);
            }
        }
        if (bail) break;
        if (n_stage) {                      // the shell's partial batch: the termination test needs the true bound
            eval_refs(lane < n_stage ? st[lane] : -1);
            n_stage = 0;
        }
        merge_pending(true);
        if (bail || R >= Rmax) break;
        const float g = ((float)R + margin - 2e-3f) * (float)gi.h;   // (conservative float version of the float64 test)
        if (g > 0.f && thrT < g * g) break;
    }

    if (!bail) {
        compact_keep();                     // the candidates within the margin of the final bound
        if (n_keep > 64 || n_keep < k) bail = true;
    }
    if (bail) {                             // hand the query to the thread-per-query kernel
        if (lane == 0) todo[atomicAdd(todo_count, 1)] = (int32_t)s;
        return;
    }

    // exact re-evaluation (the reference's float64 operation order) of the survivors: one per lane, two when there are
    // more than 32 (near-ties around the k-th distance, only possible for k close to 32)
    auto exact = [&](int i, double &dd, int &cid) {
        dd = INFINITY;
        cid = 0x7FFFFFFF;
        if (i < n_keep) {
            const int e = kj[i];
            const float4 c = __ldg(&recs[e & KNN_IDX_MASK]);
            cid = __float_as_int(c.w) & KNN_IDX_MASK;
            double ox = 0.0, oy = 0.0, oz = 0.0;
            if (PERIODIC) {
                const int code = e >> KNN_FLAG_SHIFT;
                ox = (double)knw_axis_shift(code & 3); oy = (double)knw_axis_shift((code >> 2) & 3); oz = (double)knw_axis_shift((code >> 4) & 3);
            }
            const double tx = __dsub_rn((double)me.x, __dadd_rn((double)c.x, ox));
            const double ty = __dsub_rn((double)me.y, __dadd_rn((double)c.y, oy));
            const double tz = __dsub_rn((double)me.z, __dadd_rn((double)c.z, oz));
            dd = __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz));
        }
    };
    double dd;
    int cid;
    exact(lane, dd, cid);
    const int64_t out = ((int64_t)b * P.N + my_id) * k;
    const bool by_distance = P.order != NBPC_ORDER_INDEX;
    if (n_keep > 32) {                      // the 32 smallest of two sorted runs: min(A[i], B[31 - i]) is bitonic
        double d2;
        int c2;
        exact(32 + lane, d2, c2);
        knw_sort32_pair(dd, cid, lane);
        knw_sort32_pair(d2, c2, lane);
        const double rd = __shfl_xor_sync(KNW_FULL, d2, 31);
        const int rc = __shfl_xor_sync(KNW_FULL, c2, 31);
        if ((rd < dd) | ((rd == dd) & (rc < cid))) { dd = rd; cid = rc; }
        knw_merge32_pair(dd, cid, lane);
    } else if (n_keep > k || by_distance || P.d2_out) {
        knw_sort32_pair(dd, cid, lane);     // exact (d2, index) order
    }
    if (P.d2_out && lane < k) P.d2_out[out + lane] = dd;
    if (by_distance) {
        if (lane < k) P.idx_out[out + lane] = cid;
        return;
    }
    const int ids = knw_sort32(lane < k ? cid : 0x7FFFFFFF, lane);
    if (lane < k) P.idx_out[out + lane] = ids;
}
