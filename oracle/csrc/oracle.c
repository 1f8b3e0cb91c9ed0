/* oracle.c -- plain-C restatement of the index algorithms of the hot path.  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py, never by the
 * product (neural-audio-fp_b200/ must not import oracle/).
 *
 * What it restates (the reference reaches all of it through the un-vendored faiss 1.6.5 wheel,
 * eval/utils/get_index_faiss.py:58,69-74,105-120 and eval/eval_faiss.py:147-148,211; algorithms as published):
 *   orc_flat_search      IndexFlatL2.search: squared L2, k smallest ascending, ties to the lower label, -1 / +inf padding.
 *                        fp32 direct (q - x)^2 accumulation, the formulation faiss uses for nq < 20 (the reference's
 *                        call pattern: <= 19 rows per search) -- this is the TIMED CPU baseline of the exact search.
 *   orc_kmeans           Lloyd k-means as faiss' Clustering: assign to the nearest centroid (ties -> lower id),
 *                        centroid = mean of its points, empty clusters keep their centroid (faiss splits a big
 *                        cluster instead: PARITY UNPINNED, which is why IVF-PQ is gated on hit rate).
 *   orc_pq_encode        IndexIVFPQ.add: nearest coarse centroid, residual, code_m = arg-min over 256 codewords.
 *   orc_ivfpq_search     IndexIVFPQ.search: nprobe nearest lists, per (query, list) look-up table
 *                        T[m][c] = |(q - c_l)_m - pq_m[c]|^2 in fp32, ADC distance = sum_m T[m][code_m] (fp32, m ascending),
 *                        k smallest over the probed lists by (distance, label).
 * Same results as oracle/flat_index.py and oracle/ivfpq_index.py (numpy), which tests/test_oracle_c.py checks.
 *
 * gcc -O3 -mavx2 -mfma -fopenmp -shared -fPIC oracle.c -o liboracle.so   (oracle/cbuild.py)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_version(void) { return 2; }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------- bounded top-k by (distance, label) ascending */
typedef struct {
    float d;
    int64_t i;
} cand_t;

static inline int cand_less(float d0, int64_t i0, float d1, int64_t i1) { return d0 < d1 || (d0 == d1 && i0 < i1); }

/* keeps the k best in h[0..*n) sorted ascending; insertion is O(k) but k <= 128 and almost every candidate is rejected */
static inline void topk_push(cand_t* h, int* n, int k, float d, int64_t i) {
    if (*n == k && !cand_less(d, i, h[k - 1].d, h[k - 1].i)) return;
    int p = *n < k ? (*n)++ : k - 1;
    while (p > 0 && cand_less(d, i, h[p - 1].d, h[p - 1].i)) {
        h[p] = h[p - 1];
        --p;
    }
    h[p].d = d;
    h[p].i = i;
}

/* ---------------------------------------------------------------- exact flat L2 */
static inline float l2_f32(const float* a, const float* b, int d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int j = 0;
    for (; j + 8 <= d; j += 8)
        for (int u = 0; u < 8; ++u) {
            const float t = a[j + u] - b[j + u];
            acc[u] += t * t;
        }
    float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    for (; j < d; ++j) {
        const float t = a[j] - b[j];
        s += t * t;
    }
    return s;
}

/* q (nq, d), x (n, d) row-major float32 -> D (nq, k) float32, I (nq, k) int64.  Threads split the database; every
 * thread keeps nq private top-k lists, merged at the end (the reference searches <= 19 rows at a time, so the
 * database -- not the query set -- is what has to be split). */
int orc_flat_search(const float* q, int64_t nq, const float* x, int64_t n, int d, int k, float* D, int64_t* I) {
    if (k < 1 || k > 1024 || nq < 0 || n < 0) return -1;
    int nthreads = orc_max_threads();
    cand_t* all = (cand_t*)malloc(sizeof(cand_t) * (size_t)nthreads * (size_t)(nq > 0 ? nq : 1) * (size_t)k);
    int* cnt = (int*)calloc((size_t)nthreads * (size_t)(nq > 0 ? nq : 1), sizeof(int));
    if (!all || !cnt) { free(all); free(cnt); return -2; }
#pragma omp parallel
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int t = 0, nt = 1;
#endif
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        cand_t* mine = all + (size_t)t * nq * k;
        int* mcnt = cnt + (size_t)t * nq;
        const int64_t BLK = 512;                       /* rows per block: the block stays in L1/L2 for all nq rows */
        for (int64_t b0 = lo; b0 < hi; b0 += BLK) {
            const int64_t b1 = b0 + BLK < hi ? b0 + BLK : hi;
            for (int64_t qi = 0; qi < nq; ++qi) {
                const float* qr = q + qi * d;
                cand_t* h = mine + qi * k;
                int c = mcnt[qi];
                for (int64_t r = b0; r < b1; ++r) topk_push(h, &c, k, l2_f32(qr, x + r * d, d), r);
                mcnt[qi] = c;
            }
        }
    }
    for (int64_t qi = 0; qi < nq; ++qi) {
        cand_t best[1024];
        int c = 0;
        for (int t = 0; t < nthreads; ++t) {
            const cand_t* h = all + ((size_t)t * nq + qi) * k;
            const int m = cnt[(size_t)t * nq + qi];
            for (int j = 0; j < m; ++j) topk_push(best, &c, k, h[j].d, h[j].i);
        }
        for (int j = 0; j < k; ++j) {
            D[qi * k + j] = j < c ? best[j].d : INFINITY;
            I[qi * k + j] = j < c ? best[j].i : -1;
        }
    }
    free(all);
    free(cnt);
    return 0;
}

/* ---------------------------------------------------------------- k-means (Lloyd) */
/* nearest centroid of every point, ties to the lower centroid id.  The squared distance is DEFINED as the fp32 chain
 * s = fma(t, t, s), t = x_j - c_j, j ascending (faiss computes fp32 distances as well): k-means amplifies every
 * flipped assignment over its 25 iterations, so the CUDA implementation (csrc/ivfpq.cu kmeans_assign_kernel) and this
 * one only train the same quantizers if they round the same way. */
static void assign_points(const float* x, int64_t n, int ld, int d, const float* cent, int k, int32_t* assign) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* xr = x + i * ld;
        float best = FLT_MAX;
        int bi = 0;
        for (int c = 0; c < k; ++c) {
            const float* cr = cent + (int64_t)c * d;
            float s = 0.f;
            for (int j = 0; j < d; ++j) {
                const float t = xr[j] - cr[j];
                s = fmaf(t, t, s);
            }
            if (s < best) { best = s; bi = c; }
        }
        assign[i] = bi;
    }
}

/* x: n points of dimension d with row stride ld floats; cent (k, d): initial centroids in, final centroids out */
int orc_kmeans(const float* x, int64_t n, int ld, int d, int k, float* cent, int niter, int32_t* assign_out) {
    if (n < 1 || d < 1 || k < 1) return -1;
    int32_t* assign = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    double* sums = (double*)malloc(sizeof(double) * (size_t)k * d);
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    if (!assign || !sums || !cnt) { free(assign); free(sums); free(cnt); return -2; }
    for (int it = 0; it < niter; ++it) {
        assign_points(x, n, ld, d, cent, k, assign);
        memset(sums, 0, sizeof(double) * (size_t)k * d);
        memset(cnt, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t i = 0; i < n; ++i) {              /* fixed (ascending) summation order */
            double* s = sums + (int64_t)assign[i] * d;
            const float* xr = x + i * ld;
            for (int j = 0; j < d; ++j) s[j] += xr[j];
            cnt[assign[i]]++;
        }
        for (int c = 0; c < k; ++c)
            if (cnt[c] > 0)
                for (int j = 0; j < d; ++j) cent[(int64_t)c * d + j] = (float)(sums[(int64_t)c * d + j] / (double)cnt[c]);
    }
    if (assign_out) {
        assign_points(x, n, ld, d, cent, k, assign);
        memcpy(assign_out, assign, sizeof(int32_t) * (size_t)n);
    }
    free(assign);
    free(sums);
    free(cnt);
    return 0;
}

/* ---------------------------------------------------------------- IVF-PQ add */
int orc_pq_encode(const float* x, int64_t n, int d, const float* coarse, int nlist, const float* pq, int m, int ksub,
                  int32_t* assign, uint8_t* codes) {
    if (d % m != 0 || ksub > 256) return -1;
    const int dsub = d / m;
    assign_points(x, n, d, d, coarse, nlist, assign);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* xr = x + i * d;
        const float* cr = coarse + (int64_t)assign[i] * d;
        for (int s = 0; s < m; ++s) {
            double best = DBL_MAX;
            int bi = 0;
            for (int c = 0; c < ksub; ++c) {
                const float* pr = pq + ((int64_t)s * ksub + c) * dsub;
                double acc = 0.0;
                for (int j = 0; j < dsub; ++j) {
                    const double t = ((double)xr[s * dsub + j] - (double)cr[s * dsub + j]) - (double)pr[j];
                    acc += t * t;
                }
                if (acc < best) { best = acc; bi = c; }
            }
            codes[i * m + s] = (uint8_t)bi;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- IVF-PQ search (reference formulation: LUT + ADC scan)
 * order (n): rows sorted by list (stable), loff (nlist + 1): list l = order[loff[l] .. loff[l+1]). */
int orc_ivfpq_search(const float* q, int64_t nq, int d, int k, int nprobe, const float* coarse, int nlist, const float* pq,
                     int m, int ksub, const uint8_t* codes, const int64_t* order, const int64_t* loff, float* D, int64_t* I) {
    if (d % m != 0 || k < 1 || k > 1024 || nprobe < 1) return -1;
    const int dsub = d / m;
    if (nprobe > nlist) nprobe = nlist;
    int fail = 0;
#pragma omp parallel
    {
        float* lut = (float*)malloc(sizeof(float) * (size_t)m * ksub);
        cand_t* pl = (cand_t*)malloc(sizeof(cand_t) * (size_t)nprobe);
        cand_t* best = (cand_t*)malloc(sizeof(cand_t) * (size_t)k);
        if (!lut || !pl || !best) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 4)
            for (int64_t qi = 0; qi < nq; ++qi) {
                const float* qr = q + qi * d;
                int np = 0;
                for (int l = 0; l < nlist; ++l) topk_push(pl, &np, nprobe, l2_f32(qr, coarse + (int64_t)l * d, d), l);
                int c = 0;
                for (int p = 0; p < np; ++p) {
                    const int l = (int)pl[p].i;
                    const float* cr = coarse + (int64_t)l * d;
                    for (int s = 0; s < m; ++s)
                        for (int e = 0; e < ksub; ++e) {
                            const float* pr = pq + ((int64_t)s * ksub + e) * dsub;
                            float acc = 0.f;
                            for (int j = 0; j < dsub; ++j) {
                                const float t = (qr[s * dsub + j] - cr[s * dsub + j]) - pr[j];
                                acc += t * t;
                            }
                            lut[s * ksub + e] = acc;
                        }
                    for (int64_t pos = loff[l]; pos < loff[l + 1]; ++pos) {
                        const int64_t row = order[pos];
                        const uint8_t* code = codes + row * m;
                        float dist = 0.f;
                        for (int s = 0; s < m; ++s) dist += lut[s * ksub + code[s]];
                        topk_push(best, &c, k, dist, row);
                    }
                }
                for (int j = 0; j < k; ++j) {
                    D[qi * k + j] = j < c ? best[j].d : INFINITY;
                    I[qi * k + j] = j < c ? best[j].i : -1;
                }
            }
        }
        free(lut);
        free(pl);
        free(best);
    }
    return fail ? -2 : 0;
}
