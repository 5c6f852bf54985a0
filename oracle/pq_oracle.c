/* TEST INFRASTRUCTURE ONLY — plain-C restatement (oracle) of MILLION's PQ hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  It is the checker and the CPU baseline, never the product.
 *
 *   oracle_pq_encode       scripts/utils/pq_utils.py:451-499 (sa_encode_4d_keops): fp32 squared-L2
 *                          arg-min per sub-space, first minimum wins (our tie rule; KeOps 2.2.3 is
 *                          not vendored in the reference).
 *   oracle_pq_decode       scripts/utils/pq_utils.py:501-540 (sa_decode_4d)
 *   oracle_pq_decode_attn  scripts/modeldb/bindings/Interface.cu:16-120 + Kernel.cuh:11-166,
 *                          1038-1209, 1211-1270, in the fp32 form of the invariant the reference
 *                          states for it (pq_utils.py:360-368).
 * Pinned by tests/test_oracle_golden.py against fixtures produced by the reference's own Python.
 * Build: `make -C oracle` (gcc -O3 -fopenmp) -> oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* x: (n_vec, M*d_m) fp32; cent: (M, C, d_m) fp32; codes: (n_vec, M) uint8 (C <= 256) */
void oracle_pq_encode(const float* x, const float* cent, uint8_t* codes,
                      int64_t n_vec, int M, int C, int d_m) {
    const int d = M * d_m;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n_vec; ++t) {
        for (int m = 0; m < M; ++m) {
            const float* xv = x + t * d + m * d_m;
            const float* cm = cent + (int64_t)m * C * d_m;
            float best = INFINITY;
            int bi = 0;
            for (int c = 0; c < C; ++c) {
                volatile float acc = 0.f;                /* volatile: forbid FMA contraction / reassoc */
                for (int k = 0; k < d_m; ++k) {
                    volatile float diff = xv[k] - cm[c * d_m + k];
                    volatile float sq = diff * diff;
                    acc = (k == 0) ? sq : (acc + sq);
                }
                if (acc < best) { best = acc; bi = c; }
            }
            codes[t * M + m] = (uint8_t)bi;
        }
    }
}

/* codes (n_vec, M) -> out (n_vec, M*d_m) */
void oracle_pq_decode(const uint8_t* codes, const float* cent, float* out,
                      int64_t n_vec, int M, int C, int d_m) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n_vec; ++t)
        for (int m = 0; m < M; ++m)
            memcpy(out + (t * M + m) * d_m, cent + ((int64_t)m * C + codes[t * M + m]) * d_m,
                   sizeof(float) * d_m);
}

/* q (bs, nh, d); kc/vc (bs, nh_k, nk, M); kcent/vcent (M, C, d_m); kres/vres (bs, nh_k, Lt, d), rows
 * [0, r) valid; out (bs, nh, d).  All float data fp32.  No mask. */
void oracle_pq_decode_attn(const float* q, const uint8_t* kc, const uint8_t* vc,
                           const float* kcent, const float* vcent,
                           const float* kres, const float* vres, int r, float* out,
                           int bs, int nh, int nh_k, int64_t nk, int d, int M, int C, int Lt) {
    const int d_m = d / M;
    const int G = nh / nh_k;
    const float scale = 1.0f / sqrtf((float)d);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bh = 0; bh < bs * nh; ++bh) {
        const int b = bh / nh, h = bh % nh, hk = h / G;
        const float* qv = q + (int64_t)bh * d;
        float* lut = (float*)malloc(sizeof(float) * M * C);
        float* s = (float*)malloc(sizeof(float) * (nk + r + 1));
        for (int m = 0; m < M; ++m)
            for (int c = 0; c < C; ++c) {
                float a = 0.f;
                for (int k = 0; k < d_m; ++k) a += qv[m * d_m + k] * kcent[((int64_t)m * C + c) * d_m + k];
                lut[m * C + c] = a;
            }
        const uint8_t* kcb = kc + ((int64_t)b * nh_k + hk) * nk * M;
        const uint8_t* vcb = vc + ((int64_t)b * nh_k + hk) * nk * M;
        float mx = -INFINITY;
        for (int64_t j = 0; j < nk; ++j) {
            float a = 0.f;
            for (int m = 0; m < M; ++m) a += lut[m * C + kcb[j * M + m]];
            s[j] = a * scale;
            if (s[j] > mx) mx = s[j];
        }
        const float* kr = kres + ((int64_t)b * nh_k + hk) * Lt * d;
        const float* vr = vres + ((int64_t)b * nh_k + hk) * Lt * d;
        for (int i = 0; i < r; ++i) {
            float a = 0.f;
            for (int k = 0; k < d; ++k) a += qv[k] * kr[(int64_t)i * d + k];
            s[nk + i] = a * scale;
            if (s[nk + i] > mx) mx = s[nk + i];
        }
        double l = 0.0;
        double* acc = (double*)calloc(d, sizeof(double));
        for (int64_t j = 0; j < nk; ++j) {
            const float p = expf(s[j] - mx);
            l += p;
            for (int m = 0; m < M; ++m) {
                const float* cv = vcent + ((int64_t)m * C + vcb[j * M + m]) * d_m;
                for (int k = 0; k < d_m; ++k) acc[m * d_m + k] += (double)p * cv[k];
            }
        }
        for (int i = 0; i < r; ++i) {
            const float p = expf(s[nk + i] - mx);
            l += p;
            for (int k = 0; k < d; ++k) acc[k] += (double)p * vr[(int64_t)i * d + k];
        }
        for (int k = 0; k < d; ++k) out[(int64_t)bh * d + k] = (float)(acc[k] / l);
        free(acc); free(s); free(lut);
    }
}
