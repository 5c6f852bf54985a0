// outlier.cu — the outlier side store (extension: the reference has no outlier code).
//
//   outlier_split_kernel : per head-vector, the k_out entries of largest |x| (ties: lowest dim) are zeroed in the copy that the
//                          PQ encoder (encode_tc.cu / codec_generic.cu, unchanged) will see, and recorded as (dim, delta) with
//                          delta = x[dim] - cent[m, code, k].  `code` is the arg-min of the MASKED sub-vector, evaluated here with
//                          the encoder's exact arithmetic (sub, mul, add each rounded to nearest; first minimum wins;
//                          pq_utils.py:483-494 restated), so it is the code the encoder stores.
//   outlier_apply_kernel : reconstruction x_hat[dim] += delta on top of reconstruct_kernel's output (sa_decode_4d, pq_utils.py:501-540).
// One warp per head-vector; d <= 256.
#include "common.cuh"

namespace million {

constexpr int kMaxOut = MILLION_MAX_OUTLIERS;

template <typename T>
__global__ void __launch_bounds__(128) outlier_split_kernel(const T* __restrict__ x, int64_t x_head_stride, const float* __restrict__ cent,
                                                            T* __restrict__ x_masked, uint8_t* __restrict__ out_idx, T* __restrict__ out_val,
                                                            int64_t out_head_stride, int64_t t0, int n_tokens, int d, int M, int C, int k_out) {
    __shared__ float rows[4][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int head = blockIdx.y;
    const int t = blockIdx.x * 4 + warp;
    if (t >= n_tokens) return;
    const int e = d >> 5;   // elements per lane (1..8), contiguous: lane owns dims [lane*e, lane*e + e)
    const int dm = d / M;
    const T* xr = x + head * x_head_stride + (int64_t)t * d;
    float v[8], av[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        v[u] = (u < e) ? io<T>::to_f(xr[lane * e + u]) : 0.f;
        av[u] = (u < e) ? fabsf(v[u]) : -3.f;   // below the initial best: never selected
    }
    int sel_dim[kMaxOut];
    float sel_raw[kMaxOut];
#pragma unroll
    for (int i = 0; i < kMaxOut; ++i) {
        sel_dim[i] = 0; sel_raw[i] = 0.f;
        if (i < k_out) {
            float best = -2.f;
            int bd = 0x7fffffff;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (av[u] > best) { best = av[u]; bd = lane * e + u; }   // ascending dims: strict > keeps the lowest
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int od = __shfl_xor_sync(0xffffffffu, bd, off);
                if (ob > best || (ob == best && od < bd)) { best = ob; bd = od; }
            }
            float mine = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (u < e && lane * e + u == bd) { mine = v[u]; v[u] = 0.f; av[u] = -3.f; }
            sel_dim[i] = bd;
            sel_raw[i] = __shfl_sync(0xffffffffu, mine, bd / e);
        }
    }
    T* xm = x_masked + ((int64_t)head * n_tokens + t) * d;
    float* row = rows[warp];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (u < e) { xm[lane * e + u] = io<T>::from_f(v[u]); row[lane * e + u] = v[u]; }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kMaxOut; ++i) {
        if (i < k_out) {
            const int dim = sel_dim[i], m = dim / dm, k = dim - m * dm;
            const float* cm = cent + (int64_t)m * C * dm;
            float best = INFINITY;
            int bc = 0x7fffffff;
            for (int c = lane; c < C; c += 32) {
                float acc = 0.f;
                for (int kk = 0; kk < dm; ++kk) {
                    const float diff = __fsub_rn(row[m * dm + kk], cm[c * dm + kk]);
                    const float sq = __fmul_rn(diff, diff);
                    acc = (kk == 0) ? sq : __fadd_rn(acc, sq);
                }
                if (acc < best) { best = acc; bc = c; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oc = __shfl_xor_sync(0xffffffffu, bc, off);
                if (ob < best || (ob == best && oc < bc)) { best = ob; bc = oc; }
            }
            if (lane == 0) {
                const int64_t rec = head * out_head_stride + (t0 + t) * k_out + i;
                out_idx[rec] = (uint8_t)dim;
                out_val[rec] = io<T>::from_f(__fsub_rn(sel_raw[i], cm[bc * dm + k]));
            }
        }
    }
}

int launch_outlier_split(const void* x, int x_dtype, int64_t xhs, const float* cent, void* x_masked, uint8_t* out_idx, void* out_val,
                         int64_t ohs, int64_t t0, int n_heads, int n_tokens, int d, int M, int C, int k_out, cudaStream_t stream) {
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    dim3 grid((n_tokens + 3) / 4, n_heads), block(128);
    if (x_dtype == MILLION_F16)
        outlier_split_kernel<__half><<<grid, block, 0, stream>>>((const __half*)x, xhs, cent, (__half*)x_masked, out_idx, (__half*)out_val, ohs, t0, n_tokens, d, M, C, k_out);
    else
        outlier_split_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, xhs, cent, (__nv_bfloat16*)x_masked, out_idx, (__nv_bfloat16*)out_val, ohs, t0, n_tokens, d, M, C, k_out);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

template <typename TO, typename TV>
__global__ void outlier_apply_kernel(TO* __restrict__ out, int64_t out_head_stride, const uint8_t* __restrict__ idx, const TV* __restrict__ val,
                                     int64_t store_head_stride, int64_t t0, int n_tokens, int d, int k_out) {
    const int head = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tokens) return;
    const int64_t rec = head * store_head_stride + (t0 + t) * k_out;
    TO* row = out + head * out_head_stride + (int64_t)t * d;
    for (int i = 0; i < k_out; ++i) {   // the dims of one vector are distinct
        const int dim = idx[rec + i];
        row[dim] = io<TO>::from_f(io<TO>::to_f(row[dim]) + io<TV>::to_f(val[rec + i]));
    }
}

int launch_outlier_apply(void* out, int dtype, int64_t ohs, const uint8_t* idx, const void* val, int val_dtype, int64_t shs, int64_t t0,
                         int n_heads, int n_tokens, int d, int k_out, cudaStream_t stream) {
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    dim3 grid((n_tokens + 127) / 128, n_heads), block(128);
#define MILLION_APPLY(TO, TV) outlier_apply_kernel<TO, TV><<<grid, block, 0, stream>>>((TO*)out, ohs, idx, (const TV*)val, shs, t0, n_tokens, d, k_out)
    if (val_dtype == MILLION_F16) {
        if (dtype == MILLION_F16) MILLION_APPLY(__half, __half);
        else if (dtype == MILLION_BF16) MILLION_APPLY(__nv_bfloat16, __half);
        else MILLION_APPLY(float, __half);
    } else {
        if (dtype == MILLION_F16) MILLION_APPLY(__half, __nv_bfloat16);
        else if (dtype == MILLION_BF16) MILLION_APPLY(__nv_bfloat16, __nv_bfloat16);
        else MILLION_APPLY(float, __nv_bfloat16);
    }
#undef MILLION_APPLY
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

}  // namespace million
