// attn_generic.cu — all-shapes decode-attention kernel (plain SIMT, fp32 everywhere).
//
// Same result as the reference's at::matmul + flash_decoding_split_kernel + flash_decoding_residual_kernel
// + flash_decoding_reduce_kernel chain (scripts/modeldb/bindings/Interface.cu:49-118, Kernel.cuh:11-166,
// 1038-1209, 1211-1270) in ONE launch: the LUT is built in shared memory, the fp16 window is one more
// "split", and the last CTA of every (b, kv-head) group merges the partials (atomic ticket).  It is the
// fallback for shapes the fast kernel does not cover and the on-device cross-check for it.
#include "attn_common.cuh"

namespace million {

template <typename T>
__global__ void __launch_bounds__(128) attn_generic_kernel(const AttnArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);  // M*C (coded splits) or d (window: q); not used in direct mode
    float* qs = lut + (a.direct ? 0 : a.M * a.C);     // coded splits with K outliers / direct mode: q * scale, d floats
    __shared__ float P[128];
    __shared__ float red[33];
    __shared__ float mscr[kMergeScratch];
    __shared__ int flag;

    const int split = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x;
    const int G = a.nh / a.nh_k, dm = a.d / a.M, hb = b * a.nh_k + hk;
    const int n_parts = a.n_parts;
    const bool window = (split == a.n_splits);
    const T* kcent = reinterpret_cast<const T*>(a.k_cent);
    const T* vcent = reinterpret_cast<const T*>(a.v_cent);

    int t0 = 0, t1 = 0;
    if (window) t1 = window_rows(a); else split_range(a, split, t0, t1);

    for (int g = 0; g < G; ++g) {
        const int h = hk * G + g;
        const T* q = reinterpret_cast<const T*>(a.q) + (int64_t)(b * a.nh + h) * a.d;
        __syncthreads();
        if (window) {
            for (int i = tid; i < a.d; i += 128) lut[i] = io<T>::to_f(q[i]) * a.scale_log2;
        } else if (t1 > t0) {
            if (!a.direct)
                for (int i = tid; i < a.M * a.C; i += 128) {
                    const int m = i / a.C;
                    float acc = 0.f;
                    for (int k = 0; k < dm; ++k)
                        acc = fmaf(io<T>::to_f(q[m * dm + k]), io<T>::to_f(kcent[(int64_t)i * dm + k]), acc);
                    lut[i] = acc * a.scale_log2;
                }
            if (a.k_out || a.direct)
                for (int i = tid; i < a.d; i += 128) qs[i] = io<T>::to_f(q[i]) * a.scale_log2;
        }
        __syncthreads();

        float run_m = -INFINITY, run_l = 0.f;
        float acc[2] = {0.f, 0.f};  // d <= 256
        for (int tile = t0; tile < t1; tile += 128) {
            const int j = tile + tid;
            float s = -INFINITY;
            if (j < t1) {
                if (window) {
                    const T* kr = reinterpret_cast<const T*>(a.k_res) + ((int64_t)hb * a.res_len + j) * a.d;
                    float acc_s = 0.f;
                    for (int k = 0; k < a.d; ++k) acc_s = fmaf(lut[k], io<T>::to_f(kr[k]), acc_s);
                    s = acc_s;
                } else {
                    float acc_s = 0.f;
                    if (!a.direct) {
                        for (int m = 0; m < a.M; ++m) acc_s += lut[m * a.C + k_code_at(a, hb, j, m)];
                    } else {
                        // large codebooks (nbits > 8): the same sum, sub-space by sub-space, straight from the centroids (L2 resident)
                        for (int m = 0; m < a.M; ++m) {
                            const T* c = kcent + ((int64_t)m * a.C + k_code_at(a, hb, j, m)) * dm;
                            float part = 0.f;
                            for (int k = 0; k < dm; ++k) part = fmaf(qs[m * dm + k], io<T>::to_f(c[k]), part);
                            acc_s += part;
                        }
                    }
                    if (a.k_out) {   // outlier side store: x_hat[dim] = centroid + delta
                        const int64_t rec = hb * a.ko_head_stride + (int64_t)j * a.k_out;
                        for (int i = 0; i < a.k_out; ++i)
                            acc_s = fmaf(qs[a.ko_idx[rec + i]], io<T>::to_f(reinterpret_cast<const T*>(a.ko_val)[rec + i]), acc_s);
                    }
                    s = acc_s;
                }
            }
            const float tmax = block_reduce<true>(s, red);
            const float new_m = fmaxf(run_m, tmax);
            const float alpha = exp2_safe(run_m, new_m);
            const float p = (j < t1) ? exp2_safe(s, new_m) : 0.f;
            const float psum = block_reduce<false>(p, red);
            run_l = run_l * alpha + psum;
            run_m = new_m;
            P[tid] = p;
            __syncthreads();
            const int len = min(128, t1 - tile);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = tid + u * 128;
                if (i < a.d) {
                    float o = acc[u] * alpha;
                    if (window) {
                        const T* vr = reinterpret_cast<const T*>(a.v_res) + ((int64_t)hb * a.res_len + tile) * a.d + i;
                        for (int jj = 0; jj < len; ++jj) o = fmaf(P[jj], io<T>::to_f(vr[(int64_t)jj * a.d]), o);
                    } else {
                        const int m = i / dm, k = i % dm;
                        for (int jj = 0; jj < len; ++jj) {
                            const int code = v_code_at(a, hb, tile + jj, m);
                            float vv = io<T>::to_f(vcent[((int64_t)m * a.C + code) * dm + k]);
                            if (a.v_out) {
                                const int64_t rec = hb * a.vo_head_stride + (int64_t)(tile + jj) * a.v_out;
                                for (int i2 = 0; i2 < a.v_out; ++i2)
                                    if (a.vo_idx[rec + i2] == i) vv += io<T>::to_f(reinterpret_cast<const T*>(a.vo_val)[rec + i2]);
                            }
                            o = fmaf(P[jj], vv, o);
                        }
                    }
                    acc[u] = o;
                }
            }
            __syncthreads();
        }
        float* part = a.parts + ((int64_t)(b * a.nh + h) * n_parts + split) * part_stride(a.d);
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (tid + u * 128 < a.d) part[tid + u * 128] = acc[u];
        if (tid == 0) { part[a.d] = run_m; part[a.d + 1] = run_l; }
    }
    if (last_cta_of_group(a.counters, hb, n_parts, &flag)) merge_group<T>(a, b, hk, n_parts, mscr);
}

int launch_attn_generic(const AttnArgs& a_in, int io_dtype, cudaStream_t stream) {
    AttnArgs a = a_in;
    a.n_parts = a.n_splits + 1;
    if (a.d > 256) MILLION_UNSUPPORTED("generic decode attention supports d <= 256 (got %d)", a.d);
    a.direct = (size_t)a.M * a.C * sizeof(float) > 160 * 1024;       // e.g. M=64, C=1024 (nbits 10): 256 KB of LUT do not fit
    const size_t smem = sizeof(float) * ((a.direct ? 0 : (size_t)a.M * a.C) + a.d);
    dim3 grid(a.n_splits + 1, a.nh_k, a.bs), block(128);
    if (io_dtype == MILLION_F16) {
        MILLION_CUDA_OK(cudaFuncSetAttribute(attn_generic_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MILLION_CUDA_OK(launch_kernel(attn_generic_kernel<__half>, grid, block, smem, stream, a.pdl != 0, a));
    } else {
        MILLION_CUDA_OK(cudaFuncSetAttribute(attn_generic_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MILLION_CUDA_OK(launch_kernel(attn_generic_kernel<__nv_bfloat16>, grid, block, smem, stream, a.pdl != 0, a));
    }
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

// ------------------------------------------------------------------------------------------------
// cross-rank merge of (n_parts, rows, d+2) [o_unnorm | m (natural log) | l]
template <typename T>
__global__ void lse_merge_kernel(const float* __restrict__ parts, int n_parts, int64_t n_rows, int d, T* __restrict__ out) {
    const int64_t row = blockIdx.x;
    const int stride = d + 2;
    float mstar = -INFINITY;
    for (int i = 0; i < n_parts; ++i) {
        const float* p = parts + ((int64_t)i * n_rows + row) * stride;
        if (p[d + 1] > 0.f) mstar = fmaxf(mstar, p[d]);
    }
    float den = 0.f;
    for (int i = 0; i < n_parts; ++i) {
        const float* p = parts + ((int64_t)i * n_rows + row) * stride;
        if (p[d + 1] > 0.f) den += p[d + 1] * __expf(p[d] - mstar);
    }
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < n_parts; ++i) {
            const float* p = parts + ((int64_t)i * n_rows + row) * stride;
            if (p[d + 1] > 0.f) acc += p[k] * __expf(p[d] - mstar);
        }
        out[row * d + k] = io<T>::from_f(den > 0.f ? acc / den : 0.f);
    }
}

int launch_lse_merge(const float* parts, int n_parts, int64_t n_rows, int d, void* out, int io_dtype, cudaStream_t stream) {
    if (n_rows == 0) return MILLION_OK;
    dim3 grid((unsigned)n_rows), block(128);
    if (io_dtype == MILLION_F16) lse_merge_kernel<__half><<<grid, block, 0, stream>>>(parts, n_parts, n_rows, d, (__half*)out);
    else if (io_dtype == MILLION_BF16) lse_merge_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(parts, n_parts, n_rows, d, (__nv_bfloat16*)out);
    else lse_merge_kernel<float><<<grid, block, 0, stream>>>(parts, n_parts, n_rows, d, (float*)out);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

}  // namespace million
