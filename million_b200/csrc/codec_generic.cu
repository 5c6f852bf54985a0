// codec_generic.cu — all-shapes PQ encoder (exact fp32 arg-min on CUDA cores), reconstruct, and the small
// window-maintenance kernels.
//
//   encode      : scripts/utils/pq_utils.py:451-499 sa_encode_4d_keops — fp32 squared-L2 arg-min, first
//                 minimum wins; arithmetic restated exactly (sub, mul, add each rounded to nearest, no FMA),
//                 so codes are bit-identical to the oracle, ties included.
//   reconstruct : scripts/utils/pq_utils.py:501-540 sa_decode_4d
//   window ops  : scripts/utils/pq_utils.py:304-311, scripts/utils/paged_pq_utils.py:186-199
#include "codec.cuh"

namespace million {


template <typename T, int DM>
__global__ void __launch_bounds__(128) encode_generic_kernel(const T* __restrict__ x, int64_t x_head_stride,
                                                             const float* __restrict__ cent, CodeDst dst,
                                                             int n_tokens, int d, int M, int C, int MG) {
    extern __shared__ __align__(16) float cs[];  // MG * C * DM
    const int head = blockIdx.z;
    const int m0 = blockIdx.y * MG;
    const int mg = min(MG, M - m0);
    for (int i = threadIdx.x; i < mg * C * DM; i += blockDim.x) cs[i] = cent[(int64_t)m0 * C * DM + i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tokens) return;
    const T* xr = x + head * x_head_stride + (int64_t)t * d;
    for (int mi = 0; mi < mg; ++mi) {
        float xv[DM];
#pragma unroll
        for (int k = 0; k < DM; ++k) xv[k] = io<T>::to_f(xr[(m0 + mi) * DM + k]);
        const float* cm = cs + mi * C * DM;
        float best = INFINITY;
        int bi = 0;
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < DM; ++k) {
                const float diff = __fsub_rn(xv[k], cm[c * DM + k]);
                const float sq = __fmul_rn(diff, diff);
                acc = (k == 0) ? sq : __fadd_rn(acc, sq);
            }
            if (acc < best) { best = acc; bi = c; }
        }
        dst.put(head, t, m0 + mi, bi);
    }
}

template <typename T>
static int launch_encode_t(const T* x, int64_t xhs, const float* cent, const CodeDst& dst, int n_heads, int n_tokens,
                           int d, int M, int C, cudaStream_t stream) {
    const int dm = d / M;
    int MG = 8;
    while (MG > 1 && (size_t)MG * C * dm * sizeof(float) > 64 * 1024) MG >>= 1;
    const size_t smem = (size_t)MG * C * dm * sizeof(float);
    if (smem > 200 * 1024) MILLION_UNSUPPORTED("encode: C*d_m too large (C=%d d_m=%d)", C, dm);
    dim3 grid((n_tokens + 127) / 128, (M + MG - 1) / MG, n_heads), block(128);
#define MILLION_ENC_CASE(DMV)                                                                                      \
    case DMV:                                                                                                      \
        MILLION_CUDA_OK(cudaFuncSetAttribute(encode_generic_kernel<T, DMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        encode_generic_kernel<T, DMV><<<grid, block, smem, stream>>>(x, xhs, cent, dst, n_tokens, d, M, C, MG);    \
        break;
    switch (dm) {
        MILLION_ENC_CASE(1) MILLION_ENC_CASE(2) MILLION_ENC_CASE(4) MILLION_ENC_CASE(8) MILLION_ENC_CASE(16)
        default: MILLION_UNSUPPORTED("encode: d/M = %d not in {1,2,4,8,16}", dm);
    }
#undef MILLION_ENC_CASE
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

int launch_encode_generic(const void* x, int x_dtype, int64_t xhs, const float* cent, const CodeDst& dst, int n_heads,
                          int n_tokens, int d, int M, int C, cudaStream_t stream) {
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    switch (x_dtype) {
        case MILLION_F16: return launch_encode_t((const __half*)x, xhs, cent, dst, n_heads, n_tokens, d, M, C, stream);
        case MILLION_BF16: return launch_encode_t((const __nv_bfloat16*)x, xhs, cent, dst, n_heads, n_tokens, d, M, C, stream);
        case MILLION_F32: return launch_encode_t((const float*)x, xhs, cent, dst, n_heads, n_tokens, d, M, C, stream);
    }
    MILLION_UNSUPPORTED("encode: unknown x dtype %d", x_dtype);
}

// ------------------------------------------------------------------------------------------------ reconstruct
template <int ELEM_BYTES>
__global__ void reconstruct_kernel(const void* __restrict__ codes, int code_bytes, int64_t chs, int64_t cts, int64_t cms,
                                   const unsigned char* __restrict__ cent, unsigned char* __restrict__ out,
                                   int64_t out_head_stride, int n_tokens, int d, int M, int C) {
    const int head = blockIdx.y;
    const int dm = d / M;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n_tokens * M
    if (idx >= (int64_t)n_tokens * M) return;
    const int t = (int)(idx / M), m = (int)(idx % M);
    const int64_t off = head * chs + t * cts + m * cms;
    const int code = code_bytes == 1 ? reinterpret_cast<const uint8_t*>(codes)[off] : reinterpret_cast<const uint16_t*>(codes)[off];
    const unsigned char* src = cent + ((int64_t)m * C + code) * dm * ELEM_BYTES;
    unsigned char* dstp = out + (head * out_head_stride + (int64_t)t * d + m * dm) * ELEM_BYTES;
    for (int k = 0; k < dm * ELEM_BYTES; ++k) dstp[k] = src[k];
}

int launch_reconstruct(const void* codes, int code_bytes, int64_t chs, int64_t cts, int64_t cms, const void* cent, void* out,
                       int dtype, int64_t ohs, int n_heads, int n_tokens, int d, int M, int C, cudaStream_t stream) {
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    dim3 grid((unsigned)(((int64_t)n_tokens * M + 255) / 256), n_heads), block(256);
    if (dtype == MILLION_F32)
        reconstruct_kernel<4><<<grid, block, 0, stream>>>(codes, code_bytes, chs, cts, cms, (const unsigned char*)cent, (unsigned char*)out, ohs, n_tokens, d, M, C);
    else
        reconstruct_kernel<2><<<grid, block, 0, stream>>>(codes, code_bytes, chs, cts, cms, (const unsigned char*)cent, (unsigned char*)out, ohs, n_tokens, d, M, C);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

// ------------------------------------------------------------------------------------------------ window ops
// copy n rows of d elements (elem_bytes each) per head for K and V in one launch; 16-byte vectors
__global__ void rows_copy2_kernel(unsigned char* k_dst, unsigned char* v_dst, int64_t dst_head_stride_b, int64_t dst_off_b,
                                  const unsigned char* k_src, const unsigned char* v_src, int64_t src_head_stride_b, int64_t src_off_b,
                                  int64_t bytes_per_head) {
    const int head = blockIdx.y;
    const bool isv = blockIdx.z;
    unsigned char* dstp = (isv ? v_dst : k_dst) + head * dst_head_stride_b + dst_off_b;
    const unsigned char* srcp = (isv ? v_src : k_src) + head * src_head_stride_b + src_off_b;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < bytes_per_head; i += (int64_t)gridDim.x * blockDim.x * 16)
        *reinterpret_cast<uint4*>(dstp + i) = *reinterpret_cast<const uint4*>(srcp + i);
}

int launch_rows_copy2(void* k_dst, void* v_dst, int64_t dst_hs_b, int64_t dst_off_b, const void* k_src, const void* v_src,
                      int64_t src_hs_b, int64_t src_off_b, int n_heads, int64_t bytes_per_head, cudaStream_t stream) {
    if (n_heads == 0 || bytes_per_head == 0) return MILLION_OK;
    if (bytes_per_head % 16 || dst_off_b % 16 || src_off_b % 16 || dst_hs_b % 16 || src_hs_b % 16) {
        set_error("window copy needs 16-byte aligned rows (d*elem %% 16 == 0)");
        return MILLION_ERR_INVALID;
    }
    const int threads = 128;
    int bx = (int)((bytes_per_head / 16 + threads - 1) / threads);
    if (bx > 64) bx = 64;
    dim3 grid(bx, n_heads, 2);
    rows_copy2_kernel<<<grid, threads, 0, stream>>>((unsigned char*)k_dst, (unsigned char*)v_dst, dst_hs_b, dst_off_b,
                                                    (const unsigned char*)k_src, (const unsigned char*)v_src, src_hs_b, src_off_b, bytes_per_head);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

// The new token's K and V rows -> window row (*r_dev + r_off), clamped to the window: the append of pq_utils.py:304-311 with the
// row index read on the device (CUDA-graph replay of a decode step).  Used when the attention kernel does not append itself.
__global__ void window_append_dev_kernel(unsigned char* k_win, unsigned char* v_win, int64_t win_hs_b, const unsigned char* k_new,
                                         const unsigned char* v_new, int row_bytes, const int* r_dev, int r_off, int res_len) {
    int row = (r_dev ? __ldg(r_dev) : 0) + r_off;
    row = row < 0 ? 0 : (row >= res_len ? res_len - 1 : row);
    const int head = blockIdx.x;
    const bool isv = blockIdx.y;
    unsigned char* dst = (isv ? v_win : k_win) + head * win_hs_b + (int64_t)row * row_bytes;
    const unsigned char* src = (isv ? v_new : k_new) + (int64_t)head * row_bytes;
    for (int i = threadIdx.x * 16; i < row_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}

int launch_window_append_dev(void* k_win, void* v_win, int64_t win_hs_b, const void* k_new, const void* v_new, int n_heads, int row_bytes,
                             const int* r_dev, int r_off, int res_len, cudaStream_t stream) {
    if (n_heads == 0) return MILLION_OK;
    if (row_bytes % 16 || win_hs_b % 16 || ((uintptr_t)k_win | (uintptr_t)v_win | (uintptr_t)k_new | (uintptr_t)v_new) % 16) {
        set_error("window append needs 16-byte aligned rows");
        return MILLION_ERR_INVALID;
    }
    window_append_dev_kernel<<<dim3(n_heads, 2), 32, 0, stream>>>((unsigned char*)k_win, (unsigned char*)v_win, win_hs_b, (const unsigned char*)k_new,
                                                                 (const unsigned char*)v_new, row_bytes, r_dev, r_off, res_len);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

__global__ void counter_add_kernel(int* ctr, int n, int delta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ctr[i] += delta;
}

int launch_counter_add(int* ctr, int n, int delta, cudaStream_t stream) {
    if (n == 0) return MILLION_OK;
    counter_add_kernel<<<(n + 127) / 128, 128, 0, stream>>>(ctr, n, delta);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

// ---------------------------------------------------------------- decode-token producer step: rotary embedding of q and k
// out = x * cos + rotate_half(x) * sin for the ONE new token of every sequence (q_len = 1): the expression of
// transformers' apply_rotary_pos_emb, which the reference calls in its attention forward (scripts/modeldb/models/
// modeling_llama.py:500-512) and which costs ~10 elementwise launches per layer there.  Roundings follow the elementwise
// evaluation in the I/O dtype (each product and the sum computed in fp32 and rounded to T), so the result is bit-identical to the
// torch expression.  One block per (sequence, head) row, one thread per dimension pair (i, i + d/2).
template <typename T>
__global__ void rope_qk_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ cos_, const T* __restrict__ sin_,
                               T* __restrict__ q_out, T* __restrict__ k_out, int nh, int nh_k, int d) {
    const int row = blockIdx.x, heads = nh + nh_k;
    const int b = row / heads, h = row - b * heads;
    const T* x = h < nh ? q + ((int64_t)b * nh + h) * d : k + ((int64_t)b * nh_k + (h - nh)) * d;
    T* y = h < nh ? q_out + ((int64_t)b * nh + h) * d : k_out + ((int64_t)b * nh_k + (h - nh)) * d;
    const T* c = cos_ + (int64_t)b * d;
    const T* s = sin_ + (int64_t)b * d;
    const int half_d = d >> 1;
    for (int i = threadIdx.x; i < half_d; i += blockDim.x) {
        const float x1 = io<T>::to_f(x[i]), x2 = io<T>::to_f(x[i + half_d]);
        const float a0 = io<T>::to_f(io<T>::from_f(__fmul_rn(x1, io<T>::to_f(c[i]))));
        const float b0 = io<T>::to_f(io<T>::from_f(__fmul_rn(-x2, io<T>::to_f(s[i]))));
        const float a1 = io<T>::to_f(io<T>::from_f(__fmul_rn(x2, io<T>::to_f(c[i + half_d]))));
        const float b1 = io<T>::to_f(io<T>::from_f(__fmul_rn(x1, io<T>::to_f(s[i + half_d]))));
        y[i] = io<T>::from_f(__fadd_rn(a0, b0));
        y[i + half_d] = io<T>::from_f(__fadd_rn(a1, b1));
    }
}

int launch_rope_qk(const void* q, const void* k, const void* cos_, const void* sin_, void* q_out, void* k_out, int dtype, int bs, int nh,
                   int nh_k, int d, cudaStream_t stream) {
    if (bs == 0) return MILLION_OK;
    const int rows = bs * (nh + nh_k), threads = d / 2 <= 32 ? 32 : (d / 2 <= 64 ? 64 : 128);
    if (dtype == MILLION_F16)
        rope_qk_kernel<__half><<<rows, threads, 0, stream>>>((const __half*)q, (const __half*)k, (const __half*)cos_, (const __half*)sin_,
                                                            (__half*)q_out, (__half*)k_out, nh, nh_k, d);
    else if (dtype == MILLION_BF16)
        rope_qk_kernel<__nv_bfloat16><<<rows, threads, 0, stream>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)cos_,
                                                                   (const __nv_bfloat16*)sin_, (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_out, nh, nh_k, d);
    else
        rope_qk_kernel<float><<<rows, threads, 0, stream>>>((const float*)q, (const float*)k, (const float*)cos_, (const float*)sin_,
                                                           (float*)q_out, (float*)k_out, nh, nh_k, d);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

}  // namespace million
