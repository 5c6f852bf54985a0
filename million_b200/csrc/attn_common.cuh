// attn_common.cuh — argument block, split planning and the cross-split merge shared by the generic and
// the fast decode-attention kernels.
#pragma once
#include "common.cuh"

namespace million {

// Device-side view of million_attn_params (pointers typed, strides in bytes).
struct AttnArgs {
    const void* q;
    const uint8_t* k_codes;
    const uint8_t* v_codes;
    const int64_t* v_page_ids;
    const void* k_cent;
    const void* v_cent;
    const void* k_res;
    const void* v_res;
    void* out;
    float* partial_out;   // PARTIAL_ONLY target or nullptr
    int* counters;        // (bs*nh_k) arrival tickets, zero on entry and on exit
    float* parts;         // (bs*nh, n_parts, d+2) fp32: [o_unnormalised | m (log2 units) | l]
    int64_t k_head_stride, v_head_stride, v_ld;
    int bs, nh, nh_k, d, M, C, nk, r, res_len;
    int v_layout, page_size, n_pages;
    int n_splits;         // CTAs over the coded tokens; part index n_splits = the fp16 window
    int units_per_split;  // 16-token units per split
    float scale_log2;     // log2(e)/sqrt(d)
};

// tokens [begin, end) of split `s`
__device__ __forceinline__ void split_range(const AttnArgs& a, int s, int& begin, int& end) {
    const long long b = (long long)s * a.units_per_split * 16;
    const long long e = b + (long long)a.units_per_split * 16;
    begin = (int)(b < a.nk ? b : a.nk);
    end = (int)(e < a.nk ? e : a.nk);
}

// V code of token j, sub-space m for head block hb = b*nh_k + hk  (generic path; the fast path stages tiles)
__device__ __forceinline__ int v_code_at(const AttnArgs& a, int hb, int j, int m) {
    if (a.v_layout == MILLION_V_ROWMAJOR) return a.v_codes[hb * a.v_head_stride + (int64_t)j * a.M + m];
    if (a.v_layout == MILLION_V_TRANSPOSED) return a.v_codes[hb * a.v_head_stride + (int64_t)m * a.v_ld + j];
    const int64_t page = a.v_page_ids[(int64_t)hb * a.n_pages + j / a.page_size];
    return a.v_codes[(page * a.M + m) * a.page_size + (j % a.page_size)];
}

// Merge the n_splits+1 partial states of every query head of group (b, hk) and write the result.
// Called by the last CTA of the group (after the ticket), all threads of the block participate.
// Algebra of flash_decoding_reduce_kernel (Kernel.cuh:1249-1269) on un-normalised fp32 partials.
template <typename T>
__device__ __forceinline__ void merge_group(const AttnArgs& a, int b, int hk) {
    const int G = a.nh / a.nh_k;
    const int n_parts = a.n_splits + 1;
    const int stride = a.d + 2;
    for (int g = 0; g < G; ++g) {
        const int h = hk * G + g;
        const float* base = a.parts + ((int64_t)(b * a.nh + h) * n_parts) * stride;
        float mstar = -INFINITY;
        for (int i = 0; i < n_parts; ++i) {
            const float l = __ldcg(base + (int64_t)i * stride + a.d + 1);
            const float m = __ldcg(base + (int64_t)i * stride + a.d);
            if (l > 0.f) mstar = fmaxf(mstar, m);
        }
        float den = 0.f;
        for (int i = 0; i < n_parts; ++i) {
            const float l = __ldcg(base + (int64_t)i * stride + a.d + 1);
            const float m = __ldcg(base + (int64_t)i * stride + a.d);
            if (l > 0.f) den += l * exp2f(m - mstar);
        }
        for (int k = threadIdx.x; k < a.d; k += blockDim.x) {
            float acc = 0.f;
            for (int i = 0; i < n_parts; ++i) {
                const float l = __ldcg(base + (int64_t)i * stride + a.d + 1);
                const float m = __ldcg(base + (int64_t)i * stride + a.d);
                if (l > 0.f) acc += __ldcg(base + (int64_t)i * stride + k) * exp2f(m - mstar);
            }
            if (a.partial_out) {
                a.partial_out[(int64_t)(b * a.nh + h) * stride + k] = acc;
            } else {
                reinterpret_cast<T*>(a.out)[(int64_t)(b * a.nh + h) * a.d + k] =
                    io<T>::from_f(den > 0.f ? acc / den : 0.f);
            }
        }
        if (a.partial_out && threadIdx.x == 0) {
            a.partial_out[(int64_t)(b * a.nh + h) * stride + a.d] = mstar * kLn2;  // natural-log units
            a.partial_out[(int64_t)(b * a.nh + h) * stride + a.d + 1] = den;
        }
    }
}

// Arrival ticket: returns true in every thread of the LAST CTA of group `grp` (of `expected` CTAs).
__device__ __forceinline__ bool last_cta_of_group(int* counters, int grp, int expected, int* smem_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(counters + grp, 1);
        const int last = (t == expected - 1);
        if (last) counters[grp] = 0;  // leave the counter clean for the next launch
        *smem_flag = last;
    }
    __syncthreads();
    const bool last = (*smem_flag != 0);
    if (last) __threadfence();
    return last;
}

}  // namespace million
