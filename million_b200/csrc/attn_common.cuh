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
    int n_splits;         // CTAs over the coded tokens
    int n_parts;          // part SLOTS per head in `parts` (stride); how many are valid is decided per group
    int ws_parts;         // part slots per head that fit in the caller's workspace
    int auto_splits;      // the caller left the split count to us (flat scheduling allowed)
    int flat, flat_per, flat_ug;   // flat scheduling: CTA c owns 64-token units [c*flat_per, (c+1)*flat_per) of the (group, unit) space
    int units_per_split;  // 16-token units per split
    float scale_log2;     // log2(e)/sqrt(d)
    unsigned long long* dbg_timing;   // optional (million_debug_set_timing_buffer): 8 globaltimer stamps per CTA
};

__device__ __forceinline__ void dbg_stamp(const AttnArgs& a, int slot) {
    if (a.dbg_timing && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        a.dbg_timing[cta * 8 + slot] = t;
    }
}

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

// Merge the first n_parts partial states of every query head of group (b, hk) and write the result.
// Called by the last CTA of the group (after the ticket), all threads of the block participate.
// Algebra of flash_decoding_reduce_kernel (Kernel.cuh:1249-1269) on un-normalised fp32 partials.
// All heads are merged together in two rounds of independent loads (one for (m, l), one for o): a serial chain of L2
// round trips per head and per part used to put 10 us on the tail of every batch-1 launch.
// `scr` = shared scratch of kMergeScratch floats.
constexpr int kMergeScratch = 3 * 2048 + 64;
template <typename T>
__device__ __forceinline__ void merge_group(const AttnArgs& a, int b, int hk, int n_parts, float* scr) {
    const int G = a.nh / a.nh_k;
    const int slots = a.n_parts;
    const int stride = a.d + 2;
    float* mm = scr;               // [gc][n_parts] running max of each part
    float* ll = scr + 2048;        // [gc][n_parts] denominators
    float* ww = scr + 4096;        // [gc][n_parts] merge weights
    float* hd = scr + 6144;        // [gc][2]: m*, den
    int gc_max = 2048 / n_parts;
    if (gc_max < 1) gc_max = 1;    // n_parts <= 1024 is enforced by the host
    for (int g0 = 0; g0 < G; g0 += gc_max) {
        const int gc = min(gc_max, G - g0);
        const int h0 = hk * G + g0;
        const float* base = a.parts + ((int64_t)(b * a.nh + h0) * slots) * stride;   // head g, part i at (g*slots + i)*stride
        for (int idx = threadIdx.x; idx < gc * n_parts; idx += blockDim.x) {
            const int64_t off = (int64_t)((idx / n_parts) * slots + idx % n_parts) * stride;
            mm[idx] = __ldcg(base + off + a.d);
            ll[idx] = __ldcg(base + off + a.d + 1);
        }
        __syncthreads();
        if (threadIdx.x < gc) {
            const int g = threadIdx.x;
            float mstar = -INFINITY;
            for (int i = 0; i < n_parts; ++i)
                if (ll[g * n_parts + i] > 0.f) mstar = fmaxf(mstar, mm[g * n_parts + i]);
            float den = 0.f;
            for (int i = 0; i < n_parts; ++i) {
                const float l = ll[g * n_parts + i];
                const float w = (l > 0.f) ? exp2f(mm[g * n_parts + i] - mstar) : 0.f;
                ww[g * n_parts + i] = w;
                den += l * w;
            }
            hd[2 * g] = mstar;
            hd[2 * g + 1] = den;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < gc * a.d; idx += blockDim.x) {
            const int g = idx / a.d, k = idx % a.d;
            const float* src = base + (int64_t)g * slots * stride + k;
            const float* w = ww + g * n_parts;
            float acc = 0.f;
            int i = 0;
            for (; i + 8 <= n_parts; i += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (int64_t)(i + u) * stride);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = fmaf(v[u], w[i + u], acc);
            }
            for (; i < n_parts; ++i) acc = fmaf(__ldcg(src + (int64_t)i * stride), w[i], acc);
            const int h = h0 + g;
            if (a.partial_out) {
                a.partial_out[(int64_t)(b * a.nh + h) * stride + k] = acc;
                if (k == 0) {
                    a.partial_out[(int64_t)(b * a.nh + h) * stride + a.d] = hd[2 * g] * kLn2;   // natural-log units
                    a.partial_out[(int64_t)(b * a.nh + h) * stride + a.d + 1] = hd[2 * g + 1];
                }
            } else {
                const float den = hd[2 * g + 1];
                reinterpret_cast<T*>(a.out)[(int64_t)(b * a.nh + h) * a.d + k] = io<T>::from_f(den > 0.f ? acc / den : 0.f);
            }
        }
        __syncthreads();
    }
}

// Arrival ticket: returns true in every thread of the LAST CTA of group `grp` (of `expected` CTAs).
// The block's partial-state writes are ordered before the ticket by bar.sync + an acq_rel RMW at gpu scope (cheaper than a
// full __threadfence in every thread); the last CTA's loads (ld.cg after the second bar.sync) observe all of them.
__device__ __forceinline__ bool last_cta_of_group(int* counters, int grp, int expected, int* smem_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        int t;
        asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(t) : "l"(counters + grp) : "memory");
        const int last = (t == expected - 1);
        if (last) counters[grp] = 0;  // leave the counter clean for the next launch
        *smem_flag = last;
    }
    __syncthreads();
    return *smem_flag != 0;
}

}  // namespace million
