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
    float* partial_out;   // PARTIAL_ONLY target or nullptr; the P2PState block in the fused split-KV instantiations (P2P = true)
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
    // outlier side store (0 / nullptr = disabled): records (dim u8, delta io dtype) per coded token, head strides in records
    int k_out, v_out;
    const uint8_t* ko_idx; const void* ko_val; int64_t ko_head_stride;
    const uint8_t* vo_idx; const void* vo_val; int64_t vo_head_stride;
    // fused window append: the new token's key/value (bs, nh_k, d), stored into window row r-1 by the CTA that owns that row
    const void* k_new; const void* v_new;
    const int* r_dev;                 // device-resident window length: r = min(*r_dev + r, res_len) (graph replay), or nullptr
    int code_bytes;                   // 1: uint8 codes (C <= 256); 2: uint16 codes (nbits2dtype for nbits > 8, C <= 65536) — generic kernel only
    int direct;                       // generic kernel: the M x C LUT does not fit shared memory, scores come from the centroids
    int pdl;                          // host only: launch with the programmatic-serialization attribute (MILLION_ATTN_PDL)
    int p2p;                          // host only: fused split-KV instantiation (partial_out = P2PState)
    uint32_t zero;                    // always 0, unknown to the compiler: orders the outlier-record prefetches (attn_fast_helpers.cuh rec_take)
#ifdef MILLION_DEBUG
    int dbg_mode;                     // ablation switch (million_debug_set_mode): bit 0 skips QK gathers, bit 1 skips PV, bit 2 tile loads
    unsigned long long* dbg_timing;   // optional (million_debug_set_timing_buffer): 64 words per CTA, see dbg_stamp
#endif
};

// window rows valid for this launch (see million_attn_params.r_dev)
__device__ __forceinline__ int window_rows(const AttnArgs& a) {
    if (a.r_dev == nullptr) return a.r;
    const int r = __ldg(a.r_dev) + a.r;
    return r < a.res_len ? r : a.res_len;
}

// Debug instrumentation exists only in -DMILLION_DEBUG builds (tools/attn_phases.py, tools/ablate.py): product builds carry
// neither the globals nor the run-time branches.
#ifdef MILLION_DEBUG
// 64 words per CTA: raw globaltimer stamps [piece 0..3][slot 0..15] (stores only: a stamp must not wait on memory)
__device__ __forceinline__ void dbg_stamp(unsigned long long* dbg_timing, int piece, int slot) {
    if (dbg_timing && threadIdx.x == 0 && piece < 4) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg_timing[(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 64 + piece * 16 + slot] = t;
    }
}
__device__ __forceinline__ void dbg_stamp(const AttnArgs& a, int slot, int piece = 0) { dbg_stamp(a.dbg_timing, piece, slot); }
#define MILLION_DBG_MODE(a, bit) ((a).dbg_mode & (bit))
#else
__device__ __forceinline__ void dbg_stamp(const AttnArgs&, int, int = 0) {}
#define MILLION_DBG_MODE(a, bit) 0
#endif

// tokens [begin, end) of split `s`
__device__ __forceinline__ void split_range(const AttnArgs& a, int s, int& begin, int& end) {
    const long long b = (long long)s * a.units_per_split * 16;
    const long long e = b + (long long)a.units_per_split * 16;
    begin = (int)(b < a.nk ? b : a.nk);
    end = (int)(e < a.nk ? e : a.nk);
}

// code at BYTE offset `base` + element index `idx` of a one- or two-byte code array (strides of the ABI are in bytes)
__device__ __forceinline__ int code_load(const uint8_t* p, int64_t base, int64_t idx, int code_bytes) {
    return code_bytes == 2 ? (int)reinterpret_cast<const uint16_t*>(p + base)[idx] : (int)p[base + idx];
}
// K code of token j, sub-space m (row-major)
__device__ __forceinline__ int k_code_at(const AttnArgs& a, int hb, int j, int m) {
    return code_load(a.k_codes, hb * a.k_head_stride, (int64_t)j * a.M + m, a.code_bytes);
}
// V code of token j, sub-space m for head block hb = b*nh_k + hk  (generic path; the fast path stages tiles)
__device__ __forceinline__ int v_code_at(const AttnArgs& a, int hb, int j, int m) {
    if (a.v_layout == MILLION_V_ROWMAJOR) return code_load(a.v_codes, hb * a.v_head_stride, (int64_t)j * a.M + m, a.code_bytes);
    if (a.v_layout == MILLION_V_TRANSPOSED) return code_load(a.v_codes, hb * a.v_head_stride + (int64_t)m * a.v_ld, j, a.code_bytes);
    const int64_t page = a.v_page_ids[(int64_t)hb * a.n_pages + j / a.page_size];
    return code_load(a.v_codes, 0, (page * a.M + m) * a.page_size + (j % a.page_size), a.code_bytes);
}

// Merge the first n_parts partial states of every query head of group (b, hk) and write the result.
// Called by the last CTA of the group (after the ticket), all threads of the block participate.
// Algebra of flash_decoding_reduce_kernel (Kernel.cuh:1249-1269) on un-normalised fp32 partials.
// All heads are merged together in two rounds of independent loads (one for (m, l), one for o): a serial chain of L2
// round trips per head and per part used to put 10 us on the tail of every batch-1 launch.
// `scr` = shared scratch of kMergeScratch floats.
constexpr int kMergeScratch = 3 * 2048 + 64;
// floats per partial-state row [o (d) | m | l] in the kernels' workspace, padded to 16 bytes so that rows can be bulk-copied
__host__ __device__ __forceinline__ constexpr int part_stride(int d) { return (d + 2 + 3) & ~3; }

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Symmetric buffer of the split-KV exchange (same layout on every rank; million_splitkv_symmetric_bytes):
//   [0, 1024)        per-source-rank flags of million_splitkv_push_merge: rank g at byte 128*g (u32 sequence number)
//   [1024, ...)      recv[parity 2][world][rows][d+2] fp32 (million_splitkv_push_merge)
//   [p2p_ll_offset)  the same layout in tagged 8-byte cells (exchange fused into the attention kernel)
constexpr int kP2PRecvOff = 1024;
// The exchange fused into the attention kernel uses a second receive area right behind the first: the same
// [parity][world][rows][d+2] layout in 8-byte cells {fp32 value, u32 sequence number}.  A value and the tag that validates it
// travel in ONE 8-byte store, so the receiver polls the data itself: no release fence (a round trip over NVLink before a flag
// may be published), no flag — the latency of the exchange is one store's flight time (the idea of NCCL's LL protocol).
__host__ __device__ __forceinline__ constexpr int64_t p2p_ll_offset(int world, int64_t rows, int d) {
    return kP2PRecvOff + (int64_t)2 * world * rows * (d + 2) * 4;
}
__device__ __forceinline__ void st_tagged(float* cell, float v, unsigned tag) {
    asm volatile("st.volatile.global.v2.b32 [%0], {%1, %2};" ::"l"(cell), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// device-memory block of the split-KV protocol (million_splitkv_state_init): read only on the merge path
struct P2PState {
    unsigned counter; int ticket; int err; int spin_limit;   // spin_limit: polls (100 ns apart) before a wait gives up; 0 = default
    int rank, world, rows, pad1;     // rows = bs * nh of the calls this block is used with
    unsigned char* peer[8];
};
constexpr int kP2PDefaultSpins = 1 << 23;    // ~0.8 s

struct MergeArgs {
    int nh, nh_k, n_parts, d;
    const float* parts;
    float* partial_out;
    void* out;
    float* big;            // optional large shared staging buffer lent by the caller (dead tables), big_floats floats
    int big_floats;
    unsigned long long* bar;   // 8 bytes of shared memory for the mbarrier of the staged path
#ifdef MILLION_DEBUG
    unsigned long long* dbg_timing;
    int dbg_piece;
#endif
};
#ifdef MILLION_DEBUG
__device__ __forceinline__ void dbg_stamp_m(const MergeArgs& a, int slot) { dbg_stamp(a.dbg_timing, a.dbg_piece, slot); }
#else
__device__ __forceinline__ void dbg_stamp_m(const MergeArgs&, int) {}
#endif
// Out of line on purpose: it runs once per group, and inlined into the attention kernels its registers and code perturb the
// allocation and layout of their main loops (measured: +7 us per launch at batch 8 for the same loop rate).
// P2P = true (split-KV across GPUs fused into the attention launch, MILLION_ATTN_FUSED_SPLITKV): a.partial_out is the P2PState
// block; the merged rows of THIS group go straight into every peer's receive buffer, a per-(rank, group) flag is published, and
// the same CTA waits for the peers' flags of its group and writes the final rows — no cross-group ticket on the critical path,
// no second launch.  A separate instantiation on purpose: code in the plain merge perturbs the attention kernels that call it
// (measured in round 1: +2.5 % per launch at batch 8).
template <typename T, bool P2P>
__device__ __noinline__ void merge_group_impl(const MergeArgs a, int b, int hk, int n_parts, float* scr) {
    const int G = a.nh / a.nh_k;
    constexpr bool p2p = P2P;
    P2PState* const ps = reinterpret_cast<P2PState*>(a.partial_out);
    unsigned p2p_seq = 0, p2p_par = 0;
    int p2p_rank = 0, p2p_world = 0, p2p_rows = 0;
    size_t p2p_slot = 0;
    if constexpr (P2P) {
        p2p_rows = ps->rows;
        p2p_slot = (size_t)p2p_rows * (a.d + 2);   // the counter moves only after every group of this launch is done: all group-last CTAs read the same value
        p2p_seq = *reinterpret_cast<volatile unsigned*>(&ps->counter) + 1;
        p2p_par = p2p_seq & 1;
        p2p_rank = ps->rank; p2p_world = ps->world;
    }
    // the peers' buffer addresses, read once: inside the store loop below every load would sit behind the previous volatile store
    // (a chain of L2 round trips that grows with the world size)
    unsigned char* p2p_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if constexpr (P2P) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < p2p_world) p2p_peer[r] = ps->peer[r];
    }
    const int slots = a.n_parts;
    const int stride = part_stride(a.d);    // workspace rows
    const int ostride = a.d + 2;            // rows of partial_out (public layout)
    float* mm = scr;               // [gc][n_parts] running max of each part
    float* ll = scr + 2048;        // [gc][n_parts] denominators
    float* ww = scr + 4096;        // [gc][n_parts] merge weights
    float* hd = scr + 6144;        // [gc][2]: m*, den
#ifndef MILLION_MERGE_DIRECT_PARTS
#define MILLION_MERGE_DIRECT_PARTS 8     // A/B (us per launch at 32K, batch 8 / 4 / 2 / 1): staged merge only 97.3 / 55.9 / 32.0 / 22.5; direct up to 4 parts 93.6 / 55.8 / 32.1 / 22.5; up to 8: 94.0 / 52.2 / 32.0 / 22.4; up to 12: 94.3 / 52.6 / 31.3 / -; up to 20: 95.4 / 53.8 / 32.0 / 22.8
#endif
    if constexpr (!P2P) {
        // Few parts (large batches: 2-3 CTAs per group): no staging, no barriers — every thread requests the few values of its
        // outputs and the parts' (m, l) pairs at once (one L2 round trip) and merges them in registers.
        if (n_parts <= MILLION_MERGE_DIRECT_PARTS) {
            for (int i = threadIdx.x; i < G * a.d; i += blockDim.x) {
                const int g = i / a.d, k = i - g * a.d, h = hk * G + g;
                const float* base = a.parts + ((int64_t)(b * a.nh + h) * slots) * stride;
                float o[MILLION_MERGE_DIRECT_PARTS > 0 ? MILLION_MERGE_DIRECT_PARTS : 1], m[MILLION_MERGE_DIRECT_PARTS > 0 ? MILLION_MERGE_DIRECT_PARTS : 1],
                    l[MILLION_MERGE_DIRECT_PARTS > 0 ? MILLION_MERGE_DIRECT_PARTS : 1];
#pragma unroll
                for (int p = 0; p < MILLION_MERGE_DIRECT_PARTS; ++p) {
                    const bool on = p < n_parts;
                    o[p] = on ? __ldcg(base + (int64_t)p * stride + k) : 0.f;
                    m[p] = on ? __ldcg(base + (int64_t)p * stride + a.d) : -INFINITY;
                    l[p] = on ? __ldcg(base + (int64_t)p * stride + a.d + 1) : 0.f;
                }
                float mstar = -INFINITY;
#pragma unroll
                for (int p = 0; p < MILLION_MERGE_DIRECT_PARTS; ++p)
                    if (l[p] > 0.f) mstar = fmaxf(mstar, m[p]);
                float acc = 0.f, den = 0.f;
#pragma unroll
                for (int p = 0; p < MILLION_MERGE_DIRECT_PARTS; ++p) {
                    const float w = l[p] > 0.f ? exp2f(m[p] - mstar) : 0.f;
                    acc = fmaf(o[p], w, acc);
                    den = fmaf(l[p], w, den);
                }
                if (a.partial_out) {
                    a.partial_out[(int64_t)(b * a.nh + h) * ostride + k] = acc;
                    if (k == 0) {
                        a.partial_out[(int64_t)(b * a.nh + h) * ostride + a.d] = mstar * kLn2;   // natural-log units
                        a.partial_out[(int64_t)(b * a.nh + h) * ostride + a.d + 1] = den;
                    }
                } else {
                    reinterpret_cast<T*>(a.out)[(int64_t)(b * a.nh + h) * a.d + k] = io<T>::from_f(den > 0.f ? acc / den : 0.f);
                }
            }
            return;
        }
    }
    int gc_max = 2048 / n_parts;
    if (gc_max < 1) gc_max = 1;    // n_parts <= 1024 is enforced by the host
    // many parts (one group spread over every SM: KV-head sharding at batch 1): take as many heads per round as the lent
    // staging buffer holds, rather than falling back to strided L2 loads (19 dependent round trips for 148 parts)
    if (a.big != nullptr && a.bar != nullptr) {
        const int fit = (int)(a.big_floats / ((int64_t)n_parts * stride));
        if (fit >= 1 && fit < gc_max) gc_max = fit;
    }
    for (int g0 = 0; g0 < G; g0 += gc_max) {
        const int gc = min(gc_max, G - g0);
        const int h0 = hk * G + g0;
        const float* base = a.parts + ((int64_t)(b * a.nh + h0) * slots) * stride;   // head g, part i at (g*slots + i)*stride
        // Staged path (the caller lent a big shared buffer and everything fits): one TMA bulk copy per head brings all its
        // partial rows to shared memory, so the merge costs ONE L2 round trip and a handful of instructions.  It runs on a
        // different SM every launch (its code is cold) and it is the serial tail of the launch: per-element copies
        // (cp.async, ld.cg) spent 2-3 us here just being issued.
        const bool staged = a.big != nullptr && a.bar != nullptr && (int64_t)gc * n_parts * stride <= a.big_floats;
        if (staged) {
            const uint32_t big_s = (uint32_t)__cvta_generic_to_shared(a.big), bar_s = (uint32_t)__cvta_generic_to_shared(a.bar);
            if (n_parts < 8) {
                // few rows (large batches: 2-3 parts per group, merged in mid-kernel): 16-byte cp.async, a warp per row, is
                // cheaper than setting up a bulk copy
                for (int g = 0; g < gc; ++g)
                    for (int i = threadIdx.x >> 5; i < n_parts; i += blockDim.x >> 5) {
                        const float* src = base + (int64_t)(g * slots + i) * stride;
                        const uint32_t dst = big_s + (uint32_t)((g * n_parts + i) * stride) * 4;
                        for (int c = threadIdx.x & 31; c < stride / 4; c += 32)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)c * 16), "l"(src + 4 * c) : "memory");
                    }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            } else if (threadIdx.x == 0) {
                const uint32_t bytes = (uint32_t)(n_parts * stride * 4);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
                asm volatile("fence.proxy.async;" ::: "memory");   // barrier init + earlier generic accesses of `big` / other CTAs' rows
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes * (uint32_t)gc) : "memory");
                for (int g = 0; g < gc; ++g)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(big_s + (uint32_t)g * bytes),
                                 "l"(base + (int64_t)g * slots * stride), "r"(bytes), "r"(bar_s)
                                 : "memory");
                uint32_t done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_s) : "memory");
                asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar_s) : "memory");
            }
            dbg_stamp_m(a, 11);
            __syncthreads();
            dbg_stamp_m(a, 12);
            for (int idx = threadIdx.x; idx < gc * n_parts; idx += blockDim.x) {
                mm[idx] = a.big[idx * stride + a.d];
                ll[idx] = a.big[idx * stride + a.d + 1];
            }
        } else {
            for (int idx = threadIdx.x; idx < gc * n_parts; idx += blockDim.x) {
                const int64_t off = (int64_t)((idx / n_parts) * slots + idx % n_parts) * stride;
                mm[idx] = __ldcg(base + off + a.d);
                ll[idx] = __ldcg(base + off + a.d + 1);
            }
        }
        __syncthreads();
        dbg_stamp_m(a, 9);
        // weights: one warp per head, lanes over the parts
        for (int g = threadIdx.x >> 5; g < gc; g += blockDim.x >> 5) {
            const int lane = threadIdx.x & 31;
            float mstar = -INFINITY;
            for (int i = lane; i < n_parts; i += 32)
                if (ll[g * n_parts + i] > 0.f) mstar = fmaxf(mstar, mm[g * n_parts + i]);
            mstar = warp_max(mstar);
            float den = 0.f;
            for (int i = lane; i < n_parts; i += 32) {
                const float l = ll[g * n_parts + i];
                const float w = (l > 0.f) ? exp2f(mm[g * n_parts + i] - mstar) : 0.f;
                ww[g * n_parts + i] = w;
                den += l * w;
            }
            den = warp_sum(den);
            if (lane == 0) {
                hd[2 * g] = mstar;
                hd[2 * g + 1] = den;
            }
        }
        __syncthreads();
        dbg_stamp_m(a, 10);
        int g = 0, k = threadIdx.x;
        while (k >= a.d) { k -= a.d; ++g; }
        const int dg = blockDim.x / a.d, dk = blockDim.x - dg * a.d;     // one block stride in (head, dim) steps
        for (; g < gc; g += dg, k += dk) {
            if (k >= a.d) { k -= a.d; ++g; if (g >= gc) break; }
            const float* w = ww + g * n_parts;
            float acc = 0.f;
            if (staged) {
                const float* src = a.big + (int64_t)g * n_parts * stride + k;
#pragma unroll 6
                for (int i = 0; i < n_parts; ++i) acc = fmaf(src[i * stride], w[i], acc);
            } else {
                const float* src = base + (int64_t)g * slots * stride + k;
                int i = 0;
                for (; i + 8 <= n_parts; i += 8) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (int64_t)(i + u) * stride);
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = fmaf(v[u], w[i + u], acc);
                }
                for (; i < n_parts; ++i) acc = fmaf(__ldcg(src + (int64_t)i * stride), w[i], acc);
            }
            const int h = h0 + g;
            if constexpr (P2P) {
                // fused split-KV: this rank's state of row (b, h) goes straight into slot [parity][rank] of EVERY rank's receive
                // buffer (NVLink stores, consecutive k = coalesced); layout of million_splitkv_push_merge
                const size_t off = ((size_t)p2p_par * p2p_world + p2p_rank) * p2p_slot + (size_t)(b * a.nh + h) * ostride;
                const int64_t ll_off = p2p_ll_offset(p2p_world, p2p_rows, a.d);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (r >= p2p_world) break;
                    float* dst = reinterpret_cast<float*>(p2p_peer[r] + ll_off) + 2 * off;       // 8-byte cells {value, sequence tag}
                    st_tagged(dst + 2 * k, acc, p2p_seq);
                    if (k == 0) { st_tagged(dst + 2 * a.d, hd[2 * g] * kLn2, p2p_seq); st_tagged(dst + 2 * (a.d + 1), hd[2 * g + 1], p2p_seq); }
                }
            } else if (a.partial_out) {
                a.partial_out[(int64_t)(b * a.nh + h) * ostride + k] = acc;
                if (k == 0) {
                    a.partial_out[(int64_t)(b * a.nh + h) * ostride + a.d] = hd[2 * g] * kLn2;   // natural-log units
                    a.partial_out[(int64_t)(b * a.nh + h) * ostride + a.d + 1] = hd[2 * g + 1];
                }
            } else {
                const float den = hd[2 * g + 1];
                reinterpret_cast<T*>(a.out)[(int64_t)(b * a.nh + h) * a.d + k] = io<T>::from_f(den > 0.f ? acc / den : 0.f);
            }
        }
        __syncthreads();
    }
    if constexpr (P2P) {
        // ---------------------------------------------------------------- wait for the peers' rows of THIS group, final merge
        // No flag and no fence: every 8-byte cell carries the sequence number of the call that wrote it, the threads below spin on
        // the cells they need (all in this rank's own receive area).  (1) the (m, l) pairs of G rows x world parts, (2) weights,
        // (3) the outputs.
        const int limit = ps->spin_limit > 0 ? ps->spin_limit : kP2PDefaultSpins;
        const float* recv = reinterpret_cast<const float*>(ps->peer[p2p_rank] + p2p_ll_offset(p2p_world, p2p_rows, a.d)) +
                            2 * ((size_t)p2p_par * p2p_world * p2p_slot);
        const int W = p2p_world, os = a.d + 2, row0 = b * a.nh + hk * G;
        int* sflag = reinterpret_cast<int*>(scr);
        float* sm_m = scr + 64;            // [G][W]   (G * W <= 8 * 64 fits the 2048-float blocks with room to spare)
        float* sm_l = scr + 2048;
        float* sm_w = scr + 4096;
        __syncthreads();                   // the group merge above is done with the scratch
        if (threadIdx.x == 0) *sflag = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < G * W; i += blockDim.x) {
            const int g = i / W, w = i - g * W;
            const float* p = recv + 2 * ((size_t)w * p2p_slot + (size_t)(row0 + g) * os);
            // (m, l) are adjacent cells: one 16-byte load brings both values and both tags
            uint4 c;
            int spins = 0;
            bool ok = true;
            for (;;) {
                asm volatile("ld.volatile.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(p + 2 * a.d) : "memory");
                if (c.y == p2p_seq && c.w == p2p_seq) break;
                __nanosleep(100);
                if (++spins > limit) { ok = false; break; }
            }
            if (!ok) { ps->err = 1; *sflag = 1; }
            sm_m[i] = __uint_as_float(c.x);
            sm_l[i] = __uint_as_float(c.z);
        }
        __syncthreads();
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            float mstar = -INFINITY, den = 0.f;
            for (int w = 0; w < W; ++w)
                if (sm_l[g * W + w] > 0.f) mstar = fmaxf(mstar, sm_m[g * W + w]);
            for (int w = 0; w < W; ++w) {
                const float l = sm_l[g * W + w];
                const float wt = l > 0.f ? __expf(sm_m[g * W + w] - mstar) : 0.f;
                sm_w[g * W + w] = wt;
                den += l * wt;
            }
            const float inv = den > 0.f ? 1.f / den : 0.f;
            for (int w = 0; w < W; ++w) sm_w[g * W + w] *= inv;
        }
        __syncthreads();
        const bool ml_timed_out = *sflag != 0;
        for (int i = threadIdx.x; i < G * a.d; i += blockDim.x) {
            const int g = i / a.d, k = i - g * a.d;
            const float* p = recv + 2 * ((size_t)(row0 + g) * os + k);
            // all `world` cells of this output are requested at once (independent loads: one L2 round trip, not `world` of them),
            // and polled again together until every tag is this call's
            unsigned bits[8], tag[8];
            bool ok = !ml_timed_out;
            int spins = 0;
            while (ok) {
                bool all = true;
#pragma unroll
                for (int w = 0; w < 8; ++w)
                    if (w < W) asm volatile("ld.volatile.global.v2.b32 {%0, %1}, [%2];" : "=r"(bits[w]), "=r"(tag[w]) : "l"(p + 2 * (size_t)w * p2p_slot) : "memory");
#pragma unroll
                for (int w = 0; w < 8; ++w)
                    if (w < W) all = all && tag[w] == p2p_seq;
                if (all) break;
                __nanosleep(100);
                if (++spins > limit) ok = false;
            }
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if (w < W) acc = fmaf(__uint_as_float(bits[w]), sm_w[g * W + w], acc);
            if (!ok) ps->err = 1;
            // a wait that gave up must not pass stale rows on as a result: poison them (and P2PState.err is set)
            reinterpret_cast<T*>(a.out)[(size_t)(row0 + g) * a.d + k] = io<T>::from_f(ok ? acc : __int_as_float(0x7fc00000));
        }
        // off the critical path: the last group of the launch advances the call counter (every group-last CTA of the NEXT launch
        // reads it after that launch's dependency on this one is resolved)
        __syncthreads();
        if (threadIdx.x == 0) {
            int t;
            asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(t) : "l"(&ps->ticket) : "memory");
            if (t == p2p_rows / G - 1) {              // groups = rows / (nh / nh_k)
                ps->ticket = 0;
                *reinterpret_cast<volatile unsigned*>(&ps->counter) = p2p_seq;
            }
        }
    }
}

template <typename T, bool P2P = false>
__device__ __forceinline__ void merge_group(const AttnArgs& a, int b, int hk, int n_parts, float* scr, float* big = nullptr, int big_floats = 0, unsigned long long* bar = nullptr, int dbg_piece = 0) {
    MergeArgs m;
    m.bar = bar;
    m.big = big; m.big_floats = big_floats;
    m.nh = a.nh; m.nh_k = a.nh_k; m.n_parts = a.n_parts; m.d = a.d; m.parts = a.parts; m.partial_out = a.partial_out; m.out = a.out;
#ifdef MILLION_DEBUG
    m.dbg_piece = dbg_piece; m.dbg_timing = a.dbg_timing;
#endif
    merge_group_impl<T, P2P>(m, b, hk, n_parts, scr);
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
