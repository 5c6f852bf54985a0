// attn_fast.cu — the B200 decode-attention kernel for the common shape d=128, M=64 (d_m=2), C=256,
// G = nh/nh_k in {1,2,4} (larger groups run as several 4-head sub-groups).
//
// What bounds this op on B200 is not HBM alone but the shared-memory gather rate (128 B/clk/SM): every coded
// token needs 64 K-LUT lookups and 64 V-codebook lookups.  tools/microbench*.cu measured on B200: LDS.64
// conflict-free 2.09 clk/warp-instr, random banks 4.9; LDS.32 1.13; and mma.sync *serialises* with LDS
// (1 mma + 8 LDS.64 costs the sum), so tensor-core reductions are out.  Hence:
//   * QK: thread-per-token; lane l visits the 16 code words of its token in an order rotated by
//     rot(l) = (l%16 + l/16) % 16, so at every step the 16 lanes of a half-warp hit 16 different
//     sub-space columns = 16 different bank pairs -> every LUT gather is conflict free, and no cross-lane
//     reduction is ever needed.  The gather address is formed by ONE PRMT: (code << 8) | column_offset.
//   * LUT entries hold all G heads of the KV group packed as fp16 (8 B for G=4) -> one gather serves 4 heads;
//     partial sums run in packed half2 over 16 entries, then flush to fp32.
//   * PV: a half-warp per token, lane owns 4 sub-spaces; V-codebook gathers are conflict free the same way;
//     p is broadcast through a 16-byte shared slot; products accumulate in half2 per 32-token tile, fp32 across.
//   * each warp streams its own 32-token code tiles with a private cp.async double buffer: no block barrier in
//     the main loop; online softmax state is per warp and merged once at the end.
//   * the fp16 window and the cross-CTA merge are fused (last split CTA / last-arrival ticket).
// Replaces Interface.cu:49-118 + Kernel.cuh:11-166, 1038-1209, 1211-1270 of the reference.
#include "attn_common.cuh"

namespace million {

namespace fast {

constexpr int kPairs = 8;                       // (QK warp, PV warp) pairs
constexpr int kWarps = 2 * kPairs;
constexpr int kThreads = kWarps * 32;
constexpr int kTile = 32;                       // tokens per tile
constexpr int kRowBytes = 64;                   // M = 64 one-byte codes
constexpr int kTileBytes = kTile * kRowBytes;   // 2 KB: one K tile per QK warp, one V tile per PV warp
constexpr int kPSlot = kTile * 8 + 32;          // 4 halves per token + the rescale factors of the tile
constexpr int kVtabBytes = 64 * 1024;
constexpr float kRescaleMargin = 6.f;           // log2 units: p <= 64 before a rescale is forced

template <int G> struct LutCfg;
template <> struct LutCfg<4> { static constexpr int bytes = 128 * 1024; };
template <> struct LutCfg<2> { static constexpr int bytes = 64 * 1024; };
template <> struct LutCfg<1> { static constexpr int bytes = 64 * 1024; };

// column of sub-space m inside a 64-entry table row (both K tables and the V table)
__host__ __device__ __forceinline__ constexpr int col_of(int m) {
    const int W = m >> 2, b = m & 3;
    return (b >> 1) * 32 + (b & 1) * 16 + W;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni MB_DONE;\n\t"
        "bra.uni MB_WAIT;\n\t"
        "MB_DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// plain shared-memory loads through the generic pointer of the dynamic smem block: the compiler is free to batch them
// (asm volatile loads would be kept in program order and serialise on the 30-cycle LDS latency)
__device__ __forceinline__ uint32_t lds32(const unsigned char* base, uint32_t off) { return *reinterpret_cast<const uint32_t*>(base + off); }
__device__ __forceinline__ uint2 lds64(const unsigned char* base, uint32_t off) { return *reinterpret_cast<const uint2*>(base + off); }
__device__ __forceinline__ uint4 lds128(const unsigned char* base, uint32_t off) { return *reinterpret_cast<const uint4*>(base + off); }
// acc_lo += fp16(packed.lo), acc_hi += fp16(packed.hi) in fp32: Blackwell FHADD (PTX add.rn.f32.f16), one op each
__device__ __forceinline__ void fhadd2(float& acc_lo, float& acc_hi, uint32_t packed) {
    const unsigned short lo = (unsigned short)(packed & 0xffffu), hi = (unsigned short)(packed >> 16);
    asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(acc_lo) : "h"(lo));
    asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(acc_hi) : "h"(hi));
}
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

}  // namespace fast

// ------------------------------------------------------------------------------------------------
// Codebook preparation (once per codebook): fp16 tables in the column order the kernel gathers with.
//   prepared[0      .. 64 KB) : kT[c][col(m)] = half2(Kcent[m][c][0], Kcent[m][c][1])
//   prepared[64 KB .. 128 KB) : vT[c][col(m)] = half2(Vcent[m][c][0], Vcent[m][c][1])
template <typename T>
__global__ void codebook_prepare_kernel(const T* __restrict__ kcent, const T* __restrict__ vcent, uint32_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 2 * 64 * 256
    if (idx >= 2 * 64 * 256) return;
    const int which = idx / (64 * 256), rem = idx % (64 * 256);
    const int m = rem / 256, c = rem % 256;
    const T* src = (which ? vcent : kcent) + ((int64_t)m * 256 + c) * 2;
    const __half2 v = __floats2half2_rn(io<T>::to_f(src[0]), io<T>::to_f(src[1]));
    out[which * 64 * 256 + c * 64 + fast::col_of(m)] = fast::as_u32(v);
}

int launch_codebook_prepare(const void* kcent, const void* vcent, int io_dtype, void* out, cudaStream_t stream) {
    dim3 grid(2 * 64 * 256 / 256), block(256);
    if (io_dtype == MILLION_F16) codebook_prepare_kernel<__half><<<grid, block, 0, stream>>>((const __half*)kcent, (const __half*)vcent, (uint32_t*)out);
    else if (io_dtype == MILLION_BF16) codebook_prepare_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)kcent, (const __nv_bfloat16*)vcent, (uint32_t*)out);
    else codebook_prepare_kernel<float><<<grid, block, 0, stream>>>((const float*)kcent, (const float*)vcent, (uint32_t*)out);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

// ------------------------------------------------------------------------------------------------
namespace fast {

// Table gathers with an absolute shared-window address: the PRMT result (code << 8 | column offset) is the register part,
// the table base is an immediate, so a gather is exactly PRMT + LDS.  Not volatile: the tables are read-only in the main
// loop and the compiler may schedule these loads freely.  kSmemBase is checked at kernel entry.
constexpr uint32_t kSmemBase = 0x400;   // dynamic shared memory starts after the 1 KB the driver reserves per CTA (no static smem)
template <uint32_t IMM>
__device__ __forceinline__ uint2 gather64(uint32_t r) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(r), "n"(IMM));
    return v;
}
template <uint32_t IMM>
__device__ __forceinline__ uint32_t gather32(uint32_t r) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(r), "n"(IMM));
    return v;
}

}  // namespace fast

// VL = 0: value codes row-major (tokens x 64 bytes); VL = 1: transposed per sub-space (paged pool or (M, ld) rows)
//
// 16 warps = 8 pairs.  In pair p the QK warp (warp p) scores the tokens of tile p, p+8, ... and hands the probabilities to
// the PV warp (warp 8+p) through a 288-byte shared slot guarded by two mbarriers (full / empty), so the LUT-gather phase
// of one tile overlaps the value-gather phase of the previous one and four warps share every scheduler.
template <typename T, int G, int VL>
__global__ void __launch_bounds__(fast::kThreads, 1) attn_fast_kernel(const AttnArgs a, const uint32_t* __restrict__ prepared, const int gsub) {
    using namespace fast;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t kLutOff = 0, kVtabOff = LutCfg<G>::bytes, kStageOff = kVtabOff + kVtabBytes;   // stage = K tiles then V tiles
    constexpr uint32_t kPbufOff = kStageOff + 2 * kPairs * kTileBytes, kBarOff = kPbufOff + kPairs * kPSlot, kMiscOff = kBarOff + kPairs * 16;
    unsigned char* lut_p = smem + kLutOff;
    unsigned char* stage_p = smem + kStageOff;
    unsigned char* pbuf_p = smem + kPbufOff;
    int* flag = reinterpret_cast<int*>(smem + kMiscOff);
    // after the main loop the stage buffers and p slots are dead: reuse them for the cross-warp combine and the merge
    float* xch = reinterpret_cast<float*>(stage_p);                      // 2 * kPairs entries of kEntry floats, then kMergeScratch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_qk = warp < kPairs;
    const int pair = warp & (kPairs - 1);
    if (smem_u32(smem) != kSmemBase) {
        if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
            printf("million_b200: dynamic shared memory starts at 0x%x, expected 0x%x\n", smem_u32(smem), kSmemBase);
        __trap();
    }
    dbg_stamp(a, 0);
    const int split = blockIdx.x;
    // blockIdx.y enumerates (kv head, 4-head sub-group) when the GQA group is larger than 4
    const int hk = blockIdx.y / gsub, sub = blockIdx.y % gsub, b = blockIdx.z;
    const int Gfull = a.nh / a.nh_k;
    const int h0 = hk * Gfull + sub * G;                                // first query head of this CTA
    const int hb = b * a.nh_k + hk;
    int t0, t1;
    split_range(a, split, t0, t1);
    const bool has_codes = t1 > t0;
    const uint32_t bar_full = smem_u32(smem + kBarOff + pair * 16), bar_empty = bar_full + 8;
    if (tid < kPairs) {
        mbar_init(smem_u32(smem + kBarOff + tid * 16), 1);
        mbar_init(smem_u32(smem + kBarOff + tid * 16 + 8), 1);
    }

    // ---------------------------------------------------------------- prologue: V table + K LUT
    if (has_codes) {
        // The K gather table kT (64 KB, L2 resident) streams through the stage area in four 16 KB chunks (cp.async double
        // buffer); the V table (already in gather order) is copied straight to its place in the background.
        const uint32_t stage0 = smem_u32(stage_p), vtab_s = smem_u32(smem + kVtabOff);
        auto load_chunk = [&](int ch) {
            const char* src = reinterpret_cast<const char*>(prepared) + ch * 16384;
            const uint32_t dst = stage0 + (ch & 1) * 16384;
#pragma unroll
            for (int i = 0; i < 16384 / 16 / kThreads; ++i) cp_async16(dst + (tid + i * kThreads) * 16, src + (tid + i * kThreads) * 16, 16);
            cp_async_commit();
        };
        load_chunk(0);
        {
            const uint4* vsrc = reinterpret_cast<const uint4*>(prepared + 64 * 256);
            for (int i = tid; i < kVtabBytes / 16; i += kThreads) cp_async16(vtab_s + i * 16, vsrc + i, 16);
            cp_async_commit();
        }
        load_chunk(1);
        // K LUT: thread owns column `col` (= one sub-space) and walks the codes.  LUT[c][col] = <q_h[m], Kcent[m][c]> for the
        // G heads of the group: fp32 products of fp16 operands, rounded once to fp16 (G=1 keeps fp32 entries).
        const int col = tid & 63;
        const int bb = ((col >> 5) << 1) | ((col >> 4) & 1), W = col & 15, m = 4 * W + bb;
        float q0[G], q1[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T* q = reinterpret_cast<const T*>(a.q) + (int64_t)(b * a.nh + h0 + g) * 128;
            q0[g] = io<T>::to_f(q[2 * m]);
            q1[g] = io<T>::to_f(q[2 * m + 1]);
        }
        for (int ch = 0; ch < 4; ++ch) {
            if (ch == 0) cp_async_wait<2>();        // pending: [chunk0, vtab, chunk1] -> chunk0 landed
            else if (ch < 3) cp_async_wait<1>();    // chunk ch landed (and the V table)
            else cp_async_wait<0>();
            __syncthreads();
            const unsigned char* src = stage_p + (ch & 1) * 16384;
#pragma unroll 4
            for (int cl = tid >> 6; cl < 64; cl += kThreads / 64) {
                const int c = ch * 64 + cl;
                const __half2 cv = as_h2(*reinterpret_cast<const uint32_t*>(src + (cl * 64 + col) * 4));
                const float c0 = __low2float(cv), c1 = __high2float(cv);
                if constexpr (G == 4) {
                    const __half2 e01 = __floats2half2_rn(fmaf(c1, q1[0], c0 * q0[0]), fmaf(c1, q1[1], c0 * q0[1]));
                    const __half2 e23 = __floats2half2_rn(fmaf(c1, q1[2], c0 * q0[2]), fmaf(c1, q1[3], c0 * q0[3]));
                    // address(m, c) = (b>>1)*64K + c*256 + ((b&1)*16 + W)*8
                    *reinterpret_cast<uint2*>(lut_p + (bb >> 1) * 65536 + c * 256 + ((bb & 1) * 16 + W) * 8) = make_uint2(as_u32(e01), as_u32(e23));
                } else if constexpr (G == 2) {
                    const __half2 e01 = __floats2half2_rn(fmaf(c1, q1[0], c0 * q0[0]), fmaf(c1, q1[1], c0 * q0[1]));
                    *reinterpret_cast<uint32_t*>(lut_p + c * 256 + col * 4) = as_u32(e01);
                } else {
                    *reinterpret_cast<float*>(lut_p + c * 256 + col * 4) = fmaf(c1, q1[0], c0 * q0[0]);
                }
            }
            __syncthreads();
            if (ch + 2 < 4) load_chunk(ch + 2);
        }
    }
    __syncthreads();
    dbg_stamp(a, 1);

    // ---------------------------------------------------------------- main loop
    const int lq = lane & 15, hw = lane >> 4;
    const int n_tiles = has_codes ? (t1 - t0 + kTile - 1) / kTile : 0;
    unsigned char* pslot = pbuf_p + pair * kPSlot;
    // QK-warp state: running max (uniform) and per-lane partial denominators.  PV-warp state: output accumulators,
    // slot s of lane (hw, lq) = sub-space 4*lq + ((s + hw) & 3), 2 dims.
    float run_m[G], run_l[G], out_o[4][G][2];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        run_m[g] = -INFINITY; run_l[g] = 0.f;
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) { out_o[sl][g][0] = 0.f; out_o[sl][g][1] = 0.f; }
    }

    if (has_codes && is_qk) {
        // ================================================================ QK warp
        const int rot = (lq + hw) & 15;
        unsigned char* ksp = stage_p + pair * kTileBytes;
        const uint32_t ks_s = smem_u32(ksp);
        const uint8_t* kbase = a.k_codes + hb * a.k_head_stride;
        uint32_t koff[16];   // byte0 = column offset for even b, byte1 = for odd b, bytes 2,3 = 0
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            const int Wl = (w + rot) & 15;
            if constexpr (G == 4) koff[w] = (uint32_t)(Wl * 8) | ((uint32_t)(Wl * 8 + 128) << 8);
            else koff[w] = (uint32_t)(Wl * 4) | ((uint32_t)(Wl * 4 + 64) << 8);   // + (bb>>1)*128 comes from the immediate
        }
        auto issue_k = [&](int tile) {
            const int tok0 = t0 + tile * kTile;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int chunk = lane + i * 32;            // 0..127: 32 tokens * 4 chunks of 16 B
                const int tok = tok0 + (chunk >> 2);
                const int ok = (tile < n_tiles && tok < t1) ? 16 : 0;
                cp_async16(ks_s + chunk * 16, kbase + (int64_t)(ok ? tok : t0) * kRowBytes + (chunk & 3) * 16, ok);
            }
            cp_async_commit();
        };
        issue_k(pair);
        int n = 0;
        for (int tile = pair; tile < n_tiles; tile += kPairs, ++n) {
            cp_async_wait<0>();
            __syncwarp();
            const int tok = t0 + tile * kTile + lane;
            const bool valid = tok < t1;
            float s[G];
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] = 0.f;
            uint32_t words[16];
#pragma unroll
            for (int w = 0; w < 16; ++w) words[w] = lds32(ksp, lane * kRowBytes + (((w + rot) & 15) << 2));   // word (w + rot) % 16 of my row
#pragma unroll
            for (int w = 0; w < 16; ++w) {
#pragma unroll
                for (int bq = 0; bq < 4; ++bq) {
                    constexpr uint32_t selc[4] = {0x6604u, 0x6615u, 0x6624u, 0x6635u};
                    if constexpr (G == 4) {
                        const uint32_t ad = __byte_perm(words[w], koff[w], selc[bq]);     // (code << 8) | column offset
                        const uint2 e = (bq >> 1) ? gather64<kSmemBase + kLutOff + 65536>(ad) : gather64<kSmemBase + kLutOff>(ad);
                        fhadd2(s[0], s[1], e.x);
                        fhadd2(s[2], s[3], e.y);
                    } else {
                        // the two half-warps use different bytes in the same step so that all 32 lanes hit 32 banks
                        const int b1 = (bq + 1) & 3;
                        const uint32_t ad = __byte_perm(words[w], koff[w], hw ? selc[b1] : selc[bq]) + (uint32_t)(((hw ? b1 : bq) >> 1) * 128);
                        const uint32_t e = gather32<kSmemBase + kLutOff>(ad);
                        if constexpr (G == 2) fhadd2(s[0], s[1], e);
                        else s[0] += __uint_as_float(e);
                    }
                }
            }
            __syncwarp();                                   // every lane has read its K row
            issue_k(tile + kPairs);                         // K(n+1) streams in while the tail of this tile is processed
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] = valid ? s[g] * a.scale_log2 : -INFINITY;

            // online softmax with a lazy max: rescale only when a score exceeds the running max by 2^kRescaleMargin
            float alpha[4] = {1.f, 1.f, 1.f, 1.f};
            bool need = false;
#pragma unroll
            for (int g = 0; g < G; ++g) need = need || (s[g] > run_m[g] + kRescaleMargin);
            if (__any_sync(0xffffffffu, need)) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float nm = fmaxf(run_m[g], warp_max(s[g]));
                    if (nm > run_m[g]) {
                        alpha[g] = exp2_safe(run_m[g], nm);
                        run_l[g] *= alpha[g];
                        run_m[g] = nm;
                    }
                }
            }
            float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int g = 0; g < G; ++g) {
                p[g] = exp2_safe(s[g], run_m[g]);
                run_l[g] += p[g];
            }
            if (n > 0) mbar_wait(bar_empty, (n - 1) & 1);   // the PV warp has consumed the previous tile's slot
            *reinterpret_cast<uint2*>(pslot + lane * 8) = make_uint2(as_u32(__floats2half2_rn(p[0], p[1])), as_u32(__floats2half2_rn(p[2], p[3])));
            if (lane == 0) *reinterpret_cast<float4*>(pslot + kTile * 8) = make_float4(alpha[0], alpha[1], alpha[2], alpha[3]);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full);
        }
        cp_async_wait<0>();
    } else if (has_codes) {
        // ================================================================ PV warp
        unsigned char* vsp = stage_p + (kPairs + pair) * kTileBytes;
        const uint32_t vs_s = smem_u32(vsp);
        const uint8_t* vbase = a.v_codes + (a.v_layout == MILLION_V_PAGED ? 0 : hb * a.v_head_stride);
        // slot s -> byte bb = (s + hw) & 3 of the lane's V code word; column offset col_of(4*lq + bb) * 4
        uint32_t voff01, voff23, vsel[4];
        {
            uint32_t o[4];
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
                const int bb = (sl + hw) & 3;
                o[sl] = (uint32_t)(col_of(4 * lq + bb) * 4);
                vsel[sl] = (uint32_t)(4 + (sl & 1)) | ((uint32_t)bb << 4) | (6u << 8) | (6u << 12);   // [off, code byte bb, 0, 0]
            }
            voff01 = o[0] | (o[1] << 8);
            voff23 = o[2] | (o[3] << 8);
        }
        auto issue_v = [&](int tile) {
            const int tok0 = t0 + tile * kTile;
            const bool in_range = tile < n_tiles;
            if constexpr (VL == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int chunk = lane + i * 32;
                    const int tok = tok0 + (chunk >> 2);
                    const int ok = (in_range && tok < t1) ? 16 : 0;
                    cp_async16(vs_s + chunk * 16, vbase + (int64_t)(ok ? tok : t0) * kRowBytes + (chunk & 3) * 16, ok);
                }
            } else {
                // transposed value codes: the tile is 64 sub-space rows of 32 tokens; row m is staged at row pi(m) = m/4 + 16*(m%4)
                // (32 bytes each) so that the word reads of the PV phase below are bank-conflict free
                const uint8_t* src0 = vbase;
                int64_t row_stride = a.v_ld;
                if (in_range) {
                    if (a.v_layout == MILLION_V_PAGED) {
                        const int64_t page = __ldg(a.v_page_ids + (int64_t)hb * a.n_pages + tok0 / a.page_size);
                        src0 = a.v_codes + page * 64 * a.page_size + (tok0 % a.page_size);
                        row_stride = a.page_size;
                    } else {
                        src0 = vbase + tok0;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int chunk = lane + i * 32;            // 0..127: 64 rows * 2 chunks of 16 tokens
                    const int m = chunk >> 1, hc = chunk & 1;
                    int ok = in_range ? (t1 - (tok0 + hc * 16)) : 0;
                    ok = ok < 0 ? 0 : (ok > 16 ? 16 : ok);
                    const uint32_t dst = vs_s + (uint32_t)((((m >> 2) + 16 * (m & 3)) * 32) + hc * 16);
                    cp_async16(dst, (ok ? src0 : vbase) + (ok ? (int64_t)m * row_stride + hc * 16 : 0), ok);
                }
            }
            cp_async_commit();
        };
        issue_v(pair);
        __half2 acc[4][G];
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
#pragma unroll
            for (int g = 0; g < G; ++g) acc[sl][g] = __float2half2_rn(0.f);
        auto flush = [&]() {
#pragma unroll
            for (int sl = 0; sl < 4; ++sl)
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float2 f = __half22float2(acc[sl][g]);
                    out_o[sl][g][0] += f.x; out_o[sl][g][1] += f.y;
                    acc[sl][g] = __float2half2_rn(0.f);
                }
        };
        int since_flush = 0, n = 0;
        for (int tile = pair; tile < n_tiles; tile += kPairs, ++n) {
            cp_async_wait<0>();                             // V(n) landed
            mbar_wait(bar_full, n & 1);                     // p(n) and the rescale factors are in the slot
            __syncwarp();
            {
                const float4 al = *reinterpret_cast<const float4*>(pslot + kTile * 8);
                const float alv[4] = {al.x, al.y, al.z, al.w};
                bool resc = false;
#pragma unroll
                for (int g = 0; g < G; ++g) resc = resc || (alv[g] != 1.f);
                if (resc) {                                 // warp-uniform: the running max moved
                    flush();
                    since_flush = 0;
#pragma unroll
                    for (int g = 0; g < G; ++g)
#pragma unroll
                        for (int sl = 0; sl < 4; ++sl) { out_o[sl][g][0] *= alv[g]; out_o[sl][g][1] *= alv[g]; }
                }
            }
            if constexpr (VL == 0) {
                // row-major codes: a half-warp per token, the lane's word holds its 4 sub-spaces of that token
#pragma unroll 4
                for (int jp = 0; jp < kTile / 2; ++jp) {
                    const int j = 2 * jp + hw;
                    const uint32_t word = lds32(vsp, j * kRowBytes + lq * 4);
                    const uint2 pk = lds64(pslot, j * 8);
                    const __half2 p01 = as_h2(pk.x), p23 = as_h2(pk.y);
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) {
                        const uint32_t ad = __byte_perm(word, sl < 2 ? voff01 : voff23, vsel[sl]);
                        const __half2 v = as_h2(gather32<kSmemBase + kVtabOff>(ad));
                        acc[sl][0] = __hfma2(__low2half2(p01), v, acc[sl][0]);
                        if constexpr (G >= 2) acc[sl][1] = __hfma2(__high2half2(p01), v, acc[sl][1]);
                        if constexpr (G == 4) {
                            acc[sl][2] = __hfma2(__low2half2(p23), v, acc[sl][2]);
                            acc[sl][3] = __hfma2(__high2half2(p23), v, acc[sl][3]);
                        }
                    }
                }
            } else {
                // transposed codes: a word holds 4 tokens of one sub-space.  Half-warp hw takes tokens 16*hw .. 16*hw+15 in
                // 4 groups of 4; the group order is rotated by lq/4 so the 32 word reads of one instruction hit 32 banks
                // (staged row pi(m) lies in bank quad 2*(lq%4) + hw, the word inside the chunk is (u + lq/4) % 4).
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int tgl = (u + (lq >> 2)) & 3;                 // token group inside my half
                    const int tg = tgl + 4 * hw;
                    uint32_t wv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) wv[k] = lds32(vsp, (lq + 16 * ((k + hw) & 3)) * 32 + hw * 16 + tgl * 4);
                    const uint4 pa = lds128(pslot, tg * 32), pb = lds128(pslot, tg * 32 + 16);
                    const uint32_t pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t sel = (uint32_t)(4 + (k & 1)) | ((uint32_t)i << 4) | 0x6600u;
                            const uint32_t ad = __byte_perm(wv[k], k < 2 ? voff01 : voff23, sel);
                            const __half2 v = as_h2(gather32<kSmemBase + kVtabOff>(ad));
                            const __half2 p01 = as_h2(pp[2 * i]), p23 = as_h2(pp[2 * i + 1]);
                            acc[k][0] = __hfma2(__low2half2(p01), v, acc[k][0]);
                            if constexpr (G >= 2) acc[k][1] = __hfma2(__high2half2(p01), v, acc[k][1]);
                            if constexpr (G == 4) {
                                acc[k][2] = __hfma2(__low2half2(p23), v, acc[k][2]);
                                acc[k][3] = __hfma2(__high2half2(p23), v, acc[k][3]);
                            }
                        }
                    }
                }
            }
            __syncwarp();                                   // every lane is done with the V tile and the p slot
            if (lane == 0) mbar_arrive(bar_empty);
            issue_v(tile + kPairs);                         // V(n+1) streams in while the QK warp scores the next tile
            if (++since_flush == 2) { flush(); since_flush = 0; }   // packed-half partial sums live for at most 2 tiles
        }
        flush();
        cp_async_wait<0>();
    }

    // ---------------------------------------------------------------- my share of the fp16 window (exact attention)
    // The r recent tokens are dealt out to the splits of the group (r/S tokens each); inside the CTA the QK warps take one
    // token at a time.  Lane owns dims 4*lane .. 4*lane+3.
    float wm[G], wl[G], wo[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) { wm[g] = -INFINITY; wl[g] = 0.f; wo[g][0] = wo[g][1] = wo[g][2] = wo[g][3] = 0.f; }
    if (is_qk) {
        const int w0 = (int)((long long)a.r * split / a.n_splits), w1 = (int)((long long)a.r * (split + 1) / a.n_splits);
        if (w0 + pair < w1) {
            float qv[G][4];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint2 qr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.q) + (int64_t)(b * a.nh + h0 + g) * 128 + 4 * lane));
                const float2 q01 = io<T>::to_f2(qr.x), q23 = io<T>::to_f2(qr.y);
                qv[g][0] = q01.x * a.scale_log2; qv[g][1] = q01.y * a.scale_log2; qv[g][2] = q23.x * a.scale_log2; qv[g][3] = q23.y * a.scale_log2;
            }
            for (int t = w0 + pair; t < w1; t += kPairs) {
                const int64_t row = ((int64_t)hb * a.res_len + t) * 128 + 4 * lane;
                const uint2 kr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.k_res) + row));
                const uint2 vr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.v_res) + row));
                const float2 k01 = io<T>::to_f2(kr.x), k23 = io<T>::to_f2(kr.y);
                const float2 v01 = io<T>::to_f2(vr.x), v23 = io<T>::to_f2(vr.y);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float sg = warp_sum(fmaf(qv[g][3], k23.y, fmaf(qv[g][2], k23.x, fmaf(qv[g][1], k01.y, qv[g][0] * k01.x))));
                    const float nm = fmaxf(wm[g], sg);
                    const float alpha = exp2_safe(wm[g], nm), pw = exp2f(sg - nm);
                    wl[g] = wl[g] * alpha + pw;
                    wo[g][0] = fmaf(pw, v01.x, wo[g][0] * alpha); wo[g][1] = fmaf(pw, v01.y, wo[g][1] * alpha);
                    wo[g][2] = fmaf(pw, v23.x, wo[g][2] * alpha); wo[g][3] = fmaf(pw, v23.y, wo[g][3] * alpha);
                    wm[g] = nm;
                }
            }
        }
    }

    __syncthreads();   // every warp is done with its stage buffers and p slots (aliased below)
    dbg_stamp(a, 2);
    // ---------------------------------------------------------------- combine the warps of this CTA -> one partial state
    // 2 * kPairs entries of [G*128 o | G m | G l]: entry p = coded tokens of pair p (o from its PV warp, m and l from its QK
    // warp), entry kPairs + p = the window tokens of QK warp p.
    // Coded layout: slot sl of lane (hw, lq) holds sub-space 4*lq + ((sl + hw) & 3); hw=1 is folded into hw=0 first.
    constexpr int kEntry = (G * 130 + 3) & ~3;   // 16-byte aligned entries
    {
        float* wx = xch + pair * kEntry;
        if (!is_qk) {
#pragma unroll
            for (int sl = 0; sl < 4; ++sl)
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        // partner (hw=1) slot (sl - 1) & 3 holds the same sub-space as my (hw=0) slot sl
                        const float theirs = __shfl_xor_sync(0xffffffffu, out_o[(sl + 3) & 3][g][k], 16);
                        if (hw == 0) wx[g * 128 + 2 * (4 * lq + sl) + k] = out_o[sl][g][k] + theirs;
                    }
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) run_l[g] = warp_sum(run_l[g]);
            float* ww = xch + (kPairs + pair) * kEntry;
#pragma unroll
            for (int g = 0; g < G; ++g)
                *reinterpret_cast<float4*>(ww + g * 128 + 4 * lane) = make_float4(wo[g][0], wo[g][1], wo[g][2], wo[g][3]);
            if (lane == 0) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    wx[G * 128 + g] = run_m[g]; wx[G * 128 + G + g] = run_l[g];
                    ww[G * 128 + g] = wm[g];    ww[G * 128 + G + g] = wl[g];
                }
            }
        }
    }
    __syncthreads();
    {
        // thread -> (dim = tid & 127, head = tid >> 7)
        const int dim = tid & 127, g = tid >> 7;
        if (g < G) {
            float mstar = -INFINITY;
#pragma unroll
            for (int e = 0; e < 2 * kPairs; ++e) mstar = fmaxf(mstar, xch[e * kEntry + G * 128 + g]);
            float o = 0.f, l = 0.f;
#pragma unroll
            for (int e = 0; e < 2 * kPairs; ++e) {
                const float* ex = xch + e * kEntry;
                const float sc = exp2_safe(ex[G * 128 + g], mstar);
                o = fmaf(ex[g * 128 + dim], sc, o);
                l = fmaf(ex[G * 128 + G + g], sc, l);
            }
            float* part = a.parts + ((int64_t)(b * a.nh + h0 + g) * a.n_parts + split) * 130;
            part[dim] = o;
            if (dim == 0) { part[128] = mstar; part[129] = l; }
        }
    }
    dbg_stamp(a, 3);
    dbg_stamp(a, 4);
    // ---------------------------------------------------------------- last CTA of the (b, hk) group merges
    const bool last = last_cta_of_group(a.counters, hb, a.n_splits * gsub, flag);
    dbg_stamp(a, 5);
    if (last) merge_group<T>(a, b, hk, xch);
    dbg_stamp(a, 6);
}

// ------------------------------------------------------------------------------------------------ launcher
template <typename T, int G, int VL>
static int launch_fast_t(const AttnArgs& a, const uint32_t* prepared, int gsub, cudaStream_t stream) {
    using namespace fast;
    const size_t smem = LutCfg<G>::bytes + kVtabBytes + 2 * kPairs * kTileBytes + kPairs * kPSlot + kPairs * 16 + 64;
    static_assert(2 * kPairs * kTileBytes + kPairs * kPSlot >= 2 * kPairs * 4 * 130 * sizeof(float), "stage + p area too small for the combine");
    static_assert(2 * kPairs * kTileBytes >= 2 * 16384, "stage area too small for the LUT-build chunks");
    static_assert(2 * kPairs * kTileBytes + kPairs * kPSlot >= kMergeScratch * sizeof(float), "stage area too small for the merge scratch");
    static bool configured = false;
    if (!configured) {
        MILLION_CUDA_OK(cudaFuncSetAttribute(attn_fast_kernel<T, G, VL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(a.n_splits, a.nh_k * gsub, a.bs), block(kThreads);
    attn_fast_kernel<T, G, VL><<<grid, block, smem, stream>>>(a, prepared, gsub);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

int launch_attn_fast(const AttnArgs& a_in, int io_dtype, const void* prepared, cudaStream_t stream, bool probe_only) {
    AttnArgs a = a_in;
    a.n_parts = a.n_splits;   // the window is dealt out to the splits, no extra part
    const int Gfull = a.nh / a.nh_k;
    if (a.d != 128 || a.M != 64 || a.C != 256) MILLION_UNSUPPORTED("fast decode attention needs d=128, M=64, C=256");
    if (!(Gfull == 1 || Gfull == 2 || Gfull % 4 == 0)) MILLION_UNSUPPORTED("fast decode attention needs nh/nh_k in {1,2,4k}");
    if (a.nk > 0 && a.v_layout == MILLION_V_PAGED && (a.page_size % 32 != 0 || ((uintptr_t)a.v_codes & 15)))
        MILLION_UNSUPPORTED("fast decode attention: paged V needs page_size %% 32 == 0 and an aligned pool");
    if (a.nk > 0 && a.v_layout == MILLION_V_TRANSPOSED && ((a.v_ld & 15) || (a.v_head_stride & 15) || ((uintptr_t)a.v_codes & 15)))
        MILLION_UNSUPPORTED("fast decode attention: transposed V needs 16-byte aligned rows");
    if (!prepared) MILLION_UNSUPPORTED("fast decode attention needs a prepared codebook (million_pq_codebook_prepare)");
    if (a.nk > 0 && (((uintptr_t)a.k_codes | (uintptr_t)a.k_head_stride) & 15))
        MILLION_UNSUPPORTED("fast decode attention needs 16-byte aligned code caches");
    if (a.nk > 0 && a.v_layout == MILLION_V_ROWMAJOR && (((uintptr_t)a.v_codes | (uintptr_t)a.v_head_stride) & 15))
        MILLION_UNSUPPORTED("fast decode attention needs 16-byte aligned code caches");
    if (a.r > 0 && (((uintptr_t)a.k_res | (uintptr_t)a.v_res) & 7)) MILLION_UNSUPPORTED("fast decode attention needs an 8-byte aligned window");
    if (probe_only) return MILLION_OK;
    const int G = Gfull >= 4 ? 4 : Gfull, gsub = Gfull >= 4 ? Gfull / 4 : 1;
    const uint32_t* prep = reinterpret_cast<const uint32_t*>(prepared);
#define MILLION_FAST_CASE(TT, GG) \
    return a.v_layout == MILLION_V_ROWMAJOR ? launch_fast_t<TT, GG, 0>(a, prep, gsub, stream) : launch_fast_t<TT, GG, 1>(a, prep, gsub, stream)
    if (io_dtype == MILLION_F16) {
        if (G == 4) MILLION_FAST_CASE(__half, 4);
        if (G == 2) MILLION_FAST_CASE(__half, 2);
        MILLION_FAST_CASE(__half, 1);
    } else {
        if (G == 4) MILLION_FAST_CASE(__nv_bfloat16, 4);
        if (G == 2) MILLION_FAST_CASE(__nv_bfloat16, 2);
        MILLION_FAST_CASE(__nv_bfloat16, 1);
    }
#undef MILLION_FAST_CASE
}

}  // namespace million
