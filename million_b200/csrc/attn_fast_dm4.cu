// attn_fast_dm4.cu — the fast decode-attention kernel for M=32 sub-spaces of d_m=4 dims ("2-bit" MILLION: d=128, C=256).
//
// Same structure as attn_fast.cu (one CTA per (batch, kv-head, split), LUT in shared memory, thread-per-token QK with
// conflict-free gathers, half-warp-per-token PV, fused window + merge); what changes is the geometry:
//   * a token's code row is 32 bytes = 8 words; lane l walks them rotated by (l%8 + l/8) % 8, and the two 8-lane groups of
//     a half-warp use bytes of different parity in the same step, so the 16 lanes of an LDS.64 wavefront still hit 16
//     different bank pairs (column of sub-space m = 4W+b is (b/2)*16 + (b%2)*8 + W, 32 columns of 8 bytes = 256-byte rows);
//   * K-LUT entries always hold 4 head slots (8 bytes; unused heads are zero), one 64 KB table;
//   * V-codebook entries are 4 halves (8 bytes): a lane owns 2 sub-spaces of its token, 2 LDS.64 gathers per token-step;
//   * value codes must be row-major (the transposed/paged layouts run the generic kernel for this shape).
// Replaces Interface.cu:49-118 + Kernel.cuh:11-166, 1038-1209, 1211-1270 for the M=32 instantiations (setup.py:10-15).
#include "attn_fast_helpers.cuh"

#ifndef MILLION_DM4_PV_UNROLL
#define MILLION_DM4_PV_UNROLL 16      // all 16 PV steps per tile: 86.3 us per launch (8B shapes, 32K, batch 8) vs 88.8 (8), 93.0 (4); flushing every tile: no gain here
#endif
#ifndef MILLION_DM4_FLUSH_TILES
#define MILLION_DM4_FLUSH_TILES 2     // tiles between two flushes of the packed-half PV sums
#endif

namespace million {

constexpr int kDm4PvUnroll = MILLION_DM4_PV_UNROLL;

namespace dm4 {
constexpr int kRow = 32;                               // bytes per token row
constexpr int kTileBytes = fast::kTile * kRow;         // 1 KB
constexpr int kLutBytes = 64 * 1024;
__host__ __device__ __forceinline__ constexpr int col_of(int m) {
    const int W = m >> 2, b = m & 3;
    return (b >> 1) * 16 + (b & 1) * 8 + W;
}
}  // namespace dm4

// prepared[0 .. 64 KB): kT[c][col(m)] = 4 halves of Kcent[m][c][0..3]; prepared[64 KB .. 128 KB): vT likewise
template <typename T>
__global__ void codebook_prepare_dm4_kernel(const T* __restrict__ kcent, const T* __restrict__ vcent, uint2* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 2 * 32 * 256
    if (idx >= 2 * 32 * 256) return;
    const int which = idx / (32 * 256), rem = idx % (32 * 256);
    const int m = rem / 256, c = rem % 256;
    const T* src = (which ? vcent : kcent) + ((int64_t)m * 256 + c) * 4;
    const __half2 lo = __floats2half2_rn(io<T>::to_f(src[0]), io<T>::to_f(src[1]));
    const __half2 hi = __floats2half2_rn(io<T>::to_f(src[2]), io<T>::to_f(src[3]));
    out[which * 32 * 256 + c * 32 + dm4::col_of(m)] = make_uint2(fast::as_u32(lo), fast::as_u32(hi));
}

int launch_codebook_prepare_dm4(const void* kcent, const void* vcent, int io_dtype, void* out, cudaStream_t stream) {
    dim3 grid(2 * 32 * 256 / 256), block(256);
    if (io_dtype == MILLION_F16) codebook_prepare_dm4_kernel<__half><<<grid, block, 0, stream>>>((const __half*)kcent, (const __half*)vcent, (uint2*)out);
    else if (io_dtype == MILLION_BF16) codebook_prepare_dm4_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)kcent, (const __nv_bfloat16*)vcent, (uint2*)out);
    else codebook_prepare_dm4_kernel<float><<<grid, block, 0, stream>>>((const float*)kcent, (const float*)vcent, (uint2*)out);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

// One segment = the coded tokens [t0, t1) of group (b, hk) (+ this CTA's share of the window), part `split` of `np`.
// VL = 0: value codes row-major (tokens x 32 bytes); VL = 1: transposed per sub-space ((M, ld) rows or the page pool)
template <typename T, int G, int VL, int OUT>
__device__ __forceinline__ void attn_dm4_segment(const AttnArgs& a, const uint32_t* __restrict__ prepared, const int gsub, const int split,
                                                 const int hk, const int sub, const int b, const int t0, const int t1, const int np) {
    using namespace fast;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t kLutOff = 0, kVtabOff = dm4::kLutBytes, kStageOff = kVtabOff + kVtabBytes;
    constexpr uint32_t kPbufOff = kStageOff + 32768 + 3072, kMiscOff = kPbufOff + kWarps * kTile * 8;   // stage area: 32 KB (LUT-build chunks) + combine spill
    unsigned char* lut_p = smem + kLutOff;
    unsigned char* stage_p = smem + kStageOff;                           // kWarps * (K tile + V tile) = 16 KB used in the main loop
    unsigned char* pbuf_p = smem + kPbufOff;
    int* flag = reinterpret_cast<int*>(smem + kMiscOff);
    float* vo_acc = reinterpret_cast<float*>(smem + kMiscOff + 1024) + (threadIdx.x >> 5) * (G * 128);   // OUT & 2: this warp's V-outlier sums [g][dim]
    float* xch = reinterpret_cast<float*>(stage_p);                      // 2 * kWarps entries, then kMergeScratch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (smem_u32(smem) != kSmemBase) {
        if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
            printf("million_b200: dynamic shared memory starts at 0x%x, expected 0x%x\n", smem_u32(smem), kSmemBase);
        __trap();
    }
    const int Gfull = a.nh / a.nh_k;
    const int h0 = hk * Gfull + sub * G;
    const int hb = b * a.nh_k + hk;
    const bool has_codes = t1 > t0;

    // ---------------------------------------------------------------- prologue: V table + K LUT
    if (has_codes) {
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
            const uint4* vsrc = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(prepared) + 65536);
            for (int i = tid; i < kVtabBytes / 16; i += kThreads) cp_async16(vtab_s + i * 16, vsrc + i, 16);
            cp_async_commit();
        }
        load_chunk(1);
        // thread owns column `col` (= one sub-space) and walks the codes: LUT[c][col] = <q_h[4m..4m+3], Kcent[m][c]> per head
        const int col = tid & 31;
        const int bb = ((col >> 4) << 1) | ((col >> 3) & 1), W = col & 7, m = 4 * W + bb;
        float qv[G][4];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T* q = reinterpret_cast<const T*>(a.q) + (int64_t)(b * a.nh + h0 + g) * 128 + 4 * m;
#pragma unroll
            for (int k = 0; k < 4; ++k) qv[g][k] = io<T>::to_f(q[k]);
        }
        for (int ch = 0; ch < 4; ++ch) {
            if (ch == 0) cp_async_wait<2>();
            else if (ch < 3) cp_async_wait<1>();
            else cp_async_wait<0>();
            __syncthreads();
            const unsigned char* src = stage_p + (ch & 1) * 16384;
#pragma unroll 4
            for (int cl = tid >> 5; cl < 64; cl += kThreads / 32) {
                const int c = ch * 64 + cl;
                const uint2 cv = *reinterpret_cast<const uint2*>(src + (cl * 32 + col) * 8);
                const float2 c01 = __half22float2(as_h2(cv.x)), c23 = __half22float2(as_h2(cv.y));
                float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int g = 0; g < G; ++g) e[g] = fmaf(c23.y, qv[g][3], fmaf(c23.x, qv[g][2], fmaf(c01.y, qv[g][1], c01.x * qv[g][0])));
                *reinterpret_cast<uint2*>(lut_p + c * 256 + col * 8) = make_uint2(as_u32(__floats2half2_rn(e[0], e[1])), as_u32(__floats2half2_rn(e[2], e[3])));
            }
            __syncthreads();
            if (ch + 2 < 4) load_chunk(ch + 2);
        }
        if constexpr ((OUT & 2) != 0) {
            for (int i = (threadIdx.x & 31); i < G * 128; i += 32) vo_acc[i] = 0.f;
        }
        if constexpr ((OUT & 1) != 0) {
            // q as a [dim][head] table for the outlier terms (misc area; flag and merge scratch there are used after the loop)
            T* qt = reinterpret_cast<T*>(smem + kMiscOff);
            for (int i = tid; i < 128 * G; i += kThreads)
                qt[i] = reinterpret_cast<const T*>(a.q)[(int64_t)(b * a.nh + h0 + (i % G)) * 128 + i / G];
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- main loop over this warp's tiles
    float run_m[G], run_l[G], out_o[2][G][4];    // slot s of lane (hw, lq) = sub-space 2*lq + ((s + lq) & 1), 4 dims
#pragma unroll
    for (int g = 0; g < G; ++g) {
        run_m[g] = -INFINITY; run_l[g] = 0.f;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int k = 0; k < 4; ++k) out_o[s][g][k] = 0.f;
    }
    const int lq = lane & 15, hw = lane >> 4;
    if (has_codes) {
        const int rot = ((lane & 7) + (lane >> 3)) & 7, hbit = (lane >> 3) & 1;
        unsigned char* ksp = stage_p + warp * 2 * dm4::kTileBytes;
        unsigned char* vsp = ksp + dm4::kTileBytes;
        unsigned char* pbuf_w = pbuf_p + warp * kTile * 8;
        const uint32_t ks_s = smem_u32(ksp), vs_s = smem_u32(vsp);
        const uint8_t* kbase = a.k_codes + hb * a.k_head_stride;
        const uint8_t* vbase = a.v_codes + (a.v_layout == MILLION_V_PAGED ? 0 : hb * a.v_head_stride);

        uint32_t koff[8];    // byte0 = column offset for an even byte, byte1 = for an odd byte (both without the (b/2)*128 part)
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int Wl = (w + rot) & 7;
            koff[w] = (uint32_t)(Wl * 8) | ((uint32_t)(Wl * 8 + 64) << 8);
        }
        // step bq uses byte (bq + hbit) & 3 of the word: selector and the (b/2)*128 part of the column offset, per lane
        uint32_t ksel[4], kimm[4];
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) {
            const int bsel = (bq + hbit) & 3;
            ksel[bq] = (uint32_t)(4 + (bsel & 1)) | ((uint32_t)bsel << 4) | 0x6600u;
            kimm[bq] = (uint32_t)((bsel >> 1) * 128);
        }
        // PV: slot s -> byte (s + lq) & 1 of the lane's 16-bit code pair; column offset of sub-space 2*lq + that byte
        uint32_t voff, vsel[2];
        {
            uint32_t o[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int bsel = (s + lq) & 1;
                o[s] = (uint32_t)(dm4::col_of(2 * lq + bsel) * 8);
                vsel[s] = (uint32_t)(4 + s) | ((uint32_t)bsel << 4) | 0x6600u;
            }
            voff = o[0] | (o[1] << 8);
        }

        const int n_tiles = (t1 - t0 + kTile - 1) / kTile;
        auto issue = [&](int tile, const uint8_t* gbase, uint32_t dst) {
            const int tok0 = t0 + tile * kTile;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int chunk = lane + i * 32;            // 0..63: 32 tokens * 2 chunks of 16 B
                const int tok = tok0 + (chunk >> 1);
                const int ok = (tile < n_tiles && tok < t1) ? 16 : 0;
                cp_async16(dst + chunk * 16, gbase + (int64_t)(ok ? tok : t0) * dm4::kRow + (chunk & 1) * 16, ok);
            }
            cp_async_commit();
        };
        // transposed value codes: the tile is 32 sub-space rows of 32 tokens; row m is staged at row pi(m) = m/2 + 16*(m%2)
        // (32 bytes each) so that the word reads of the PV phase below are bank-conflict free
        auto issue_vt = [&](int tile) {
            const int tok0 = t0 + tile * kTile;
            const bool in_range = tile < n_tiles;
            const uint8_t* src0 = vbase;
            int64_t row_stride = a.v_ld;
            if (in_range) {
                if (a.v_layout == MILLION_V_PAGED) {
                    const int64_t page = __ldg(a.v_page_ids + (int64_t)hb * a.n_pages + tok0 / a.page_size);
                    src0 = a.v_codes + page * 32 * a.page_size + (tok0 % a.page_size);
                    row_stride = a.page_size;
                } else {
                    src0 = vbase + tok0;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int chunk = lane + i * 32;            // 0..63: 32 rows * 2 chunks of 16 tokens
                const int m = chunk >> 1, hc = chunk & 1;
                int ok = in_range ? (t1 - (tok0 + hc * 16)) : 0;
                ok = ok < 0 ? 0 : (ok > 16 ? 16 : ok);
                const uint32_t dst = vs_s + (uint32_t)((((m >> 1) + 16 * (m & 1)) * 32) + hc * 16);
                cp_async16(dst, (ok ? src0 : vbase) + (ok ? (int64_t)m * row_stride + hc * 16 : 0), ok);
            }
            cp_async_commit();
        };
        auto issue_v = [&](int tile) {
            if constexpr (VL == 0) issue(tile, vbase, vs_s);
            else issue_vt(tile);
        };
        issue(warp, kbase, ks_s);
        issue_v(warp);

        __half2 acc[2][2][G];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int g = 0; g < G; ++g) { acc[s][0][g] = __float2half2_rn(0.f); acc[s][1][g] = __float2half2_rn(0.f); }
        auto flush = [&]() {
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float2 f0 = __half22float2(acc[s][0][g]), f1 = __half22float2(acc[s][1][g]);
                    out_o[s][g][0] += f0.x; out_o[s][g][1] += f0.y; out_o[s][g][2] += f1.x; out_o[s][g][3] += f1.y;
                    acc[s][0][g] = __float2half2_rn(0.f); acc[s][1][g] = __float2half2_rn(0.f);
                }
        };
        int since_flush = 0;

        // K-side outlier records of my token in the NEXT tile, fetched one tile ahead (see attn_fast.cu)
        // Raw load results only: nothing consumes them before the next iteration (a shift or an OR here would make this warp wait
        // for the HBM round trip on the spot).  k_out 1, 2, 4: one vector load each for dims and deltas; 3: byte-wise.
        uint32_t ko_dims = 0, ko_v01 = 0, ko_v23 = 0;
        // V-side records (OUT bit 1), same prefetch; applied in the QK phase where a lane is its token (see attn_fast.cu)
        const bool ko_pair1 = (OUT & 1) != 0 && ko_pair1_ok(a.k_out, a.ko_idx, a.ko_val, a.ko_head_stride);
        const bool vo_pair1 = (OUT & 2) != 0 && ko_pair1_ok(a.v_out, a.vo_idx, a.vo_val, a.vo_head_stride);
        const int ko_sh = ko_pair1 ? (lane & 1) : 0, vo_sh = vo_pair1 ? (lane & 1) : 0;   // t0 and the tile size are even: token parity = lane parity
        uint32_t vo_dims = 0, vo_v01 = 0, vo_v23 = 0;
        auto vo_fetch = [&](int tile, int dep) {
            if constexpr ((OUT & 2) != 0) {
                const int tok = t0 + tile * kTile + lane + dep;
                const bool ok = tile < n_tiles && tok < t1;
                const int tk = ok ? tok : t0;
                const int64_t rec = hb * a.vo_head_stride + (int64_t)(vo_pair1 ? (tk & ~1) : tk) * a.v_out;
                ko_load(a.vo_idx + rec, reinterpret_cast<const unsigned short*>(a.vo_val) + rec, a.v_out, vo_pair1, vo_dims, vo_v01, vo_v23);
            }
        };
        vo_fetch(warp, 0);
        auto ko_fetch = [&](int tile, int dep) {
            if constexpr ((OUT & 1) != 0) {
                const int tok = t0 + tile * kTile + lane + dep;
                const bool ok = tile < n_tiles && tok < t1;
                const int tk = ok ? tok : t0;
                const int64_t rec = hb * a.ko_head_stride + (int64_t)(ko_pair1 ? (tk & ~1) : tk) * a.k_out;
                const unsigned short* vals = reinterpret_cast<const unsigned short*>(a.ko_val) + rec;
                ko_load(a.ko_idx + rec, vals, a.k_out, ko_pair1, ko_dims, ko_v01, ko_v23);
            }
        };
        ko_fetch(warp, 0);

        for (int tile = warp; tile < n_tiles; tile += kWarps) {
            cp_async_wait<1>();          // pending [K(i), V(i)] -> K(i) landed
            __syncwarp();
            const int tok = t0 + tile * kTile + lane;
            const bool valid = tok < t1;
            // the records the previous iteration's loads brought in are taken over BEFORE the next loads are issued (rec_take /
            // rec_dep in attn_fast_helpers.cuh: the order is a data dependence, not a hope).  The parity pick of the pair loads
            // (ko_sh / vo_sh) happens at the consumers.
            uint32_t my_dims = 0, my_v01 = 0, my_v23 = 0, my_vdims = 0, my_vv01 = 0, my_vv23 = 0;
            int dep = 0;
            if constexpr ((OUT & 1) != 0) {
                my_dims = rec_take(ko_dims, a.zero); my_v01 = rec_take(ko_v01, a.zero); my_v23 = rec_take(ko_v23, a.zero);
                dep |= rec_dep(my_dims, my_v01, my_v23, a.zero);
            }
            if constexpr ((OUT & 2) != 0) {
                my_vdims = rec_take(vo_dims, a.zero); my_vv01 = rec_take(vo_v01, a.zero); my_vv23 = rec_take(vo_v23, a.zero);
                dep |= rec_dep(my_vdims, my_vv01, my_vv23, a.zero);
            }
            ko_fetch(tile + kWarps, dep);          // K and V records: both taken over before either prefetch is issued
            vo_fetch(tile + kWarps, dep);

            // ------------------------------------------------ QK: 32 conflict-free LUT gathers for my token
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t words[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) words[w] = lds32(ksp, lane * dm4::kRow + (((w + rot) & 7) << 2));
#pragma unroll
            for (int w = 0; w < 8; ++w) {
#pragma unroll
                for (int bq = 0; bq < 4; ++bq) {
                    const uint32_t ad = __byte_perm(words[w], koff[w], ksel[bq]) + kimm[bq];      // code*256 + column*8
                    const uint2 e = gather64<kSmemBase + kLutOff>(ad);
                    if constexpr (G == 1) { const unsigned short lo = (unsigned short)(e.x & 0xffffu); asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(s4[0]) : "h"(lo)); }
                    else fhadd2(s4[0], s4[1], e.x);
                    if constexpr (G == 4) fhadd2(s4[2], s4[3], e.y);
                }
            }
            if constexpr ((OUT & 1) != 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < a.k_out) {
                        const uint32_t pair = i < 2 ? my_v01 : my_v23;
                        const unsigned short hv = (unsigned short)((i & 1) ? (pair >> 16) : ((pair >> (16 * ko_sh)) & 0xffffu));
                        const float dv = valid ? io<T>::to_f(*reinterpret_cast<const T*>(&hv)) : 0.f;
                        // q_g[dim] for the G heads from the [dim][head] table in shared memory (one load)
                        const int dim = (my_dims >> (8 * (i + ko_sh))) & 0xff;
                        const unsigned char* qrow = smem + kMiscOff + dim * (2 * G);
                        if constexpr (G == 4) {
                            const uint2 w = *reinterpret_cast<const uint2*>(qrow);
                            const float2 q01 = io<T>::to_f2(w.x), q23 = io<T>::to_f2(w.y);
                            s4[0] = fmaf(q01.x, dv, s4[0]); s4[1] = fmaf(q01.y, dv, s4[1]);
                            s4[2] = fmaf(q23.x, dv, s4[2]); s4[3] = fmaf(q23.y, dv, s4[3]);
                        } else if constexpr (G == 2) {
                            const float2 q01 = io<T>::to_f2(*reinterpret_cast<const uint32_t*>(qrow));
                            s4[0] = fmaf(q01.x, dv, s4[0]); s4[1] = fmaf(q01.y, dv, s4[1]);
                        } else {
                            s4[0] = fmaf(io<T>::to_f(*reinterpret_cast<const T*>(qrow)), dv, s4[0]);
                        }
                    }
            }
            __syncwarp();
            issue(tile + kWarps, kbase, ks_s);
            float s[G];
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] = valid ? s4[g] * a.scale_log2 : -INFINITY;

            // ------------------------------------------------ online softmax (lazy max)
            bool need = false;
#pragma unroll
            for (int g = 0; g < G; ++g) need = need || (s[g] > run_m[g] + kRescaleMargin);
            if (__any_sync(0xffffffffu, need)) {
                flush();
                since_flush = 0;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float nm = fmaxf(run_m[g], warp_max(s[g]));
                    if (nm > run_m[g]) {
                        const float alpha = exp2_safe(run_m[g], nm);
                        run_l[g] *= alpha;
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                            for (int k = 0; k < 4; ++k) out_o[sl][g][k] *= alpha;
                        if constexpr ((OUT & 2) != 0) {      // warp-uniform branch
                            __syncwarp();
                            for (int i = lane; i < 128; i += 32) vo_acc[g * 128 + i] *= alpha;
                            __syncwarp();
                        }
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
            if constexpr ((OUT & 2) != 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < a.v_out) {
                        const uint32_t pair = i < 2 ? my_vv01 : my_vv23;
                        const unsigned short hv = (unsigned short)((i & 1) ? (pair >> 16) : ((pair >> (16 * vo_sh)) & 0xffffu));
                        const float dv = valid ? io<T>::to_f(*reinterpret_cast<const T*>(&hv)) : 0.f;
                        const int dim = (my_vdims >> (8 * (i + vo_sh))) & 0xff;
#pragma unroll
                        for (int g = 0; g < G; ++g) atomicAdd(vo_acc + g * 128 + dim, p[g] * dv);
                    }
            }
            *reinterpret_cast<uint2*>(pbuf_w + lane * 8) = make_uint2(as_u32(__floats2half2_rn(p[0], p[1])), as_u32(__floats2half2_rn(p[2], p[3])));

            cp_async_wait<1>();          // pending [V(i), K(i+1)] -> V(i) landed
            __syncwarp();

            // ------------------------------------------------ PV: a half-warp per token, lane owns 2 sub-spaces (4 dims each)
            if constexpr (VL == 1) {
                // transposed codes: a word holds 4 tokens of one sub-space.  Half-warp hw takes tokens 16*hw .. 16*hw+15 in 4 groups
                // of 4; the group order is rotated by lq/4 so the 32 word reads of one instruction hit 32 banks (staged row of
                // sub-space 2*lq+s is lq + 16*s: bank 8*(lq%4) + 4*hw + group).
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int tgl = (u + (lq >> 2)) & 3;
                    const int tg = tgl + 4 * hw;
                    const uint32_t wa = lds32(vsp, lq * 32 + hw * 16 + tgl * 4);            // sub-space 2*lq
                    const uint32_t wb = lds32(vsp, (16 + lq) * 32 + hw * 16 + tgl * 4);     // sub-space 2*lq + 1
                    const uint32_t ws[2] = {(lq & 1) ? wb : wa, (lq & 1) ? wa : wb};        // slot sl <-> sub-space 2*lq + ((sl + lq) & 1)
                    const uint4 pa = lds128(pbuf_w, tg * 32), pb = lds128(pbuf_w, tg * 32 + 16);
                    const uint32_t pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __half2 p01 = as_h2(pp[2 * i]), p23 = as_h2(pp[2 * i + 1]);
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) {
                            const uint32_t sel = (uint32_t)(4 + sl) | ((uint32_t)i << 4) | 0x6600u;
                            const uint32_t ad = __byte_perm(ws[sl], voff, sel);
                            const uint2 v = gather64<kSmemBase + kVtabOff>(ad);
                            const __half2 v01 = as_h2(v.x), v23 = as_h2(v.y);
                            acc[sl][0][0] = __hfma2(__low2half2(p01), v01, acc[sl][0][0]);
                            acc[sl][1][0] = __hfma2(__low2half2(p01), v23, acc[sl][1][0]);
                            if constexpr (G >= 2) {
                                acc[sl][0][1] = __hfma2(__high2half2(p01), v01, acc[sl][0][1]);
                                acc[sl][1][1] = __hfma2(__high2half2(p01), v23, acc[sl][1][1]);
                            }
                            if constexpr (G == 4) {
                                acc[sl][0][2] = __hfma2(__low2half2(p23), v01, acc[sl][0][2]);
                                acc[sl][1][2] = __hfma2(__low2half2(p23), v23, acc[sl][1][2]);
                                acc[sl][0][3] = __hfma2(__high2half2(p23), v01, acc[sl][0][3]);
                                acc[sl][1][3] = __hfma2(__high2half2(p23), v23, acc[sl][1][3]);
                            }
                        }
                    }
                }
            } else
#pragma unroll kDm4PvUnroll
            for (int jp = 0; jp < kTile / 2; ++jp) {
                const int j = 2 * jp + hw;
                const uint32_t pair = *reinterpret_cast<const unsigned short*>(vsp + j * dm4::kRow + lq * 2);
                const uint2 pk = lds64(pbuf_w, j * 8);
                const __half2 p01 = as_h2(pk.x), p23 = as_h2(pk.y);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const uint32_t ad = __byte_perm(pair, voff, vsel[sl]);
                    const uint2 v = gather64<kSmemBase + kVtabOff>(ad);
                    const __half2 v01 = as_h2(v.x), v23 = as_h2(v.y);
                    acc[sl][0][0] = __hfma2(__low2half2(p01), v01, acc[sl][0][0]);
                    acc[sl][1][0] = __hfma2(__low2half2(p01), v23, acc[sl][1][0]);
                    if constexpr (G >= 2) {
                        acc[sl][0][1] = __hfma2(__high2half2(p01), v01, acc[sl][0][1]);
                        acc[sl][1][1] = __hfma2(__high2half2(p01), v23, acc[sl][1][1]);
                    }
                    if constexpr (G == 4) {
                        acc[sl][0][2] = __hfma2(__low2half2(p23), v01, acc[sl][0][2]);
                        acc[sl][1][2] = __hfma2(__low2half2(p23), v23, acc[sl][1][2]);
                        acc[sl][0][3] = __hfma2(__high2half2(p23), v01, acc[sl][0][3]);
                        acc[sl][1][3] = __hfma2(__high2half2(p23), v23, acc[sl][1][3]);
                    }
                }
            }
            __syncwarp();
            issue_v(tile + kWarps);
            if (++since_flush == MILLION_DM4_FLUSH_TILES) { flush(); since_flush = 0; }
        }
        flush();
        cp_async_wait<0>();
    }

    // ---------------------------------------------------------------- my share of the fp16 window (exact attention)
    // The r recent tokens are dealt out to the splits of the group (r/S tokens each, one token per warp at a time), so no
    // CTA carries a long serial tail.  Lane owns dims 4*lane .. 4*lane+3.
    float wm[G], wl[G], wo[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) { wm[g] = -INFINITY; wl[g] = 0.f; wo[g][0] = wo[g][1] = wo[g][2] = wo[g][3] = 0.f; }
    {
        const int rw = window_rows(a);
        const int w0 = (int)((long long)rw * split / np), w1 = (int)((long long)rw * (split + 1) / np);
        const int t_new = (a.k_new != nullptr && rw - 1 >= w0 && rw - 1 < w1) ? rw - 1 : -1;   // only the split that owns the row: the batched loads below also visit (masked) rows beyond w1   // fused window append, as in attn_fast.cu
        if (w0 + warp < w1) {
            float qv[G][4];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint2 qr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.q) + (int64_t)(b * a.nh + h0 + g) * 128 + 4 * lane));
                const float2 q01 = io<T>::to_f2(qr.x), q23 = io<T>::to_f2(qr.y);
                qv[g][0] = q01.x * a.scale_log2; qv[g][1] = q01.y * a.scale_log2; qv[g][2] = q23.x * a.scale_log2; qv[g][3] = q23.y * a.scale_log2;
            }
            for (int t = w0 + warp; t < w1; t += kWarps) {
                const int64_t row = ((int64_t)hb * a.res_len + t) * 128 + 4 * lane;
                uint2 kr, vr;
                if (t == t_new) {
                    kr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.k_new) + (int64_t)hb * 128 + 4 * lane));
                    vr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.v_new) + (int64_t)hb * 128 + 4 * lane));
                    if (sub == 0) {
                        *reinterpret_cast<uint2*>(reinterpret_cast<T*>(const_cast<void*>(a.k_res)) + row) = kr;
                        *reinterpret_cast<uint2*>(reinterpret_cast<T*>(const_cast<void*>(a.v_res)) + row) = vr;
                    }
                } else {
                    kr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.k_res) + row));
                    vr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.v_res) + row));
                }
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
    // ---------------------------------------------------------------- combine the warps of this CTA -> one partial state
    // 2 * kWarps entries of [G*128 o | G m | G l]: entry w = coded tokens of warp w, entry kWarps + w = its window tokens.
    // Coded layout: slot sl of lane (hw, lq) holds sub-space 4*lq + ((sl + hw) & 3); hw=1 is folded into hw=0 first.
    constexpr int kEntry = (G * 130 + 3) & ~3;   // 16-byte aligned entries
    {
#pragma unroll
        for (int g = 0; g < G; ++g) run_l[g] = warp_sum(run_l[g]);
        float* wx = xch + warp * kEntry;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // both half-warps use the same slot -> sub-space map; fold hw=1 into hw=0
                    const float theirs = __shfl_xor_sync(0xffffffffu, out_o[sl][g][k], 16);
                    if (hw == 0) {
                        float v = out_o[sl][g][k] + theirs;
                        if constexpr ((OUT & 2) != 0) { if (t1 > t0) v += vo_acc[g * 128 + 4 * (2 * lq + ((sl + lq) & 1)) + k]; }
                        wx[g * 128 + 4 * (2 * lq + ((sl + lq) & 1)) + k] = v;
                    }
                }
        float* ww = xch + (kWarps + warp) * kEntry;
#pragma unroll
        for (int g = 0; g < G; ++g)
            *reinterpret_cast<float4*>(ww + g * 128 + 4 * lane) = make_float4(wo[g][0], wo[g][1], wo[g][2], wo[g][3]);
        if (lane == 0) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                wx[G * 128 + g] = run_m[g]; wx[G * 128 + G + g] = run_l[g];
                ww[G * 128 + g] = wm[g];   ww[G * 128 + G + g] = wl[g];
            }
        }
    }
    __syncthreads();
    {
        // thread -> (dim = tid & 127, head half = tid >> 7)
        constexpr int GH = G >= 2 ? G / 2 : 1;
        const int dim = tid & 127, hsel = tid >> 7;
        if (G >= 2 || hsel == 0) {
#pragma unroll
            for (int gi = 0; gi < GH; ++gi) {
                const int g = (G >= 2 ? hsel * GH : 0) + gi;
                float mstar = -INFINITY;
#pragma unroll
                for (int e = 0; e < 2 * kWarps; ++e) mstar = fmaxf(mstar, xch[e * kEntry + G * 128 + g]);
                float o = 0.f, l = 0.f;
#pragma unroll
                for (int e = 0; e < 2 * kWarps; ++e) {
                    const float* ex = xch + e * kEntry;
                    const float sc = exp2_safe(ex[G * 128 + g], mstar);
                    o = fmaf(ex[g * 128 + dim], sc, o);
                    l = fmaf(ex[G * 128 + G + g], sc, l);
                }
                float* part = a.parts + ((int64_t)(b * a.nh + h0 + g) * a.n_parts + split) * part_stride(128);
                part[dim] = o;
                if (dim == 0) { part[128] = mstar; part[129] = l; }
            }
        }
    }
    // ---------------------------------------------------------------- last CTA of the (b, hk) group merges
    const bool last = last_cta_of_group(a.counters, hb, np * gsub, flag);
    if (last) merge_group<T>(a, b, hk, np, xch);
}

template <typename T, int G, int VL, int OUT>
__global__ void __launch_bounds__(fast::kThreads, 1) attn_fast_dm4_kernel(const AttnArgs a, const uint32_t* __restrict__ prepared, const int gsub) {
    pdl_launch_dependents();
    pdl_wait();            // PDL here only hides the launch latency (attn_fast.cu also overlaps its table copies and first tiles)
    if (!a.flat) {
        int t0, t1;
        split_range(a, blockIdx.x, t0, t1);
        attn_dm4_segment<T, G, VL, OUT>(a, prepared, gsub, blockIdx.x, blockIdx.y / gsub, blockIdx.y % gsub, blockIdx.z, t0, t1, a.n_splits);
        return;
    }
    // flat scheduling: see attn_fast_kernel (attn_fast.cu)
    const int ug = a.flat_ug, per = a.flat_per, vg = ug + fast::kFlatPad;
    const long long total = (long long)a.bs * a.nh_k * vg;
    const long long run0 = (long long)blockIdx.x * per;
    const long long run1 = (run0 + per < total) ? run0 + per : total;
    for (int grp = (int)(run0 / vg); grp < a.bs * a.nh_k && (long long)grp * vg < run1; ++grp) {
        const long long real0 = (long long)grp * vg + fast::kFlatPad, real1 = real0 + ug;
        const long long s0 = run0 > real0 ? run0 : real0, s1 = run1 < real1 ? run1 : real1;
        if (s1 <= s0) continue;
        const int first = (int)(real0 / per), last = (int)((real1 - 1) / per);
        const int us = (int)(s0 - real0), ue = (int)(s1 - real0);
        const int t1 = (ue * 64 < a.nk) ? ue * 64 : a.nk;
        attn_dm4_segment<T, G, VL, OUT>(a, prepared, 1, (int)blockIdx.x - first, grp % a.nh_k, 0, grp / a.nh_k, us * 64, t1, last - first + 1);
        __syncthreads();
    }
}

template <typename T, int G, int VL, int OUT>
static int launch_dm4_t(const AttnArgs& a, const uint32_t* prepared, int gsub, cudaStream_t stream) {
    using namespace fast;
    const size_t smem = dm4::kLutBytes + kVtabBytes + 32768 + 3072 + kWarps * kTile * 8 + (OUT ? 1024 : 256) + ((OUT & 2) ? kWarps * G * 128 * 4 : 0);
    static_assert(32768 + 3072 >= 2 * kWarps * 4 * 130 * sizeof(float) + 64, "stage area too small for the combine");
    static_assert(32768 + 3072 >= kMergeScratch * sizeof(float), "stage area too small for the merge scratch");
    static SmemAttrOnce configured = {};
    MILLION_CUDA_OK(ensure_dynamic_smem(configured, attn_fast_dm4_kernel<T, G, VL, OUT>, smem));
    dim3 grid(a.n_splits, a.nh_k * gsub, a.bs), block(kThreads);
    if (a.flat) grid = dim3((unsigned)(((long long)a.bs * a.nh_k * (a.flat_ug + kFlatPad) + a.flat_per - 1) / a.flat_per), 1, 1);
    MILLION_CUDA_OK(launch_kernel(attn_fast_dm4_kernel<T, G, VL, OUT>, grid, block, smem, stream, a.pdl != 0, a, prepared, gsub));
    return MILLION_OK;
}

int launch_attn_fast_dm4(const AttnArgs& a, int io_dtype, int G, int gsub, const void* prepared, cudaStream_t stream) {
    const uint32_t* prep = reinterpret_cast<const uint32_t*>(prepared);
    const int kv = (a.nk > 0 && a.k_out ? 1 : 0) | (a.nk > 0 && a.v_out ? 2 : 0);     // outlier side stores: bit 0 K, bit 1 V
    const bool vl = a.v_layout != MILLION_V_ROWMAJOR;
#define MILLION_DM4_CASE(TT, GG)                                                  \
    switch (kv) {                                                                 \
        case 0: return vl ? launch_dm4_t<TT, GG, 1, 0>(a, prep, gsub, stream) : launch_dm4_t<TT, GG, 0, 0>(a, prep, gsub, stream); \
        case 1: return vl ? launch_dm4_t<TT, GG, 1, 1>(a, prep, gsub, stream) : launch_dm4_t<TT, GG, 0, 1>(a, prep, gsub, stream); \
        case 2: return vl ? launch_dm4_t<TT, GG, 1, 2>(a, prep, gsub, stream) : launch_dm4_t<TT, GG, 0, 2>(a, prep, gsub, stream); \
        default: return vl ? launch_dm4_t<TT, GG, 1, 3>(a, prep, gsub, stream) : launch_dm4_t<TT, GG, 0, 3>(a, prep, gsub, stream); \
    }
    if (io_dtype == MILLION_F16) {
        if (G == 4) MILLION_DM4_CASE(__half, 4)
        if (G == 2) MILLION_DM4_CASE(__half, 2)
        MILLION_DM4_CASE(__half, 1)
    }
    if (G == 4) MILLION_DM4_CASE(__nv_bfloat16, 4)
    if (G == 2) MILLION_DM4_CASE(__nv_bfloat16, 2)
    MILLION_DM4_CASE(__nv_bfloat16, 1)
#undef MILLION_DM4_CASE
}

}  // namespace million
