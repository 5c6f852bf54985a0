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
#include "attn_fast_helpers.cuh"

// Tiles between two flushes of the packed-half PV sums into fp32.  1 = every tile (16 terms per lane): measured FASTER than 2
// (97.5 vs 99.6 us per launch at batch 8, A/B on one box: the flush code replaces the counter's branch) and it halves the
// life of the half sums; tools/precision_budget.py shows the sums are not what limits accuracy either way.
#ifndef MILLION_PV_UNROLL
#define MILLION_PV_UNROLL 8      // PV loop (8 steps of 2 tokens per half-warp) fully unrolled: 95.7 us per launch at batch 8 vs 96.8 (4), 97.4 (2), 105.9 (1) — profiles/r02_ab_variants.txt
#endif
#ifndef MILLION_PV_FLUSH_TILES
#define MILLION_PV_FLUSH_TILES 1
#endif

namespace million {

constexpr int kPvUnroll = MILLION_PV_UNROLL;

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

// One segment = the coded tokens [t0, t1) of group (b, hk) (+ this CTA's share of the window), written as part `split` of
// `np` parts of that group.
// OUT bit 0: K-side outlier records (a.k_out <= 4 per token) are added to the scores.  OUT bit 1 (G <= 2 only: the G = 4 kernel
// has no shared memory left): V-side records — in the QK phase a lane IS its token, so right after the softmax weights it adds
// p_g * delta to a per-warp fp32 accumulator [g][dim] in shared memory (red.shared.add.f32; dims of one token are distinct, lanes
// collide rarely), rescaled with the warp's running max and folded into the warp's state at the end.  OUT = 0 is the plain
// path, untouched.
// P2P = 1: the group's last CTA also runs the cross-GPU exchange of the split-KV partitioning (merge_group<T, true>).
template <typename T, int G, int VL, int OUT, int P2P>
__device__ __forceinline__ void attn_fast_segment(const AttnArgs& a, const uint32_t* __restrict__ prepared, const int gsub, const int split,
                                                  const int hk, const int sub, const int b, const int t0, const int t1, const int np, const int piece) {
    using namespace fast;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t kLutOff = 0, kVtabOff = LutCfg<G>::bytes, kStageOff = kVtabOff + kVtabBytes;
    constexpr uint32_t kPbufOff = kStageOff + kWarps * kStageBytes, kMiscOff = kPbufOff + kWarps * kTile * 8;
    unsigned char* lut_p = smem + kLutOff;
    unsigned char* stage_p = smem + kStageOff;                           // kWarps * (K tile + V tile)
    unsigned char* pbuf_p = smem + kPbufOff;                             // kWarps * kTile * 8 bytes (4 halves per token)
    int* flag = reinterpret_cast<int*>(smem + kMiscOff);
    float* vo_acc = reinterpret_cast<float*>(smem + kMiscOff + 1024) + (threadIdx.x >> 5) * (G * 128);   // OUT & 2: this warp's [g][dim]
    // after the main loop the stage buffers and p slots are dead: reuse them for the cross-warp combine and the merge
    float* xch = reinterpret_cast<float*>(stage_p);                      // 2 * kWarps * G * 130 floats, then kMergeScratch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (smem_u32(smem) != kSmemBase) {
        if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
            printf("million_b200: dynamic shared memory starts at 0x%x, expected 0x%x\n", smem_u32(smem), kSmemBase);
        __trap();
    }
    dbg_stamp(a, 0, piece);
    const int Gfull = a.nh / a.nh_k;
    const int h0 = hk * Gfull + sub * G;                                // first query head of this CTA
    const int hb = b * a.nh_k + hk;
    const bool has_codes = t1 > t0;

    // ---------------------------------------------------------------- per-lane constants and tile streaming (used from the prologue on)
    const int lq = lane & 15, hw = lane >> 4;
    const int rot = (lq + hw) & 15;
    unsigned char* ksp = stage_p + warp * kStageBytes;
    unsigned char* vsp = ksp + kTile * kRowBytes;
    unsigned char* pbuf_w = pbuf_p + warp * kTile * 8;
    const uint32_t ks_s = smem_u32(ksp), vs_s = smem_u32(vsp);
    const uint8_t* kbase = a.k_codes + hb * a.k_head_stride;
    const uint8_t* vbase = a.v_codes + (a.v_layout == MILLION_V_PAGED ? 0 : hb * a.v_head_stride);

    // per-lane gather constants
    uint32_t koff[16];   // QK: byte0 = column offset for even b, byte1 = for odd b, bytes 2,3 = 0
#pragma unroll
    for (int w = 0; w < 16; ++w) {
        const int Wl = (w + rot) & 15;
        if constexpr (G == 4) koff[w] = (uint32_t)(Wl * 8) | ((uint32_t)(Wl * 8 + 128) << 8);
        else koff[w] = (uint32_t)(Wl * 4) | ((uint32_t)(Wl * 4 + 64) << 8);   // + (bb>>1)*128 comes from the immediate
    }
    // PV: slot s -> byte bb = (s + hw) & 3 of the lane's V code word; column offset col_of(4*lq + bb) * 4
    uint32_t voff01, voff23, vsel[4];
    {
        uint32_t o[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int bb = (s + hw) & 3;
            o[s] = (uint32_t)(col_of(4 * lq + bb) * 4);
            // result = [off (from voff, byte 4 + (s&1)), code (byte bb of the word), 0, 0]
            vsel[s] = (uint32_t)(4 + (s & 1)) | ((uint32_t)bb << 4) | (6u << 8) | (6u << 12);
        }
        voff01 = o[0] | (o[1] << 8);
        voff23 = o[2] | (o[3] << 8);
    }

    const int n_tiles = (t1 - t0 + kTile - 1) / kTile;
    // one K buffer and one V buffer per warp; K(i+1) is requested right after QK(i), V(i+1) right after PV(i)
    auto issue = [&](int tile, const uint8_t* gbase, uint32_t dst) {
        const int tok0 = t0 + tile * kTile;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int chunk = lane + i * 32;            // 0..127: 32 tokens * 4 chunks of 16 B
            const int tok = tok0 + (chunk >> 2);
            const int ok = (tile < n_tiles && tok < t1 && !MILLION_DBG_MODE(a, 4)) ? 16 : 0;
            cp_async16(dst + chunk * 16, gbase + (int64_t)(ok ? tok : t0) * kRowBytes + (chunk & 3) * 16, ok);
        }
        cp_async_commit();
    };
    // transposed value codes: the tile is 64 sub-space rows of 32 tokens; row m is staged at row pi(m) = m/4 + 16*(m%4)
    // (32 bytes each) so that the word reads of the PV phase below are bank-conflict free
    auto issue_vt = [&](int tile) {
        const int tok0 = t0 + tile * kTile;
        const bool in_range = tile < n_tiles;
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
        cp_async_commit();
    };
    auto issue_v = [&](int tile) {
        if constexpr (VL == 0) issue(tile, vbase, vs_s);
        else issue_vt(tile);
    };

    // ---------------------------------------------------------------- prologue: V table + K LUT
    if (has_codes) {
        // The V table (already in gather order) is copied to its place in the background, and this warp's first K and V tiles
        // are requested right away: their HBM latency hides behind the LUT build (the stage area is not used for anything else).
        {
            const uint32_t vtab_s = smem_u32(smem + kVtabOff);
            const uint4* vsrc = reinterpret_cast<const uint4*>(prepared + 64 * 256);
            for (int i = tid; i < kVtabBytes / 16; i += kThreads) cp_async16(vtab_s + i * 16, vsrc + i, 16);
            cp_async_commit();
        }
        issue(warp, kbase, ks_s);
        issue_v(warp);
        pdl_wait();   // codes and tables are not written by the stream predecessor; q, the window and the workspace may be
        // K LUT: thread owns column `col` (= one sub-space) and walks the codes.  LUT[c][col] = <q_h[m], Kcent[m][c]> for the
        // G heads of the group: fp32 products of fp16 operands, rounded once to fp16 (G=1 keeps fp32 entries).  The gather table
        // kT (64 KB, L2 resident) is read straight into registers, 32 coalesced loads in flight per thread (staging it through
        // shared memory cost four block barriers and kept the stage area busy).
        const int col = tid & 63;
        const int bb = ((col >> 5) << 1) | ((col >> 4) & 1), W = col & 15, m = 4 * W + bb;
        float q0[G], q1[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T* q = reinterpret_cast<const T*>(a.q) + (int64_t)(b * a.nh + h0 + g) * 128;
            q0[g] = io<T>::to_f(q[2 * m]);
            q1[g] = io<T>::to_f(q[2 * m + 1]);
        }
        #ifndef MILLION_LUT_BATCH
#define MILLION_LUT_BATCH 32
#endif
        constexpr int kRowsPerPass = kThreads / 64, kBatch = MILLION_LUT_BATCH;
#pragma unroll 1
        for (int c0 = tid >> 6; c0 < 256; c0 += kRowsPerPass * kBatch) {
            uint32_t raw[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) raw[u] = __ldg(prepared + (c0 + u * kRowsPerPass) * 64 + col);
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int c = c0 + u * kRowsPerPass;
                const __half2 cv = as_h2(raw[u]);
                const float c0f = __low2float(cv), c1f = __high2float(cv);
                // fp32 products of exact fp16 operands, rounded once to fp16 (a packed-half build is 0.7 us faster per CTA but
                // costs a second rounding, visible on peaked score distributions)
                if constexpr (G == 4) {
                    const __half2 e01 = __floats2half2_rn(fmaf(c1f, q1[0], c0f * q0[0]), fmaf(c1f, q1[1], c0f * q0[1]));
                    const __half2 e23 = __floats2half2_rn(fmaf(c1f, q1[2], c0f * q0[2]), fmaf(c1f, q1[3], c0f * q0[3]));
                    // address(m, c) = (b>>1)*64K + c*256 + ((b&1)*16 + W)*8
                    *reinterpret_cast<uint2*>(lut_p + (bb >> 1) * 65536 + c * 256 + ((bb & 1) * 16 + W) * 8) = make_uint2(as_u32(e01), as_u32(e23));
                } else if constexpr (G == 2) {
                    const __half2 e01 = __floats2half2_rn(fmaf(c1f, q1[0], c0f * q0[0]), fmaf(c1f, q1[1], c0f * q0[1]));
                    *reinterpret_cast<uint32_t*>(lut_p + c * 256 + col * 4) = as_u32(e01);
                } else {
                    *reinterpret_cast<float*>(lut_p + c * 256 + col * 4) = fmaf(c1f, q1[0], c0f * q0[0]);
                }
            }
        }
        cp_async_wait<2>();    // pending: [V table, K(0), V(0)] -> my part of the V table has landed
        if constexpr ((OUT & 2) != 0) {
            for (int i = lane; i < G * 128; i += 32) vo_acc[i] = 0.f;
        }
        if constexpr ((OUT & 1) != 0) {
            // q as a [dim][head] table for the outlier terms.  It takes the whole misc area (with G = 4 this kernel then uses
            // exactly 227 KB); the ticket flag and the merge mbarrier in there are only touched after the main loop.
            T* qt = reinterpret_cast<T*>(smem + kMiscOff);
            for (int i = tid; i < 128 * G; i += kThreads)
                qt[i] = reinterpret_cast<const T*>(a.q)[(int64_t)(b * a.nh + h0 + (i % G)) * 128 + i / G];
        }
    }
    if (!has_codes) pdl_wait();
    __syncthreads();
    dbg_stamp(a, 1, piece);

    // ---------------------------------------------------------------- main loop over this warp's tiles
    WarpState<G> st;
    state_init(st);

    if (has_codes) {
        __half2 acc[4][G];
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
#pragma unroll
            for (int g = 0; g < G; ++g) acc[sl][g] = __float2half2_rn(0.f);
        int since_flush = 0;

        // K-side outlier records of my token in the NEXT tile, fetched one tile ahead (straight from HBM into registers)
        // Raw load results only: nothing consumes them before the next iteration (a shift or an OR here would make this warp wait
        // for the HBM round trip on the spot).  k_out 1, 2, 4: one vector load each for dims and deltas; 3: byte-wise.
        uint32_t ko_dims = 0, ko_v01 = 0, ko_v23 = 0;
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

            // ------------------------------------------------ QK: 64 conflict-free LUT gathers for my token
            float s[G];
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] = 0.f;
            auto qk_gathers = [&](const uint32_t word, const uint32_t kf) {
#pragma unroll
                for (int bq = 0; bq < 4; ++bq) {
                    if constexpr (G == 4) {
                        constexpr uint32_t selc[4] = {0x6604u, 0x6615u, 0x6624u, 0x6635u};
                        const uint32_t ad = __byte_perm(word, kf, selc[bq]);     // (code << 8) | column offset
                        const uint2 e = (bq >> 1) ? gather64<kSmemBase + kLutOff + 65536>(ad) : gather64<kSmemBase + kLutOff>(ad);
                        fhadd2(s[0], s[1], e.x);
                        fhadd2(s[2], s[3], e.y);
                    } else {
                        // the two half-warps use different bytes in the same step so that all 32 lanes hit 32 banks
                        constexpr uint32_t selc[4] = {0x6604u, 0x6615u, 0x6624u, 0x6635u};
                        const int b1 = (bq + 1) & 3;
                        const uint32_t ad = __byte_perm(word, kf, hw ? selc[b1] : selc[bq]) + (uint32_t)(((hw ? b1 : bq) >> 1) * 128);
                        const uint32_t e = gather32<kSmemBase + kLutOff>(ad);
                        if constexpr (G == 2) fhadd2(s[0], s[1], e);
                        else s[0] += __uint_as_float(e);
                    }
                }
            };
            {
                uint32_t words[16];
#pragma unroll
                for (int w = 0; w < 16; ++w) words[w] = lds32(ksp, lane * kRowBytes + (((w + rot) & 15) << 2));   // word (w + rot) % 16 of my row
                if (!MILLION_DBG_MODE(a, 1))
#pragma unroll
                    for (int w = 0; w < 16; ++w) qk_gathers(words[w], koff[w]);
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
                            s[0] = fmaf(q01.x, dv, s[0]); s[1] = fmaf(q01.y, dv, s[1]);
                            s[2] = fmaf(q23.x, dv, s[2]); s[3] = fmaf(q23.y, dv, s[3]);
                        } else if constexpr (G == 2) {
                            const float2 q01 = io<T>::to_f2(*reinterpret_cast<const uint32_t*>(qrow));
                            s[0] = fmaf(q01.x, dv, s[0]); s[1] = fmaf(q01.y, dv, s[1]);
                        } else {
                            s[0] = fmaf(io<T>::to_f(*reinterpret_cast<const T*>(qrow)), dv, s[0]);
                        }
                    }
            }
            __syncwarp();                                   // every lane has read its K row
            issue(tile + kWarps, kbase, ks_s);              // K(i+1) streams in during the PV phase
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] = valid ? s[g] * a.scale_log2 : -INFINITY;

            // ------------------------------------------------ online softmax (lazy max: rescale only when needed)
            bool need = false;
#pragma unroll
            for (int g = 0; g < G; ++g) need = need || (s[g] > st.m[g] + kRescaleMargin);
            if (__any_sync(0xffffffffu, need)) {
                // flush the packed-half PV accumulators first: they are relative to the old max
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float2 f = __half22float2(acc[sl][g]);
                        st.o[sl][g][0] += f.x; st.o[sl][g][1] += f.y;
                        acc[sl][g] = __float2half2_rn(0.f);
                    }
                since_flush = 0;
                float nm[G];
#pragma unroll
                for (int g = 0; g < G; ++g) nm[g] = fmaxf(st.m[g], warp_max(s[g]));
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (nm[g] > st.m[g]) {
                        const float alpha = exp2_fast(st.m[g], nm[g]);
                        st.l[g] *= alpha;
#pragma unroll
                        for (int sl = 0; sl < 4; ++sl) { st.o[sl][g][0] *= alpha; st.o[sl][g][1] *= alpha; }
                        if constexpr ((OUT & 2) != 0) {      // warp-uniform branch: the V-outlier sums follow the same max
                            __syncwarp();
                            for (int i = lane; i < 128; i += 32) vo_acc[g * 128 + i] *= alpha;
                            __syncwarp();
                        }
                        st.m[g] = nm[g];
                    }
                }
            }
            float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int g = 0; g < G; ++g) {
                p[g] = exp2_fast(s[g], st.m[g]);
                st.l[g] += p[g];
            }
            if constexpr ((OUT & 2) != 0) {
                // V-side outlier records of my token: o_g[dim] += p_g * delta (fp32, shared-memory reduction)
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
            // p for the PV phase: 4 halves (8 bytes) per token
            // row-major V: slot = token with its two low bits swapped (tokens t and t+2 adjacent); transposed V: natural order
            const int p_slot = VL == 0 ? ((lane & ~3) | ((lane & 1) << 1) | ((lane >> 1) & 1)) : lane;
            *reinterpret_cast<uint2*>(pbuf_w + p_slot * 8) = make_uint2(as_u32(__floats2half2_rn(p[0], p[1])), as_u32(__floats2half2_rn(p[2], p[3])));

            cp_async_wait<1>();          // pending [V(i), K(i+1)] -> V(i) landed
            __syncwarp();

            // ------------------------------------------------ PV: lane owns 4 sub-spaces (slot k -> 4*lq + ((k + hw) & 3))
            if constexpr (VL == 0) {
                // row-major codes: a half-warp per token, the lane's word holds its 4 sub-spaces of that token
                // The half-warp takes tokens 4*jq + hw and 4*jq + hw + 2 of every group of four (their rows lie 16 banks away
                // from the other half-warp's: conflict free); the p slots of those two tokens are adjacent (see p_slot), so ONE
                // 16-byte broadcast load brings both.
                if (!MILLION_DBG_MODE(a, 2))
#pragma unroll kPvUnroll
                for (int jq = 0; jq < kTile / 4; ++jq) {
                    const uint4 pk = lds128(pbuf_w, (4 * jq + 2 * hw) * 8);
                    uint32_t word[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) word[u] = lds32(vsp, (4 * jq + hw + 2 * u) * kRowBytes + lq * 4);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const __half2 p01 = as_h2(u ? pk.z : pk.x), p23 = as_h2(u ? pk.w : pk.y);
#pragma unroll
                        for (int sl = 0; sl < 4; ++sl) {
                            const uint32_t ad = __byte_perm(word[u], sl < 2 ? voff01 : voff23, vsel[sl]);
                            const __half2 v = as_h2(gather32<kSmemBase + kVtabOff>(ad));
                            acc[sl][0] = __hfma2(__low2half2(p01), v, acc[sl][0]);
                            if constexpr (G >= 2) acc[sl][1] = __hfma2(__high2half2(p01), v, acc[sl][1]);
                            if constexpr (G == 4) {
                                acc[sl][2] = __hfma2(__low2half2(p23), v, acc[sl][2]);
                                acc[sl][3] = __hfma2(__high2half2(p23), v, acc[sl][3]);
                            }
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
                    const uint4 pa = lds128(pbuf_w, tg * 32), pb = lds128(pbuf_w, tg * 32 + 16);
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
            __syncwarp();                                   // every lane is done with the V tile and the p slots
            issue_v(tile + kWarps);                          // V(i+1) streams in during the next QK phase
            if (++since_flush == MILLION_PV_FLUSH_TILES) {   // packed-half partial sums live for MILLION_PV_FLUSH_TILES tiles (16 terms per lane each)
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float2 f = __half22float2(acc[sl][g]);
                        st.o[sl][g][0] += f.x; st.o[sl][g][1] += f.y;
                        acc[sl][g] = __float2half2_rn(0.f);
                    }
                since_flush = 0;
            }
        }
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float2 f = __half22float2(acc[sl][g]);
                st.o[sl][g][0] += f.x; st.o[sl][g][1] += f.y;
            }
        cp_async_wait<0>();
    }

    dbg_stamp(a, 2, piece);
    // ---------------------------------------------------------------- my share of the fp16 window (exact attention)
    // The r recent tokens are dealt out to the splits of the group (r/S tokens each, one token per warp at a time), so no
    // CTA carries a long serial tail.  Lane owns dims 4*lane .. 4*lane+3.
    float wm[G], wl[G], wo[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) { wm[g] = -INFINITY; wl[g] = 0.f; wo[g][0] = wo[g][1] = wo[g][2] = wo[g][3] = 0.f; }
    {
        const int rw = window_rows(a);
        const int w0 = (int)((long long)rw * split / np), w1 = (int)((long long)rw * (split + 1) / np);
        // fused append: window row rw-1 is the token being decoded; it is read from k_new / v_new and stored into the window here
        // (by the one warp of the one split that owns that row), replacing the copy launch of pq_utils.py:304-311
        const int t_new = (a.k_new != nullptr && rw - 1 >= w0 && rw - 1 < w1) ? rw - 1 : -1;   // only the split that owns the row: the batched loads below also visit (masked) rows beyond w1
        if (w0 + warp < w1) {
            float qv[G][4];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint2 qr = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.q) + (int64_t)(b * a.nh + h0 + g) * 128 + 4 * lane));
                const float2 q01 = io<T>::to_f2(qr.x), q23 = io<T>::to_f2(qr.y);
                qv[g][0] = q01.x * a.scale_log2; qv[g][1] = q01.y * a.scale_log2; qv[g][2] = q23.x * a.scale_log2; qv[g][3] = q23.y * a.scale_log2;
            }
            // kWinBatch tokens per round: all their row loads are in flight together (a dependent load per token put ~1 us of loaded-memory
            // latency on every one of them: 6.5 us per CTA at 32K, batch 8), then one softmax update per head for the whole round
#ifndef MILLION_WIN_BATCH
#define MILLION_WIN_BATCH 4
#endif
            constexpr int kWinBatch = MILLION_WIN_BATCH;
            for (int tb = w0 + warp; tb < w1; tb += kWinBatch * kWarps) {
                uint2 kr[kWinBatch], vr[kWinBatch];
#pragma unroll
                for (int u = 0; u < kWinBatch; ++u) {
                    const int t = tb + u * kWarps;
                    const int64_t row = ((int64_t)hb * a.res_len + (t < w1 ? t : tb)) * 128 + 4 * lane;
                    if (t == t_new) {
                        kr[u] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.k_new) + (int64_t)hb * 128 + 4 * lane));
                        vr[u] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.v_new) + (int64_t)hb * 128 + 4 * lane));
                        if (sub == 0) {
                            *reinterpret_cast<uint2*>(reinterpret_cast<T*>(const_cast<void*>(a.k_res)) + row) = kr[u];
                            *reinterpret_cast<uint2*>(reinterpret_cast<T*>(const_cast<void*>(a.v_res)) + row) = vr[u];
                        }
                    } else {
                        kr[u] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.k_res) + row));
                        vr[u] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(a.v_res) + row));
                    }
                }
                float sg[kWinBatch][G];
#pragma unroll
                for (int u = 0; u < kWinBatch; ++u) {
                    const float2 k01 = io<T>::to_f2(kr[u].x), k23 = io<T>::to_f2(kr[u].y);
#pragma unroll
                    for (int g = 0; g < G; ++g) sg[u][g] = fmaf(qv[g][3], k23.y, fmaf(qv[g][2], k23.x, fmaf(qv[g][1], k01.y, qv[g][0] * k01.x)));
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int u = 0; u < kWinBatch; ++u)
#pragma unroll
                        for (int g = 0; g < G; ++g) sg[u][g] += __shfl_xor_sync(0xffffffffu, sg[u][g], off);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float nm = wm[g];
#pragma unroll
                    for (int u = 0; u < kWinBatch; ++u) {
                        if (tb + u * kWarps >= w1) sg[u][g] = -INFINITY;
                        nm = fmaxf(nm, sg[u][g]);
                    }
                    const float alpha = exp2_fast(wm[g], nm);
                    wl[g] *= alpha;
                    wo[g][0] *= alpha; wo[g][1] *= alpha; wo[g][2] *= alpha; wo[g][3] *= alpha;
#pragma unroll
                    for (int u = 0; u < kWinBatch; ++u) {
                        const float pw = exp2_fast(sg[u][g], nm);
                        const float2 v01 = io<T>::to_f2(vr[u].x), v23 = io<T>::to_f2(vr[u].y);
                        wl[g] += pw;
                        wo[g][0] = fmaf(pw, v01.x, wo[g][0]); wo[g][1] = fmaf(pw, v01.y, wo[g][1]);
                        wo[g][2] = fmaf(pw, v23.x, wo[g][2]); wo[g][3] = fmaf(pw, v23.y, wo[g][3]);
                    }
                    wm[g] = nm;
                }
            }
        }
    }

    dbg_stamp(a, 7, piece);
    __syncthreads();   // every warp is done with its stage buffers and p slots (aliased below)
    dbg_stamp(a, 3, piece);
    // ---------------------------------------------------------------- combine the warps of this CTA -> one partial state
    // 2 * kWarps entries of [G*128 o | G m | G l]: entry w = coded tokens of warp w, entry kWarps + w = its window tokens.
    // Coded layout: slot sl of lane (hw, lq) holds sub-space 4*lq + ((sl + hw) & 3); hw=1 is folded into hw=0 first.
    constexpr int kEntry = (G * 130 + 3) & ~3;   // 16-byte aligned entries
    {
        const int lq = lane & 15, hw = lane >> 4;
#pragma unroll
        for (int g = 0; g < G; ++g) st.l[g] = warp_sum(st.l[g]);
        float* wx = xch + warp * kEntry;
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    // partner (hw=1) slot (sl - 1) & 3 holds the same sub-space as my (hw=0) slot sl
                    const float theirs = __shfl_xor_sync(0xffffffffu, st.o[(sl + 3) & 3][g][k], 16);
                    if (hw == 0) {
                        float v = st.o[sl][g][k] + theirs;
                        if constexpr ((OUT & 2) != 0) { if (has_codes) v += vo_acc[g * 128 + 2 * (4 * lq + sl) + k]; }
                        wx[g * 128 + 2 * (4 * lq + sl) + k] = v;
                    }
                }
        float* ww = xch + (kWarps + warp) * kEntry;
#pragma unroll
        for (int g = 0; g < G; ++g)
            *reinterpret_cast<float4*>(ww + g * 128 + 4 * lane) = make_float4(wo[g][0], wo[g][1], wo[g][2], wo[g][3]);
        if (lane == 0) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                wx[G * 128 + g] = st.m[g]; wx[G * 128 + G + g] = st.l[g];
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
    dbg_stamp(a, 4, piece);
    // ---------------------------------------------------------------- last CTA of the (b, hk) group merges
    const bool last = last_cta_of_group(a.counters, hb, np * gsub, flag);
    dbg_stamp(a, 5, piece);
    if (last) merge_group<T, P2P != 0>(a, b, hk, np, xch, reinterpret_cast<float*>(lut_p), (LutCfg<G>::bytes + kVtabBytes) / 4,   // LUT + V table: both dead, adjacent
                            
                             reinterpret_cast<unsigned long long*>(smem + kMiscOff + 160), piece);   // the K LUT is dead by now
    dbg_stamp(a, 6, piece);
}

// VL = 0: value codes row-major (tokens x 64 bytes); VL = 1: transposed per sub-space (paged pool or (M, ld) rows)
template <typename T, int G, int VL, int OUT, int P2P>
__global__ void __launch_bounds__(fast::kThreads, 1) attn_fast_kernel(const AttnArgs a, const uint32_t* __restrict__ prepared, const int gsub) {
    pdl_launch_dependents();   // a PDL successor may take the SMs this grid leaves; it waits for our completion before it reads q
    if (!a.flat) {
        // grid (splits, kv heads x 4-head sub-groups, batch): one segment per CTA
        int t0, t1;
        split_range(a, blockIdx.x, t0, t1);
        attn_fast_segment<T, G, VL, OUT, P2P>(a, prepared, gsub, blockIdx.x, blockIdx.y / gsub, blockIdx.y % gsub, blockIdx.z, t0, t1, a.n_splits, 0);
        return;
    }
    // Flat scheduling: the (group, 64-token unit) space is cut into runs of equal COST, one per CTA, so every SM gets the
    // same amount of work whatever bs*nh_k is (64 groups on 148 SMs would otherwise use 128).  A run may cross a group
    // boundary: the CTA then works the two pieces one after the other (LUT rebuilt), which costs a second prologue and
    // epilogue — every group is therefore preceded by kFlatPad virtual units in the space that is cut, so a CTA that
    // starts a new group inside its run gets that many fewer real units.  Part index inside a group = CTA index minus
    // the group's first CTA; everything is recomputed from (flat_per, flat_ug).
    const int ug = a.flat_ug, per = a.flat_per, vg = ug + fast::kFlatPad;
    const long long total = (long long)a.bs * a.nh_k * vg;
    const long long run0 = (long long)blockIdx.x * per;
    const long long run1 = (run0 + per < total) ? run0 + per : total;
    int piece = 0;
    for (int grp = (int)(run0 / vg); grp < a.bs * a.nh_k && (long long)grp * vg < run1; ++grp) {
        const long long real0 = (long long)grp * vg + fast::kFlatPad, real1 = real0 + ug;   // this group's real units in the cut space
        const long long s0 = run0 > real0 ? run0 : real0, s1 = run1 < real1 ? run1 : real1;
        if (s1 <= s0) continue;
        const int first = (int)(real0 / per), last = (int)((real1 - 1) / per);
        const int us = (int)(s0 - real0), ue = (int)(s1 - real0);
        const int t1 = (ue * 64 < a.nk) ? ue * 64 : a.nk;
        attn_fast_segment<T, G, VL, OUT, P2P>(a, prepared, 1, (int)blockIdx.x - first, grp % a.nh_k, 0, grp / a.nh_k, us * 64, t1, last - first + 1, piece++);
        __syncthreads();   // the next piece reuses every shared buffer
    }
}

// ------------------------------------------------------------------------------------------------ launcher
template <typename T, int G, int VL, int OUT, int P2P = 0>
static int launch_fast_t(const AttnArgs& a, const uint32_t* prepared, int gsub, cudaStream_t stream) {
    using namespace fast;
    static_assert((OUT & 2) == 0 || G <= 2, "V-side outlier accumulators need shared memory the G = 4 kernel does not have");
    const size_t smem = LutCfg<G>::bytes + kVtabBytes + kWarps * kStageBytes + kWarps * kTile * 8 + (OUT ? 1024 : 256) + ((OUT & 2) ? kWarps * G * 128 * 4 : 0);
    static_assert(kWarps * kStageBytes + kWarps * kTile * 8 >= 2 * kWarps * 4 * 130 * sizeof(float), "stage + p area too small for the combine");
    static_assert(kWarps * kStageBytes + kWarps * kTile * 8 >= kMergeScratch * sizeof(float), "stage area too small for the merge scratch");
    static SmemAttrOnce configured = {};
    MILLION_CUDA_OK(ensure_dynamic_smem(configured, attn_fast_kernel<T, G, VL, OUT, P2P>, smem));
    dim3 grid(a.n_splits, a.nh_k * gsub, a.bs), block(kThreads);
    if (a.flat) grid = dim3((unsigned)(((long long)a.bs * a.nh_k * (a.flat_ug + kFlatPad) + a.flat_per - 1) / a.flat_per), 1, 1);
    MILLION_CUDA_OK(launch_kernel(attn_fast_kernel<T, G, VL, OUT, P2P>, grid, block, smem, stream, a.pdl != 0, a, prepared, gsub));
    return MILLION_OK;
}

int launch_attn_fast_dm4(const AttnArgs& a, int io_dtype, int G, int gsub, const void* prepared, cudaStream_t stream);

int launch_attn_fast(const AttnArgs& a_in, int io_dtype, const void* prepared, cudaStream_t stream, bool probe_only) {
    AttnArgs a = a_in;
    a.n_parts = a.n_splits;   // the window is dealt out to the splits, no extra part
    a.flat = 0;
    const int Gfull = a.nh / a.nh_k;
    if (a.d != 128 || (a.M != 64 && a.M != 32) || a.C != 256 || a.code_bytes != 1) MILLION_UNSUPPORTED("fast decode attention needs d=128, M in {32, 64}, C=256, one-byte codes");
    const bool dm4 = a.M == 32;
    if (!(Gfull == 1 || Gfull == 2 || Gfull % 4 == 0)) MILLION_UNSUPPORTED("fast decode attention needs nh/nh_k in {1,2,4k}");
    if (a.nk > 0 && a.v_layout == MILLION_V_PAGED && (a.page_size % 32 != 0 || ((uintptr_t)a.v_codes & 15)))
        MILLION_UNSUPPORTED("fast decode attention: paged V needs page_size %% 32 == 0 and an aligned pool");
    if (a.nk > 0 && a.v_layout == MILLION_V_TRANSPOSED && ((a.v_ld & 15) || (a.v_head_stride & 15) || ((uintptr_t)a.v_codes & 15)))
        MILLION_UNSUPPORTED("fast decode attention: transposed V needs 16-byte aligned rows");
    if (a.nk > 0 && a.v_out > 4) MILLION_UNSUPPORTED("fast decode attention: V-side outliers need v_out <= 4");
    if (a.nk > 0 && a.v_out > 0 && a.v_out != 3 &&
        ((uintptr_t)a.vo_idx % a.v_out || a.vo_head_stride % a.v_out || (uintptr_t)a.vo_val % (2 * a.v_out)))
        MILLION_UNSUPPORTED("fast decode attention: the V-side outlier store must be aligned to one token's records");
    if (a.nk > 0 && a.k_out > 4) MILLION_UNSUPPORTED("fast decode attention: K-side outliers need k_out <= 4");
    if (a.nk > 0 && a.k_out > 0 && a.k_out != 3 &&
        ((uintptr_t)a.ko_idx % a.k_out || a.ko_head_stride % a.k_out || (uintptr_t)a.ko_val % (2 * a.k_out)))
        MILLION_UNSUPPORTED("fast decode attention: the K-side outlier store must be aligned to one token's records");
    if (a.p2p && (Gfull != 4 || dm4 || a.v_layout != MILLION_V_ROWMAJOR || a.k_out || a.v_out))
        MILLION_UNSUPPORTED("fused split-KV exchange: compiled for M=64, nh/nh_k = 4, row-major value codes, no side store (use MILLION_ATTN_PARTIAL_ONLY + million_splitkv_push_merge)");
    if (!prepared) MILLION_UNSUPPORTED("fast decode attention needs a prepared codebook (million_pq_codebook_prepare)");
    if (a.nk > 0 && (((uintptr_t)a.k_codes | (uintptr_t)a.k_head_stride) & 15))
        MILLION_UNSUPPORTED("fast decode attention needs 16-byte aligned code caches");
    if (a.nk > 0 && a.v_layout == MILLION_V_ROWMAJOR && (((uintptr_t)a.v_codes | (uintptr_t)a.v_head_stride) & 15))
        MILLION_UNSUPPORTED("fast decode attention needs 16-byte aligned code caches");
    if ((a.r > 0 || a.r_dev) && (((uintptr_t)a.k_res | (uintptr_t)a.v_res | (uintptr_t)a.k_new | (uintptr_t)a.v_new) & 7))
        MILLION_UNSUPPORTED("fast decode attention needs an 8-byte aligned window");
    if (probe_only) return MILLION_OK;
    int G = Gfull >= 4 ? 4 : Gfull, gsub = Gfull >= 4 ? Gfull / 4 : 1;
    if (!dm4 && a.nk > 0 && a.v_out > 0 && G == 4) {
        // V-side outlier records need 2 KB of shared-memory accumulators per warp, which the 4-heads-per-CTA kernel (128 KB of K
        // LUT) does not have: such calls run as 2-head sub-groups (64 KB LUT) — the codes are streamed twice, still two orders of
        // magnitude faster than the all-shapes kernel
        G = 2;
        gsub = Gfull / 2;
    }
    if (a.auto_splits && gsub == 1 && a.nk > 0) {
        // flat scheduling pays when a CTA's share of tokens dwarfs the (up to two) prologues it runs
        int sms = sm_count();
        if (sms <= 0) sms = 148;
        const long long groups = (long long)a.bs * a.nh_k;
        const int ug = (a.nk + 63) / 64;
        const long long total = groups * (ug + fast::kFlatPad);            // cost space: see attn_fast_kernel
        const long long ncta = groups * ug < sms ? groups * ug : sms;
        const int per = (int)((total + ncta - 1) / ncta);
        const int max_parts = (ug + per - 1) / per + 1;
        const double P = fast::kFlatPad * 64.0;   // prologue + epilogue of one piece, in tokens
        const long long ctas = groups * a.n_splits;
        const double split_cost = (double)((ctas + sms - 1) / sms) * ((double)a.units_per_split * 16 + P);
        const double flat_cost = (double)per * 64 + P;   // every CTA runs at least one prologue/epilogue
        if (flat_cost < 0.93 * split_cost && max_parts <= a.ws_parts && max_parts <= 1024) {
            a.flat = 1; a.flat_per = per; a.flat_ug = ug; a.n_parts = max_parts;
        }
    }
    if (dm4) return launch_attn_fast_dm4(a, io_dtype, G, gsub, prepared, stream);
    const uint32_t* prep = reinterpret_cast<const uint32_t*>(prepared);
    if (a.p2p) {   // checked above: G = 4 (one 4-head sub-group), row-major V, no side store
        return io_dtype == MILLION_F16 ? launch_fast_t<__half, 4, 0, 0, 1>(a, prep, gsub, stream) : launch_fast_t<__nv_bfloat16, 4, 0, 0, 1>(a, prep, gsub, stream);
    }
    const int kv = (a.nk > 0 && a.k_out ? 1 : 0) | (a.nk > 0 && a.v_out ? 2 : 0);
#define MILLION_FAST_CASE(TT, GG) \
    return a.v_layout == MILLION_V_ROWMAJOR ? (kv & 1 ? launch_fast_t<TT, GG, 0, 1>(a, prep, gsub, stream) : launch_fast_t<TT, GG, 0, 0>(a, prep, gsub, stream)) \
                                            : (kv & 1 ? launch_fast_t<TT, GG, 1, 1>(a, prep, gsub, stream) : launch_fast_t<TT, GG, 1, 0>(a, prep, gsub, stream))
    // V-side records (alone or with K-side ones): G <= 2 only
#define MILLION_FAST_CASE_V(TT, GG) \
    return a.v_layout == MILLION_V_ROWMAJOR ? (kv == 3 ? launch_fast_t<TT, GG, 0, 3>(a, prep, gsub, stream) : launch_fast_t<TT, GG, 0, 2>(a, prep, gsub, stream)) \
                                            : (kv == 3 ? launch_fast_t<TT, GG, 1, 3>(a, prep, gsub, stream) : launch_fast_t<TT, GG, 1, 2>(a, prep, gsub, stream))
    if (io_dtype == MILLION_F16) {
        if (kv & 2) { if (G == 2) MILLION_FAST_CASE_V(__half, 2); MILLION_FAST_CASE_V(__half, 1); }
        if (G == 4) MILLION_FAST_CASE(__half, 4);
        if (G == 2) MILLION_FAST_CASE(__half, 2);
        MILLION_FAST_CASE(__half, 1);
    } else {
        if (kv & 2) { if (G == 2) MILLION_FAST_CASE_V(__nv_bfloat16, 2); MILLION_FAST_CASE_V(__nv_bfloat16, 1); }
        if (G == 4) MILLION_FAST_CASE(__nv_bfloat16, 4);
        if (G == 2) MILLION_FAST_CASE(__nv_bfloat16, 2);
        MILLION_FAST_CASE(__nv_bfloat16, 1);
    }
#undef MILLION_FAST_CASE
#undef MILLION_FAST_CASE_V
}

}  // namespace million
