// encode_tc.cu — PQ encoder whose distance contraction runs on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces sa_encode_4d_keops (scripts/utils/pq_utils.py:451-499): code[t, m] = argmin_c sum_k (x[t,m,k] - C[m,c,k])^2.
//
// For one sub-space m the 128 x 256 distance tile of a 128-token block is ONE tcgen05.mma (M=128, N=256, K=16):
//     A[t]  = [ x_0 .. x_{dm-1},  1,     1,      1,     0 .. ]   (16 bytes per token, written by CUDA cores)
//     B[c]  = [ -2c_0 .. -2c_{dm-1}, n_hi, n_mid, n_lo, 0 .. ]   (16 bytes per centroid, resident in shared memory)
//     D[t][c] = |c|^2 - 2 x.c   (fp32 in TMEM; |c|^2 split into three terms of the input format)
// The second K chunk (k = 8..15) of both operands is a shared block of zeros reached through the descriptor's
// leading-dimension offset, so an operand row really is 16 bytes.  D lives in TMEM (2 x 256 columns, double buffered:
// the MMA of item i+1 runs while the CUDA cores reduce item i).
//
// The epilogue reads D with tcgen05.ld and does NOT trust it for the final answer: |c|^2 - 2x.c cancels badly when x
// is close to c, and the reference's arithmetic is the direct (x-c)^2 in fp32.  D is used as a FILTER: every centroid
// within eps = 2^-17 (|x|^2 + max|c|^2) of the row minimum is a candidate (min via FMNMX3, candidate bits via
// FADD + funnel shift); one candidate -> that is the code; several (about 1e-4 of the cases) or none -> the exact fp32
// formula of the oracle is evaluated over the candidates, first minimum wins.  Codes are therefore bit-identical to
// the generic encoder and the oracle, ties included.
//
// Shapes: C = 256, d/M in {2, 4}, x in fp16/bf16, centroids representable in x's format (checked by the prepare step).
#include <type_traits>

#include "codec.cuh"

namespace million {
namespace tc {

constexpr int kThreads = 256;
constexpr int kTokTile = 128;
constexpr int kMG = 16;                    // sub-spaces per CTA: 16 B tiles of 4 KB stay resident
constexpr int kBTileBytes = 256 * 16;      // 4 KB per sub-space
constexpr int kATileBytes = kTokTile * 16; // 2 KB per item
constexpr int kZeroBytes = 4096;
constexpr float kEpsScale = 7.62939453125e-06f;   // 2^-17

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / tcgen05 wrappers (PTX ISA 8.6+, sm_100a)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] = A[smem] * B[smem]^T, one K=16 step, no accumulation
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, no swizzle: 8-row x 16-byte core matrices; SBO = bytes between row blocks, LBO = bytes between K chunks
__device__ __forceinline__ uint64_t make_desc(uint32_t start, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((start >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version for sm_100
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// 64 consecutive fp32 columns of my TMEM lane
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
          "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
          "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// row minimum of 64 values and, relative to thr = min(best_so_far, that minimum) + eps, the candidate bits
// (bit 31 of word 0 = first column).  Returns the chunk minimum.  Four independent min chains and eight independent
// bit-collection chains (8 columns each): a single 32-long dependent chain per word made the epilogue latency bound.
__device__ __forceinline__ float reduce_chunk(const float (&v)[64], float best_so_far, float eps, uint32_t& w0, uint32_t& w1) {
    float m0 = v[0], m1 = v[16], m2 = v[32], m3 = v[48];
#pragma unroll
    for (int i = 1; i + 1 < 16; i += 2) {
        m0 = min3(m0, v[i], v[i + 1]);
        m1 = min3(m1, v[16 + i], v[17 + i]);
        m2 = min3(m2, v[32 + i], v[33 + i]);
        m3 = min3(m3, v[48 + i], v[49 + i]);
    }
    const float m = fminf(min3(m0, v[15], m1), min3(min3(m2, v[31], v[47]), m3, v[63]));
    const float thr = fminf(best_so_far, m) + eps;
    uint32_t q[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = __funnelshift_l(__float_as_uint(v[8 * c + i] - thr), q[c], 1);   // sign(v - thr) = 1 <=> v < thr
    }
    w0 = (q[0] << 24) | (q[1] << 16) | (q[2] << 8) | q[3];
    w1 = (q[4] << 24) | (q[5] << 16) | (q[6] << 8) | q[7];
    return m;
}

struct EncArgs {
    const void* x;
    int64_t x_head_stride;
    const unsigned char* prep;   // B tiles (M * 4 KB) | cmax2 (M floats) | flag
    const float* cent;           // (M, 256, DM) fp32 — exact re-check
    CodeDst dst;
    int n_heads, n_tokens, d, M, tiles_per_head, total_tiles;
};

// exact oracle arithmetic for one centroid
template <int DM>
__device__ __forceinline__ float exact_dist(const float (&xv)[DM], const float* c) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < DM; ++k) {
        const float diff = __fsub_rn(xv[k], c[k]);
        const float sq = __fmul_rn(diff, diff);
        acc = (k == 0) ? sq : __fadd_rn(acc, sq);
    }
    return acc;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Two warpgroups (4 warps = 128 threads = the 128 token rows of a tile); warpgroup g takes the sub-spaces j = g, g+2, ...
// of every tile.  An item (tile, j) is issued as TWO N=128 MMAs into the two 128-column halves of the group's TMEM
// region, each with its own mbarrier: while the CUDA cores reduce one half, the tensor core already works on the other
// half or on the next item, so the MMA latency is off the critical path.
template <typename T, int DM>
__global__ void __launch_bounds__(kThreads, 1) encode_tc_kernel(const EncArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Bs = smem;                                   // kMG * 4 KB
    unsigned char* As = Bs + kMG * kBTileBytes;                 // 2 warpgroups * 2 slots * 2 KB
    unsigned char* Zs = As + 4 * kATileBytes;                   // 4 KB of zeros (second K chunk of both operands)
    uint64_t* bars = reinterpret_cast<uint64_t*>(Zs + kZeroBytes);   // 2 warpgroups * 2 halves
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    float* cmax_s = reinterpret_cast<float*>(tmem_slot + 2);    // kMG floats
    uint64_t* bar_b = reinterpret_cast<uint64_t*>(cmax_s + kMG);    // arrival of the B tiles (one TMA bulk copy)

    const int tid = threadIdx.x, warp = tid >> 5;
    const int wg = warp >> 2, row = tid & 127;                  // warpgroup, token row inside the tile
    const int m0 = blockIdx.y * kMG;
    constexpr uint32_t kOne = std::is_same<T, __half>::value ? 0x3C00u : 0x3F80u;
    constexpr uint32_t kFmt = std::is_same<T, __half>::value ? 0u : 1u;
    // instruction descriptor: D fp32, A/B format, both K-major, N = 128, M = 128
    constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

    // ---- one-time setup
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    if (tid == 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
        mbar_init(smem_u32(bar_b), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // The B operand — this CTA's kMG codebook tiles [-2c | |c|^2 split], 64 KB contiguous in the prepared buffer — comes in
        // with ONE TMA bulk copy (cp.async.bulk, async proxy end to end: no generic-proxy stores, no proxy fence for B).  The
        // tensor-map form of TMA has nothing to add for a contiguous block; the A rows cannot be copied at all, they are
        // synthesised ([x, 1, 1, 1, 0..] per token).
        constexpr uint32_t kBBytes = kMG * kBTileBytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar_b)), "r"(kBBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(Bs)),
                     "l"(a.prep + (int64_t)m0 * kBTileBytes), "r"(kBBytes), "r"(smem_u32(bar_b))
                     : "memory");
    }
    {
        uint4* z = reinterpret_cast<uint4*>(Zs);
        for (int i = tid; i < kZeroBytes / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
        if (tid < kMG) cmax_s[tid] = reinterpret_cast<const float*>(a.prep + (int64_t)a.M * kBTileBytes)[m0 + tid];
    }
    fence_async_proxy();        // generic-proxy writes (zeros) -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(smem_u32(bar_b), 0);   // the B tiles have landed
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t zs = smem_u32(Zs);
    const uint32_t bar0 = smem_u32(&bars[wg * 2]), bar1 = bar0 + 8;
    const uint32_t dcol = tmem_base + wg * 256;
    const uint32_t taddr = dcol + ((uint32_t)((warp & 3) * 32) << 16);   // my TMEM lane quarter
    unsigned char* Aw = As + wg * 2 * kATileBytes;

    const int n_my_tiles = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_items = n_my_tiles * (kMG / 2);

    // state of the item being built / issued next
    int nx_tile = blockIdx.x, nx_jj = 0, nx_head = 0, nx_tok = 0;
    bool nx_live = false;
    uint32_t nx_xw[DM / 2];
    auto locate = [&]() {     // (head, token) of my row in tile nx_tile — once per tile, the only integer divisions
        nx_head = nx_tile / a.tiles_per_head;
        nx_tok = (nx_tile - nx_head * a.tiles_per_head) * kTokTile + row;
        nx_live = nx_tile < a.total_tiles && nx_tok < a.n_tokens;
    };
    auto fetch_x = [&]() {    // my token's values for sub-space m0 + wg + 2*nx_jj
        const uint32_t* xp = reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(a.x) + nx_head * a.x_head_stride +
                                                               (int64_t)(nx_live ? nx_tok : 0) * a.d + (m0 + wg + 2 * nx_jj) * DM);
#pragma unroll
        for (int i = 0; i < DM / 2; ++i) nx_xw[i] = nx_live ? __ldg(xp + i) : 0u;
    };
    auto build_and_issue_half0 = [&](int slot) {   // A row -> smem slot, group barrier, MMA half 0
        uint4 rowv;
        if constexpr (DM == 2) rowv = make_uint4(nx_xw[0], kOne | (kOne << 16), kOne, 0u);
        else rowv = make_uint4(nx_xw[0], nx_xw[1], kOne | (kOne << 16), kOne);
        *reinterpret_cast<uint4*>(Aw + slot * kATileBytes + row * 16) = rowv;
        fence_async_proxy();
        tc_fence_before();
        named_bar_sync(1 + wg, 128);
        tc_fence_after();
        if (row == 0) {
            const uint32_t as = smem_u32(Aw + slot * kATileBytes), bs = smem_u32(Bs + (wg + 2 * nx_jj) * kBTileBytes);
            umma_f16(dcol, make_desc(as, zs - as, 128), make_desc(bs, zs - bs, 128), idesc);
            umma_commit(bar0);
        }
    };
    auto issue_half1 = [&](int slot) {
        tc_fence_before();
        named_bar_sync(1 + wg, 128);
        tc_fence_after();
        if (row == 0) {
            const uint32_t as = smem_u32(Aw + slot * kATileBytes), bs = smem_u32(Bs + (wg + 2 * nx_jj) * kBTileBytes + 128 * 16);
            umma_f16(dcol + 128, make_desc(as, zs - as, 128), make_desc(bs, zs - bs, 128), idesc);
            umma_commit(bar1);
        }
    };
    auto advance = [&]() {
        if (++nx_jj == kMG / 2) { nx_jj = 0; nx_tile += gridDim.x; locate(); }
    };

    if (n_items > 0) {
        locate();
        fetch_x();
        build_and_issue_half0(0);
        issue_half1(0);
    }
    for (int i = 0; i < n_items; ++i) {
        // ---- the item whose MMAs are in flight becomes the current one
        const int head = nx_head, tok = nx_tok, j = wg + 2 * nx_jj;
        const bool live = nx_live;
        float xv[DM];
        {
            const float2 f0 = io<T>::to_f2(nx_xw[0]);
            xv[0] = f0.x; xv[1] = f0.y;
            if constexpr (DM == 4) { const float2 f1 = io<T>::to_f2(nx_xw[1]); xv[2] = f1.x; xv[3] = f1.y; }
        }
        float x2 = 0.f;
#pragma unroll
        for (int k = 0; k < DM; ++k) x2 = fmaf(xv[k], xv[k], x2);
        const float eps = (x2 + cmax_s[j]) * kEpsScale;
        const bool has_next = i + 1 < n_items;
        if (has_next) { advance(); fetch_x(); }      // operands of the next item are requested early

        uint32_t cw[8];
        float mn[4];
        // ---- half 0: columns 0..127
        mbar_wait(bar0, i & 1);
        tc_fence_after();
        float best = INFINITY;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float v[64];
            tmem_ld64(taddr + ch * 64, v);
            mn[ch] = reduce_chunk(v, best, eps, cw[2 * ch], cw[2 * ch + 1]);
            best = fminf(best, mn[ch]);
        }
        if (has_next) build_and_issue_half0((i + 1) & 1);   // buffer 0 is free again: next item's first half starts now
        // ---- half 1: columns 128..255
        mbar_wait(bar1, i & 1);
        tc_fence_after();
#pragma unroll
        for (int ch = 2; ch < 4; ++ch) {
            float v[64];
            tmem_ld64(taddr + ch * 64, v);
            mn[ch] = reduce_chunk(v, best, eps, cw[2 * ch], cw[2 * ch + 1]);
            best = fminf(best, mn[ch]);
        }
        if (has_next) issue_half1((i + 1) & 1);

        int cnt = 0, code = 0;
#pragma unroll
        for (int ch = 3; ch >= 0; --ch) {
            if (mn[ch] > best + eps) { cw[2 * ch] = 0; cw[2 * ch + 1] = 0; }     // filtered against a looser threshold
            cnt += __popc(cw[2 * ch]) + __popc(cw[2 * ch + 1]);
            if (cw[2 * ch + 1]) code = 64 * ch + 32 + __clz(cw[2 * ch + 1]);
            if (cw[2 * ch]) code = 64 * ch + __clz(cw[2 * ch]);                  // lowest candidate column wins
        }
        if (cnt != 1 && live) {
            // rare: several centroids inside the filter width (or a degenerate all-zero row): exact fp32 arg-min
            const float* cm = a.cent + (int64_t)(m0 + j) * 256 * DM;
            float bestd = INFINITY;
            int bi = 0;
            for (int w = 0; w < 8; ++w) {
                uint32_t bits = cnt == 0 ? 0xffffffffu : cw[w];
                while (bits) {
                    const int lz = __clz(bits);
                    bits &= ~(0x80000000u >> lz);
                    const int c = 32 * w + lz;
                    float cv[DM];
#pragma unroll
                    for (int k = 0; k < DM; ++k) cv[k] = __ldg(cm + c * DM + k);
                    const float dd = exact_dist<DM>(xv, cv);
                    if (dd < bestd) { bestd = dd; bi = c; }
                }
            }
            code = bi;
        }
        if (live) a.dst.put(head, tok, m0 + j, code);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ prepare
// B tiles: row c of sub-space m = [-2c_0 .. -2c_{dm-1}, n_hi, n_mid, n_lo, 0 ..] in T; cmax2[m] = max_c |c|^2; flag = 0 when a
// centroid component (or its double) is not representable in T.
template <typename T, int DM>
__global__ void encode_prepare_kernel(const float* __restrict__ cent, unsigned char* __restrict__ out, int M) {
    const int m = blockIdx.x, c = threadIdx.x;   // 256 threads
    __shared__ float red[256];
    const float* cv = cent + ((int64_t)m * 256 + c) * DM;
    T row[8];
    double n2 = 0.0;
    bool exact = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) row[k] = io<T>::from_f(0.f);
#pragma unroll
    for (int k = 0; k < DM; ++k) {
        const float v = cv[k];
        row[k] = io<T>::from_f(-2.f * v);
        exact = exact && (io<T>::to_f(row[k]) == -2.f * v) && (io<T>::to_f(io<T>::from_f(v)) == v);
        n2 += (double)v * (double)v;
    }
    const T hi = io<T>::from_f((float)n2);
    const double r1 = n2 - (double)io<T>::to_f(hi);
    const T mid = io<T>::from_f((float)r1);
    const double r2 = r1 - (double)io<T>::to_f(mid);
    const T lo = io<T>::from_f((float)r2);
    row[DM] = hi; row[DM + 1] = mid; row[DM + 2] = lo;
    *reinterpret_cast<uint4*>(out + ((int64_t)m * 256 + c) * 16) = *reinterpret_cast<const uint4*>(row);
    red[c] = (float)n2;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (c < s) red[c] = fmaxf(red[c], red[c + s]);
        __syncthreads();
    }
    if (c == 0) reinterpret_cast<float*>(out + (int64_t)M * tc::kBTileBytes)[m] = red[0];
    if (!exact || !isfinite((float)n2)) atomicExch(reinterpret_cast<int*>(out + (int64_t)M * tc::kBTileBytes + (int64_t)M * 4), 0);
}

}  // namespace tc

int64_t encode_tc_prepared_bytes(int d, int M, int C) {
    const int dm = M > 0 ? d / M : 0;
    if (C != 256 || !(dm == 2 || dm == 4) || M % tc::kMG) return 0;
    return (int64_t)M * tc::kBTileBytes + (int64_t)M * 4 + 16;
}

int launch_encode_tc_prepare(const float* cent, int x_dtype, int d, int M, int C, void* out, cudaStream_t stream) {
    if (encode_tc_prepared_bytes(d, M, C) == 0 || !(x_dtype == MILLION_F16 || x_dtype == MILLION_BF16))
        MILLION_UNSUPPORTED("tensor-core encoder: needs C=256, d/M in {2,4}, M %% 16 == 0, fp16/bf16 input");
    const int dm = d / M;
    int one = 1;
    MILLION_CUDA_OK(cudaMemcpyAsync((char*)out + (int64_t)M * tc::kBTileBytes + (int64_t)M * 4, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
    unsigned char* o = (unsigned char*)out;
    if (x_dtype == MILLION_F16) {
        if (dm == 2) tc::encode_prepare_kernel<__half, 2><<<M, 256, 0, stream>>>(cent, o, M);
        else tc::encode_prepare_kernel<__half, 4><<<M, 256, 0, stream>>>(cent, o, M);
    } else {
        if (dm == 2) tc::encode_prepare_kernel<__nv_bfloat16, 2><<<M, 256, 0, stream>>>(cent, o, M);
        else tc::encode_prepare_kernel<__nv_bfloat16, 4><<<M, 256, 0, stream>>>(cent, o, M);
    }
    MILLION_CUDA_OK(cudaGetLastError());
    // one-time, synchronous: are the centroids representable in the input format (else the filter bound does not hold)?
    int flag = 0;
    MILLION_CUDA_OK(cudaMemcpyAsync(&flag, (char*)out + (int64_t)M * tc::kBTileBytes + (int64_t)M * 4, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MILLION_CUDA_OK(cudaStreamSynchronize(stream));
    if (!flag) MILLION_UNSUPPORTED("tensor-core encoder: centroids are not exactly representable in the input format");
    return MILLION_OK;
}

template <typename T, int DM>
static int launch_tc_t(const tc::EncArgs& a, cudaStream_t stream) {
    const size_t smem = tc::kMG * tc::kBTileBytes + 4 * tc::kATileBytes + tc::kZeroBytes + 32 + 16 + tc::kMG * 4 + 8 + 64;
    static SmemAttrOnce configured = {};
    MILLION_CUDA_OK(ensure_dynamic_smem(configured, tc::encode_tc_kernel<T, DM>, smem));
    const int ygroups = a.M / tc::kMG;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    int gx = sms / ygroups;
    if (gx < 1) gx = 1;
    if (gx > a.total_tiles) gx = a.total_tiles;
    dim3 grid(gx, ygroups), block(tc::kThreads);
    tc::encode_tc_kernel<T, DM><<<grid, block, smem, stream>>>(a);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

int launch_encode_tc(const void* x, int x_dtype, int64_t xhs, const float* cent, const void* prepared, const CodeDst& dst,
                     int n_heads, int n_tokens, int d, int M, int C, cudaStream_t stream, bool probe_only) {
    if (encode_tc_prepared_bytes(d, M, C) == 0) MILLION_UNSUPPORTED("tensor-core encoder: needs C=256, d/M in {2,4}, M %% 16 == 0");
    if (!(x_dtype == MILLION_F16 || x_dtype == MILLION_BF16)) MILLION_UNSUPPORTED("tensor-core encoder: fp16/bf16 input only");
    if (!prepared) MILLION_UNSUPPORTED("tensor-core encoder needs a prepared codebook (million_pq_encoder_prepare)");
    if (((uintptr_t)x & 15) || ((xhs * 2) & 15) || ((d * 2) & 15)) MILLION_UNSUPPORTED("tensor-core encoder needs 16-byte aligned rows");
    if (dst.code_bytes != 1) MILLION_UNSUPPORTED("tensor-core encoder writes 1-byte codes");
    if (probe_only) return MILLION_OK;
    tc::EncArgs a;
    a.x = x; a.x_head_stride = xhs; a.prep = (const unsigned char*)prepared; a.cent = cent; a.dst = dst;
    a.n_heads = n_heads; a.n_tokens = n_tokens; a.d = d; a.M = M;
    a.tiles_per_head = (n_tokens + tc::kTokTile - 1) / tc::kTokTile;
    a.total_tiles = a.tiles_per_head * n_heads;
    const int dm = d / M;
    if (x_dtype == MILLION_F16) return dm == 2 ? launch_tc_t<__half, 2>(a, stream) : launch_tc_t<__half, 4>(a, stream);
    return dm == 2 ? launch_tc_t<__nv_bfloat16, 2>(a, stream) : launch_tc_t<__nv_bfloat16, 4>(a, stream);
}

}  // namespace million
