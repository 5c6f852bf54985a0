// encode_grid.cu — exact PQ encoder for two-dimensional sub-spaces (d/M = 2: the "4-bit" configuration M=64, d=128).
//
// The reference finds codes by brute force: 256 distances per (vector, sub-space) (scripts/utils/pq_utils.py:483-494, pykeops
// arg-min).  The tensor-core encoder (encode_tc.cu) does the same 256 distances with tcgen05 and is bound by reading them back
// from tensor memory.  In TWO dimensions the arg-min needs none of that: a 64 x 64 grid over the centroids' bounding box stores,
// per cell, the (at most 15) centroids that can be nearest to ANY point of the cell; a vector looks up its cell and evaluates
// only those — with the encoder's exact arithmetic (sub, mul, add each rounded to nearest), in ascending centroid order with a
// strict '<', so the code is bit-identical to the brute-force result, ties included.
//
//   candidate rule (conservative): c is listed for cell Q iff  dmin^2(c, Q) <= min_c' dmax^2(c', Q) * (1 + 1e-5)
//     - the true nearest centroid of any p in Q has distance <= min_c' dmax^2(c', Q);
//     - fp32 evaluation perturbs a squared distance by < 4e-7 relative, far inside the 1e-5 slack;
//     - cells are inflated by 1e-3 of their size, 100x the fp32 error of the cell-index computation.
//   points outside the (25 % enlarged) bounding box, NaNs, and cells with more than 15 candidates take the full 256 scan.
//
// Work split: a CTA keeps the tables of TWO sub-spaces in shared memory (2 x 64 KB) and streams a slice of the vectors; the 32
// sub-space groups of one slice run side by side, so every 32-byte sector of the input is read from HBM once and from L2 by the
// four groups that share it.
#include <type_traits>

#include "codec.cuh"

namespace million {

constexpr int kGrid = 64;                         // cells per axis
constexpr int kGridCells = kGrid * kGrid;
constexpr int kGridHdr = 32;                      // bytes: lo_x, lo_y, inv_x, inv_y, hi_x, hi_y, 0, 0
constexpr int kGridStride = kGridHdr + kGridCells * 16;
#ifndef MILLION_GRID_BATCH
#define MILLION_GRID_BATCH 3      // candidates evaluated per shared-memory round trip (A/B: 3 -> 191 us, 4 -> 196, 8 -> 198, one by one -> 199; profiles/r02_ab_variants.txt)
#endif
#ifndef MILLION_GRID_ILP
#define MILLION_GRID_ILP 1
#endif
#ifndef MILLION_GRID_THREADS
#define MILLION_GRID_THREADS 1024   // 32 warps per SM hide the dependent table loads better than 16 (217 -> 199 us per 8 x 32768 vectors)
#endif
constexpr int kGridThreads = MILLION_GRID_THREADS;

int64_t encode_grid_prepared_bytes(int d, int M, int C) { return (M > 0 && d == 2 * M && C >= 2 && C <= 256 && M % 2 == 0) ? (int64_t)M * kGridStride : 0; }

__global__ void __launch_bounds__(256) encode_grid_prepare_kernel(const float* __restrict__ cent, unsigned char* __restrict__ out, int C) {
    __shared__ float cx[256], cy[256];
    __shared__ float red[4][8];
    __shared__ float hdr[8];
    const int m = blockIdx.x, tid = threadIdx.x;
    if (tid < C) { cx[tid] = cent[((int64_t)m * C + tid) * 2]; cy[tid] = cent[((int64_t)m * C + tid) * 2 + 1]; }
    __syncthreads();
    float v[4] = {tid < C ? cx[tid] : INFINITY, tid < C ? -cx[tid] : INFINITY, tid < C ? cy[tid] : INFINITY, tid < C ? -cy[tid] : INFINITY};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] = fminf(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
        if ((tid & 31) == 0) red[k][tid >> 5] = v[k];
    }
    __syncthreads();
    if (tid == 0) {
        float mn[4];
        for (int k = 0; k < 4; ++k) { mn[k] = red[k][0]; for (int w = 1; w < 8; ++w) mn[k] = fminf(mn[k], red[k][w]); }
        float x0 = mn[0], x1 = -mn[1], y0 = mn[2], y1 = -mn[3];
        float wx = x1 - x0, wy = y1 - y0;
        if (!(wx > 0.f)) wx = 1.f;
        if (!(wy > 0.f)) wy = 1.f;
        const float lox = x0 - 0.25f * wx, hix = x1 + 0.25f * wx, loy = y0 - 0.25f * wy, hiy = y1 + 0.25f * wy;
        hdr[0] = lox; hdr[1] = loy; hdr[2] = (float)kGrid / (hix - lox); hdr[3] = (float)kGrid / (hiy - loy);
        hdr[4] = hix; hdr[5] = hiy; hdr[6] = 0.f; hdr[7] = 0.f;
        float* oh = reinterpret_cast<float*>(out + (int64_t)m * kGridStride);
        for (int k = 0; k < 8; ++k) oh[k] = hdr[k];
    }
    __syncthreads();
    const double lox = hdr[0], loy = hdr[1], sx = 1.0 / (double)hdr[2], sy = 1.0 / (double)hdr[3];
    uint4* tab = reinterpret_cast<uint4*>(out + (int64_t)m * kGridStride + kGridHdr);
    for (int cell = tid; cell < kGridCells; cell += blockDim.x) {
        const int ix = cell % kGrid, iy = cell / kGrid;
        const double qx0 = lox + (ix - 1e-3) * sx, qx1 = lox + (ix + 1 + 1e-3) * sx;
        const double qy0 = loy + (iy - 1e-3) * sy, qy1 = loy + (iy + 1 + 1e-3) * sy;
        double ub = 1e300;
        for (int c = 0; c < C; ++c) {
            const double ax = fmax(fabs((double)cx[c] - qx0), fabs((double)cx[c] - qx1));
            const double ay = fmax(fabs((double)cy[c] - qy0), fabs((double)cy[c] - qy1));
            ub = fmin(ub, ax * ax + ay * ay);
        }
        const double lim = ub * (1.0 + 1e-5) + 1e-30;
        // 1. radius test -> R (ascending ids).  2. drop every c in R that some c' in R beats at all four corners of the cell:
        //    g(p) = |p-c|^2 - (1+eps)|p-c'|^2 is concave, so its minimum over the rectangle is at a corner; g > 0 there means c'
        //    is strictly nearer than c (beyond any fp32 rounding) everywhere in the cell.
        constexpr int kMaxR = 96;
        unsigned char R[kMaxR];
        int nr = 0;
        for (int c = 0; c < C; ++c) {
            const double dx = fmax(fmax(qx0 - (double)cx[c], (double)cx[c] - qx1), 0.0);
            const double dy = fmax(fmax(qy0 - (double)cy[c], (double)cy[c] - qy1), 0.0);
            if (dx * dx + dy * dy <= lim) {
                if (nr < kMaxR) R[nr] = (unsigned char)c;
                ++nr;
            }
        }
        unsigned char e[16];
        for (int k = 0; k < 16; ++k) e[k] = 0;
        int cnt = 0;
        if (nr > kMaxR) {
            cnt = 255;
        } else {
            const double px[4] = {qx0, qx1, qx0, qx1}, py[4] = {qy0, qy0, qy1, qy1};
            for (int a = 0; a < nr; ++a) {
                const double ax = cx[R[a]], ay = cy[R[a]];
                bool dominated = false;
                for (int b = 0; b < nr && !dominated; ++b) {
                    if (b == a) continue;
                    const double bx = cx[R[b]], by = cy[R[b]];
                    bool all = true;
                    for (int k = 0; k < 4; ++k) {
                        const double da = (px[k] - ax) * (px[k] - ax) + (py[k] - ay) * (py[k] - ay);
                        const double db = (px[k] - bx) * (px[k] - bx) + (py[k] - by) * (py[k] - by);
                        if (!(da > db * (1.0 + 1e-5) + 1e-30)) { all = false; break; }
                    }
                    dominated = all;
                }
                if (!dominated) {
                    if (cnt < 15) e[1 + cnt] = R[a];
                    ++cnt;
                }
            }
        }
        e[0] = cnt <= 15 ? (unsigned char)cnt : 255;     // 255: too many candidates -> full scan
        uint4 w;
        w.x = e[0] | (e[1] << 8) | (e[2] << 16) | ((uint32_t)e[3] << 24);
        w.y = e[4] | (e[5] << 8) | (e[6] << 16) | ((uint32_t)e[7] << 24);
        w.z = e[8] | (e[9] << 8) | (e[10] << 16) | ((uint32_t)e[11] << 24);
        w.w = e[12] | (e[13] << 8) | (e[14] << 16) | ((uint32_t)e[15] << 24);
        tab[cell] = w;
    }
}

int launch_encode_grid_prepare(const float* cent, int d, int M, int C, void* out, cudaStream_t stream) {
    if (encode_grid_prepared_bytes(d, M, C) == 0) MILLION_UNSUPPORTED("grid encoder needs d/M = 2, even M, C <= 256 (d=%d M=%d C=%d)", d, M, C);
    encode_grid_prepare_kernel<<<M, 256, 0, stream>>>(cent, (unsigned char*)out, C);
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

__device__ __forceinline__ float dist2_exact(float px, float py, float2 c) {
    const float dx = __fsub_rn(px, c.x), dy = __fsub_rn(py, c.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__device__ __forceinline__ int grid_code(float px, float py, const float* __restrict__ hdr, const uint4* __restrict__ tab,
                                         const float2* __restrict__ cs, int C) {
    float best = INFINITY;
    int bi = 0;
    const bool inside = px >= hdr[0] && px < hdr[4] && py >= hdr[1] && py < hdr[5];     // false for NaN
    uint4 e = make_uint4(255u, 0u, 0u, 0u);
    if (inside) {
        int ix = (int)((px - hdr[0]) * hdr[2]), iy = (int)((py - hdr[1]) * hdr[3]);
        ix = min(max(ix, 0), kGrid - 1);
        iy = min(max(iy, 0), kGrid - 1);
        e = tab[iy * kGrid + ix];
    }
    const int cnt = e.x & 0xff;
    if (cnt == 255) {
        for (int c = 0; c < C; ++c) {
            const float dd = dist2_exact(px, py, cs[c]);
            if (dd < best) { best = dd; bi = c; }
        }
        return bi;
    }
    const uint32_t w[4] = {e.x, e.y, e.z, e.w};
#if MILLION_GRID_BATCH > 0
    // Candidates in batches: the centroid loads and the distance arithmetic of a batch are independent of each other (ids past
    // `cnt` are 0: a valid, ignored lookup), only the select chain is serial — one shared-memory round trip per batch instead
    // of one per candidate (the kernel waits on exactly those: stall_short_scoreboard 33 %, stall_wait 29 %).
#pragma unroll
    for (int base = 0; base < 15; base += MILLION_GRID_BATCH) {
        if (base >= cnt) break;
        float dd[MILLION_GRID_BATCH];
        int cc[MILLION_GRID_BATCH];
#pragma unroll
        for (int j = 0; j < MILLION_GRID_BATCH; ++j) {
            const int i = base + 1 + j;
            cc[j] = i < 16 ? (int)((w[(i >> 2) & 3] >> (8 * (i & 3))) & 0xff) : 0;
            dd[j] = dist2_exact(px, py, cs[cc[j]]);
        }
#pragma unroll
        for (int j = 0; j < MILLION_GRID_BATCH; ++j)       // ascending centroid ids: strict '<' keeps the first minimum
            if (base + 1 + j <= cnt && dd[j] < best) { best = dd[j]; bi = cc[j]; }
    }
    return bi;
#else
#pragma unroll
    for (int i = 1; i < 16; ++i) {
        if (i > cnt) break;
        const int c = (w[i >> 2] >> (8 * (i & 3))) & 0xff;       // ascending centroid ids: strict '<' keeps the first minimum
        const float dd = dist2_exact(px, py, cs[c]);
        if (dd < best) { best = dd; bi = c; }
    }
    return bi;
#endif
}

template <typename T>
__global__ void __launch_bounds__(kGridThreads, 1) encode_grid_kernel(const T* __restrict__ x, int64_t x_head_stride, const float* __restrict__ cent,
                                                                      const unsigned char* __restrict__ prepared, CodeDst dst, int n_tokens,
                                                                      long long n_vec, long long per_cta, int d, int C) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint4* tab = reinterpret_cast<uint4*>(smem);                                  // [2][4096]
    float2* cs = reinterpret_cast<float2*>(smem + 2 * kGridCells * 16);           // [2][256]
    float* hdr = reinterpret_cast<float*>(smem + 2 * kGridCells * 16 + 2 * 256 * 8);   // [2][8]
    const int grp = blockIdx.x, tid = threadIdx.x;
    for (int s = 0; s < 2; ++s) {
        const unsigned char* src = prepared + (int64_t)(2 * grp + s) * kGridStride;
        const uint32_t dst_s = (uint32_t)__cvta_generic_to_shared(tab + s * kGridCells);
        for (int i = tid; i < kGridCells; i += kGridThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s + i * 16), "l"(src + kGridHdr + (int64_t)i * 16) : "memory");
        if (tid < 8) hdr[s * 8 + tid] = reinterpret_cast<const float*>(src)[tid];
        for (int c = tid; c < 256; c += kGridThreads)
            cs[s * 256 + c] = c < C ? reinterpret_cast<const float2*>(cent)[(int64_t)(2 * grp + s) * C + c] : make_float2(INFINITY, INFINITY);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long v0 = (long long)blockIdx.y * per_cta, v1 = min(n_vec, v0 + per_cta);
    // each thread walks its vectors with the NEXT one's 4 input elements already in flight (the loads are 256-byte strided across
    // the warp: their L2 latency would otherwise sit in front of every lookup chain)
    using Raw = typename std::conditional<sizeof(T) == 2, uint2, float4>::type;
    // (head, token) of a vector index advance incrementally: a 64-bit division per vector was a third of the loop's instructions
    auto row = [&](int head, int t) { return reinterpret_cast<const Raw*>(x + head * x_head_stride + (int64_t)t * d + 4 * grp); };
    auto advance = [&](int& head, int& t) {
        t += kGridThreads;
        if (t >= n_tokens) { const int q = t / n_tokens; head += q; t -= q * n_tokens; }     // 32-bit, and only when a head boundary is crossed
    };
    auto unpack = [](const Raw& r, float (&p)[4]) {
        if constexpr (sizeof(T) == 2) {
            const float2 a = io<T>::to_f2(r.x), b = io<T>::to_f2(r.y);
            p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y;
        } else {
            p[0] = r.x; p[1] = r.y; p[2] = r.z; p[3] = r.w;
        }
    };
#if MILLION_GRID_ILP == 2
    // two vectors per thread and iteration: four independent lookup chains (cell -> candidate ids -> centroid coordinates are
    // dependent shared-memory loads; with 16-32 warps per SM the kernel waits on them: stall_short_scoreboard 33 %)
    {
        long long va = v0 + tid;
        int ha = (int)(va / n_tokens), ta = (int)(va - (long long)ha * n_tokens);
        int hb = ha, tb = ta;
        advance(hb, tb);
        Raw ra = {}, rb = {};
        if (va < v1) ra = *row(ha, ta);
        if (va + kGridThreads < v1) rb = *row(hb, tb);
        while (va < v1) {
            const long long vb = va + kGridThreads, vna = va + 2 * kGridThreads, vnb = va + 3 * kGridThreads;
            int hna = hb, tna = tb;
            advance(hna, tna);
            int hnb = hna, tnb = tna;
            advance(hnb, tnb);
            Raw na = {}, nb = {};
            if (vna < v1) na = *row(hna, tna);
            if (vnb < v1) nb = *row(hnb, tnb);
            float pa[4], pb[4];
            unpack(ra, pa);
            unpack(rb, pb);
            const int a0 = grid_code(pa[0], pa[1], hdr, tab, cs, C);
            const int b0 = grid_code(pb[0], pb[1], hdr, tab, cs, C);
            const int a1 = grid_code(pa[2], pa[3], hdr + 8, tab + kGridCells, cs + 256, C);
            const int b1 = grid_code(pb[2], pb[3], hdr + 8, tab + kGridCells, cs + 256, C);
            dst.put2(ha, ta, 2 * grp, a0, a1);
            if (vb < v1) dst.put2(hb, tb, 2 * grp, b0, b1);
            ra = na; rb = nb;
            va = vna; ha = hna; ta = tna; hb = hnb; tb = tnb;
        }
        return;
    }
#endif
    long long v = v0 + tid;
    int head = (int)(v / n_tokens), t = (int)(v - (long long)head * n_tokens);
    Raw raw = {};
    if (v < v1) raw = *row(head, t);
    while (v < v1) {
        const long long vn = v + kGridThreads;
        int hn = head, tn = t;
        advance(hn, tn);
        Raw nxt = {};
        if (vn < v1) nxt = *row(hn, tn);
        float p[4];
        if constexpr (sizeof(T) == 2) {
            const float2 a = io<T>::to_f2(raw.x), b = io<T>::to_f2(raw.y);
            p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y;
        } else {
            p[0] = raw.x; p[1] = raw.y; p[2] = raw.z; p[3] = raw.w;
        }
        const int c0 = grid_code(p[0], p[1], hdr, tab, cs, C);
        const int c1 = grid_code(p[2], p[3], hdr + 8, tab + kGridCells, cs + 256, C);
        dst.put2(head, t, 2 * grp, c0, c1);
        raw = nxt;
        v = vn; head = hn; t = tn;
    }
}

int launch_encode_grid(const void* x, int x_dtype, int64_t xhs, const float* cent, const void* prepared, const CodeDst& dst, int n_heads,
                       int n_tokens, int d, int M, int C, cudaStream_t stream, bool probe_only) {
    if (encode_grid_prepared_bytes(d, M, C) == 0) MILLION_UNSUPPORTED("grid encoder needs d/M = 2, even M, C <= 256 (d=%d M=%d C=%d)", d, M, C);
    if (!prepared) MILLION_UNSUPPORTED("grid encoder needs the tables of million_pq_encoder_grid_prepare");
    const int eb = x_dtype == MILLION_F32 ? 4 : 2;
    if (((uintptr_t)x % (4 * eb)) || ((xhs * eb) % (4 * eb)) || ((int64_t)d * eb) % (4 * eb)) MILLION_UNSUPPORTED("grid encoder needs 4-element aligned rows");
    if (probe_only) return MILLION_OK;
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    const long long n_vec = (long long)n_heads * n_tokens;
    const int groups = M / 2;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    // slices: enough CTAs to fill the machine a few times over, never fewer than ~2K vectors per CTA (the 132 KB prologue)
    long long slices = (n_vec + 2047) / 2048;
    const long long max_slices = (8LL * sms + groups - 1) / groups;
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    const long long per_cta = (n_vec + slices - 1) / slices;
    const size_t smem = 2 * kGridCells * 16 + 2 * 256 * 8 + 2 * 8 * 4;
    dim3 grid(groups, (unsigned)slices), block(kGridThreads);
#define MILLION_GRID_CASE(TT)                                                                                                          \
    MILLION_CUDA_OK(cudaFuncSetAttribute(encode_grid_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    encode_grid_kernel<TT><<<grid, block, smem, stream>>>((const TT*)x, xhs, cent, (const unsigned char*)prepared, dst, n_tokens, n_vec, per_cta, d, C)
    if (x_dtype == MILLION_F16) { MILLION_GRID_CASE(__half); }
    else if (x_dtype == MILLION_BF16) { MILLION_GRID_CASE(__nv_bfloat16); }
    else if (x_dtype == MILLION_F32) { MILLION_GRID_CASE(float); }
    else MILLION_UNSUPPORTED("grid encoder: unknown x dtype %d", x_dtype);
#undef MILLION_GRID_CASE
    MILLION_CUDA_OK(cudaGetLastError());
    return MILLION_OK;
}

}  // namespace million
