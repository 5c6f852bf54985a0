// tools/microbench6.cu — attainable rate of the decode kernel's gather mixes with compiler-scheduled (non-volatile) loads
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int ITERS = 512;
constexpr uint32_t kBase = 0x400;
__device__ __forceinline__ uint32_t g32(uint32_t r) { uint32_t v; asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(r), "n"(kBase + 131072)); return v; }
__device__ __forceinline__ uint2 g64(uint32_t r) { uint2 v; asm("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(r), "n"(kBase)); return v; }
__device__ __forceinline__ void fhadd2(float& lo, float& hi, uint32_t p) {
    const unsigned short l = (unsigned short)(p & 0xffffu), h = (unsigned short)(p >> 16);
    asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(lo) : "h"(l)); asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(hi) : "h"(h));
}
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
// MODE 0: QK only (16 words x 4 gathers, FHADD); 1: QK integer adds; 2: PV only (16 x 4 gathers x 4 HFMA2); 3: QK(FHADD) then PV; 4: QK(int) then PV;
// 5: fused per step QK(FHADD)+PV; 6: fused QK(int)+PV
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, const uint32_t* in, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < (131072 + 65536 + 16384) / 4; i += blockDim.x) sm[i] = in[i & 1023] * (i + 1);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lq = lane & 15, hw = lane >> 4, rot = (lq + hw) & 15;
    const unsigned char* ksp = smem + 131072 + 65536 + (warp & 3) * 4096;   // K tile 2 KB + V tile 2 KB (shared by warp pairs here)
    const unsigned char* vsp = ksp + 2048;
    uint32_t koff[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) { const int Wl = (w + rot) & 15; koff[w] = (uint32_t)(Wl * 8) | ((uint32_t)(Wl * 8 + 128) << 8); }
    const uint32_t voff01 = (uint32_t)(lq * 4) | ((uint32_t)(lq * 4 + 64) << 8), voff23 = (uint32_t)(lq * 4 + 128) | ((uint32_t)(lq * 4 + 192) << 8);
    uint32_t vsel[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) vsel[s] = (uint32_t)(4 + (s & 1)) | ((uint32_t)((s + hw) & 3) << 4) | 0x6600u;
    float s[4] = {0, 0, 0, 0}; uint32_t fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0};
    __half2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[i][g] = __float2half2_rn(0.f);
    uint32_t pk0 = in[lane], pk1 = in[lane + 32];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        constexpr bool QK = MODE != 2, PV = MODE >= 2, INT = (MODE == 1 || MODE == 4 || MODE == 6), FUSED = MODE >= 5;
        uint32_t words[16];
        if (QK) {
#pragma unroll
            for (int w = 0; w < 16; ++w) words[w] = *reinterpret_cast<const uint32_t*>(ksp + lane * 64 + (((w + rot) & 15) << 2)) + it;
        }
        auto qk_step = [&](int w) {
#pragma unroll
            for (int bq = 0; bq < 4; ++bq) {
                constexpr uint32_t selc[4] = {0x6604u, 0x6615u, 0x6624u, 0x6635u};
                const uint32_t ad = __byte_perm(words[w], koff[w], selc[bq]);
                const uint2 e = g64(ad + (bq >> 1) * 65536);
                if (INT) { fa[w >> 2] += e.x; fb[w >> 2] += e.y; } else { fhadd2(s[0], s[1], e.x); fhadd2(s[2], s[3], e.y); }
            }
        };
        auto pv_step = [&](int jp) {
            const int j = 2 * jp + hw;
            const uint32_t word = *reinterpret_cast<const uint32_t*>(vsp + j * 64 + lq * 4) + it;
            const __half2 p01 = as_h2(pk0 + jp), p23 = as_h2(pk1 + jp);
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
                const uint32_t ad = __byte_perm(word, sl < 2 ? voff01 : voff23, vsel[sl]);
                const __half2 v = as_h2(g32(ad));
                acc[sl][0] = __hfma2(__low2half2(p01), v, acc[sl][0]); acc[sl][1] = __hfma2(__high2half2(p01), v, acc[sl][1]);
                acc[sl][2] = __hfma2(__low2half2(p23), v, acc[sl][2]); acc[sl][3] = __hfma2(__high2half2(p23), v, acc[sl][3]);
            }
        };
        if (FUSED) {
#pragma unroll
            for (int w = 0; w < 16; ++w) { qk_step(w); pv_step(w); }
        } else {
            if (QK) {
#pragma unroll
                for (int w = 0; w < 16; ++w) qk_step(w);
            }
            if (PV) {
#pragma unroll 4
                for (int jp = 0; jp < 16; ++jp) pv_step(jp);
            }
        }
        if (INT) { s[0] += (float)((fa[0] & 0xffff) + (fa[1] & 0xffff) + (fa[2] & 0xffff) + (fa[3] & 0xffff)); s[1] += (float)((fa[0] >> 16) + (fa[1] >> 16) + (fa[2] >> 16) + (fa[3] >> 16));
                   s[2] += (float)((fb[0] & 0xffff) + (fb[1] & 0xffff) + (fb[2] & 0xffff) + (fb[3] & 0xffff)); s[3] += (float)((fb[0] >> 16) + (fb[1] >> 16) + (fb[2] >> 16) + (fb[3] >> 16));
                   fa[0] = fa[1] = fa[2] = fa[3] = fb[0] = fb[1] = fb[2] = fb[3] = 0; }
        pk0 += __float_as_uint(s[0]) & 1; pk1 += __float_as_uint(s[2]) & 1;
    }
    const long long t1 = clock64();
    float r = s[0] + s[1] + s[2] + s[3];
    for (int i = 0; i < 4; ++i) for (int g = 0; g < 4; ++g) r += __low2float(acc[i][g]) + __high2float(acc[i][g]);
    if (r == 1.2345f) out[0] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, float* out, uint32_t* in, long long* cyc, int sms, int warps) {
    const int smem = 131072 + 65536 + 16384;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE><<<sms, warps * 32, smem>>>(out, in, cyc); cudaDeviceSynchronize();
    k<MODE><<<sms, warps * 32, smem>>>(out, in, cyc); cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("%-26s warps=%2d: %.2f clk per token per SM  (%s)\n", name, warps, avg / ((double)ITERS * 32 * warps), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, 16); uint32_t* in; cudaMalloc(&in, 4096); cudaMemset(in, 0x3c, 4096); long long* cyc; cudaMalloc(&cyc, 256 * 8);
    for (int w : {8, 12, 16}) {
        run<0>("QK (FHADD)", out, in, cyc, sms, w); run<1>("QK (int adds)", out, in, cyc, sms, w); run<2>("PV", out, in, cyc, sms, w);
        run<3>("QK(FHADD) ; PV", out, in, cyc, sms, w); run<4>("QK(int) ; PV", out, in, cyc, sms, w);
        run<5>("fused QK(FHADD)+PV", out, in, cyc, sms, w); run<6>("fused QK(int)+PV", out, in, cyc, sms, w);
    }
    return 0;
}
