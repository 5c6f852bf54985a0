// tools/microbench4.cu — clock64-based issue rates (independent of the SM clock), 8 vs 16 warps per SM, plus the PV / QK instruction mixes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
__device__ __forceinline__ uint32_t lds32a(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64a(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
template <int OP>
__global__ void k(float* out, const uint32_t* in, long long* cyc) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = in[i & 1023];
    __syncthreads();
    uint32_t x = in[threadIdx.x & 31], y = in[(threadIdx.x & 31) + 32];
    float f[16]; uint32_t u[16]; __half2 h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { f[i] = i + 1.f; u[i] = x + i; h[i] = __float2half2_rn((float)i); }
    const __half2 p = *reinterpret_cast<__half2*>(&y), v = *reinterpret_cast<__half2*>(&x);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) f[i] = fmaf(f[i], 1.0001f, 0.5f);                                              // FFMA independent
            if (OP == 1) { unsigned short lo = (unsigned short)(x & 0xffff); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[i]) : "h"(lo)); }
            if (OP == 3) h[i] = __hfma2(__low2half2(p), v, h[i]);
            if (OP == 6) u[i] = __byte_perm(u[i], y, 0x6604 + (i & 3));
            if (OP == 7) u[i] = u[i] + y + x;
            if (OP == 11) { u[i] += lds32a(sbase + ((u[i] & 0x3f) << 7)); }                              // LDS.32 conflict-free, dependent address
            if (OP == 12) { uint2 t = lds64a(sbase * 2 - (uint32_t)__cvta_generic_to_shared(sm) + ((u[i] & 0x3f) << 8)); u[i] += t.x + t.y; }
        }
        if (OP == 20) {   // PV-like block: 4 x (1 PRMT + 1 LDS.32 + 4 HFMA2)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t ad = __byte_perm(u[i], y, 0x6604 + i) & 0xfffc;
                const uint32_t e = lds32a(sbase + (ad & 0x1f80));
                const __half2 ev = *reinterpret_cast<const __half2*>(&e);
                h[4 * i] = __hfma2(__low2half2(p), ev, h[4 * i]); h[4 * i + 1] = __hfma2(__high2half2(p), ev, h[4 * i + 1]);
                h[4 * i + 2] = __hfma2(__low2half2(v), ev, h[4 * i + 2]); h[4 * i + 3] = __hfma2(__high2half2(v), ev, h[4 * i + 3]);
                u[i] += 0x01010101u;
            }
        }
        if (OP == 21) {   // QK-like block: 4 x (1 PRMT + 1 LDS.64 + 4 FHADD)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t ad = __byte_perm(u[i], y, 0x6604 + i);
                const uint2 e = lds64a(sbase * 2 - (uint32_t)__cvta_generic_to_shared(sm) + (ad & 0x3f00));
                unsigned short a0 = e.x & 0xffff, a1 = e.x >> 16, a2 = e.y & 0xffff, a3 = e.y >> 16;
                asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[0]) : "h"(a0)); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[1]) : "h"(a1));
                asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[2]) : "h"(a2)); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[3]) : "h"(a3));
                u[i] += 0x01010101u;
            }
        }
        if (OP == 22) {   // both blocks
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t ad = __byte_perm(u[i], y, 0x6604 + i) & 0xfffc;
                const uint32_t e = lds32a(sbase + (ad & 0x1f80));
                const __half2 ev = *reinterpret_cast<const __half2*>(&e);
                h[4 * i] = __hfma2(__low2half2(p), ev, h[4 * i]); h[4 * i + 1] = __hfma2(__high2half2(p), ev, h[4 * i + 1]);
                h[4 * i + 2] = __hfma2(__low2half2(v), ev, h[4 * i + 2]); h[4 * i + 3] = __hfma2(__high2half2(v), ev, h[4 * i + 3]);
                const uint32_t ad2 = __byte_perm(u[i + 4], y, 0x6604 + i);
                const uint2 e2 = lds64a(sbase * 2 - (uint32_t)__cvta_generic_to_shared(sm) + (ad2 & 0x3f00));
                unsigned short a0 = e2.x & 0xffff, a1 = e2.x >> 16, a2 = e2.y & 0xffff, a3 = e2.y >> 16;
                asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[0]) : "h"(a0)); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[1]) : "h"(a1));
                asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[2]) : "h"(a2)); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[3]) : "h"(a3));
                u[i] += 0x01010101u; u[i + 4] += 0x01010101u;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += f[i] + u[i] + __low2float(h[i]) + __high2float(h[i]);
    if (s == 1.2345f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, float* out, uint32_t* in, long long* cyc, int sms, int warps, double per_iter) {
    cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<OP><<<sms, warps * 32, 65536>>>(out, in, cyc); cudaDeviceSynchronize();
    k<OP><<<sms, warps * 32, 65536>>>(out, in, cyc); cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("%-28s warps=%2d: %8.0f clk, %.2f clk per warp-iter-unit/SM -> %.2f units/clk/SM  (%s)\n", name, warps, avg, avg / (ITERS * per_iter * warps) * 1.0,
           ITERS * per_iter * warps / avg, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, 16); uint32_t* in; cudaMalloc(&in, 4096); cudaMemset(in, 0x3c, 4096); long long* cyc; cudaMalloc(&cyc, 256 * 8);
    for (int w : {4, 8, 16}) {
        run<0>("FFMA (16 indep)", out, in, cyc, sms, w, 16); run<1>("FHADD", out, in, cyc, sms, w, 16); run<3>("HFMA2", out, in, cyc, sms, w, 16);
        run<6>("PRMT", out, in, cyc, sms, w, 16); run<7>("IADD3", out, in, cyc, sms, w, 16);
        run<11>("LDS.32 dep addr", out, in, cyc, sms, w, 16); run<12>("LDS.64 dep addr", out, in, cyc, sms, w, 16);
        run<20>("PV block (4 gathers)", out, in, cyc, sms, w, 4); run<21>("QK block (4 gathers)", out, in, cyc, sms, w, 4); run<22>("PV+QK blocks (4+4)", out, in, cyc, sms, w, 4);
    }
    return 0;
}
