// tools/microbench3.cu — issue rate of the arithmetic instructions the decode kernel leans on (per SM, 16 warps resident)
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int OP>
__global__ void k(float* out, const uint32_t* in) {
    uint32_t x = in[threadIdx.x], y = in[threadIdx.x + 32];
    float f[8]; uint32_t u[8]; __half2 h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = i + 1.f; u[i] = x + i; h[i] = __float2half2_rn((float)i); }
    const __half2 p = *reinterpret_cast<__half2*>(&y), v = *reinterpret_cast<__half2*>(&x);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) f[i] = fmaf(f[i], 1.0001f, f[(i + 1) & 7]);                                   // FFMA
            if (OP == 1) { unsigned short lo = (unsigned short)(x & 0xffff); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[i]) : "h"(lo)); }   // FHADD
            if (OP == 2) { unsigned short hi = (unsigned short)(x >> 16); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(f[i]) : "h"(hi)); }      // FHADD .H1
            if (OP == 3) h[i] = __hfma2(__low2half2(p), v, h[i]);                                       // HFMA2 .H0_H0
            if (OP == 4) h[i] = __hfma2(p, v, h[i]);                                                    // HFMA2 plain
            if (OP == 5) h[i] = __hadd2(h[i], v);                                                       // HADD2
            if (OP == 6) u[i] = __byte_perm(u[i], y, 0x6604 + i);                                       // PRMT
            if (OP == 7) u[i] = u[i] + y + u[(i + 1) & 7];                                              // IADD3
            if (OP == 8) u[i] = __dp2a_lo((int)x, (int)y, (int)u[i]);                                   // IDP.2A
            if (OP == 9) f[i] = f[i] + f[(i + 1) & 7];                                                  // FADD
            if (OP == 10) { float2 t = __half22float2(h[i]); f[i] += t.x; }                             // HADD2.F32 + FADD
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + u[i] + __low2float(h[i]) + __high2float(h[i]);
    if (s == 1.2345f) out[0] = s;
}
template <int OP> void run(const char* name, float* out, uint32_t* in, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<sms, 512>>>(out, in); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<sms, 512>>>(out, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = 16.0 * ITERS * 8;   // warp-instructions per SM
    printf("%-22s %.3f ms -> %.2f warp-instr/clk/SM (@1.965 GHz)\n", name, ms, instr / (ms * 1e-3 * 1.965e9));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, 16); uint32_t* in; cudaMalloc(&in, 4096); cudaMemset(in, 0x3c, 4096);
    run<0>("FFMA", out, in, sms); run<9>("FADD", out, in, sms); run<1>("FHADD", out, in, sms); run<2>("FHADD .H1", out, in, sms);
    run<3>("HFMA2 p.H0_H0", out, in, sms); run<4>("HFMA2", out, in, sms); run<5>("HADD2", out, in, sms);
    run<6>("PRMT", out, in, sms); run<7>("IADD3", out, in, sms); run<8>("IDP.2A", out, in, sms); run<10>("HADD2.F32+FADD (2 ops)", out, in, sms);
    return 0;
}
