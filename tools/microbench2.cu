// tools/microbench2.cu — follow-up: are HMMA, LDS and ALU issue independent on B200?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int ITERS = 2048;
// MODE bits: 1 = mma (4 per iter), 2 = ffma (NF per iter), 4 = lds64 (8 per iter, addresses loop-invariant+xor), 8 = lds32
template <int MODE, int NF>
__global__ void k(float* out) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i & 0x3ff;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float c[4][4] = {};
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
    float f[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    uint32_t acc = 0;
    uint32_t off = (lane % 16) * 8 + (threadIdx.x >> 5) * 128;   // conflict-free bank pairs
    uint32_t off4 = lane * 4 + (threadIdx.x >> 5) * 128;
    for (int i = 0; i < ITERS; ++i) {
        if (MODE & 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) mma16816(c[j], a, b);
        }
        if (MODE & 2) {
#pragma unroll
            for (int j = 0; j < NF; ++j) f[j & 7] = fmaf(f[j & 7], 1.0001f, 0.5f);
        }
        if (MODE & 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint2 v = *reinterpret_cast<const uint2*>(sm + ((off + j * 4096) & 0x1ffff));
                acc ^= v.x + v.y;
            }
            off += (acc & 0x380) ;   // data-dependent row change, keeps bank
        }
        if (MODE & 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t v = *reinterpret_cast<const uint32_t*>(sm + ((off4 + j * 4096) & 0x1ffff));
                acc ^= v;
            }
            off4 += (acc & 0x380);
        }
    }
    float s = acc;
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    for (int j = 0; j < 8; ++j) s += f[j];
    if (s == 123.456f) out[0] = s;
}
template <typename F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int MODE, int NF> void run(const char* name, float* out, int sms) {
    auto kern = k<MODE, NF>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (int warps : {8, 16}) {
        float ms = time_ms([&] { kern<<<sms, warps * 32, 128 * 1024>>>(out); });
        double clk = ms * 1e-3 * 1.965e9 / ITERS;   // clocks per loop iteration per SM (all warps)
        printf("%-28s warps=%2d: %.3f ms, %.1f clk/iter/SM (%.2f clk per warp-iter)\n", name, warps, ms, clk, clk / warps);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, 16);
    run<1, 0>("4 mma", out, sms);
    run<2, 32>("32 ffma", out, sms);
    run<3, 32>("4 mma + 32 ffma", out, sms);
    run<4, 0>("8 lds64", out, sms);
    run<8, 0>("8 lds32", out, sms);
    run<5, 0>("4 mma + 8 lds64", out, sms);
    run<9, 0>("4 mma + 8 lds32", out, sms);
    run<6, 32>("32 ffma + 8 lds64", out, sms);
    run<7, 32>("4 mma + 32 ffma + 8 lds64", out, sms);
    run<12, 0>("8 lds64 + 8 lds32", out, sms);
    return 0;
}
