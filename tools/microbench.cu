// tools/microbench.cu — B200 micro-measurements that decide the decode-attention design:
//   hmma      : mma.sync.m16n8k16 f16->f32 issue rate per SM
//   lds64_cf  : LDS.64 gather, conflict-free by construction (16 lanes -> 16 distinct bank pairs)
//   lds64_rnd : LDS.64 gather, random bank pairs
//   lds32_cf / lds32_rnd : same for 4-byte gathers
//   mix       : lds64_cf + hmma interleaved (do the pipes overlap?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int ITERS = 4096;

__global__ void k_hmma(float* out) {
    float c[4][4] = {};
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
    a[0] += threadIdx.x;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816(c[j], a, b);
    }
    float s = 0;
    for (int j = 0; j < 4; ++j) for (int k = 0; k < 4; ++k) s += c[j][k];
    if (s == 123.f) out[0] = s;
}

template <int BYTES, bool RANDOM, bool WITH_MMA>
__global__ void k_lds(float* out, const uint32_t* rnd) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int n_words = 128 * 1024 / 4;
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // row index: pseudo-random per lane and iteration (data dependent like a code byte); bank: lane-owned or random
    uint32_t state = rnd[threadIdx.x + blockIdx.x * blockDim.x];
    uint64_t acc = 0;
    float c[4] = {};
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
    constexpr int SLOTS_PER_ROW = 128 / BYTES;         // one 128-byte row = 32 banks
    constexpr int ROWS = 128 * 1024 / 128;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            state = state * 1664525u + 1013904223u;
            const uint32_t row = (state >> 8) % ROWS;
            const uint32_t slot = RANDOM ? ((state >> 20) % SLOTS_PER_ROW) : (lane % SLOTS_PER_ROW);
            const unsigned char* p = sm + row * 128 + slot * BYTES;
            if (BYTES == 8) { uint2 v = *reinterpret_cast<const uint2*>(p); acc += v.x ^ v.y; if (WITH_MMA) { a[0] = v.x; a[2] = v.y; } }
            else { uint32_t v = *reinterpret_cast<const uint32_t*>(p); acc += v; if (WITH_MMA) a[0] = v; }
            if (WITH_MMA && (j & 1)) mma16816(c, a, b);
        }
    }
    if (acc == 0x1234567u || c[0] == 77.f) out[0] = (float)acc;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock attr %.0f MHz\n", p.name, sms, khz / 1e3);
    float* out; cudaMalloc(&out, 16);
    uint32_t* rnd; cudaMalloc(&rnd, sms * 1024 * 4);
    { uint32_t* h = new uint32_t[sms * 1024]; for (int i = 0; i < sms * 1024; ++i) h[i] = i * 2654435761u + 12345u; cudaMemcpy(rnd, h, sms * 1024 * 4, cudaMemcpyHostToDevice); }
    const double ghz = 1.9;  // nominal; ratios matter
    for (int warps : {4, 8, 16}) {
        float ms = time_ms([&] { k_hmma<<<sms, warps * 32>>>(out); });
        double mmas = (double)sms * warps * ITERS * 4;
        printf("hmma  warps/SM=%2d: %.3f ms  -> %.2f MMA(m16n8k16)/ns chip, %.3f MMA/clk/SM @%.1fGHz, %.1f TFLOP/s\n", warps, ms,
               mmas / (ms * 1e6), mmas / (ms * 1e-3) / sms / (ghz * 1e9), ghz, mmas * 4096 / (ms * 1e-3) / 1e12);
    }
    auto run_lds = [&](const char* name, auto kern, int bytes) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        for (int warps : {8, 16}) {
            float ms = time_ms([&] { kern<<<sms, warps * 32, 128 * 1024>>>(out, rnd); });
            double ops = (double)sms * warps * ITERS * 8;   // warp-level LDS instructions
            printf("%-10s warps/SM=%2d: %.3f ms -> %.3f clk per warp-LDS per SM @%.1fGHz, %.1f B/clk/SM\n", name, warps, ms,
                   (ms * 1e-3) * ghz * 1e9 / (ops / sms), ghz, ops / sms * 32 * bytes / ((ms * 1e-3) * ghz * 1e9));
        }
    };
    run_lds("lds64_cf", k_lds<8, false, false>, 8);
    run_lds("lds64_rnd", k_lds<8, true, false>, 8);
    run_lds("lds32_cf", k_lds<4, false, false>, 4);
    run_lds("lds32_rnd", k_lds<4, true, false>, 4);
    run_lds("mix64+mma", k_lds<8, false, true>, 8);
    return 0;
}
