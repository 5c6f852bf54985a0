// common.cuh — device/host utilities shared by the million_b200 kernels (sm_100a only).
// Replaces the reference's core/{Scalar,Vector,Reduction,DeviceOps}.cuh helper layer
// (scripts/modeldb/bindings/core/*.cuh): vector loads, shuffle reductions (instead of
// single_thread_reduce, DeviceOps.cuh), dtype traits.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/million_b200.h"

namespace million {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MILLION_CUDA_OK(expr)                                         \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) return ::million::cuda_fail(_e, #expr); \
    } while (0)

#define MILLION_REQUIRE(cond, ...)              \
    do {                                        \
        if (!(cond)) {                          \
            ::million::set_error(__VA_ARGS__);  \
            return MILLION_ERR_INVALID;         \
        }                                       \
    } while (0)

#define MILLION_UNSUPPORTED(...)            \
    do {                                    \
        ::million::set_error(__VA_ARGS__);  \
        return MILLION_ERR_UNSUPPORTED;     \
    } while (0)

int sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: one flag per device, so a process that touches a
// second GPU (device_map='auto', tests iterating over devices) configures the kernel there too.  Racing threads at worst set
// the attribute twice.
struct SmemAttrOnce { unsigned char done[64]; };
template <typename F>
inline cudaError_t ensure_dynamic_smem(SmemAttrOnce& st, F func, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && st.done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && tracked) st.done[dev] = 1;
    return e;
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// launch_dependents: the next kernel of the stream (if launched with the programmatic-serialization attribute) may start its
// prologue on SMs this grid no longer occupies.  wait: returns once the previous grid has completed and its writes are visible
// (immediately when the launch carried no such dependency).  Everything a kernel reads BEFORE its wait must not be produced by
// its stream predecessor.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---------------------------------------------------------------- dtype traits
template <typename T> struct io;
template <> struct io<__half> {
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
    static __device__ __forceinline__ float2 to_f2(uint32_t packed) {
        return __half22float2(*reinterpret_cast<const __half2*>(&packed));
    }
};
template <> struct io<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
    static __device__ __forceinline__ float2 to_f2(uint32_t packed) {
        return make_float2(__uint_as_float(packed << 16), __uint_as_float(packed & 0xffff0000u));
    }
};
template <> struct io<float> {
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide reductions for blocks of up to 32 warps; `scratch` holds >= 33 floats
template <bool kMax>
__device__ __forceinline__ float block_reduce(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = kMax ? warp_max(v) : warp_sum(v);
    __syncthreads();  // protect scratch from the previous use
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : (kMax ? -INFINITY : 0.f);
    r = kMax ? warp_max(r) : warp_sum(r);
    return r;
}

// exp2 that maps (-inf) - (-inf) to 0 instead of NaN
__device__ __forceinline__ float exp2_safe(float x, float ref) {
    return (ref == -INFINITY) ? 0.f : exp2f(x - ref);
}

// same, one MUFU.EX2 (ex2.approx.ftz: 2 ulp, arguments below -126 flush to 0) instead of exp2f's range handling: for the
// inner loops, where x - ref <= a few units by construction
__device__ __forceinline__ float exp2_fast(float x, float ref) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x - ref));
    return (ref == -INFINITY) ? 0.f : r;
}

// ---------------------------------------------------------------- streaming loads
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

}  // namespace million
