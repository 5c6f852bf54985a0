// TEST INFRASTRUCTURE ONLY — C launcher around the REFERENCE's own CUDA kernels, compiled from where they lie
// (-I $REFERENCE/scripts/modeldb/bindings; nothing of the reference is copied into this repo).
// Instantiates the production trio of scripts/modeldb/bindings/Kernel.cuh for f16 / u8 codes, Lt=d=128, M=64, C=256:
//   flash_decoding_split_kernel (:11-166), flash_decoding_residual_kernel (:1038-1209), flash_decoding_reduce_kernel (:1211-1270)
// launched exactly as Interface.cu:62-115 does (grid/block shapes, default stream -> here: the given stream).
// The LUT (Interface.cu:49-50, at::matmul) is computed by the caller with torch and passed in.
#include "Kernel.cuh"

template <int Ns>
static int launch(const __half* lut, const uint8_t* kc, const uint8_t* vc, const __half* vcent, const __half* q, const __half* kres,
                  const __half* vres, int r, __half* pout, __half* plse, __half* out, int bs, int nh, int nh_k, int nk, cudaStream_t st) {
    constexpr int Lt = 128, d = 128, M = 64, C = 256;
    const int Ls = (nk + Ns - 1) / Ns;
    flash_decoding_split_kernel<__half, uint8_t, Ns, Lt, d, M, C><<<dim3(bs, nh, Ns), dim3(Lt), 0, st>>>(lut, kc, vc, vcent, pout, plse, bs, nh, nh_k, nk, Ls);
    flash_decoding_residual_kernel<__half, Ns, Lt, d><<<dim3(bs, nh), dim3(d), 0, st>>>(q, kres, vres, r, pout, plse, bs, nh, nh_k);
    cudaMemsetAsync(out, 0, sizeof(__half) * bs * nh * d, st);   // Interface.cu:104 torch::zeros
    flash_decoding_reduce_kernel<__half, Ns, d><<<dim3(bs, nh), dim3(d), 0, st>>>(pout, plse, out, bs, nh);
    return (int)cudaGetLastError();
}

extern "C" int ref_flash_decoding_f16u8_Lt128d128M64C256(int Ns, const void* lut, const void* kc, const void* vc, const void* vcent,
                                                        const void* q, const void* kres, const void* vres, int r, void* pout, void* plse,
                                                        void* out, int bs, int nh, int nh_k, int nk, void* stream) {
#define ARGS (const __half*)lut, (const uint8_t*)kc, (const uint8_t*)vc, (const __half*)vcent, (const __half*)q, (const __half*)kres, \
             (const __half*)vres, r, (__half*)pout, (__half*)plse, (__half*)out, bs, nh, nh_k, nk, (cudaStream_t)stream
    switch (Ns) {
        case 16: return launch<16>(ARGS);
        case 32: return launch<32>(ARGS);
    }
    return -1;
}
