// codec.cuh — destination descriptor shared by the encoders (row-major / transposed / paged code stores).
#pragma once
#include "common.cuh"

namespace million {

struct CodeDst {
    void* base;
    int code_bytes;
    int64_t head_stride, token_stride, m_stride, t0;  // elements
    const int64_t* page_ids;                          // non-null => paged pool (pages, M, page_size) uint8
    int64_t page_ids_head_stride;
    int page_size, M;

    __device__ __forceinline__ int64_t offset(int head, int t, int m) const {
        if (page_ids) {
            // token indices fit 32 bits (nk is an int32 in the attention call): 32-bit division instead of two 64-bit ones
            const uint32_t tt = (uint32_t)(t0 + t), chunk = tt / (uint32_t)page_size, in_page = tt - chunk * (uint32_t)page_size;
            const int64_t page = page_ids[head * page_ids_head_stride + chunk];
            return (page * M + m) * page_size + in_page;
        }
        return head * head_stride + (t0 + t) * token_stride + m * m_stride;
    }
    // codes of the adjacent sub-spaces m (even) and m + 1: one address computation (one block-table lookup), and one 2-byte
    // store where the two codes are adjacent in memory
    __device__ __forceinline__ void put2(int head, int t, int m, int c0, int c1) const {
        const int64_t off = offset(head, t, m);
        const int64_t step = page_ids ? (int64_t)page_size : m_stride;
        if (code_bytes == 1) {
            uint8_t* p = reinterpret_cast<uint8_t*>(base) + off;
            if (step == 1 && (reinterpret_cast<uintptr_t>(p) & 1) == 0) {
                *reinterpret_cast<uint16_t*>(p) = (uint16_t)(c0 | (c1 << 8));
            } else {
                p[0] = (uint8_t)c0;
                p[step] = (uint8_t)c1;
            }
        } else {
            uint16_t* p = reinterpret_cast<uint16_t*>(base) + off;
            p[0] = (uint16_t)c0;
            p[step] = (uint16_t)c1;
        }
    }
    __device__ __forceinline__ void put(int head, int t, int m, int code) const {
        const int64_t off = offset(head, t, m);
        if (code_bytes == 1) reinterpret_cast<uint8_t*>(base)[off] = (uint8_t)code;
        else reinterpret_cast<uint16_t*>(base)[off] = (uint16_t)code;
    }
};

int launch_encode_generic(const void* x, int x_dtype, int64_t xhs, const float* cent, const CodeDst& dst, int n_heads,
                          int n_tokens, int d, int M, int C, cudaStream_t stream);

}  // namespace million
