// dispatch_stubs.cu — temporary: fast-path hooks until attn_fast.cu / encode_tc.cu land.
#include "attn_common.cuh"
#include "codec.cuh"
namespace million {

int encode_dispatch(const void* x, int x_dtype, int64_t xhs, const float* cent, void* codes, int code_bytes, int64_t chs,
                    int64_t cts, int64_t cms, int64_t t0, const int64_t* page_ids, int64_t pihs, int page_size, int n_heads,
                    int n_tokens, int d, int M, int C, int impl, cudaStream_t stream) {
    if (impl == MILLION_IMPL_FAST) { set_error("fast encoder not built"); return MILLION_ERR_UNSUPPORTED; }
    CodeDst dst{codes, code_bytes, chs, cts, cms, t0, page_ids, pihs, page_size, M};
    return launch_encode_generic(x, x_dtype, xhs, cent, dst, n_heads, n_tokens, d, M, C, stream);
}
}
