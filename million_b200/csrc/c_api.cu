// c_api.cu — extern "C" entry points of libmillion_b200.so (include/million_b200.h): argument checks,
// implementation choice, launches.  No torch, no exceptions, no host synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "attn_common.cuh"
#include "codec.cuh"

namespace million {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return MILLION_ERR_CUDA;
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached[dev] = n;
    }
    return cached[dev];
}

int launch_attn_generic(const AttnArgs& a, int io_dtype, cudaStream_t stream);
int launch_attn_fast(const AttnArgs& a, int io_dtype, const void* prepared, cudaStream_t stream, bool probe_only);
int launch_codebook_prepare(const void* kcent, const void* vcent, int io_dtype, void* out, cudaStream_t stream);
int launch_codebook_prepare_dm4(const void* kcent, const void* vcent, int io_dtype, void* out, cudaStream_t stream);
int launch_lse_merge(const float* parts, int n_parts, int64_t n_rows, int d, void* out, int io_dtype, cudaStream_t stream);
int launch_reconstruct(const void* codes, int code_bytes, int64_t chs, int64_t cts, int64_t cms, const void* cent, void* out,
                       int dtype, int64_t ohs, int n_heads, int n_tokens, int d, int M, int C, cudaStream_t stream);
int launch_rows_copy2(void* k_dst, void* v_dst, int64_t dst_hs_b, int64_t dst_off_b, const void* k_src, const void* v_src,
                      int64_t src_hs_b, int64_t src_off_b, int n_heads, int64_t bytes_per_head, cudaStream_t stream);
int64_t encode_tc_prepared_bytes(int d, int M, int C);
int launch_outlier_split(const void* x, int x_dtype, int64_t xhs, const float* cent, void* x_masked, uint8_t* out_idx, void* out_val,
                         int64_t ohs, int64_t t0, int n_heads, int n_tokens, int d, int M, int C, int k_out, cudaStream_t stream);
int launch_outlier_apply(void* out, int dtype, int64_t ohs, const uint8_t* idx, const void* val, int val_dtype, int64_t shs, int64_t t0,
                         int n_heads, int n_tokens, int d, int k_out, cudaStream_t stream);
int launch_encode_tc_prepare(const float* cent, int x_dtype, int d, int M, int C, void* out, cudaStream_t stream);
int launch_encode_tc(const void* x, int x_dtype, int64_t xhs, const float* cent, const void* prepared, const CodeDst& dst,
                     int n_heads, int n_tokens, int d, int M, int C, cudaStream_t stream, bool probe_only);

int64_t encode_grid_prepared_bytes(int d, int M, int C);
int launch_encode_grid_prepare(const float* cent, int d, int M, int C, void* out, cudaStream_t stream);
int launch_encode_grid(const void* x, int x_dtype, int64_t xhs, const float* cent, const void* prepared, const CodeDst& dst, int n_heads,
                       int n_tokens, int d, int M, int C, cudaStream_t stream, bool probe_only);

static int encode_dispatch(const void* x, int x_dtype, int64_t xhs, const float* cent, const void* prepared, void* codes, int code_bytes,
                           int64_t chs, int64_t cts, int64_t cms, int64_t t0, const int64_t* page_ids, int64_t pihs, int page_size,
                           int n_heads, int n_tokens, int d, int M, int C, int impl, cudaStream_t stream) {
    CodeDst dst{codes, code_bytes, chs, cts, cms, t0, page_ids, pihs, page_size, M};
    if (impl == MILLION_IMPL_GENERIC) return launch_encode_generic(x, x_dtype, xhs, cent, dst, n_heads, n_tokens, d, M, C, stream);
    if (impl == MILLION_IMPL_FAST) return launch_encode_tc(x, x_dtype, xhs, cent, prepared, dst, n_heads, n_tokens, d, M, C, stream, false);
    if (impl == MILLION_IMPL_GRID) {
        if (code_bytes != 1) MILLION_UNSUPPORTED("grid encoder writes one-byte codes");
        return launch_encode_grid(x, x_dtype, xhs, cent, prepared, dst, n_heads, n_tokens, d, M, C, stream, false);
    }
    // AUTO.  Two-dimensional sub-spaces: `prepared` holds the candidate-grid tables (million_pq_encoder_auto_prepare builds
    // those for this shape) and the exact grid encoder runs — 4x the tensor-core encoder, bit-identical codes.
    if (prepared != nullptr && code_bytes == 1 && encode_grid_prepared_bytes(d, M, C) > 0) {
        if (launch_encode_grid(x, x_dtype, xhs, cent, prepared, dst, n_heads, n_tokens, d, M, C, stream, true) == MILLION_OK)
            return launch_encode_grid(x, x_dtype, xhs, cent, prepared, dst, n_heads, n_tokens, d, M, C, stream, false);
        return launch_encode_generic(x, x_dtype, xhs, cent, dst, n_heads, n_tokens, d, M, C, stream);   // the tables are not tensor-core tiles
    }
    // otherwise the tensor-core encoder, which pays off from a few hundred vectors on
    if ((int64_t)n_heads * n_tokens >= 256 &&
        launch_encode_tc(x, x_dtype, xhs, cent, prepared, dst, n_heads, n_tokens, d, M, C, stream, true) == MILLION_OK)
        return launch_encode_tc(x, x_dtype, xhs, cent, prepared, dst, n_heads, n_tokens, d, M, C, stream, false);
    return launch_encode_generic(x, x_dtype, xhs, cent, dst, n_heads, n_tokens, d, M, C, stream);
}

#ifdef MILLION_DEBUG
static unsigned long long* g_dbg_timing = nullptr;
static int g_dbg_mode = 0;
#endif
int launch_window_append_dev(void* k_win, void* v_win, int64_t win_hs_b, const void* k_new, const void* v_new, int n_heads, int row_bytes,
                             const int* r_dev, int r_off, int res_len, cudaStream_t stream);
int launch_counter_add(int* ctr, int n, int delta, cudaStream_t stream);
int launch_rope_qk(const void* q, const void* k, const void* cos_, const void* sin_, void* q_out, void* k_out, int dtype, int bs, int nh,
                   int nh_k, int d, cudaStream_t stream);
static inline int elem_bytes(int dtype) { return dtype == MILLION_F32 ? 4 : 2; }

}  // namespace million

using namespace million;

extern "C" {

int million_abi_version(void) { return MILLION_ABI_VERSION; }
const char* million_last_error(void) { return g_err; }

int million_device_info(int* sms, int* major, int* minor) {
    int dev = 0;
    MILLION_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    MILLION_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return MILLION_OK;
}

int64_t million_pq_encoder_prepared_bytes(int d, int M, int C) { return encode_tc_prepared_bytes(d, M, C); }

int million_pq_encoder_prepare(const float* cent, int x_dtype, int d, int M, int C, void* prepared, million_stream_t stream) {
    MILLION_REQUIRE(cent && prepared, "encoder_prepare: null pointer");
    return launch_encode_tc_prepare(cent, x_dtype, d, M, C, prepared, (cudaStream_t)stream);
}

int64_t million_pq_encoder_grid_prepared_bytes(int d, int M, int C) { return encode_grid_prepared_bytes(d, M, C); }

int64_t million_pq_encoder_auto_prepared_bytes(int d, int M, int C) {
    const int64_t g = encode_grid_prepared_bytes(d, M, C);
    return g > 0 ? g : encode_tc_prepared_bytes(d, M, C);
}

int million_pq_encoder_auto_prepare(const float* cent, int x_dtype, int d, int M, int C, void* prepared, million_stream_t stream) {
    MILLION_REQUIRE(cent && prepared, "encoder_auto_prepare: null pointer");
    if (encode_grid_prepared_bytes(d, M, C) > 0) return launch_encode_grid_prepare(cent, d, M, C, prepared, (cudaStream_t)stream);
    return launch_encode_tc_prepare(cent, x_dtype, d, M, C, prepared, (cudaStream_t)stream);
}

int million_pq_encoder_grid_prepare(const float* cent, int d, int M, int C, void* prepared, million_stream_t stream) {
    MILLION_REQUIRE(cent && prepared, "encoder_grid_prepare: null pointer");
    return launch_encode_grid_prepare(cent, d, M, C, prepared, (cudaStream_t)stream);
}

int million_pq_encode(const void* x, int x_dtype, int64_t x_head_stride, const float* cent, const void* prepared, void* codes, int code_bytes,
                      int64_t codes_head_stride, int64_t codes_token_stride, int64_t codes_m_stride, int64_t t0,
                      int n_heads, int n_tokens, int d, int M, int C, int impl, million_stream_t stream) {
    MILLION_REQUIRE(n_heads >= 0 && n_tokens >= 0, "encode: negative sizes");
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    MILLION_REQUIRE(x && cent && codes, "encode: null pointer");
    MILLION_REQUIRE(M > 0 && d > 0 && d % M == 0 && C > 1, "encode: need d %% M == 0 and C > 1 (d=%d M=%d C=%d)", d, M, C);
    MILLION_REQUIRE(code_bytes == 1 || code_bytes == 2, "encode: code_bytes must be 1 or 2");
    MILLION_REQUIRE(C <= (code_bytes == 1 ? 256 : 65536), "encode: C=%d does not fit %d-byte codes", C, code_bytes);
    return encode_dispatch(x, x_dtype, x_head_stride, cent, prepared, codes, code_bytes, codes_head_stride, codes_token_stride,
                           codes_m_stride, t0, nullptr, 0, 0, n_heads, n_tokens, d, M, C, impl, (cudaStream_t)stream);
}

int million_pq_encode_paged(const void* x, int x_dtype, int64_t x_head_stride, const float* cent, const void* prepared, uint8_t* page_pool,
                            const int64_t* page_ids, int64_t page_ids_head_stride, int page_size, int64_t t0, int n_heads,
                            int n_tokens, int d, int M, int C, int impl, million_stream_t stream) {
    MILLION_REQUIRE(n_heads >= 0 && n_tokens >= 0, "encode_paged: negative sizes");
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    MILLION_REQUIRE(x && cent && page_pool && page_ids, "encode_paged: null pointer");
    MILLION_REQUIRE(M > 0 && d > 0 && d % M == 0 && C > 1 && C <= 256, "encode_paged: bad d/M/C (%d/%d/%d)", d, M, C);
    MILLION_REQUIRE(page_size > 0, "encode_paged: page_size must be positive");
    return encode_dispatch(x, x_dtype, x_head_stride, cent, prepared, page_pool, 1, 0, 0, 0, t0, page_ids, page_ids_head_stride,
                           page_size, n_heads, n_tokens, d, M, C, impl, (cudaStream_t)stream);
}

int million_pq_decode(const void* codes, int code_bytes, int64_t chs, int64_t cts, int64_t cms, const void* cent, void* out,
                      int dtype, int64_t out_head_stride, int n_heads, int n_tokens, int d, int M, int C,
                      million_stream_t stream) {
    MILLION_REQUIRE(n_heads >= 0 && n_tokens >= 0, "decode: negative sizes");
    if (n_heads == 0 || n_tokens == 0) return MILLION_OK;
    MILLION_REQUIRE(codes && cent && out, "decode: null pointer");
    MILLION_REQUIRE(M > 0 && d % M == 0, "decode: need d %% M == 0");
    MILLION_REQUIRE(code_bytes == 1 || code_bytes == 2, "decode: code_bytes must be 1 or 2");
    MILLION_REQUIRE(dtype >= MILLION_F16 && dtype <= MILLION_F32, "decode: bad dtype");
    return launch_reconstruct(codes, code_bytes, chs, cts, cms, cent, out, dtype, out_head_stride, n_heads, n_tokens, d, M, C,
                              (cudaStream_t)stream);
}

int million_pq_outlier_split(const void* x, int x_dtype, int64_t x_head_stride, const float* cent, void* x_masked, uint8_t* out_idx,
                             void* out_val, int64_t out_head_stride, int64_t t0, int n_heads, int n_tokens, int d, int M, int C,
                             int k_out, million_stream_t stream) {
    MILLION_REQUIRE(x && cent && x_masked && out_idx && out_val, "outlier_split: null pointer");
    MILLION_REQUIRE(x_dtype == MILLION_F16 || x_dtype == MILLION_BF16, "outlier_split: x must be f16 or bf16");
    MILLION_REQUIRE(n_heads >= 0 && n_tokens >= 0 && t0 >= 0, "outlier_split: bad sizes");
    MILLION_REQUIRE(M > 0 && d % M == 0 && d % 32 == 0 && d <= 256 && d / M <= 16 && C > 1, "outlier_split: needs d %% 32 == 0, d <= 256, d/M <= 16 (d=%d M=%d)", d, M);
    MILLION_REQUIRE(k_out >= 1 && k_out <= MILLION_MAX_OUTLIERS && k_out < d, "outlier_split: k_out must be in [1, %d]", MILLION_MAX_OUTLIERS);
    return launch_outlier_split(x, x_dtype, x_head_stride, cent, x_masked, out_idx, out_val, out_head_stride, t0, n_heads, n_tokens, d, M, C,
                                k_out, (cudaStream_t)stream);
}

int million_pq_outlier_apply(void* out, int dtype, int64_t out_head_stride, const uint8_t* idx, const void* val, int val_dtype,
                             int64_t store_head_stride, int64_t t0, int n_heads, int n_tokens, int d, int k_out, million_stream_t stream) {
    MILLION_REQUIRE(out && idx && val, "outlier_apply: null pointer");
    MILLION_REQUIRE(dtype >= MILLION_F16 && dtype <= MILLION_F32 && (val_dtype == MILLION_F16 || val_dtype == MILLION_BF16), "outlier_apply: bad dtype");
    MILLION_REQUIRE(n_heads >= 0 && n_tokens >= 0 && t0 >= 0 && d > 0 && k_out >= 1 && k_out <= MILLION_MAX_OUTLIERS, "outlier_apply: bad sizes");
    return launch_outlier_apply(out, dtype, out_head_stride, idx, val, val_dtype, store_head_stride, t0, n_heads, n_tokens, d, k_out,
                                (cudaStream_t)stream);
}

int64_t million_pq_codebook_prepared_bytes(int d, int M, int C) { return (d == 128 && (M == 64 || M == 32) && C == 256) ? 2 * 64 * 1024 : 0; }

int million_pq_codebook_prepare(const void* k_cent, const void* v_cent, int dtype, int d, int M, int C, void* prepared,
                                million_stream_t stream) {
    MILLION_REQUIRE(k_cent && v_cent && prepared, "codebook_prepare: null pointer");
    MILLION_REQUIRE(dtype >= MILLION_F16 && dtype <= MILLION_F32, "codebook_prepare: bad dtype");
    if (million_pq_codebook_prepared_bytes(d, M, C) == 0) MILLION_UNSUPPORTED("codebook_prepare: only d=128, M in {32, 64}, C=256");
    if (M == 32) return launch_codebook_prepare_dm4(k_cent, v_cent, dtype, prepared, (cudaStream_t)stream);
    return launch_codebook_prepare(k_cent, v_cent, dtype, prepared, (cudaStream_t)stream);
}

// Cost of the fp16 window and of the per-CTA prologue (LUT build) in units of 16 coded tokens (measured on B200)
static const int kWindowUnits = 0, kPrologueUnits = 44;   // ~4.5 us of per-CTA prologue + epilogue at ~0.1 us per unit

int million_pq_decode_attn_default_splits(int bs, int nh_k, int nk) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int groups = bs * nh_k > 0 ? bs * nh_k : 1;
    const int units = (nk + 15) / 16 + kWindowUnits;
    // One CTA per SM at a time.  Pick the split count S that minimises waves(S) * (units/S + prologue): more, smaller
    // splits fill partial waves better (64 groups: S=2 uses 128 of 148 SMs, S=9 fills 3.9 of 4 waves) but pay the LUT
    // build once per CTA.  Never fewer than 128 coded tokens per split.
    int max_s = (nk + 127) / 128;
    if (max_s < 1) max_s = 1;
    if (max_s > 192) max_s = 192;   // one group on 148 SMs (KV-head sharding at batch 1) wants one split per SM
    int best_s = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= max_s; ++s) {
        const int ctas = groups * s;
        const int waves = (ctas + sms - 1) / sms;
        const double cost = (double)waves * ((double)((units + s - 1) / s) + kPrologueUnits);
        if (cost < best_cost * 0.995) { best_cost = cost; best_s = s; }
    }
    return best_s;
}

int64_t million_pq_decode_attn_workspace_bytes(int bs, int nh, int nh_k, int d, int max_splits) {
    const int64_t counters = ((int64_t)bs * nh_k * 4 + 255) / 256 * 256;
    return counters + (int64_t)bs * nh * (max_splits + 2) * part_stride(d) * 4;   // +2: window part (generic) / boundary pieces (flat)
}

int million_pq_decode_attn(const million_attn_params* p, million_stream_t stream) {
    MILLION_REQUIRE(p != nullptr, "attn: null params");
    MILLION_REQUIRE(p->struct_size == sizeof(million_attn_params), "attn: struct_size %u != %zu (ABI mismatch)",
                    p->struct_size, sizeof(million_attn_params));
    MILLION_REQUIRE(p->io_dtype == MILLION_F16 || p->io_dtype == MILLION_BF16, "attn: io_dtype must be f16 or bf16");
    MILLION_REQUIRE(p->bs >= 0 && p->nh > 0 && p->nh_k > 0 && p->nh % p->nh_k == 0, "attn: bad head counts");
    const int code_bytes = p->code_bytes == 0 ? 1 : p->code_bytes;
    MILLION_REQUIRE(code_bytes == 1 || code_bytes == 2, "attn: code_bytes must be 1 (uint8) or 2 (uint16)");
    MILLION_REQUIRE(p->M > 0 && p->d % p->M == 0 && p->C > 1 && p->C <= (code_bytes == 1 ? 256 : 65536), "attn: bad d/M/C (%d/%d/%d) for %d-byte codes", p->d, p->M, p->C, code_bytes);
    MILLION_REQUIRE(p->nk >= 0 && p->r >= 0 && p->r <= p->res_len, "attn: bad nk/r (nk=%d r=%d res_len=%d)", p->nk, p->r, p->res_len);
    if (p->bs == 0) return MILLION_OK;
    const bool partial_only = (p->flags & MILLION_ATTN_PARTIAL_ONLY) != 0;
    const bool fused_splitkv = (p->flags & MILLION_ATTN_FUSED_SPLITKV) != 0;
    MILLION_REQUIRE(!(partial_only && fused_splitkv), "attn: PARTIAL_ONLY and FUSED_SPLITKV exclude each other");
    if (fused_splitkv) MILLION_REQUIRE(p->out != nullptr, "attn: fused split-KV writes the merged result to `out`");
    if (fused_splitkv) MILLION_REQUIRE(p->p2p_state != nullptr && ((uintptr_t)p->p2p_state & 15) == 0, "attn: fused split-KV needs the (16-byte aligned) state block of million_splitkv_state_init");
    MILLION_REQUIRE(p->q && p->k_cent && p->v_cent && p->workspace, "attn: null pointer");
    MILLION_REQUIRE(partial_only ? (p->partial != nullptr) : (p->out != nullptr), "attn: missing output pointer");
    MILLION_REQUIRE(p->nk == 0 || (p->k_codes && p->v_codes), "attn: null code pointer");
    MILLION_REQUIRE((p->r == 0 && p->r_dev == nullptr) || (p->k_res && p->v_res), "attn: null residual pointer");
    MILLION_REQUIRE((p->k_new == nullptr) == (p->v_new == nullptr), "attn: k_new and v_new go together");
    MILLION_REQUIRE(p->k_new == nullptr || p->r >= 1, "attn: with k_new/v_new, r counts the new token (r >= 1)");
    MILLION_REQUIRE(p->k_new == nullptr || (p->d * 2) % 16 == 0, "attn: fused window append needs 16-byte rows");
    MILLION_REQUIRE(p->v_layout >= MILLION_V_ROWMAJOR && p->v_layout <= MILLION_V_PAGED, "attn: bad v_layout");
    if (p->v_layout == MILLION_V_PAGED && p->nk > 0) {
        MILLION_REQUIRE(p->v_page_ids && p->page_size > 0, "attn: paged V needs page ids and page_size");
        MILLION_REQUIRE((int64_t)p->n_pages * p->page_size >= p->nk, "attn: %d pages of %d < nk=%d", p->n_pages, p->page_size, p->nk);
    }
    if (p->nk > 0) MILLION_REQUIRE(p->k_head_stride >= (int64_t)p->nk * p->M * code_bytes, "attn: k_head_stride too small");

    int S = p->n_splits > 0 ? p->n_splits : million_pq_decode_attn_default_splits(p->bs, p->nh_k, p->nk);
    MILLION_REQUIRE(p->workspace_bytes >= million_pq_decode_attn_workspace_bytes(p->bs, p->nh, p->nh_k, p->d, S),
                    "attn: workspace too small for %d splits", S);
    MILLION_REQUIRE(((uintptr_t)p->workspace & 15) == 0, "attn: workspace must be 16-byte aligned");

    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = p->q; a.k_codes = p->k_codes; a.v_codes = p->v_codes; a.v_page_ids = p->v_page_ids;
    a.k_cent = p->k_cent; a.v_cent = p->v_cent; a.k_res = p->k_res; a.v_res = p->v_res;
    a.out = p->out; a.partial_out = partial_only ? p->partial : nullptr;
    a.counters = reinterpret_cast<int*>(p->workspace);
    a.parts = reinterpret_cast<float*>(reinterpret_cast<char*>(p->workspace) + ((int64_t)p->bs * p->nh_k * 4 + 255) / 256 * 256);
    a.k_head_stride = p->k_head_stride; a.v_head_stride = p->v_head_stride; a.v_ld = p->v_ld;
    a.bs = p->bs; a.nh = p->nh; a.nh_k = p->nh_k; a.d = p->d; a.M = p->M; a.C = p->C; a.nk = p->nk; a.r = p->r;
    a.res_len = p->res_len; a.v_layout = p->v_layout; a.page_size = p->page_size; a.n_pages = p->n_pages;
    a.n_splits = S;
    a.auto_splits = p->n_splits <= 0;
    a.ws_parts = (int)((p->workspace_bytes - (((int64_t)p->bs * p->nh_k * 4 + 255) / 256 * 256)) / ((int64_t)p->bs * p->nh * part_stride(p->d) * 4));
    const int units = (p->nk + 15) / 16;
    a.units_per_split = ((units + S - 1) / S + 3) / 4 * 4;   // splits start on multiples of 64 tokens (tiles never straddle a page)
    if (a.units_per_split < 4) a.units_per_split = 4;
    MILLION_REQUIRE(S + 1 <= 1024, "attn: at most 1023 splits");
    a.scale_log2 = kLog2e / sqrtf((float)p->d);
    if (fused_splitkv) { a.partial_out = reinterpret_cast<float*>(p->p2p_state); a.p2p = 1; }
    a.pdl = (p->flags & MILLION_ATTN_PDL) != 0;
    if (p->nk > 0 && p->k_out > 0 && p->k_out_idx && p->k_out_val) {
        MILLION_REQUIRE(p->k_out <= MILLION_MAX_OUTLIERS && p->d <= 256, "attn: k_out %d > %d or d > 256", p->k_out, MILLION_MAX_OUTLIERS);
        MILLION_REQUIRE(p->k_out_head_stride >= (int64_t)p->nk * p->k_out, "attn: k_out_head_stride too small");
        a.k_out = p->k_out; a.ko_idx = p->k_out_idx; a.ko_val = p->k_out_val; a.ko_head_stride = p->k_out_head_stride;
    }
    if (p->nk > 0 && p->v_out > 0 && p->v_out_idx && p->v_out_val) {
        MILLION_REQUIRE(p->v_out <= MILLION_MAX_OUTLIERS && p->d <= 256, "attn: v_out %d > %d or d > 256", p->v_out, MILLION_MAX_OUTLIERS);
        MILLION_REQUIRE(p->v_out_head_stride >= (int64_t)p->nk * p->v_out, "attn: v_out_head_stride too small");
        a.v_out = p->v_out; a.vo_idx = p->v_out_idx; a.vo_val = p->v_out_val; a.vo_head_stride = p->v_out_head_stride;
    }
    a.k_new = p->k_new; a.v_new = p->v_new; a.r_dev = p->r_dev;
    a.code_bytes = code_bytes;
#ifdef MILLION_DEBUG
    a.dbg_timing = g_dbg_timing;
    a.dbg_mode = g_dbg_mode;
#endif

    cudaStream_t st = (cudaStream_t)stream;
    const bool fast_ok = p->impl != MILLION_IMPL_GENERIC &&
                         (p->impl == MILLION_IMPL_FAST || launch_attn_fast(a, p->io_dtype, p->prepared_codebook, st, true) == MILLION_OK);
    if (fast_ok) return launch_attn_fast(a, p->io_dtype, p->prepared_codebook, st, false);   // appends the new token itself
    if (fused_splitkv) MILLION_UNSUPPORTED("attn: the fused split-KV exchange exists in the fast kernel only (M=64, nh/nh_k = 4, row-major V)");
    if (a.k_new) {
        // the all-shapes kernel reads the window only: append with a copy launch first (same device-resident row index)
        int rc = launch_window_append_dev(const_cast<void*>(p->k_res), const_cast<void*>(p->v_res), (int64_t)p->res_len * p->d * 2, p->k_new, p->v_new,
                                          p->bs * p->nh_k, p->d * 2, p->r_dev, p->r - 1, p->res_len, st);
        if (rc != MILLION_OK) return rc;
        a.k_new = a.v_new = nullptr;
    }
    return launch_attn_generic(a, p->io_dtype, st);
}

#ifdef MILLION_DEBUG
/* debug hooks (only in -DMILLION_DEBUG builds, not part of the public header): ablation switch and per-CTA phase time stamps */
void million_debug_set_mode(int m) { g_dbg_mode = m; }
void million_debug_set_timing_buffer(void* buf) { g_dbg_timing = reinterpret_cast<unsigned long long*>(buf); }
#endif

int million_counter_add(int32_t* ctr, int n, int delta, million_stream_t stream) {
    MILLION_REQUIRE(ctr != nullptr && n >= 0, "counter_add: bad arguments");
    return launch_counter_add(ctr, n, delta, (cudaStream_t)stream);
}

int million_rope_qk(const void* q, const void* k, const void* cos_, const void* sin_, void* q_out, void* k_out, int dtype, int bs, int nh,
                    int nh_k, int d, million_stream_t stream) {
    MILLION_REQUIRE(q && k && cos_ && sin_ && q_out && k_out, "rope_qk: null pointer");
    MILLION_REQUIRE(dtype >= MILLION_F16 && dtype <= MILLION_F32, "rope_qk: bad dtype");
    MILLION_REQUIRE(bs >= 0 && nh > 0 && nh_k > 0 && d > 0 && d % 2 == 0, "rope_qk: bad shape (d must be even)");
    return launch_rope_qk(q, k, cos_, sin_, q_out, k_out, dtype, bs, nh, nh_k, d, (cudaStream_t)stream);
}

int million_lse_merge(const float* parts, int n_parts, int64_t n_rows, int d, void* out, int io_dtype, million_stream_t stream) {
    MILLION_REQUIRE(parts && out && n_parts > 0 && n_rows >= 0 && d > 0, "lse_merge: bad arguments");
    MILLION_REQUIRE(io_dtype >= MILLION_F16 && io_dtype <= MILLION_F32, "lse_merge: bad dtype");
    return launch_lse_merge(parts, n_parts, n_rows, d, out, io_dtype, (cudaStream_t)stream);
}

int million_window_append(void* k_win, void* v_win, int64_t win_head_stride, const void* k_src, const void* v_src,
                          int64_t src_head_stride, int n_heads, int r0, int n, int d, int dtype, million_stream_t stream) {
    MILLION_REQUIRE(k_win && v_win && k_src && v_src, "window_append: null pointer");
    MILLION_REQUIRE(n_heads >= 0 && r0 >= 0 && n >= 0 && d > 0, "window_append: bad sizes");
    const int eb = elem_bytes(dtype);
    return launch_rows_copy2(k_win, v_win, win_head_stride * eb, (int64_t)r0 * d * eb, k_src, v_src, src_head_stride * eb, 0,
                             n_heads, (int64_t)n * d * eb, (cudaStream_t)stream);
}

int million_window_shift(void* k_win, void* v_win, int64_t win_head_stride, int n_heads, int shift, int rem, int d, int dtype,
                         million_stream_t stream) {
    MILLION_REQUIRE(k_win && v_win, "window_shift: null pointer");
    MILLION_REQUIRE(n_heads >= 0 && shift > 0 && rem >= 0 && d > 0, "window_shift: bad sizes");
    const int eb = elem_bytes(dtype);
    // chunks of at most `shift` rows never overlap their own source; stream order keeps them sequential
    for (int done = 0; done < rem; done += shift) {
        const int n = (rem - done) < shift ? (rem - done) : shift;
        int rc = launch_rows_copy2(k_win, v_win, win_head_stride * eb, (int64_t)done * d * eb, k_win, v_win, win_head_stride * eb,
                                   (int64_t)(done + shift) * d * eb, n_heads, (int64_t)n * d * eb, (cudaStream_t)stream);
        if (rc != MILLION_OK) return rc;
    }
    return MILLION_OK;
}

}  // extern "C"
