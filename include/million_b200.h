/* million_b200.h — C ABI of libmillion_b200.so: the B200-native (sm_100a) implementation of
 * MILLION's product-quantized KV-cache hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one native/third-party call of the
 * reference (paths relative to the reference checkout, Zhaohui-Xu/MILLION):
 *
 *   million_pq_encode            scripts/utils/pq_utils.py:451-499  sa_encode_4d_keops (pykeops arg-min),
 *                                call sites pq_utils.py:235-236, 292-293; paged_pq_utils.py:161-167, 241-242
 *   million_pq_encode_paged      + the page write of scripts/utils/dynamic_paged_pq_utils.py:768-811
 *   million_pq_decode            scripts/utils/pq_utils.py:501-540  sa_decode_4d (torch.gather)
 *   million_pq_decode_attn       scripts/modeldb/bindings/Interface.cu:16-120
 *                                flash_decoding_allocated_buffer<...> (at::matmul LUT + split + residual +
 *                                reduce kernels, bindings/Kernel.cuh:11-166, 1038-1209, 1211-1270), bound
 *                                to Python at bindings/bindings.template.cpp:50-62; and the paged variant
 *                                the live cache asks for at scripts/utils/paged_pq_utils.py:547, 621-635
 *   million_lse_merge            scripts/modeldb/bindings/Kernel.cuh:1211-1270 flash_decoding_reduce_kernel
 *                                (used here for the cross-GPU split-KV merge; the reference has no multi-GPU)
 *   million_window_append        scripts/utils/pq_utils.py:304-311 (copy of the new k/v into the fp16 window)
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current CUDA device unless
 *     its name ends in _host; no torch types; no exceptions cross this boundary.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *     The reference launches on the legacy default stream (Interface.cu:65,88,109); callers here pass the
 *     current torch stream.
 *   - return value: MILLION_OK or a million_status; the reference printed and exit()ed on a launch error
 *     (Interface.cu:3-11), we return the code and keep a thread-local message (million_last_error()).
 *   - inputs are borrowed, outputs/scratch are caller-owned.  Nothing is allocated by the library except
 *     small per-device constants on first use.
 *   - there is no CPU fallback: without a CUDA device every compute call returns MILLION_ERR_CUDA.
 */
#ifndef MILLION_B200_H_
#define MILLION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MILLION_ABI_VERSION 8

typedef void* million_stream_t; /* cudaStream_t */

enum million_dtype { MILLION_F16 = 0, MILLION_BF16 = 1, MILLION_F32 = 2 };

enum million_status {
    MILLION_OK = 0,
    MILLION_ERR_INVALID = 1,     /* bad argument (null pointer, inconsistent sizes) */
    MILLION_ERR_UNSUPPORTED = 2, /* shape/dtype combination not compiled */
    MILLION_ERR_CUDA = 3         /* CUDA runtime error; see million_last_error() */
};

/* value_codes layouts understood by million_pq_decode_attn / written by the encoders */
enum million_v_layout {
    MILLION_V_ROWMAJOR = 0,   /* (bs, nh_k, nk, M)           pq_utils.py:117-125 */
    MILLION_V_TRANSPOSED = 1, /* (bs, nh_k, M, ld>=nk)       paged_pq_utils.py:77-80 (live PagedPQCache) */
    MILLION_V_PAGED = 2       /* pool (pages, M, page_size) + block table (bs, nh_k, n_pages) int64
                                 dynamic_paged_pq_utils.py:46-48, 453-456; paged_pq_utils.py:621-635 */
};

/* which implementation to run; AUTO picks the fastest one that supports the shape */
enum million_impl {
    MILLION_IMPL_AUTO = 0,    /* encode: expects the tables of million_pq_encoder_auto_prepare in `prepared_encoder` */
    MILLION_IMPL_GENERIC = 1, /* plain SIMT kernels, every shape */
    MILLION_IMPL_FAST = 2,    /* decode: conflict-free shared-memory LUT gathers (attn_fast.cu); encode: tcgen05 distances */
    MILLION_IMPL_GRID = 3     /* encode only, d/M = 2: exact candidate-grid search (encode_grid.cu); `prepared_encoder` must then
                                 point at the tables of million_pq_encoder_grid_prepare */
};

int million_abi_version(void);
const char* million_last_error(void);
/* sm_count / cc of the current device; MILLION_ERR_CUDA when there is none */
int million_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Encoder: codes[v, m] = argmin_c sum_k (fp32(x[v, m*d_m + k]) - cent[m, c, k])^2, first minimum wins.
 * Replaces sa_encode_4d_keops (pq_utils.py:451-499).
 *
 *   x         n_heads blocks of n_tokens rows of d values (x_dtype), block stride x_head_stride ELEMENTS
 *             (so a slice [:, :, :64, :] of a (bs, nh_k, Lt, d) window can be encoded in place)
 *   cent      (M, C, d/M) fp32, contiguous (the reference casts to fp32 too, pq_utils.py:484)
 *   codes     element (head, t, m) is written at
 *                 codes + head*codes_head_stride + (t0 + t)*codes_token_stride + m*codes_m_stride   (ELEMENTS)
 *             row-major cache (bs,nh_k,cap,M): token_stride=M, m_stride=1; transposed (bs,nh_k,M,cap):
 *             token_stride=1, m_stride=cap.  code_bytes = 1 (C<=256) or 2 (C<=65536, nbits2dtype).
 */
/* Optional operand tiles for the tensor-core (tcgen05) encoder: the codebook as [-2c | |c|^2 split] rows in the
 * input format, per-sub-space max |c|^2, and a representability check (synchronous, one-time per codebook and input
 * dtype).  Returns MILLION_ERR_UNSUPPORTED when the shape is not covered (needs C=256, d/M in {2,4}, M % 16 == 0,
 * fp16/bf16 input) or the centroids are not exactly representable in the input format; callers then pass NULL and
 * the exact CUDA-core encoder runs. */
int64_t million_pq_encoder_prepared_bytes(int d, int M, int C);
int million_pq_encoder_prepare(const float* cent, int x_dtype, int d, int M, int C, void* prepared, million_stream_t stream);

/* Tables of the candidate-grid encoder (two-dimensional sub-spaces, d = 2M, C <= 256): per sub-space a 64 x 64 grid over the
 * centroids with, per cell, every centroid that can be nearest to a point of the cell.  Exact: codes are bit-identical to the
 * brute-force arg-min.  Asynchronous on `stream`, no host synchronisation; 0 bytes / MILLION_ERR_UNSUPPORTED for other shapes.
 * Pass the buffer as `prepared_encoder` together with impl = MILLION_IMPL_GRID. */
int64_t million_pq_encoder_grid_prepared_bytes(int d, int M, int C);
int million_pq_encoder_grid_prepare(const float* cent, int d, int M, int C, void* prepared, million_stream_t stream);

/* The tables MILLION_IMPL_AUTO wants in `prepared_encoder` for a shape: the candidate grid where it applies (d = 2M, C <= 256,
 * one-byte codes), else the tensor-core tiles (same return codes as million_pq_encoder_prepare).  A host that passes these with
 * impl = AUTO gets the fastest exact encoder without knowing which one that is; NULL runs the exact CUDA-core encoder. */
int64_t million_pq_encoder_auto_prepared_bytes(int d, int M, int C);
int million_pq_encoder_auto_prepare(const float* cent, int x_dtype, int d, int M, int C, void* prepared, million_stream_t stream);

int million_pq_encode(const void* x, int x_dtype, int64_t x_head_stride,
                      const float* cent, const void* prepared_encoder /* may be NULL */,
                      void* codes, int code_bytes,
                      int64_t codes_head_stride, int64_t codes_token_stride, int64_t codes_m_stride,
                      int64_t t0,
                      int n_heads, int n_tokens, int d, int M, int C,
                      int impl, million_stream_t stream);

/* Same arg-min, written into a page pool (pages, M, page_size) uint8 through a block table:
 * token (t0+t) of head `head` goes to pool[page_ids[head*page_ids_head_stride + (t0+t)/page_size], m, (t0+t)%page_size].
 * Format: dynamic_paged_pq_utils.py:46-48, 768-811. */
int million_pq_encode_paged(const void* x, int x_dtype, int64_t x_head_stride,
                            const float* cent, const void* prepared_encoder /* may be NULL */,
                            uint8_t* page_pool, const int64_t* page_ids, int64_t page_ids_head_stride,
                            int page_size, int64_t t0,
                            int n_heads, int n_tokens, int d, int M, int C,
                            int impl, million_stream_t stream);

/* Outlier split (extension; no reference counterpart).  Per head-vector: the k_out entries of
 * largest |x| (ties: lowest dim, listed largest first) are zeroed in x_masked — which the caller then encodes with
 * million_pq_encode[_paged] — and recorded as (dim, delta) with delta = x[dim] - cent[m, code[m], k] rounded to x_dtype,
 * code[m] being the arg-min of the MASKED sub-vector (same arithmetic as the encoder, so it equals the stored code).
 *   x_masked   (n_heads, n_tokens, d) x_dtype, contiguous
 *   out_idx    record (head, t0 + t, i) at out_idx + head*out_head_stride + (t0 + t)*k_out + i   (uint8; d <= 256)
 *   out_val    same addressing, x_dtype */
int million_pq_outlier_split(const void* x, int x_dtype, int64_t x_head_stride, const float* cent,
                             void* x_masked, uint8_t* out_idx, void* out_val, int64_t out_head_stride, int64_t t0,
                             int n_heads, int n_tokens, int d, int M, int C, int k_out, million_stream_t stream);

/* out[head, t, idx[head, t0 + t, i]] += val[head, t0 + t, i]: the reconstruction with the side store applied
 * (after million_pq_decode).  out (n_heads, n_tokens, d) with head stride in ELEMENTS, dtype f16/bf16/f32 (val: f16/bf16
 * as val_dtype). */
int million_pq_outlier_apply(void* out, int dtype, int64_t out_head_stride, const uint8_t* idx, const void* val, int val_dtype,
                             int64_t store_head_stride, int64_t t0, int n_heads, int n_tokens, int d, int k_out,
                             million_stream_t stream);

/* Reconstruct: out[v, m*d_m + k] = cent[m, codes[v, m], k].  Replaces sa_decode_4d (pq_utils.py:501-540).
 * cent/out share `dtype` (the reference returns C's dtype).  codes addressed like in million_pq_encode. */
int million_pq_decode(const void* codes, int code_bytes,
                      int64_t codes_head_stride, int64_t codes_token_stride, int64_t codes_m_stride,
                      const void* cent, void* out, int dtype, int64_t out_head_stride,
                      int n_heads, int n_tokens, int d, int M, int C,
                      million_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Decode attention of ONE new query token per (batch, head) over nk PQ-coded tokens plus the r most
 * recent tokens kept in fp16/bf16 ("residual window"), no mask:
 *     s_j = scale * sum_m <q_h[m], Kcent[m, Kcode[b,hk,j,m]]>   (j <  nk)   hk = h / (nh/nh_k)
 *     s_j = scale * <q_h, Kres[b,hk,j-nk]>                       (j >= nk)
 *     o_h = softmax(s) . [decode(Vcode) ; Vres]
 * One launch does what Interface.cu:16-120 does with one cuBLAS call + three kernels.
 */
typedef struct million_attn_params {
    uint32_t struct_size; /* = sizeof(million_attn_params) */
    int32_t io_dtype;     /* MILLION_F16 | MILLION_BF16: dtype of q, cents, residuals, out */
    int32_t impl;         /* million_impl */
    int32_t flags;        /* MILLION_ATTN_* */

    int32_t bs, nh, nh_k, d, M, C;
    int32_t nk; /* quantized tokens per head, >= 0 */
    int32_t r;  /* valid residual rows, 0 <= r <= res_len (the reference requires r > 0) */

    const void* q; /* (bs, nh, d) — the reference's (bs, nh, 1, d) */
    const uint8_t* k_codes; /* (bs, nh_k, nk, M) row-major, uint8 (or uint16: code_bytes = 2) */
    int64_t k_head_stride;  /* BYTES between consecutive (b, hk) blocks (>= nk*M; lets a preallocated cache be used) */

    int32_t v_layout;       /* million_v_layout */
    int32_t page_size;      /* V_PAGED */
    const uint8_t* v_codes; /* ROWMAJOR/TRANSPOSED: first byte of head (0,0); PAGED: the page pool */
    int64_t v_head_stride;  /* BYTES between (b, hk) blocks (unused for PAGED) */
    int64_t v_ld;           /* TRANSPOSED: bytes between consecutive m rows */
    const int64_t* v_page_ids; /* PAGED: (bs, nh_k, n_pages) */
    int32_t n_pages;
    int32_t res_len; /* rows allocated per head in k_res/v_res (Lt) */

    const void* k_cent; /* (M, C, d/M) io_dtype */
    const void* v_cent;
    const void* k_res; /* (bs, nh_k, res_len, d) io_dtype; rows [0, r) valid */
    const void* v_res;

    void* out; /* (bs, nh, d) io_dtype; NULL allowed with MILLION_ATTN_PARTIAL_ONLY */

    /* scratch, caller-owned: million_pq_decode_attn_workspace_bytes(...) bytes, 16-byte aligned.  The first
     * 4*bs*nh_k bytes are arrival counters that must be zero before the FIRST call; the kernel leaves them
     * zero again. */
    void* workspace;
    int64_t workspace_bytes;
    int32_t n_splits; /* CTAs per (b, kv-head) over the token range; 0 = choose from the SM count */

    /* split-KV across GPUs: with MILLION_ATTN_PARTIAL_ONLY the merged but UN-normalised state of this rank
     * is written here as fp32 (bs, nh, d+2) = [o_unnormalised (d) | running max m | denominator l] and `out`
     * is not touched.  Merge ranks with million_lse_merge. */
    float* partial;

    /* fp16 gather tables made by million_pq_codebook_prepare() from (k_cent, v_cent); required by the FAST
     * implementation (AUTO falls back to GENERIC when NULL).  Must be re-made whenever the codebooks change. */
    const void* prepared_codebook;

    /* Outlier side store (extension; the reference has no outlier code).  Per coded
     * token and KV head, k_out (K) / v_out (V) records (dim:uint8, delta:io_dtype): the reconstruction the attention runs
     * on is decode(code) with x_hat[dim] += delta.  Layout (bs, nh_k, >= nk, k_out), head stride in RECORDS.
     * k_out = v_out = 0 (or NULL pointers) = disabled: the path is bit-identical to the one without the store. */
    int32_t k_out, v_out;
    const uint8_t* k_out_idx;
    const void* k_out_val;
    int64_t k_out_head_stride;
    const uint8_t* v_out_idx;
    const void* v_out_val;
    int64_t v_out_head_stride;

    /* Split-KV across GPUs fused into this launch (MILLION_ATTN_FUSED_SPLITKV): instead of returning the rank's partial state,
     * the last CTA of every (b, kv-head) group stores the group's rows straight into every peer's symmetric buffer over NVLink
     * as 8-byte cells {value, sequence number} (data and validity in one store: no fence, no flag), spins on the peers' cells of
     * ITS group (bounded) and writes the merged rows to `out` — one launch per layer, no cross-group serialisation, exchange
     * latency = one store's flight time.  Same buffers as million_splitkv_push_merge
     * (rows = bs * nh); `p2p_state` is the million_splitkv_state_bytes() block prepared once by million_splitkv_state_init.
     * Compiled for the fast kernel at M = 64, nh/nh_k = 4, row-major value codes, no side store; MILLION_ERR_UNSUPPORTED otherwise.
     * If a wait gives up (dead peer), the affected rows are NaN and the error word of the state block is set. */
    void* p2p_state;

    /* Window append fused into the attention launch (pq_utils.py:304-311 followed by :313-327, one launch instead of a copy
     * kernel + the attention): k_new / v_new (bs, nh_k, d) io_dtype are the key/value of the token being decoded.  The kernel
     * stores them into row r-1 of k_res / v_res and attends over them, so `r` COUNTS the new token.  NULL = the caller has
     * appended already (million_window_append). */
    const void* k_new;
    const void* v_new;
    /* Device-resident window length, for CUDA-graph replay of a decode step (the reference passes r by value,
     * pq_utils.py:92): when non-NULL the kernel uses r = min(*r_dev + r, res_len), read at kernel start, so a captured
     * launch with k_new/v_new and r = 1 means "append at row *r_dev, attend over *r_dev + 1 rows".  Bump the counter after the
     * last layer of a step with million_counter_add.  Host-side checks on r then apply to the offset only. */
    const int32_t* r_dev;
    /* Width of the codes in k_codes / v_codes / the page pool: 0 or 1 = uint8 (C <= 256), 2 = uint16 (the reference's
     * nbits2dtype for nbits > 8, pq_utils.py:542-552; C <= 65536).  Strides stay in BYTES.  Two-byte codes run on the all-shapes
     * kernel (the reference compiles no kernel for them at all: pq_utils.py:58-59). */
    int32_t code_bytes;
    int32_t reserved0;
} million_attn_params;

#define MILLION_MAX_OUTLIERS 8

#define MILLION_ATTN_PARTIAL_ONLY 1
#define MILLION_ATTN_FUSED_SPLITKV 2
/* Programmatic dependent launch: the kernel may start while its stream predecessor drains — table copies and the first code
 * tiles are requested at once, everything else (q, k_new/v_new, the window, the workspace, r_dev) is read after the
 * predecessor has completed.  Requirement: the predecessor does not WRITE k_codes / v_codes / page table / prepared_codebook
 * / outlier stores of this call.  Capturable in CUDA graphs. */
#define MILLION_ATTN_PDL 4

/* Codebook preparation for the FAST decode-attention kernel: both codebooks as fp16 pairs in the kernel's gather
 * order (d=128, M in {32, 64}, C=256 only: returns 0 bytes / MILLION_ERR_UNSUPPORTED otherwise).  One tiny launch; callers
 * cache the result per codebook (DynamicPQCache.set_cent does). */
int64_t million_pq_codebook_prepared_bytes(int d, int M, int C);
int million_pq_codebook_prepare(const void* k_cent, const void* v_cent, int dtype, int d, int M, int C, void* prepared,
                                million_stream_t stream);

int64_t million_pq_decode_attn_workspace_bytes(int bs, int nh, int nh_k, int d, int max_splits);
/* the split count AUTO would use (so callers can size the workspace) */
int million_pq_decode_attn_default_splits(int bs, int nh_k, int nk);
int million_pq_decode_attn(const million_attn_params* p, million_stream_t stream);

/* Merge n_parts partial states (n_parts, bs*nh, d+2) fp32 [o_unnorm | m | l] -> out (bs*nh, d) io_dtype.
 * Same algebra as flash_decoding_reduce_kernel (Kernel.cuh:1249-1269); empty parts (l = 0) are skipped. */
int million_lse_merge(const float* parts, int n_parts, int64_t n_rows, int d, void* out, int io_dtype,
                      million_stream_t stream);

/* Split-KV across GPUs without a collective library: every rank stores its partial state into the slot it owns in every
 * peer's symmetric buffer (NVLink P2P stores), publishes a flag, waits for all sources and merges — one launch, graph
 * capturable.  `peer_bases_host`: host array of `world` device pointers, the symmetric buffers of all ranks mapped into
 * this process (e.g. torch.distributed._symmetric_memory buffer_ptrs), each million_splitkv_symmetric_bytes() big and
 * zero-initialised once.  `state`: 16 bytes of zero-initialised device memory private to this rank (or the block of
 * million_splitkv_state_init).  rows <= 512: every block of the kernel waits for the peers, so the grid must be co-resident.
 * flags: 0 or MILLION_ATTN_PDL (`local_partial` is read after the stream predecessor has completed).
 * A wait that gives up (default ~0.8 s, million_splitkv_set_timeout) writes NaN rows and sets the error word state[2]. */
int64_t million_splitkv_symmetric_bytes(int world, int64_t rows, int d);
/* Protocol state of one rank: [calls completed u32 | ticket i32 | error i32 | spin limit i32 | rank i32 | world i32 | rows i32 | pad | peer pointers x 8].
 * million_splitkv_state_init zeroes it and records rank, world, rows (= bs * nh of the calls it will serve) and the peers'
 * symmetric buffers (asynchronous on `stream`);
 * million_splitkv_push_merge only needs its first 16 bytes. */
int64_t million_splitkv_state_bytes(void);
int million_splitkv_state_init(void* state, void* const* peer_bases_host, int rank, int world, int64_t rows, million_stream_t stream);
/* how long a wait for the peers may last before it gives up (0 = the default) */
int million_splitkv_set_timeout(void* state, int64_t microseconds, million_stream_t stream);
int million_splitkv_push_merge(const float* local_partial, void* const* peer_bases_host, int rank, int world, int64_t rows, int d,
                               void* out, int io_dtype, void* state, int flags, million_stream_t stream);

/* window[(head), r0 + i, :] = src[(head), i, :] for i < n — pq_utils.py:304-311, paged_pq_utils.py:377-378.
 * Both K and V in one launch.  Strides in ELEMENTS. */
int million_window_append(void* k_win, void* v_win, int64_t win_head_stride,
                          const void* k_src, const void* v_src, int64_t src_head_stride,
                          int n_heads, int r0, int n, int d, int dtype, million_stream_t stream);

/* ctr[i] += delta for i < n (device int32 counters, e.g. the r_dev of million_attn_params): one tiny launch, graph capturable. */
int million_counter_add(int32_t* ctr, int n, int delta, million_stream_t stream);

/* Producer step of a decode token: rotary position embedding of the new token's query and key rows in ONE launch.
 * Replaces transformers' apply_rotary_pos_emb as called by the reference's attention forward
 * (scripts/modeldb/models/modeling_llama.py:500-512; ~10 elementwise launches per layer):
 *     out = x * cos + rotate_half(x) * sin,   rotate_half(x) = cat(-x[d/2:], x[:d/2])
 * q (bs, nh, d), k (bs, nh_k, d), cos / sin (bs, d), all contiguous, dtype = F16 | BF16 | F32; q_out / k_out may alias q / k.
 * Each product and the sum are rounded to the dtype like the elementwise torch evaluation: results are bit-identical to it.
 * k_out can be passed straight to million_pq_decode_attn as k_new (the window append is fused there). */
int million_rope_qk(const void* q, const void* k, const void* cos, const void* sin, void* q_out, void* k_out,
                    int dtype, int bs, int nh, int nh_k, int d, million_stream_t stream);

/* window[:, 0:rem, :] = window[:, shift:shift+rem, :] (paged flush, paged_pq_utils.py:186-199) */
int million_window_shift(void* k_win, void* v_win, int64_t win_head_stride, int n_heads, int shift, int rem,
                         int d, int dtype, million_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MILLION_B200_H_ */
