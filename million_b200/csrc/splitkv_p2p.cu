// splitkv_p2p.cu — cross-GPU merge of split-KV partial states over NVLink peer memory, fused into ONE small kernel:
// every rank stores its (rows, d+2) fp32 state [o un-normalised | m | l] straight into the receive slot it owns in every
// peer's symmetric buffer (P2P stores), publishes a sequence flag with release semantics at system scope, waits for the
// flags of all sources and applies the log-sum-exp merge (flash_decoding_reduce_kernel algebra,
// scripts/modeldb/bindings/Kernel.cuh:1249-1269).  Replaces NCCL all_gather + merge (2 launches + ~15 us of collective
// latency per layer) and, unlike NCCL on this pool, can be captured in a CUDA graph: the sequence number lives in device
// memory.  The reference has no multi-GPU path (scripts/modeldb/main_pq.py:74).
//
// Symmetric buffer layout: attn_common.cuh (flags of this kernel in [0, 1024), recv area at kP2PRecvOff).
// Double buffering by sequence parity is enough: a rank can publish call n+1 only after it finished call n (stream order),
// and a peer can start call n+2 only after it has seen everybody's call n+1.
#include <string.h>

#include "attn_common.cuh"

namespace million {

constexpr int kMaxWorld = 8;
struct P2PArgs {
    unsigned char* peer[kMaxWorld];   // symmetric buffer of every rank, mapped into this process
    const float* local;               // (rows, d+2) this rank's state
    void* out;                        // (rows, d)
    unsigned* counter;                // device: calls completed so far
    int* ticket;                      // device: zero between launches
    int* err;                         // device: set to 1 if a wait timed out
    const int* spin_limit;            // device: polls before a wait gives up (0 = default)
    int64_t rows;
    int rank, world, d;
};

template <typename T>
__global__ void __launch_bounds__(128) splitkv_push_merge_kernel(const P2PArgs a) {
    __shared__ unsigned seq_s;
    __shared__ int gave_up;
    __shared__ float wts[kMaxWorld], hd[2];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, stride = a.d + 2;
    pdl_launch_dependents();
    pdl_wait();                       // `local` is the stream predecessor's output
    if (tid == 0) { seq_s = *reinterpret_cast<volatile unsigned*>(a.counter) + 1; gave_up = 0; }
    __syncthreads();
    const unsigned seq = seq_s, par = seq & 1;
    const size_t slot_floats = (size_t)a.rows * stride;

    // ---- push my row into slot [par][rank] of every rank (my own included)
    const float* src = a.local + row * stride;
    for (int g = 0; g < a.world; ++g) {
        float* dst = reinterpret_cast<float*>(a.peer[g] + kP2PRecvOff) + ((size_t)par * a.world + a.rank) * slot_floats + row * stride;
        for (int i = tid; i < stride; i += blockDim.x) dst[i] = src[i];
    }
    // Ordering: my block's stores -> bar.sync -> acq_rel ticket at gpu scope -> (last block) release stores of the flags at
    // system scope -> the peers' acquire loads.  Causality is transitive across the two scopes, so no per-thread system fence.
    __shared__ int is_last;
    __syncthreads();
    if (tid == 0) {
        int t;
        asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(t) : "l"(a.ticket) : "memory");
        is_last = (t == (int)gridDim.x - 1);
        if (is_last) {
            *a.ticket = 0;
            *reinterpret_cast<volatile unsigned*>(a.counter) = seq;
        }
    }
    __syncthreads();
    if (is_last && tid < a.world) st_release_sys(reinterpret_cast<unsigned*>(a.peer[tid] + 128 * a.rank), seq);   // every block of this rank has pushed
    // ---- wait for the state of every source rank (bounded: a dead peer must not hang the GPU)
    if (tid < a.world) {
        const unsigned* f = reinterpret_cast<const unsigned*>(a.peer[a.rank] + 128 * tid);
        const int limit = *a.spin_limit > 0 ? *a.spin_limit : kP2PDefaultSpins;
        int spins = 0;
        while ((int)(ld_acquire_sys(f) - seq) < 0) {
            __nanosleep(100);
            if (++spins > limit) { *a.err = 1; gave_up = 1; break; }
        }
    }
    __syncthreads();

    // ---- merge
    const float* base = reinterpret_cast<const float*>(a.peer[a.rank] + kP2PRecvOff) + (size_t)par * a.world * slot_floats + row * stride;
    if (tid == 0) {
        float mstar = -INFINITY;
        for (int g = 0; g < a.world; ++g) {
            const float* p = base + (size_t)g * slot_floats;
            if (__ldcg(p + a.d + 1) > 0.f) mstar = fmaxf(mstar, __ldcg(p + a.d));
        }
        float den = 0.f;
        for (int g = 0; g < a.world; ++g) {
            const float* p = base + (size_t)g * slot_floats;
            const float l = __ldcg(p + a.d + 1);
            const float w = l > 0.f ? __expf(__ldcg(p + a.d) - mstar) : 0.f;
            wts[g] = w;
            den += l * w;
        }
        hd[0] = den;
    }
    __syncthreads();
    const float den = hd[0];
    for (int k = tid; k < a.d; k += blockDim.x) {
        float acc = 0.f;
        for (int g = 0; g < a.world; ++g) acc = fmaf(__ldcg(base + (size_t)g * slot_floats + k), wts[g], acc);
        // a wait that gave up must not pass stale or partial rows on as a result: poison them (the error word is set as well)
        reinterpret_cast<T*>(a.out)[row * a.d + k] = io<T>::from_f(gave_up ? __int_as_float(0x7fc00000) : (den > 0.f ? acc / den : 0.f));
    }
}

int launch_splitkv_push_merge(const P2PArgs& a, int io_dtype, bool pdl, cudaStream_t stream) {
    if (a.rows == 0) return MILLION_OK;
    dim3 grid((unsigned)a.rows), block(128);
    if (io_dtype == MILLION_F16) MILLION_CUDA_OK(launch_kernel(splitkv_push_merge_kernel<__half>, grid, block, 0, stream, pdl, a));
    else if (io_dtype == MILLION_BF16) MILLION_CUDA_OK(launch_kernel(splitkv_push_merge_kernel<__nv_bfloat16>, grid, block, 0, stream, pdl, a));
    else MILLION_CUDA_OK(launch_kernel(splitkv_push_merge_kernel<float>, grid, block, 0, stream, pdl, a));
    return MILLION_OK;
}

}  // namespace million

using namespace million;

extern "C" {

int64_t million_splitkv_symmetric_bytes(int world, int64_t rows, int d) {
    return p2p_ll_offset(world, rows, d) + (int64_t)2 * world * rows * (d + 2) * 8;      // flags | fp32 receive area | tagged 8-byte cells
}

// every block of the exchange kernel spins until all ranks have published: the whole grid must be co-resident
// (148 SMs x 16 blocks of 128 threads on B200; a quarter of that leaves room for whatever else is running)
static const int64_t kMaxExchangeRows = 512;

int64_t million_splitkv_state_bytes(void) { return (int64_t)sizeof(P2PState); }

int million_splitkv_set_timeout(void* state, int64_t microseconds, million_stream_t stream) {
    MILLION_REQUIRE(state != nullptr && microseconds >= 0, "splitkv_set_timeout: bad arguments");
    int64_t spins = microseconds * 10;          // one poll per 100 ns
    if (spins > 0x7fffffff) spins = 0x7fffffff;
    const int v = (int)spins;
    MILLION_CUDA_OK(cudaMemcpyAsync(&reinterpret_cast<P2PState*>(state)->spin_limit, &v, sizeof v, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return MILLION_OK;
}

int million_splitkv_state_init(void* state, void* const* peer_bases_host, int rank, int world, int64_t rows, million_stream_t stream) {
    MILLION_REQUIRE(state && peer_bases_host, "splitkv_state_init: null pointer");
    MILLION_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "splitkv_state_init: world must be 1..8");
    P2PState h;
    memset(&h, 0, sizeof h);
    MILLION_REQUIRE(rows > 0 && rows < (1ll << 31), "splitkv_state_init: bad rows");
    h.rank = rank; h.world = world; h.rows = (int)rows;
    for (int g = 0; g < world; ++g) h.peer[g] = (unsigned char*)peer_bases_host[g];
    // pageable source: the copy is staged by the runtime before the call returns, so the stack object may go away
    MILLION_CUDA_OK(cudaMemcpyAsync(state, &h, sizeof h, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return MILLION_OK;
}

/* state: device memory of 16 bytes, zero-initialised by the caller once: [counter u32 | ticket i32 | err i32 | spin limit i32] */
int million_splitkv_push_merge(const float* local_partial, void* const* peer_bases_host, int rank, int world, int64_t rows, int d,
                               void* out, int io_dtype, void* state, int flags, million_stream_t stream) {
    MILLION_REQUIRE(local_partial && peer_bases_host && out && state, "splitkv_push_merge: null pointer");
    MILLION_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "splitkv_push_merge: world must be 1..8");
    MILLION_REQUIRE(rows >= 0 && d > 0, "splitkv_push_merge: bad sizes");
    MILLION_REQUIRE(rows <= kMaxExchangeRows, "splitkv_push_merge: at most %lld (batch, head) rows per call (every block waits for the peers: the grid must be co-resident)", (long long)kMaxExchangeRows);
    MILLION_REQUIRE(io_dtype >= MILLION_F16 && io_dtype <= MILLION_F32, "splitkv_push_merge: bad dtype");
    P2PArgs a;
    for (int g = 0; g < kMaxWorld; ++g) a.peer[g] = g < world ? (unsigned char*)peer_bases_host[g] : nullptr;
    a.local = local_partial; a.out = out;
    a.counter = (unsigned*)state; a.ticket = (int*)state + 1; a.err = (int*)state + 2; a.spin_limit = (const int*)state + 3;
    a.rows = rows; a.rank = rank; a.world = world; a.d = d;
    return launch_splitkv_push_merge(a, io_dtype, (flags & MILLION_ATTN_PDL) != 0, (cudaStream_t)stream);
}

}  // extern "C"
