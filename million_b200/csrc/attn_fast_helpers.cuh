// attn_fast_helpers.cuh — small device helpers shared by the fast decode-attention kernels (attn_fast.cu: M=64, d_m=2;
// attn_fast_dm4.cu: M=32, d_m=4).
#pragma once
#include "attn_common.cuh"

namespace million {

namespace fast {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTile = 32;                       // tokens per warp tile
constexpr int kRowBytes = 64;                   // M = 64 one-byte codes
constexpr int kStageBytes = 2 * kTile * kRowBytes;   // one K tile + one V tile per warp
constexpr int kVtabBytes = 64 * 1024;
#ifndef MILLION_FLAT_PAD
#define MILLION_FLAT_PAD 24
#endif
constexpr int kFlatPad = MILLION_FLAT_PAD;      // cost of one more (prologue + epilogue) in 64-token units
#ifndef MILLION_RESCALE_MARGIN
#define MILLION_RESCALE_MARGIN 6.f
#endif
constexpr float kRescaleMargin = MILLION_RESCALE_MARGIN;   // log2 units: p <= 64 before a rescale is forced

template <int G> struct LutCfg;
template <> struct LutCfg<4> { static constexpr int bytes = 128 * 1024; };
template <> struct LutCfg<2> { static constexpr int bytes = 64 * 1024; };
template <> struct LutCfg<1> { static constexpr int bytes = 64 * 1024; };

// column of sub-space m inside a 64-entry table row (both K tables and the V table)
__host__ __device__ __forceinline__ constexpr int col_of(int m) {
    const int W = m >> 2, b = m & 3;
    return (b >> 1) * 32 + (b & 1) * 16 + W;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// plain shared-memory loads through the generic pointer of the dynamic smem block: the compiler is free to batch them
// (asm volatile loads would be kept in program order and serialise on the 30-cycle LDS latency)
__device__ __forceinline__ uint32_t lds32(const unsigned char* base, uint32_t off) { return *reinterpret_cast<const uint32_t*>(base + off); }
__device__ __forceinline__ uint2 lds64(const unsigned char* base, uint32_t off) { return *reinterpret_cast<const uint2*>(base + off); }
__device__ __forceinline__ uint4 lds128(const unsigned char* base, uint32_t off) { return *reinterpret_cast<const uint4*>(base + off); }
// acc_lo += fp16(packed.lo), acc_hi += fp16(packed.hi) in fp32: Blackwell FHADD (PTX add.rn.f32.f16), one op each
__device__ __forceinline__ void fhadd2(float& acc_lo, float& acc_hi, uint32_t packed) {
    const unsigned short lo = (unsigned short)(packed & 0xffffu), hi = (unsigned short)(packed >> 16);
    asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(acc_lo) : "h"(lo));
    asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(acc_hi) : "h"(hi));
}
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

// K-side outlier records of one token (k_out <= 4): dims packed one per byte, deltas two per word.  Every width is ONE load whose
// destination IS the caller's 32-bit variable (ld.u8 / ld.u16 zero-extend into a .b32 register): a conversion or move behind the
// load would make the warp wait for the HBM round trip on the spot instead of one tile later (measured: +20 % per launch).
// pair1 (k_out == 1 on an even-aligned store): the caller points at the EVEN token of the lane's pair and the two records come in
// with the loads of the k_out == 2 case — two neighbouring lanes read the same halfword / word; the consumers then select by token
// parity.  The single-record byte + halfword loads were 20 % slower per launch than two records per token.
__device__ __forceinline__ void ko_load(const uint8_t* idx, const unsigned short* vals, int k_out, bool pair1, uint32_t& dims, uint32_t& v01, uint32_t& v23) {
    // EVERY path defines all three registers with a load.  Leaving v23 untouched where it is unused (k_out <= 2) made ptxas
    // restore it with a register move placed after the loads; that move waits on the loads' scoreboard (shared by all of these
    // loads), i.e. on the loads just issued — one exposed DRAM round trip per tile (M=32, 2 records: 120 -> 165 us per launch;
    // found with tools/sass_sb_check.py).  The extra halfword load hits the sector the neighbouring load brings in.
    if (k_out == 2 || pair1) {
        asm("ld.global.nc.u16 %0, [%1];" : "=r"(dims) : "l"(idx));
        asm("ld.global.nc.u32 %0, [%1];" : "=r"(v01) : "l"(vals));
        asm("ld.global.nc.u16 %0, [%1];" : "=r"(v23) : "l"(vals));           // unused value
    } else if (k_out == 4) {
        asm("ld.global.nc.u32 %0, [%1];" : "=r"(dims) : "l"(idx));
        asm("ld.global.nc.u32 %0, [%1];" : "=r"(v01) : "l"(vals));
        asm("ld.global.nc.u32 %0, [%1+4];" : "=r"(v23) : "l"(vals));
    } else if (k_out == 1) {
        asm("ld.global.nc.u8 %0, [%1];" : "=r"(dims) : "l"(idx));
        asm("ld.global.nc.u16 %0, [%1];" : "=r"(v01) : "l"(vals));
        asm("ld.global.nc.u8 %0, [%1];" : "=r"(v23) : "l"(vals));            // unused value
    } else {   // 3 records: byte-wise (the packing below does wait for the loads: the slow but complete case)
        dims = (uint32_t)__ldg(idx) | ((uint32_t)__ldg(idx + 1) << 8) | ((uint32_t)__ldg(idx + 2) << 16);
        v01 = (uint32_t)__ldg(vals) | ((uint32_t)__ldg(vals + 1) << 16);
        v23 = __ldg(vals + 2);
    }
}

// Hand-over of a prefetched record from the load registers to the registers the tile's consumers read, and the ordering token
// for the NEXT prefetch.  All record loads of a warp share one hardware scoreboard, and a scoreboard wait is a wait for
// EVERYTHING outstanding on it: whatever touches a load register after the next prefetch has been issued waits for that
// prefetch's round trip too.  ptxas is free to sink plain copies below the next loads (it did: 120 -> 146-165 us per launch
// on the M=32 kernel with two records per token), so the copies are XORs with a zero the compiler cannot see (z = AttnArgs::zero),
// and the next prefetch's token index adds rec_dep() of them — a true data dependence: wait, copy, then issue.
__device__ __forceinline__ uint32_t rec_take(uint32_t loaded, uint32_t z) { return loaded ^ z; }
__device__ __forceinline__ int rec_dep(uint32_t a, uint32_t b, uint32_t c, uint32_t z) { return (int)((a | b | c) & z); }

// which records of a store can be fetched as aligned pairs (kernel-uniform)
__device__ __forceinline__ bool ko_pair1_ok(int k_out, const uint8_t* idx, const void* val, int64_t head_stride) {
    return k_out == 1 && ((reinterpret_cast<uintptr_t>(idx) | (uintptr_t)head_stride) & 1) == 0 && (reinterpret_cast<uintptr_t>(val) & 3) == 0;
}
}  // namespace fast

namespace fast {

// Per-warp online-softmax state (uniform across the lanes of the warp) and accumulators.
template <int G>
struct WarpState {
    float m[G];          // running max, log2 units (scaled logits)
    float l[G];          // per-LANE partial denominator (summed across lanes at the end)
    float o[4][G][2];    // per-lane fp32 output accumulators: slot s -> sub-space 4*l' + ((s + hw) & 3), 2 dims
};

template <int G>
__device__ __forceinline__ void state_init(WarpState<G>& st) {
#pragma unroll
    for (int g = 0; g < G; ++g) { st.m[g] = -INFINITY; st.l[g] = 0.f; }
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int g = 0; g < G; ++g) { st.o[b][g][0] = 0.f; st.o[b][g][1] = 0.f; }
}

// Table gathers with an absolute shared-window address: the PRMT result (code << 8 | column offset) is the register part,
// the table base is an immediate, so a gather is exactly PRMT + LDS.  Not volatile: the tables are read-only in the main
// loop and the compiler may schedule these loads freely.  kSmemBase is checked at kernel entry.
constexpr uint32_t kSmemBase = 0x400;   // dynamic shared memory starts after the 1 KB the driver reserves per CTA (no static smem)
template <uint32_t IMM>
__device__ __forceinline__ uint2 gather64(uint32_t r) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(r), "n"(IMM));
    return v;
}
template <uint32_t IMM>
__device__ __forceinline__ uint32_t gather32(uint32_t r) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(r), "n"(IMM));
    return v;
}

}  // namespace fast

}  // namespace million
