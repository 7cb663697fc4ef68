// mrp_device.cuh — shared device-side types and helpers for the sm_100a feature kernels.
//
// Design in one paragraph (details in DESIGN.md): the reference evaluates, per snapshot, O(window)
// loops over three rolling deques (1800 mids/spreads, 1500 tick directions, 1500 trade directions;
// feature_processor.cpp:48-108,215-269).  Here every window quantity slides: snapshot t's value is
// snapshot t-1's plus the ticks that arrived minus the ticks that left the window, so each tick / trade
// is read from HBM once (kernels_window.cu).  The time axis of a stream is cut into row segments; one
// CTA streams through one segment, warmed by a halo of one window.
// All arithmetic is fp64, compiled with -fmad=false so a*b+c is rounded exactly as the reference's
// x86-64 SSE2 code rounds it; explicit fma() is used only inside scan carries where the reference
// has no counterpart.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mrp_fg.h"

namespace mrp {

struct DevInputs {
    int64_t n_rows;  // total rows of the batch == level / column stride
    const uint64_t* ts;
    const double* bid_px;
    const double* ask_px;
    const int32_t* bid_sz;
    const int32_t* ask_sz;
    const double* buy;
    const double* sell;
    const int32_t* adds;
    const uint32_t* kcur;
    const uint32_t* jcur;
    const double* mid;
    const double* spread;
    const int8_t* tdir;
};

struct DevOutputs {
    double* raw;     // [19][n_rows]
    double* norm;    // [19][n_rows]
    int8_t* bdir;    // [L][n_rows] or null
    int8_t* adir;    // [L][n_rows] or null
    int32_t* counts; // [MRP_FG_NUM_COUNTS][n_rows] or null
};

// One CTA's unit of work on the tick / trade / snapshot axes: rows [row_begin,row_end) of one stream.
struct DevSegment {
    int64_t row_begin, row_end;      // global row indices
    int64_t stream_row0, stream_row1; // the stream's global row range
    int64_t tick0, n_ticks;          // the stream's ticks: global offset and count
    int64_t trade0, n_trades;        // the stream's trades
};

// One CTA's unit of work in the normaliser: a tile of `n` rows starting at global row `row`.
struct DevTile {
    int64_t row;          // global first row of the tile
    int64_t stream_row0;  // global first row of the stream
    int32_t n;            // rows in this tile (<= tile_rows)
    int32_t index;        // tile index within the stream
    int32_t first_id;     // global id of the stream's tile 0
    int32_t stream;       // stream index within the batch
};


struct DevParams {
    int32_t L, mid_window, dir_window, spread_divisor, norm_window, return_lag, stress_levels;
    int32_t emit_counts;
    double ewm_alpha, ewm_decay;            // alpha = 2/(span+1), decay = 1-alpha
    double inv_spread_divisor;              // 1/spread_divisor
    double ofi_alpha, ofi_decay, ofi_min_total_vol;
    double stress_alpha, stress_decay, stress_min_size, stress_distance_decay;
    const double* ofi_weight;   // [L]  exp(-i*depth_decay), computed by the host libm
    const double* ewm_pow;      // [win_ring+1]    decay^n
    const double* ewm_ipow;     // [513]           decay^-n
    const double* ofi_pow;      // [257]           ofi_decay^n
    double ofi_pw2[6], stress_pw2[6];  // decay^(1,2,4,8,16,32) again, as kernel-parameter constants (the scan steps)
    const double* stress_pow;   // [257]           stress_decay^n
    const double* inv_cnt_norm; // [norm_window+1] 1/n, [0] = 0
    const double* clog2c;       // [max(dir_window, mid_window)+2]  c*log2(c), [0] = 0
    const double* inv_cnt;      // [max(mid_window, dir_window)+2]  1/n, [0] = 0
    // the window constants of the steady state (full windows), so that the common row needs no table look-up
    double inv_ret_full, pow_ret_full;      // 1/(mid_window-1), decay^(mid_window-1)
    double inv_dir_full, clog2c_dir_full;   // 1/dir_window, dir_window*log2(dir_window)
    // window kernel: entries of the tick / trade prefix rings and the most new ticks / trades one chunk may bring
    int32_t win_ring, win_chunk_ticks, win_trade_ring, win_chunk_trades;
    int32_t ema_warm_rows;      // halo rows that warm the two row-axis EMAs
    int32_t tile_rows;          // normaliser tile (divides norm_window, or == norm_window)
    int32_t tiles_back;         // norm_window / tile_rows
};

#ifndef MRP_BOOK_THREADS
#define MRP_BOOK_THREADS 512   // measured at 50 M rows, one CTA per SM: 384 threads (146 registers) 4.79 ms, 416 5.04, 448 4.82, 512 (123 registers, no
                               // spills) 4.38, 512 with three stages 4.45, 576 5.39, 640 5.08, 768 5.61.  (Round 1's kernel needed 168 registers: 384 x 1
                               // 4.88 ms against 256 x 2 at 128 registers 4.99 ms; the leaner round-2 kernel fits sixteen warps.)
#endif
constexpr int kThreads = MRP_BOOK_THREADS;   // CTA size of the book pass (= rows per chunk)
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v, int d) {
    return __shfl_up_sync(0xffffffffu, v, d);
}
__device__ __forceinline__ bool finite_d(double x) { return fabs(x) <= 1.7976931348623157e308; }

// -x*log2(x) terms in the reference's order (feature_processor.cpp:225-228, 265-267).
__device__ __forceinline__ double entropy_term(double acc, double p) {
    if (p > 0) acc -= p * log2(p);
    return acc;
}


// 1/sqrt(v) for a positive normal v: the seed plus one third-order step (the library's own fast path)
__device__ __forceinline__ double rsqrt_pos(double v) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
    const double e = fma(-v * r, r, 1.0);           // 1 - v r^2
    const double p = fma(e, 0.375, 0.5);
    return fma(p, e * r, r);                         // r (1 + e/2 + 3 e^2/8)
}
// positive, normal and far from the range ends: the inputs rsqrt_pos is exact enough for
__device__ __forceinline__ bool rsqrt_pos_ok(double v) { return (uint32_t)(__double2hiint(v) - 0x03000000) < 0x79000000u; }

// ---- bulk asynchronous copies (TMA unit, SASS UBLKCP) and the mbarriers that track them
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
// the same with an L2 eviction policy (createpolicy): evict_first for data nobody reads again, evict_last for data that is
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, const void* src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

}  // namespace mrp
