// kernels_norm.cu — the rolling z-score (FeatureNormalizer) and its write epilogue.
//
// Reference (feature_normalizer.cpp:32-83): running sum / sum-of-squares per feature, add the new row,
// evict the row 30 000 back, z = (x - sum/n) / sqrt(sum2/n - (sum/n)^2) (std = 1 when the variance is not
// positive).  Those running sums random-walk in rounding error as the stream grows.
//
// k_normalize evaluates the SAME window sums without any long-lived accumulator.  The row axis of a stream
// is cut into tiles of `tile_rows` rows, tile_rows dividing norm_window (30 000 = 30 x 1000).  For a row t
// in tile T:
//      window sum(t) = [totals of tiles T-30 .. T-1] + sum_{i <= t in tile} (x_i - x_{i-W})
// so a warp needs 30 small totals of the tile's predecessors plus one in-tile scan of (x - x_lag, x^2 - x_lag^2).
// Tiles are handed out by an atomic ticket in time order; each warp first publishes its own column total,
// only then waits for the predecessors' totals of the same column — predecessors hold smaller tickets, hence
// are resident: no deadlock, no serial carry chain.  The lagged tile was streamed ~30 tickets earlier and is still in the
// 126 MB L2, so HBM sees each raw value once (read) and each z-score once (write).
// After the ticket is taken the ten warps of a CTA never meet again: every warp owns whole columns, a lane owns
// four consecutive rows per 128-row step (one 256-bit access: the warp touches 1 KB contiguously), and the only
// cross-lane traffic is one warp scan per step.
//
// Non-finite raw values follow the reference's own arithmetic (a NaN poisons the running sums for ever, an
// infinity until it is evicted and then for ever): k_normalize records where each column first goes
// non-finite, and k_norm_poison — which returns at once on clean data — rewrites the affected rows.
//
// k_normalize_faithful re-enacts the reference's add-then-evict chain in its exact rounding order
// (one thread per stream x feature; serial, for validation only).
#include <algorithm>
#include <cstdlib>

#include "mrp_device.cuh"

namespace mrp {

constexpr int kNormWarps = 10;
constexpr int kNormThreads = kNormWarps * 32;
constexpr int kNF = MRP_FG_NUM_FEATURES;
constexpr int kRPL = 4;               // consecutive rows per lane: one 256-bit access, a warp covers 1 KB contiguously
constexpr int kStep = 32 * kRPL;      // rows per warp step

// A tile total doubles as its own "published" flag: the table is preset to an all-ones pattern (a NaN no
// arithmetic produces), a total is one aligned 8-byte store, and a reader polls the value itself — no separate
// flag, no fences on either side.
constexpr unsigned long long kUnpublished = 0xffffffffffffffffull;
__device__ __forceinline__ unsigned long long peek_total(const double* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ double poll_total(const double* p, unsigned long long u) {  // u: an earlier peek of *p
    while (u == kUnpublished) {
        __nanosleep(20);
        u = peek_total(p);
    }
    return __longlong_as_double((long long)u);
}
__device__ __forceinline__ double publishable(double s) {  // keep computed NaNs distinguishable from "unpublished"
    return s == s ? s : __longlong_as_double(0x7ff8000000000000ll);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Four consecutive doubles of a lane.  ALIGN = 32: one 256-bit access (sm_100 LDG.256 / STG.256: the warp's 32 lanes
// touch 8 full lines); ALIGN = 16: two 128-bit accesses; otherwise scalar.
template <int ALIGN>
__device__ __forceinline__ void load4(const double* __restrict__ p, int i, int n, double (&o)[kRPL]) {
    if (ALIGN == 32 && i + kRPL <= n) {
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p + i));
    } else if (ALIGN == 16 && i + kRPL <= n) {
        const double2 a = *reinterpret_cast<const double2*>(p + i), b = *reinterpret_cast<const double2*>(p + i + 2);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < kRPL; ++k) o[k] = i + k < n ? p[i + k] : 0.0;
    }
}
template <int ALIGN>
__device__ __forceinline__ void store4(double* __restrict__ p, int i, int n, const double (&o)[kRPL]) {
    if (ALIGN == 32 && i + kRPL <= n) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "d"(o[0]), "d"(o[1]), "d"(o[2]), "d"(o[3]) : "memory");
    } else if (ALIGN == 16 && i + kRPL <= n) {
        *reinterpret_cast<double2*>(p + i) = make_double2(o[0], o[1]);
        *reinterpret_cast<double2*>(p + i + 2) = make_double2(o[2], o[3]);
    } else {
#pragma unroll
        for (int k = 0; k < kRPL; ++k)
            if (i + k < n) p[i + k] = o[k];
    }
}
__device__ __forceinline__ int common_align(const void* a, const void* b, int w_rows) {
    const uintptr_t m = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | ((uintptr_t)w_rows * 8u);
    return (m & 31) == 0 ? 32 : ((m & 15) == 0 ? 16 : 8);
}

// non-finite bookkeeping, per (stream, column): first stream-local row holding NaN / +inf / -inf
struct NormPoison {
    unsigned int first_nan, first_pinf, first_ninf;
};

// ---- pass 1 of one (tile, column): the tile total, published as soon as the column has been read once
template <int ALIGN>
__device__ __forceinline__ void tile_total(const double* __restrict__ x, const int n, const int64_t tl0, double* __restrict__ tot,
                                           NormPoison* __restrict__ q, int* __restrict__ any_poison, const int lane) {
    double s = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0;
    unsigned int fn = 0xffffffffu, fp = 0xffffffffu, fm = 0xffffffffu;
    for (int i0 = 0; i0 < n; i0 += 2 * kStep) {
        const int ia = i0 + lane * kRPL, ib = ia + kStep;
        double a[kRPL], b[kRPL];
        load4<ALIGN>(x, ia, n, a);
        load4<ALIGN>(x, ib, n, b);
        uint32_t m = 0u;  // largest |high word|: >= 0x7ff00000 means an infinity or a NaN
#pragma unroll
        for (int k = 0; k < kRPL; ++k) {
            s += a[k];
            s2 = fma(a[k], a[k], s2);
            t1 += b[k];
            t2 = fma(b[k], b[k], t2);
            m = max(m, max((uint32_t)__double2hiint(a[k]) & 0x7fffffffu, (uint32_t)__double2hiint(b[k]) & 0x7fffffffu));
        }
        if (m >= 0x7ff00000u) {  // rare: record where the column first goes non-finite (stream-local rows)
#pragma unroll
            for (int k = 0; k < 2 * kRPL; ++k) {
                const double y = k < kRPL ? a[k] : b[k - kRPL];
                const int i = (k < kRPL ? ia : ib) + (k & (kRPL - 1));
                if (i >= n) continue;
                const unsigned int r = (unsigned int)(tl0 + i);
                if (y != y) fn = min(fn, r);
                else if (y == __longlong_as_double(0x7ff0000000000000ll)) fp = min(fp, r);
                else if (y == __longlong_as_double(0xfff0000000000000ll)) fm = min(fm, r);
            }
        }
    }
    if ((fn & fp & fm) != 0xffffffffu) {
        if (fn != 0xffffffffu) atomicMin(&q->first_nan, fn);
        if (fp != 0xffffffffu) atomicMin(&q->first_pinf, fp);
        if (fm != 0xffffffffu) atomicMin(&q->first_ninf, fm);
        *any_poison = 1;
    }
    s += t1;
    s2 += t2;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) {
        __stcg(tot, publishable(s));
        __stcg(tot + kNF, publishable(s2));
    }
}

// ---- pass 2 of one (tile, column): the window sum slides along the tile
//        sum(t) = [totals of the preceding tiles inside the window] + sum_{i<=t in tile} (x_i - x_{i-W})
template <int ALIGN>
__device__ __forceinline__ void tile_zscores(const double* __restrict__ x, double* __restrict__ z, const int n, const bool has_lag,
                                             const int W, const int64_t tl0, const double* __restrict__ pred_tot, const int nb,
                                             const double* __restrict__ inv_n, int32_t* __restrict__ cnt_out, const int lane) {
    const double* __restrict__ xl = x - W;
    // the first step's values are in flight while the predecessors' totals are collected
    double v[kRPL], l[kRPL] = {0.0, 0.0, 0.0, 0.0};
    load4<ALIGN>(x, lane * kRPL, n, v);
    if (has_lag) load4<ALIGN>(xl, lane * kRPL, n, l);
    double run1 = 0.0, run2 = 0.0;
    {
        double h1 = 0.0, h2 = 0.0;
        for (int j0 = 0; j0 < nb; j0 += 32) {
            if (j0 + lane < nb) {
                const double* __restrict__ tp = pred_tot + (int64_t)(j0 + lane) * 2 * kNF;
                const unsigned long long ua = peek_total(tp), ub = peek_total(tp + kNF);  // both in flight together
                h1 += poll_total(tp, ua);
                h2 += poll_total(tp + kNF, ub);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            h1 += __shfl_xor_sync(0xffffffffu, h1, d);
            h2 += __shfl_xor_sync(0xffffffffu, h2, d);
        }
        run1 = h1;
        run2 = h2;
    }
    const double invW = __ldg(inv_n + W);
    for (int i0 = 0; i0 < n; i0 += kStep) {
        const int ib = i0 + lane * kRPL;
        double vn[kRPL] = {0.0, 0.0, 0.0, 0.0}, ln[kRPL] = {0.0, 0.0, 0.0, 0.0};
        if (i0 + kStep < n) {  // next step's values: in flight during this step's arithmetic
            load4<ALIGN>(x, ib + kStep, n, vn);
            if (has_lag) load4<ALIGN>(xl, ib + kStep, n, ln);
        }
        // a tile without lag lies inside the first window of its stream (tl0 + n <= W): 1/n needs no clamp there
        const double* __restrict__ inv_fill = inv_n + (has_lag ? 0 : tl0 + ib + 1);
        double p1[kRPL], p2[kRPL];
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int k = 0; k < kRPL; ++k) {
            a1 += v[k] - l[k];
            a2 += fma(v[k], v[k], -(l[k] * l[k]));
            p1[k] = a1;
            p2[k] = a2;
        }
        double s1 = a1, s2 = a2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double u1 = shfl_up_d(s1, d), u2 = shfl_up_d(s2, d);
            if (lane >= d) { s1 += u1; s2 += u2; }
        }
        // exclusive offsets by shuffle, not by subtraction: a NaN in this lane's rows must not reach the rows before it
        const double x1 = shfl_up_d(s1, 1), x2 = shfl_up_d(s2, 1);
        const double b1 = run1 + (lane ? x1 : 0.0), b2 = run2 + (lane ? x2 : 0.0);
        double zz[kRPL];
#pragma unroll
        for (int k = 0; k < kRPL; ++k) {
            const double e1 = b1 + p1[k], e2 = b2 + p2[k];
            double inv = invW;
            if (!has_lag) inv = __ldg(inv_fill + k);  // the window is still filling: n = rows so far (feature_normalizer.cpp:71)
            const double mean = e1 * inv;                       // :62
            const double variance = e2 * inv - mean * mean;     // :63
            // :67 — std = 1 unless the variance is positive.  Fast path: positive, normal, far from the range ends
            double rstd;
            if ((uint32_t)(__double2hiint(variance) - 0x03000000) < 0x79000000u) rstd = rsqrt_pos(variance);
            else rstd = variance > 0.0 ? rsqrt(variance) : 1.0;
            zz[k] = (v[k] - mean) * rstd;                       // :68
        }
        store4<ALIGN>(z, ib, n, zz);
        if (cnt_out) {
#pragma unroll
            for (int k = 0; k < kRPL; ++k)
                if (ib + k < n) {
                    const int64_t tl = tl0 + ib + k;
                    cnt_out[ib + k] = tl + 1 >= W ? W : (int)(tl + 1);
                }
        }
        run1 += __shfl_sync(0xffffffffu, s1, 31);
        run2 += __shfl_sync(0xffffffffu, s2, 31);
#pragma unroll
        for (int k = 0; k < kRPL; ++k) { v[k] = vn[k]; l[k] = ln[k]; }
    }
}

// One CTA per tile (atomic ticket, time order), ten warps, each owning the columns w and w + 10.  All of the
// CTA's column totals are published first (one read of the tile), then the z-scores are produced from a second read
// that the L2 serves; after the ticket the warps never meet again.
__global__ void __launch_bounds__(kNormThreads, 3)
k_normalize(const DevOutputs out, const int64_t N, const DevTile* __restrict__ tiles, double* __restrict__ tile_tot,
            int* __restrict__ flags, unsigned int* __restrict__ ticket, NormPoison* __restrict__ poison, int* __restrict__ any_poison,
            const DevParams P) {
    __shared__ int sId;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sId = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int id = sId;
    const DevTile T = tiles[id];
    const int n = T.n, W = P.norm_window;
    const int64_t tl0 = (int64_t)T.index * P.tile_rows;  // stream-local index of the tile's first row
    for (int c = warp; c < kNF; c += kNormWarps) {
        const double* __restrict__ x = out.raw + (int64_t)c * N + T.row;
        double* tot = tile_tot + (int64_t)id * 2 * kNF + c;
        NormPoison* q = poison + (int64_t)T.stream * kNF + c;
        const int al = common_align(x, x, 0);
        if (al == 32) tile_total<32>(x, n, tl0, tot, q, any_poison, lane);
        else if (al == 16) tile_total<16>(x, n, tl0, tot, q, any_poison, lane);
        else tile_total<8>(x, n, tl0, tot, q, any_poison, lane);
    }
    const int nb = T.index < P.tiles_back ? T.index : P.tiles_back;  // preceding tiles inside the window
    const bool has_lag = T.index >= P.tiles_back;
    for (int c = warp; c < kNF; c += kNormWarps) {
        const double* __restrict__ x = out.raw + (int64_t)c * N + T.row;
        double* __restrict__ z = out.norm + (int64_t)c * N + T.row;
        const double* pred = tile_tot + (int64_t)(id - nb) * 2 * kNF + c;
        int32_t* cnt_out = (c == 0 && out.counts) ? out.counts + (int64_t)MRP_C_NORM_N * N + T.row : nullptr;
        const int al = common_align(x, z, W);
        if (al == 32) tile_zscores<32>(x, z, n, has_lag, W, tl0, pred, nb, P.inv_cnt_norm, cnt_out, lane);
        else if (al == 16) tile_zscores<16>(x, z, n, has_lag, W, tl0, pred, nb, P.inv_cnt_norm, cnt_out, lane);
        else tile_zscores<8>(x, z, n, has_lag, W, tl0, pred, nb, P.inv_cnt_norm, cnt_out, lane);
    }
}

// ================================================================================================
// k_normalize_runs — the same two passes with the tile streamed through shared memory by the copy engine.
//
// Blocks of 224 rows are fetched by ONE bulk asynchronous copy per array (cp.async.bulk + mbarrier, issued by lane 0 two
// blocks ahead), a lane owns 7 CONSECUTIVE rows of the block (odd run length: 64-bit shared accesses at lane stride 7 are
// conflict-free), so the in-block prefix is seven serial adds in registers and the warp scan is paid once per 224 rows
// instead of once per 128; z-scores leave through shared memory and one bulk store per block.  No lane ever issues a
// global load for the data: ~45 instructions per z-score instead of ~110.
// Source / destination addresses are rounded to the 16 bytes the copy engine wants; the element shift that results
// (0 or 1) is added to the slot index, and the at most two edge rows of a block that a 16-byte store cannot cover are
// written by lane 0.
// ================================================================================================
#ifndef MRP_NORM_RUN_WARPS
#define MRP_NORM_RUN_WARPS 10
#endif
#ifndef MRP_NORM_RUN_CTAS
#define MRP_NORM_RUN_CTAS 2
#endif
constexpr int kRunWarps = MRP_NORM_RUN_WARPS;   // warps per CTA (= tile); each owns the columns w, w + kRunWarps, ...
constexpr int kRun = 7;                  // rows per lane per block
constexpr int kBlk = 32 * kRun;          // 224
constexpr int kBlkSlot = kBlk + 2;       // + the alignment shift, padded to 16 bytes
constexpr int kSlots = 6;                // pass 1 uses all six as a ring; pass 2: x = 0,1  lag = 2,3  z = 4,5
struct __align__(16) NormWarpStage {
    double slot[kSlots][kBlkSlot];
    uint64_t full[kSlots];               // pass 1: one per slot; pass 2: full[0..1] (a block's x and lag complete together)
};

// lane 0: request rows [0, rows) starting at src into slot (16-byte rounding on both ends); returns the bytes requested
__device__ __forceinline__ uint32_t block_bytes(const double* src, int rows) {
    const uintptr_t p0 = reinterpret_cast<uintptr_t>(src), p1 = p0 + (uintptr_t)rows * 8u;
    return (uint32_t)(((p1 + 15) & ~(uintptr_t)15) - (p0 & ~(uintptr_t)15));
}
__device__ __forceinline__ void block_load(double* slot, const double* src, int rows, uint64_t* bar) {
    const uintptr_t q0 = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15;
    bulk_g2s(slot, reinterpret_cast<const void*>(q0), block_bytes(src, rows), bar);
}
__device__ __forceinline__ void block_load_hint(double* slot, const double* src, int rows, uint64_t* bar, uint64_t pol) {
    const uintptr_t q0 = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15;
    bulk_g2s_hint(slot, reinterpret_cast<const void*>(q0), block_bytes(src, rows), bar, pol);
}
#ifndef MRP_NORM_HINTS
#define MRP_NORM_HINTS 3
#endif
__device__ __forceinline__ int block_shift(const double* src) { return (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1u); }

// per-warp staging state: bit i = the phase parity the next wait on full[i] expects
struct RunPipe {
    uint32_t phase;
    __device__ __forceinline__ void wait(NormWarpStage& S, int i) {
        mbar_wait(&S.full[i], (phase >> i) & 1u);
        phase ^= 1u << i;
    }
};

// pass 1: tile totals of this warp's columns (c0, c0 + 10, ...), one after the other through ONE ring of six slots, so six
// blocks (10 KB per warp) are in flight from HBM and only the very first block's latency is exposed
__device__ __forceinline__ void run_totals(NormWarpStage& S, RunPipe& rp, const double* __restrict__ raw, const int64_t N, const int64_t row,
                                           const int c0, const int n, const int64_t tl0, double* __restrict__ tot0,
                                           NormPoison* __restrict__ q0, int* __restrict__ any_poison, const int lane) {
    const int nblk = (n + kBlk - 1) / kBlk;
    const int ncol = (kNF - c0 + kRunWarps - 1) / kRunWarps;
    const int total = nblk * ncol;
    const double* __restrict__ x0 = raw + (int64_t)c0 * N + row;
    const int64_t col_stride = (int64_t)kRunWarps * N;
    int ij = 0;              // block and column of the next request
    int64_t icol = 0;
    auto issue = [&](int sl) {
        if (lane == 0) {
            const int rows = min(kBlk, n - ij * kBlk);
            const double* src = x0 + icol + (int64_t)ij * kBlk;
            fence_proxy_async_smem();   // the lanes' reads of this slot (ordered before this point by __syncwarp) precede the engine's write
            mbar_expect_tx(&S.full[sl], block_bytes(src, rows));
            block_load(S.slot[sl], src, rows, &S.full[sl]);
        }
        if (++ij == nblk) { ij = 0; icol += col_stride; }
    };
#pragma unroll
    for (int t = 0; t < kSlots; ++t)
        if (t < total) issue(t);
    int sl = 0, t = 0;
    for (int ci = 0; ci < ncol; ++ci) {
        const double* __restrict__ x = x0 + (int64_t)ci * col_stride;
        double s = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0;
        unsigned int fn = 0xffffffffu, fp = 0xffffffffu, fm = 0xffffffffu;
        for (int j = 0; j < nblk; ++j, ++t) {
            rp.wait(S, sl);
            const int rows = min(kBlk, n - j * kBlk);
            const double* __restrict__ xs = S.slot[sl] + block_shift(x + (int64_t)j * kBlk) + lane * kRun;
            double v[kRun];
            if (rows == kBlk) {
#pragma unroll
                for (int k = 0; k < kRun; ++k) v[k] = xs[k];
            } else {
#pragma unroll
                for (int k = 0; k < kRun; ++k) v[k] = lane * kRun + k < rows ? xs[k] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kRun; ++k) {
                if (k & 1) { t1 += v[k]; t2 = fma(v[k], v[k], t2); }
                else { s += v[k]; s2 = fma(v[k], v[k], s2); }
            }
            // an infinity or a NaN among the values makes the sum of squares non-finite (and it stays so)
            if (!(fabs(s2 + t2) <= 1.7976931348623157e308)) {  // rare: record where the column first goes non-finite
#pragma unroll
                for (int k = 0; k < kRun; ++k) {
                    const unsigned int r = (unsigned int)(tl0 + j * kBlk + lane * kRun + k);   // stream-local row
                    const double y = v[k];
                    if (y != y) fn = min(fn, r);
                    else if (y == __longlong_as_double(0x7ff0000000000000ll)) fp = min(fp, r);
                    else if (y == __longlong_as_double(0xfff0000000000000ll)) fm = min(fm, r);
                }
            }
            __syncwarp();   // every lane has read the slot
            if (t + kSlots < total) issue(sl);
            if (++sl == kSlots) sl = 0;
        }
        if ((fn & fp & fm) != 0xffffffffu) {
            NormPoison* q = q0 + ci * kRunWarps;
            if (fn != 0xffffffffu) atomicMin(&q->first_nan, fn);
            if (fp != 0xffffffffu) atomicMin(&q->first_pinf, fp);
            if (fm != 0xffffffffu) atomicMin(&q->first_ninf, fm);
            *any_poison = 1;
        }
        s += t1;
        s2 += t2;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            s2 += __shfl_xor_sync(0xffffffffu, s2, d);
        }
        if (lane == 0) {
            double* tot = tot0 + ci * kRunWarps;
            __stcg(tot, publishable(s));
            __stcg(tot + kNF, publishable(s2));
        }
    }
}

// pass 2: z-scores of one column of the tile
template <bool LAG>
__device__ __forceinline__ void run_zscores(NormWarpStage& S, RunPipe& rp, const double* __restrict__ x, double* __restrict__ z, const int n,
                                            const int W, const int64_t tl0, const double* __restrict__ pred_tot, const int nb,
                                            const double* __restrict__ inv_n, int32_t* __restrict__ cnt_out, const int lane) {
    const double* __restrict__ xl = x - W;
    const int nblk = (n + kBlk - 1) / kBlk;
    auto issue = [&](int j) {
        if (lane == 0) {
            const int rows = min(kBlk, n - j * kBlk);
            const double* src = x + (int64_t)j * kBlk;
            const int s = j & 1;
            uint32_t bytes = block_bytes(src, rows);
            if (LAG) bytes += block_bytes(src - W, rows);
            fence_proxy_async_smem();   // as in pass 1: generic-proxy reads of the slots before the engine's writes
            mbar_expect_tx(&S.full[s], bytes);
#if MRP_NORM_HINTS & 4
            block_load_hint(S.slot[s], src, rows, &S.full[s], l2_policy_evict_last());   // read again as the lag, 15 tiles on
#else
            block_load(S.slot[s], src, rows, &S.full[s]);
#endif
#if MRP_NORM_HINTS & 1
            if (LAG) block_load_hint(S.slot[2 + s], src - W, rows, &S.full[s], l2_policy_evict_first());   // last use of these rows
#else
            if (LAG) block_load(S.slot[2 + s], src - W, rows, &S.full[s]);
#endif
        }
    };
    // the first blocks are in flight while the predecessors' totals are collected
    if (nblk > 0) issue(0);
    if (nblk > 1) issue(1);
    double run1 = 0.0, run2 = 0.0;
    {
        double h1 = 0.0, h2 = 0.0;
        for (int j0 = 0; j0 < nb; j0 += 32) {
            if (j0 + lane < nb) {
                const double* __restrict__ tp = pred_tot + (int64_t)(j0 + lane) * 2 * kNF;
                const unsigned long long ua = peek_total(tp), ub = peek_total(tp + kNF);  // both in flight together
                h1 += poll_total(tp, ua);
                h2 += poll_total(tp + kNF, ub);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            h1 += __shfl_xor_sync(0xffffffffu, h1, d);
            h2 += __shfl_xor_sync(0xffffffffu, h2, d);
        }
        run1 = h1;
        run2 = h2;
    }
    const double invW = __ldg(inv_n + W);
    const int r0 = lane * kRun;    // this lane's first row of a block
    for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        rp.wait(S, st);
        const int rows = min(kBlk, n - j * kBlk);
        const double* __restrict__ xs = S.slot[st] + block_shift(x + (int64_t)j * kBlk) + r0;
        const double* __restrict__ ls = S.slot[2 + st] + block_shift(xl + (int64_t)j * kBlk) + r0;
        double v[kRun], l[kRun], p1[kRun], p2[kRun];
        if (rows == kBlk) {
#pragma unroll
            for (int k = 0; k < kRun; ++k) {
                v[k] = xs[k];
                if (LAG) l[k] = ls[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < kRun; ++k) {
                const bool ok = r0 + k < rows;
                v[k] = ok ? xs[k] : 0.0;
                if (LAG) l[k] = ok ? ls[k] : 0.0;
            }
        }
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int k = 0; k < kRun; ++k) {
            if (LAG) {
                a1 += v[k] - l[k];
                a2 += fma(v[k], v[k], -(l[k] * l[k]));
            } else {
                a1 += v[k];
                a2 = fma(v[k], v[k], a2);
            }
            p1[k] = a1;
            p2[k] = a2;
        }
        __syncwarp();   // every lane has read the stage
        if (j + 2 < nblk) issue(j + 2);
        double s1 = a1, s2 = a2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double u1 = shfl_up_d(s1, d), u2 = shfl_up_d(s2, d);
            if (lane >= d) { s1 += u1; s2 += u2; }
        }
        // exclusive offsets by shuffle, not by subtraction: a NaN in this lane's rows must not reach the rows before it
        const double x1 = shfl_up_d(s1, 1), x2 = shfl_up_d(s2, 1);
        const double b1 = run1 + (lane ? x1 : 0.0), b2 = run2 + (lane ? x2 : 0.0);
        // the z slot of two blocks ago must have left shared memory
        if (j >= 2) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
        }
        double* __restrict__ zg = z + (int64_t)j * kBlk;
        const int zsh = block_shift(zg);
        double* __restrict__ zs = S.slot[4 + st] + zsh;
        // a tile without lag lies inside the first window of its stream (tl0 + n <= W): 1/n needs no clamp there
        const double* __restrict__ inv_fill = inv_n + (LAG ? 0 : tl0 + (int64_t)j * kBlk + r0 + 1);
        double var[kRun], mean[kRun];
        uint32_t okmin = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < kRun; ++k) {
            const double e1 = b1 + p1[k], e2 = b2 + p2[k];
            double inv = invW;
            if (!LAG) inv = __ldg(inv_fill + k);  // the window is still filling: n = rows so far (feature_normalizer.cpp:71)
            mean[k] = e1 * inv;                             // :62
            var[k] = e2 * inv - mean[k] * mean[k];          // :63
            // fast path of :67 — positive, normal, far from the range ends
            okmin = min(okmin, (uint32_t)((uint32_t)(__double2hiint(var[k]) - 0x03000000) < 0x79000000u));
        }
        if (okmin != 0u) {
#pragma unroll
            for (int k = 0; k < kRun; ++k) zs[r0 + k] = (v[k] - mean[k]) * rsqrt_pos(var[k]);   // :68
        } else {   // some variance is zero, negative, NaN or extreme: std = 1 unless the variance is positive (:67)
#pragma unroll
            for (int k = 0; k < kRun; ++k) {
                const double rstd = var[k] > 0.0 ? rsqrt(var[k]) : 1.0;
                zs[r0 + k] = (v[k] - mean[k]) * rstd;
            }
        }
        fence_proxy_async_smem();   // the copy engine reads what the lanes just wrote
        __syncwarp();
        if (lane == 0) {
            // rows [ra, ra + cnt) are covered by 16-byte units; at most one row on either side is stored directly
            const int ra = zsh, cnt = (rows - ra) & ~1;
            if (ra && rows > 0) zg[0] = zs[0];
#if MRP_NORM_HINTS & 2
            if (cnt > 0) bulk_s2g_hint(zg + ra, zs + ra, (uint32_t)cnt * 8u, l2_policy_evict_first());   // nobody reads z here
#else
            if (cnt > 0) bulk_s2g(zg + ra, zs + ra, (uint32_t)cnt * 8u);
#endif
            bulk_commit();
            if (ra + cnt < rows) zg[rows - 1] = zs[rows - 1];
        }
        if (cnt_out) {
#pragma unroll
            for (int k = 0; k < kRun; ++k)
                if (r0 + k < rows) {
                    const int64_t tl = tl0 + (int64_t)j * kBlk + r0 + k;
                    cnt_out[(int64_t)j * kBlk + r0 + k] = tl + 1 >= W ? W : (int)(tl + 1);
                }
        }
        run1 += __shfl_sync(0xffffffffu, s1, 31);
        run2 += __shfl_sync(0xffffffffu, s2, 31);
    }
    if (lane == 0) bulk_wait_read<0>();   // shared memory may be reused (and the CTA may exit) only after the stores have read it
    __syncwarp();
}

__global__ void __launch_bounds__(kRunWarps * 32, MRP_NORM_RUN_CTAS)
k_normalize_runs(const DevOutputs out, const int64_t N, const DevTile* __restrict__ tiles, double* __restrict__ tile_tot,
                 unsigned int* __restrict__ ticket, NormPoison* __restrict__ poison, int* __restrict__ any_poison, const DevParams P) {
    extern __shared__ __align__(16) unsigned char norm_dyn[];
    __shared__ int sId;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    NormWarpStage& S = reinterpret_cast<NormWarpStage*>(norm_dyn)[warp];
    if (tid == 0) sId = (int)atomicAdd(ticket, 1u);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kSlots; ++i) mbar_init(&S.full[i], 1);
        mbar_init_fence();
    }
    __syncthreads();
    RunPipe rp{0u};
    const int id = sId;
    const DevTile T = tiles[id];
    const int n = T.n, W = P.norm_window;
    const int64_t tl0 = (int64_t)T.index * P.tile_rows;  // stream-local index of the tile's first row
    run_totals(S, rp, out.raw, N, T.row, warp, n, tl0, tile_tot + (int64_t)id * 2 * kNF + warp, poison + (int64_t)T.stream * kNF + warp,
               any_poison, lane);
    const int nb = T.index < P.tiles_back ? T.index : P.tiles_back;  // preceding tiles inside the window
    const bool has_lag = T.index >= P.tiles_back;
    for (int c = warp; c < kNF; c += kRunWarps) {
        const double* __restrict__ x = out.raw + (int64_t)c * N + T.row;
        double* __restrict__ z = out.norm + (int64_t)c * N + T.row;
        const double* pred = tile_tot + (int64_t)(id - nb) * 2 * kNF + c;
        int32_t* cnt_out = (c == 0 && out.counts) ? out.counts + (int64_t)MRP_C_NORM_N * N + T.row : nullptr;
        if (has_lag) run_zscores<true>(S, rp, x, z, n, W, tl0, pred, nb, P.inv_cnt_norm, cnt_out, lane);
        else run_zscores<false>(S, rp, x, z, n, W, tl0, pred, nb, P.inv_cnt_norm, cnt_out, lane);
    }
}

// Rewrites the z-scores of columns that went non-finite, following the reference's running-sum arithmetic
// (feature_normalizer.cpp:37-51,61-69):  a NaN makes sum / sum2 NaN for ever; an infinity makes sum = +-inf and
// sum2 = +inf (variance NaN -> std 1 -> z = x - (+-inf)) until a NaN, an infinity of the other sign or its own
// eviction W rows later turns the sums into NaN for ever.  One CTA per (stream, column); exits at once when clean.
//
// "For ever" reaches beyond a time chunk / time shard: `carry` ([n_streams][19][3] first NaN / +inf / -inf in ABSOLUTE stream
// rows, LLONG_MAX = none; may be null) brings in what happened before the bound slice, whose first row is absolute row
// slice_row0[stream]; the merged state is written back for the slice that follows.
constexpr long long kNoRow = 0x7fffffffffffffffll;
__global__ void k_norm_poison(const DevOutputs out, const int64_t N, const int64_t* __restrict__ row_off, const NormPoison* __restrict__ poison,
                              const int* __restrict__ any_poison, long long* __restrict__ carry, const long long* __restrict__ slice_row0,
                              const DevParams P) {
    const int sidx = blockIdx.x / kNF, c = blockIdx.x % kNF;
    long long cN = kNoRow, cP = kNoRow, cM = kNoRow;
    long long base = 0;
    if (carry) {
        const long long* cr = carry + ((int64_t)sidx * kNF + c) * 3;
        cN = cr[0]; cP = cr[1]; cM = cr[2];
        base = slice_row0[sidx];
    }
    if (*any_poison == 0 && cN == kNoRow && cP == kNoRow && cM == kNoRow) return;
    const NormPoison q = poison[(int64_t)sidx * kNF + c];
    const unsigned int none = 0xffffffffu;
    // slice-local rows (negative = before the slice)
    long long tNl = q.first_nan == none ? kNoRow : (long long)q.first_nan;
    long long tPl = q.first_pinf == none ? kNoRow : (long long)q.first_pinf, tMl = q.first_ninf == none ? kNoRow : (long long)q.first_ninf;
    if (cN != kNoRow && cN - base < tNl) tNl = cN - base;
    if (cP != kNoRow && cP - base < tPl) tPl = cP - base;
    if (cM != kNoRow && cM - base < tMl) tMl = cM - base;
    __syncthreads();   // every thread has read the carry before it is updated
    if (carry && threadIdx.x == 0) {
        long long* cr = carry + ((int64_t)sidx * kNF + c) * 3;
        cr[0] = tNl == kNoRow ? kNoRow : tNl + base;
        cr[1] = tPl == kNoRow ? kNoRow : tPl + base;
        cr[2] = tMl == kNoRow ? kNoRow : tMl + base;
    }
    if (tNl == kNoRow && tPl == kNoRow && tMl == kNoRow) return;
    const int64_t r0 = row_off[sidx], r1 = row_off[sidx + 1];
    const long long n = r1 - r0;
    const long long tN = tNl < n ? tNl : n;
    const long long tP = tPl < n ? tPl : n, tM = tMl < n ? tMl : n;
    const long long tI = tP < tM ? tP : tM;                 // first infinity
    const long long tO = tP < tM ? tM : tP;                 // first infinity of the other sign (n if none)
    long long tZ = tN;                                      // sums are NaN from here on
    if (tI < n) {
        if (tO < tZ) tZ = tO;
        if (tI + P.norm_window < tZ) tZ = tI + P.norm_window;
    }
    const double sig_inf = tP < tM ? __longlong_as_double(0x7ff0000000000000ll) : __longlong_as_double(0xfff0000000000000ll);
    const double* __restrict__ x = out.raw + (int64_t)c * N + r0;
    double* __restrict__ z = out.norm + (int64_t)c * N + r0;
    long long start = tI < tZ ? tI : tZ;
    if (start < 0) start = 0;
    for (long long t = start + threadIdx.x; t < n; t += blockDim.x) {
        if (t >= tZ) z[t] = __longlong_as_double(0x7ff8000000000000ll);
        else z[t] = x[t] - sig_inf;                       // mean = +-inf, std = 1
    }
}

// One thread per (stream, feature): the reference's own chain, feature_normalizer.cpp:37-51,61-69.
__global__ void k_normalize_faithful(const DevOutputs out, const int64_t N, const int64_t* __restrict__ row_off,
                                     const int n_streams, const DevParams P) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_streams * kNF) return;
    const int s = idx / kNF, c = idx % kNF;
    const int64_t r0 = row_off[s], r1 = row_off[s + 1];
    const double* __restrict__ x = out.raw + (int64_t)c * N;
    double* __restrict__ z = out.norm + (int64_t)c * N;
    const int64_t W = P.norm_window;
    double sum = 0.0, sum2 = 0.0;
    for (int64_t t = r0; t < r1; ++t) {
        const double val = x[t];
        sum += val;
        sum2 += val * val;
        int64_t size = t - r0 + 1;
        if (size > W) {
            const double old = x[t - W];
            sum += -old;
            sum2 += -(old * old);
            size = W;
        }
        const double mean = sum / (double)size;
        const double variance = (sum2 / (double)size) - (mean * mean);
        const double stddev = (variance > 0.0) ? sqrt(variance) : 1.0;
        z[t] = (val - mean) / stddev;
    }
}

size_t norm_poison_bytes(int n_streams) { return (size_t)n_streams * kNF * sizeof(NormPoison) + 16; }

cudaError_t launch_normalize(const DevOutputs& out, int64_t N, const DevTile* tiles, int ntiles, double* tile_tot, int* flags,
                             void* poison, const int64_t* row_off, int n_streams, long long* carry, const long long* slice_row0,
                             const DevParams& P, cudaStream_t st) {
    // the ticket counter at flags[ntiles]; tile totals preset to "unpublished"; poison table preset to "none" + the any-flag
    cudaError_t e = cudaMemsetAsync(flags + ntiles, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(tile_tot, 0xff, (size_t)ntiles * 2 * kNF * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const size_t pb = (size_t)n_streams * kNF * sizeof(NormPoison);
    e = cudaMemsetAsync(poison, 0xff, pb, st);
    if (e != cudaSuccess) return e;
    int* any = reinterpret_cast<int*>(static_cast<unsigned char*>(poison) + pb);
    e = cudaMemsetAsync(any, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    static const bool use_runs = [] {   // MRP_FG_NORM_KERNEL=scan selects the register-scan kernel (kept for comparison)
        const char* v = std::getenv("MRP_FG_NORM_KERNEL");
        return !(v && v[0] == 's');
    }();
    if (use_runs) {
        constexpr int kBytes = (int)sizeof(NormWarpStage) * kRunWarps;
        static bool ready[64];  // the opt-in is per device (idempotent; a race only repeats the call)
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (!ready[dev & 63]) {
            e = cudaFuncSetAttribute(k_normalize_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes);
            if (e != cudaSuccess) return e;
            ready[dev & 63] = true;
        }
        k_normalize_runs<<<ntiles, kRunWarps * 32, kBytes, st>>>(out, N, tiles, tile_tot, reinterpret_cast<unsigned int*>(flags + ntiles),
                                                               static_cast<NormPoison*>(poison), any, P);
    } else {
        k_normalize<<<ntiles, kNormThreads, 0, st>>>(out, N, tiles, tile_tot, flags, reinterpret_cast<unsigned int*>(flags + ntiles),
                                                        static_cast<NormPoison*>(poison), any, P);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_norm_poison<<<n_streams * kNF, 256, 0, st>>>(out, N, row_off, static_cast<const NormPoison*>(poison), any, carry, slice_row0, P);
    return cudaGetLastError();
}

// the fix-up pass on its own: applies a carry to z-scores that are already there (a time shard that learns later what its
// predecessors saw); also the way to read the merged state back
cudaError_t launch_norm_poison(const DevOutputs& out, int64_t N, void* poison, const int64_t* row_off, int n_streams, long long* carry,
                               const long long* slice_row0, const DevParams& P, cudaStream_t st) {
    const size_t pb = (size_t)n_streams * kNF * sizeof(NormPoison);
    const int* any = reinterpret_cast<const int*>(static_cast<unsigned char*>(poison) + pb);
    k_norm_poison<<<n_streams * kNF, 256, 0, st>>>(out, N, row_off, static_cast<const NormPoison*>(poison), any, carry, slice_row0, P);
    return cudaGetLastError();
}

// ================================================================================================
// Global standardisation of the z-scored columns — what the reference's Python stage does next with sklearn's StandardScaler
// (regime_classifier/python/classifier.py:35-37, data_checks.py:78-80: mean and population standard deviation of every column over ALL
// rows of ALL dates, NaN cells skipped).  Moments are additive, so ranks all-reduce 19 x {n, sum, sum of squares} once and every
// rank scales its own rows.  Deterministic: block b of column c always sums the same rows in the same order.
// ================================================================================================
constexpr int kMomRows = 8192, kMomThreads = 256;

__global__ void __launch_bounds__(kMomThreads)
k_col_moments(const double* __restrict__ x, const int64_t col_stride, const int64_t N, double* __restrict__ part) {
    __shared__ double sh[3][kMomThreads / 32];
    const int c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * kMomRows;
    const int64_t r1 = r0 + kMomRows < N ? r0 + kMomRows : N;
    const double* __restrict__ col = x + (int64_t)c * col_stride;
    double cnt = 0.0, s = 0.0, s2 = 0.0;
    for (int64_t r = r0 + tid; r < r1; r += kMomThreads) {
        const double v = col[r];
        if (v == v) {   // NaN cells are missing values for the scaler
            cnt += 1.0;
            s += v;
            s2 = fma(v, v, s2);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        s += __shfl_xor_sync(0xffffffffu, s, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) { sh[0][warp] = cnt; sh[1][warp] = s; sh[2][warp] = s2; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0, q = 0.0;
        for (int w = 0; w < kMomThreads / 32; ++w) { a += sh[0][w]; b += sh[1][w]; q += sh[2][w]; }
        double* o = part + ((int64_t)blockIdx.x * kNF + c) * 3;
        o[0] = a; o[1] = b; o[2] = q;
    }
}

// out = (x - mean) / scale per column; ms = [19][2] {mean, scale} on the device
__global__ void __launch_bounds__(256)
k_standardize(const double* __restrict__ x, double* __restrict__ out, const int64_t N, const double* __restrict__ ms) {
    const int c = blockIdx.y;
    const double mean = ms[c * 2], scale = ms[c * 2 + 1];
    const double* __restrict__ xi = x + (int64_t)c * N;
    double* __restrict__ xo = out + (int64_t)c * N;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x) xo[r] = (xi[r] - mean) / scale;
}

// dst[i] = src[i] - base: the cursors of a packed halo count from its own first tick / trade
__global__ void k_rebase(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const int64_t n, const uint32_t base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i] - base;
}
cudaError_t launch_rebase(const uint32_t* src, uint32_t* dst, int64_t n, uint32_t base, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_rebase<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, st>>>(src, dst, n, base);
    return cudaGetLastError();
}

int col_moment_blocks(int64_t N) { return (int)((N + kMomRows - 1) / kMomRows); }

// column c of the matrix starts at x + c * col_stride; N rows of it are summed
cudaError_t launch_col_moments(const double* x, int64_t col_stride, int64_t N, double* part, cudaStream_t st) {
    if (N == 0) return cudaSuccess;
    k_col_moments<<<dim3((unsigned)col_moment_blocks(N), kNF), kMomThreads, 0, st>>>(x, col_stride, N, part);
    return cudaGetLastError();
}

cudaError_t launch_standardize(const double* x, double* out, int64_t N, const double* mean_scale_dev, cudaStream_t st) {
    if (N == 0) return cudaSuccess;
    const unsigned gx = (unsigned)std::min<int64_t>((N + 255) / 256, 148 * 16);
    k_standardize<<<dim3(gx, kNF), 256, 0, st>>>(x, out, N, mean_scale_dev);
    return cudaGetLastError();
}

cudaError_t launch_normalize_faithful(const DevOutputs& out, int64_t N, const int64_t* row_off, int n_streams,
                                      const DevParams& P, cudaStream_t st) {
    const int total = n_streams * kNF;
    k_normalize_faithful<<<(total + 31) / 32, 32, 0, st>>>(out, N, row_off, n_streams, P);
    return cudaGetLastError();
}

}  // namespace mrp
