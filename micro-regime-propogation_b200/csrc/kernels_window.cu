// kernels_window.cu — the cursor-addressed rolling-window features (tick axis + trade axis), one kernel.
//
//   tick axis : ewm_volatility, realized_variance, directional_volatility, spread_volatility,
//               tick_direction_entropy      (ref: feature_processor.cpp:48-107, :215-229; feature_engine.cpp:130-150)
//   trade axis: reversal_rate, aggressor_bias, shannon_entropy
//                                           (ref: feature_processor.cpp:231-247, :254-269; feature_engine.cpp:87-93)
//
// The reference re-walks a 1800-tick (1500-trade) deque for every snapshot.  Here every window quantity is a
// DIFFERENCE OF TWO PREFIXES kept in shared-memory rings, so a snapshot costs a handful of shared loads, whatever the
// window length, and every tick / trade is read from HBM exactly once:
//
//   per tick k (the 50 ms sample M_k, s_k), with x_k = log(M_k / M_{k-1})^2 and d_k = s_k - pivot:
//        QA[k+1] = QA[k] + x_k            QS[k+1] = QS[k] +- x_k   (+ when the midprice rose)
//        QD1[k+1] = QD1[k] + d_k          QD2[k+1] = QD2[k] + d_k^2
//        QG[k+1] = (1-alpha) QG[k] + alpha x_k                      (a plain EMA over all ticks)
//        QC[k+1] = QC[k] + (did not rise | moved << 16)             (two 16-bit counters, wrapping)
//   a snapshot with tick cursor e, s = max(e - 1800, 0), cnt = e - s - 1 returns:
//        sum r^2 = QA[e] - QA[s+1],  up - down = QS[e] - QS[s+1],  spread sums = QD*[e] - QD*[s],
//        restarted EW variance (feature_processor.cpp:64-79) = QG[e] - (1-alpha)^cnt QG[s+1] + (1-alpha)^cnt x_{s+1},
//        direction counts = QC[e] - QC[e-1500].
//   per trade j: QT[j+1] = QT[j] + (pos, neg, reversal, dir) as four wrapping 16-bit counters in one 64-bit word.
//
// One CTA streams through one row segment in time order; the prefixes restart at the segment's first window (a halo
// of 1802 ticks / 1502 trades — no halo ROWS), so their magnitude stays within ~100 window sums and a difference
// keeps ~14 digits.  Work is split by warp role (288 threads, one CTA per SM: rings + staging take 215 KB), ONE CTA barrier per chunk of
// <= 128 rows / 1408 ticks plus named barriers inside the roles:
//   warps 0-3  producers: the chunk's new ticks, 11 consecutive ticks per thread out of a staging buffer the copy engine
//              filled (cp.async.bulk + mbarrier, requested one chunk ahead; no thread issues a global load for a tick):
//              one reciprocal per two ticks and a 6-term log1p series instead of two logs, thread-local prefixes in
//              registers, one warp scan + one cross-warp pass per quantity, results into the rings;
//   warps 4-7  consumers: plan the next chunk (cursors, how many rows fit, request its ticks), gather the PREVIOUS chunk's
//              prefix values while the producers work on this one, finalise those rows from registers;
//   warp 8     the trade prefixes (10 trades per lane and pass; a chunk's ~256 trades are one pass): everything but the ring
//              stores happens before the "previous chunk has been gathered" barrier.
// A snapshot that brings more new ticks / trades than a chunk holds simply advances the rings in several steps.
// Three roles execute three instruction streams at once through one 32 KB instruction cache, so the trade warp runs rolled loops,
// the ring-store loop exists once (the rings are padded by one run; the rare overhang is copied to the head afterwards) and the
// cold paths (library square root, count columns, direct_tick_row) are out of line.  The rows' finalisation, on the other hand, is
// deliberately ONE basic block (selects, branch-free roots, inline quotients): its latency, not its size, was on the critical path.
//
// Ticks the fast formulation cannot represent — a non-positive / non-finite midprice, a 50 ms move of 2^-10 relative or
// more (series range), a non-finite spread — contribute nothing to the prefixes and are flagged; every snapshot whose
// window contains a flagged tick is evaluated by direct_tick_row(), a plain transcription of the reference loops with
// true logs.  There is no CPU path.
#include "mrp_device.cuh"

namespace mrp {

#ifndef MRP_WIN_PROD_WARPS
#define MRP_WIN_PROD_WARPS 4
#endif
#ifndef MRP_WIN_RUN
#define MRP_WIN_RUN 11
#endif
#ifndef MRP_WIN_WHATIF
#define MRP_WIN_WHATIF 0   // timing experiments only (bit 0: no tick math, 1: no finalisation, 2: no warp scan, 3: no ring stores) — results are WRONG when set
#endif
#ifndef MRP_WIN_SPLIT_SPREAD
#define MRP_WIN_SPLIT_SPREAD 0   // 1: the spread prefixes (QD1, QD2) are produced by warps of their own beside the midprice producers
#endif
constexpr bool kSplit = MRP_WIN_SPLIT_SPREAD != 0;
constexpr int kProdWarps = MRP_WIN_PROD_WARPS, kConsWarps = 4;
constexpr int kProd = kProdWarps * 32;             // (midprice) producer threads
constexpr int kSpread = kSplit ? kProd : 0;        // spread producer threads: the same ticks per thread as the midprice producers
constexpr int kFirstCons = kProd + kSpread;        // first consumer thread
constexpr int kCons = kConsWarps * 32;             // consumer threads == most rows of one chunk
constexpr int kWinThreads = kFirstCons + kCons + 32;    // + the trade warp
constexpr int kRun = MRP_WIN_RUN;                  // consecutive ticks per producer thread (odd: conflict-free 64-bit shared accesses); long
                                                   // runs keep the per-thread scan overhead (~300 instructions) small next to the tick arithmetic
constexpr int kTradeRun = 10;                      // consecutive trades per lane and pass of the trade warp: a chunk of 128 rows brings
                                                   // ~260 trades, one pass of 320
#ifndef MRP_WIN_DEPTH
#define MRP_WIN_DEPTH 1
#endif
#ifndef MRP_WIN_PROD_REGS
#define MRP_WIN_PROD_REGS 1      // producers: 1 = the thread's run in registers, fully unrolled; 0 = short rolled loops over the staging buffer
#endif
#ifndef MRP_WIN_ORDER
#define MRP_WIN_ORDER 1          // consumers: 0 = gather, finalise, plan; 1 = plan, gather, finalise; 2 = finalise (previous iteration's rows), plan, gather
#endif
#ifndef MRP_WIN_MATH_UNROLL
#define MRP_WIN_MATH_UNROLL 2    // tick pairs per trip of the producers' arithmetic loop
#endif
#ifndef MRP_WIN_STORE_UNROLL
#define MRP_WIN_STORE_UNROLL 1   // ticks per trip of the producers' ring loop
#endif
#ifndef MRP_WIN_SLOG_SMEM
#define MRP_WIN_SLOG_SMEM 1          // the c*log2(c) table in shared memory (from global memory through the L1: +0.35 ms at 50 M rows)
#endif
// Build-time switches of the organisation, all measured at 50 M rows (profiles/r2b_windows_experiments.md).
// The defaults are the fastest combination; none of the others is faster, which is the finding: the kernel is neither
// copy-latency-bound (kDepth 1 -> 2 -> 3: 6.16 / 6.17 / 6.8 ms) nor instruction-cache-bound (the hot loops were cut from 44 KB to
// 25-31 KB, the SM's instruction-cache hit rate went from 68 % to 97.7-99.9 % and the GPC-level cache from 91 % to 2-24 % of its request
// peak: same time), and its time follows neither its instruction count nor its longest warp.  (Times below are from before the
// finalisation became one basic block, which took the default from 6.02 to 5.30 ms.)
//   MRP_WIN_DEPTH        chunks planned (and requested from the copy engine) ahead of the producers
//   MRP_WIN_PROD_REGS    1: a producer thread keeps its 11-tick run in registers, everything unrolled; 0: short rolled loops
//                        that overwrite the staged midprices with x in place (7 KB less code, but +20 % instructions: 8.2 ms
//                        rolled, 6.7 ms with the arithmetic loop unrolled)
//   MRP_WIN_ORDER        consumers: 1 = plan, gather, finalise; 0 = gather, finalise, plan (rings released 800 cycles earlier: slower);
//                        2 = finalise the previous iteration's rows first, then plan and gather (7.04 ms: the ring stores wait for the gather)
//   MRP_WIN_PROD_WARPS / MRP_WIN_RUN   4 x 11 (default) 6.16 ms; 5 x 9 6.70; 6 x 7 6.48; 7 x 7 6.38; 8 x 5 6.45
//   MRP_WIN_SPLIT_SPREAD 1: four more warps produce the spread prefixes (QD1, QD2), the midprice producers' loop shrinks from 1027
//                        to 827 instructions — and the kernel goes from 6.02 to 6.22 ms (460 -> 501 M warp-instructions per 8 M rows at
//                        41 -> 43 % issue-active): it is the TOTAL instruction count at ~0.45 instructions per scheduler-cycle that
//                        sets the time, not the longest warp
[[maybe_unused]] constexpr int kMathUnroll = MRP_WIN_MATH_UNROLL, kStoreUnroll = MRP_WIN_STORE_UNROLL;
constexpr int kDepth = MRP_WIN_DEPTH;              // chunks in flight ahead of the one being produced
constexpr int kStages = kDepth + 1;                // staging buffers
constexpr int kPlans = 8;                          // plan slots: chunk q lives in slot q & 7 (planned in iteration q - kDepth, gathered in q + 1)
static_assert(kDepth >= 1 && kDepth + 3 <= kPlans, "plan slots");
constexpr int kRingPad = ((MRP_WIN_RUN + 1) / 2) * 2;   // entries past the end of each tick ring (a whole thread run; even: 16-byte alignment of what follows)
constexpr int kChunkTrades = 512;                  // most new trades of one chunk (a chunk of 128 rows brings ~260)
constexpr int kTradeStage = kChunkTrades + 32;     // staged trade bytes of a chunk (chunk_trades + 16-byte rounding on both ends)
// the log1p series covers relative moves below 2^-10 per tick (high word 0x3f500000)

// 1/a for a normal, positive a (callers flag other inputs themselves): seed + two Newton steps
__device__ __forceinline__ double rcp_pos(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}

// log(1+u) for |u| <= 2^-10: truncation below u^5/6 = 2^-52 relative.
__device__ __forceinline__ double log1p_small(double u) {
    double q = 1.0 / 5.0;
    q = fma(q, u, -1.0 / 4.0);
    q = fma(q, u, 1.0 / 3.0);
    q = fma(q, u, -0.5);
    q = fma(q, u, 1.0);
    return q * u;
}

__device__ __forceinline__ uint32_t sub_floor0(uint32_t a, uint32_t w) { return a > w ? a - w : 0u; }

// sqrt(v) for v >= 0: the reciprocal-root seed, one coupled Newton step, one correction of the root (the remainder v - s^2 is
// exact in an fma) — within an ulp of the correctly rounded root for a third of the library routine's instructions; inputs
// outside the fast range (zero, denormal, huge, NaN) take the library root
__device__ __noinline__ double sqrt_library(double v) { return sqrt(v); }   // out of line: the hot loops stay small
__device__ __forceinline__ double sqrt_fast(double v) {
    if (!rsqrt_pos_ok(v)) return sqrt_library(v);
    const double r = rsqrt_pos(v);
    const double s = v * r;
    return fma(fma(-s, s, v), 0.5 * r, s);
}
// The same root without a branch: zero gives zero, every other input outside the fast range raises `rare` (the caller recomputes
// those with the library root in ONE out-of-line call after all its roots) — three roots, two reciprocals and three quotients of
// a row then sit in one basic block and overlap, instead of running one after the other between branches and calls.
__device__ __forceinline__ double sqrt_fast_nb(double v, bool& rare) {
    const bool ok = rsqrt_pos_ok(v);
    rare |= !ok && v != 0.0;
    const double r = rsqrt_pos(ok ? v : 1.0);
    const double s = v * r;
    const double f = fma(fma(-s, s, v), 0.5 * r, s);
    return v == 0.0 ? 0.0 : f;
}
__device__ __noinline__ void sqrt_rare3(double a, double b, double c, double* out) {
    out[0] = sqrt(a);
    out[1] = sqrt(b);
    out[2] = sqrt(c);
}
__device__ __forceinline__ bool small_mag(double v, uint32_t hi_limit) {  // |v| < the double whose high word is hi_limit; NaN -> false
    return ((uint32_t)__double2hiint(v) & 0x7fffffffu) < hi_limit;
}
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- direct evaluation of one snapshot's tick-axis features: the reference loops, in the reference order
//      (feature_processor.cpp:48-107, :215-229).  Used for windows that contain a flagged tick.
struct DirectParams {
    uint32_t mid_window, dir_window;
    int spread_divisor;
    double ewm_alpha, ewm_decay;
};

// the optional integer count columns are written out of line: the hot loops of the three roles share a 32 KB instruction cache
__device__ __noinline__ void write_tick_counts(int32_t* __restrict__ counts, int64_t N, int64_t r, int n_valid, int n_up, int t_up, int t_dn,
                                               int t_tot) {
    counts[(int64_t)MRP_C_RET_VALID * N + r] = n_valid;
    counts[(int64_t)MRP_C_RET_UP * N + r] = n_up;
    counts[(int64_t)MRP_C_TICK_UP * N + r] = t_up;
    counts[(int64_t)MRP_C_TICK_DOWN * N + r] = t_dn;
    counts[(int64_t)MRP_C_TICK_TOTAL * N + r] = t_tot;
}
__device__ __noinline__ void write_trade_counts(int32_t* __restrict__ counts, int64_t N, int64_t r, int pos, int neg, int rev, int nd, int isum) {
    counts[(int64_t)MRP_C_TRADE_POS * N + r] = pos;
    counts[(int64_t)MRP_C_TRADE_NEG * N + r] = neg;
    counts[(int64_t)MRP_C_TRADE_REVERSALS * N + r] = rev;
    counts[(int64_t)MRP_C_TRADE_TOTAL * N + r] = nd;
    counts[(int64_t)MRP_C_TRADE_DIR_SUM * N + r] = isum;
}
__device__ __forceinline__ void write_tick_row(const DevOutputs& out, int64_t N, int64_t r, double ewm, double rv, double dv, double sv,
                                               double tent, int n_valid, int n_up, int t_up, int t_dn, int t_tot) {
    out.raw[(int64_t)MRP_F_EWM_VOLATILITY * N + r] = ewm;
    out.raw[(int64_t)MRP_F_REALIZED_VARIANCE * N + r] = rv;
    out.raw[(int64_t)MRP_F_DIRECTIONAL_VOLATILITY * N + r] = dv;
    out.raw[(int64_t)MRP_F_SPREAD_VOLATILITY * N + r] = sv;
    out.raw[(int64_t)MRP_F_TICK_DIRECTION_ENTROPY * N + r] = tent;
    if (out.counts) write_tick_counts(out.counts, N, r, n_valid, n_up, t_up, t_dn, t_tot);
}

__device__ __noinline__ void direct_tick_row(const double* __restrict__ mid, const double* __restrict__ spread, uint32_t e,
                                             const DirectParams P, double* __restrict__ raw, int32_t* __restrict__ counts, int64_t N,
                                             int64_t r) {
    const uint32_t s = sub_floor0(e, P.mid_window);
    double rv = 0.0, ewm_var = 0.0, up_var = 0.0, down_var = 0.0;
    int count = 0, up_count = 0, down_count = 0;
    if (e > s) {
        double a = mid[s], la = log(a);
        for (uint32_t k = s + 1; k < e; ++k) {
            const double b = mid[k], lb = log(b);
            if (!(a == 0.0 || b == 0.0)) {                       // :56
                const double ret = lb - la, x = ret * ret;        // :57
                rv += x;                                          // :58
                ewm_var = count == 0 ? x : P.ewm_decay * ewm_var + P.ewm_alpha * x;  // :71-75
                if (ret > 0) { up_var += x; ++up_count; }         // :87-92
                else { down_var += x; ++down_count; }
                ++count;
            }
            a = b;
            la = lb;
        }
    }
    const double o_ewm = count > 0 ? sqrt(ewm_var) : 0.0;         // ewm_volatility, :79
    const double o_rv = count > 0 ? rv / count : 0.0;             // realized_variance, :61
    const double au = up_count > 0 ? up_var / up_count : 0.0, ad = down_count > 0 ? down_var / down_count : 0.0;
    const double diff = au - ad;
    const double o_dv = sqrt(fabs(diff)) * ((diff >= 0) ? 1.0 : -1.0);  // directional_volatility, :95-98
    double acc = 0.0;                                             // spread_volatility, :101-107
    for (uint32_t k = s; k < e; ++k) acc += spread[k];
    const double mean = acc / P.spread_divisor;
    double var = 0.0;
    for (uint32_t k = s; k < e; ++k) var += (spread[k] - mean) * (spread[k] - mean);
    const double o_sv = sqrt(var / P.spread_divisor);
    int up = 0, down = 0, zero = 0;                               // tick_direction_entropy, :217-229
    const uint32_t d0 = sub_floor0(e, P.dir_window);
    for (uint32_t k = d0; k < e; ++k) {
        int dir = 0;
        if (k > 0) {
            const double cur = mid[k], prev = mid[k - 1];
            dir = (cur > prev) ? 1 : ((cur < prev) ? -1 : 0);    // feature_engine.cpp:138
        }
        if (dir > 0) ++up; else if (dir < 0) ++down; else ++zero;
    }
    double ent = 0.0;
    if (e > d0) {
        const double total = up + down + zero;
        ent = entropy_term(entropy_term(entropy_term(0.0, up / total), down / total), zero / total);
    }
    DevOutputs o{};
    o.raw = raw;
    o.counts = counts;
    write_tick_row(o, N, r, o_ewm, o_rv, o_dv, o_sv, ent, count, up_count, up, down, (int)(e - d0));
}

// a / b for a positive count b with r = rcp_pos(b): product, ONE correction step (the quotient's remainder is exact in an fma) — the
// correctly rounded quotient for the small integers that meet here, without a division and without a table; quotients over the
// same count share the reciprocal
__device__ __forceinline__ double quotient_by(double a, double b, double r) {
    const double q = a * r;
    return fma(fma(-q, b, a), r, q);
}

// trade-axis features of one snapshot from its window counts (feature_processor.cpp:231-247, :254-269)
// clog2c: the host's c*log2(c) table, copied to shared memory
__device__ __forceinline__ void write_trade_row(const DevOutputs& out, int64_t N, int64_t r, const double* __restrict__ clog2c, int nd, int pos,
                                                int neg, int rev, int isum) {
    const int tot = pos + neg;                                  // :261-269
    // three quotients on two reciprocals, guarded denominators and selects: one basic block
    const double bn = (double)(nd > 0 ? nd : 1), rn = rcp_pos(bn);
    const double bt = (double)(tot > 0 ? tot : 1), rt = rcp_pos(bt);
    const double rr = nd > 1 ? quotient_by((double)rev, bn, rn) : 0.0;                                        // :239 — divides by size, not size-1
    const double ab = nd > 0 ? quotient_by((double)isum, bn, rn) : __longlong_as_double(0x7ff8000000000000ll);  // :246 — 0/0 = NaN with no trades
    // binary entropy = (T log2 T - pos log2 pos - neg log2 neg) / T
    const double ent = tot > 0 ? quotient_by((clog2c[tot] - clog2c[pos]) - clog2c[neg], bt, rt) : 0.0;
    out.raw[(int64_t)MRP_F_REVERSAL_RATE * N + r] = rr;
    out.raw[(int64_t)MRP_F_AGGRESSOR_BIAS * N + r] = ab;
    out.raw[(int64_t)MRP_F_SHANNON_ENTROPY * N + r] = ent;
    if (out.counts) write_trade_counts(out.counts, N, r, pos, neg, rev, nd, isum);
}

// What a consumer thread lifts out of the rings for its row between the two barriers of a chunk.
struct RowGather {
    double SA, SS;        // sum x, sum +-x over the returns (s, e)
    double D1, D2;        // sum d, sum d^2 over the spreads [s, e)
    double Ge, Gs1;       // the running EMA at e and at s + 1
    double xf;            // x of the window's first return (tick s + 1)
    int negs, movm;       // sign-set / moved counts over the returns (s, e)
    int negd, movd;       // the same over the direction window [d0, e)
    int pos, neg, rev, isum;   // trade window counts
    uint32_t e, je;
    long long row;        // -1: nothing gathered
    bool flagged;
};

// One snapshot's five tick-axis features from its window values (see the file header).
__device__ __forceinline__ void finish_tick_row(const DevOutputs& out, const int64_t N, const DevParams& P, const double* __restrict__ clog2c,
                                                const double pivot, const RowGather& g) {
    double ewm = 0.0, rv = 0.0, dv = 0.0, sv = 0.0, tent = 0.0;
    int cnt = 0, nUp = 0, t_up = 0, t_dn = 0, t_tot = 0;
    const uint32_t e = g.e;
    if (e > 0u) {
        // One basic block: selects instead of branches, so that the independent chains (three roots, two reciprocals, the entropy's
        // table look-ups) overlap.  Values computed for a case that does not apply are discarded by the selects.
        const uint32_t s = sub_floor0(e, (uint32_t)P.mid_window);
        cnt = (int)(e - s) - 1;
        const bool has = cnt > 0;
        nUp = has ? cnt - g.negs : 0;
        // a window of flat ticks only is exactly zero in the reference; do not let rounding residue of the prefix
        // differences through the square roots
        const bool moved = has && g.movm > 0;
        const double tot = (g.SA > 0.0 && moved) ? g.SA : 0.0;
        const bool full = cnt == P.mid_window - 1;   // the steady state: the window's constants sit in the kernel parameters
        double pw = P.pow_ret_full, inv_cnt = P.inv_ret_full;
        if (has && !full) {   // the first window of a stream only: no load in flight otherwise
            asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(pw) : "l"(P.ewm_pow + cnt));
            inv_cnt = rcp_pos((double)cnt);
        }
        rv = tot * inv_cnt;  // feature_processor.cpp:61
        // :64-79: sum_{i>=1} alpha decay^(cnt-i) x_i + decay^cnt x_1  (the restart takes x_1 with weight decay^(cnt-1))
        double w = fma(pw, g.xf - g.Gs1, g.Ge);
        if (!(tot != 0.0 && w > 0.0)) w = 0.0;
        double sumU = 0.5 * (g.SA + g.SS), sumD = 0.5 * (g.SA - g.SS);
        if (!(sumU > 0.0) || !moved) sumU = 0.0;
        if (!(sumD > 0.0) || !moved) sumD = 0.0;
        const int nDn = cnt - nUp;
        const double au = nUp > 0 ? sumU * rcp_pos((double)(nUp > 0 ? nUp : 1)) : 0.0;          // :95-98 (an empty side counts 0)
        const double ad = (has && nDn > 0) ? sumD * rcp_pos((double)(nDn > 0 ? nDn : 1)) : 0.0;
        const double diff = au - ad;
        // :101-107 with s_k = pivot + d_k: mean = (sum s)/1500, var = sum (s_k - mean)^2 / 1500
        const double n = (double)(e - s);
        const double mean = (g.D1 + n * pivot) * P.inv_spread_divisor;
        const double delta = mean - pivot;
        double ss = g.D2 - 2.0 * delta * g.D1 + n * delta * delta;
        if (!(ss > 0.0)) ss = 0.0;
        const double r0 = w, r1 = fabs(diff), r2 = ss * P.inv_spread_divisor;
        bool rare = false;
        double q0 = sqrt_fast_nb(r0, rare), q1 = sqrt_fast_nb(r1, rare), q2 = sqrt_fast_nb(r2, rare);
        if (rare) {   // denormal, huge or NaN radicands: the library root for all three
            double q[3];
            sqrt_rare3(r0, r1, r2, q);
            q0 = q[0]; q1 = q[1]; q2 = q[2];
        }
        ewm = has ? q0 : 0.0;
        dv = has ? q1 * ((diff >= 0) ? 1.0 : -1.0) : 0.0;
        sv = q2;
        const uint32_t d0 = sub_floor0(e, (uint32_t)P.dir_window);
        t_tot = (int)(e - d0);
        t_up = t_tot - g.negd;            // sign clear <=> the midprice rose
        t_dn = g.negd + g.movd - t_tot;   // sign set and moved
        const int zero = t_tot - g.movd;
        // :222-229: -sum p log2 p = (T log2 T - sum c log2 c) / T, c*log2(c) tabulated by the host
        const bool dfull = t_tot == P.dir_window;
        tent = (((dfull ? P.clog2c_dir_full : clog2c[t_tot]) - clog2c[t_up]) - clog2c[t_dn] - clog2c[zero]) *
               (dfull ? P.inv_dir_full : rcp_pos((double)t_tot));
        if (cnt < 0) cnt = 0;
    }
    write_tick_row(out, N, g.row, ewm, rv, dv, sv, tent, cnt, nUp, t_up, t_dn, t_tot);
}

// four wrapping 16-bit counters in one word: field-wise a + b
__device__ __forceinline__ unsigned long long add4x16(unsigned long long a, unsigned long long b) {
    const unsigned long long H = 0x8000800080008000ull;
    return ((a & ~H) + (b & ~H)) ^ ((a ^ b) & H);
}

#ifdef MRP_WIN_TIMERS
// experiment builds only: per-role phase timers, summed over all CTAs (read with mrp_fg_debug_win_timers)
__device__ unsigned long long g_win_tm[3][12];
// the clock is read through an asm statement with a memory clobber: a plain clock64() may be moved across the barrier next to it, which
// books a barrier wait on the phase AFTER the barrier (the first timer runs of this session did exactly that)
__device__ __forceinline__ long long tm_clock() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#define TM_DECL long long tm_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tm_last = tm_clock(); long long tm_iters = 0
#define TM(i) do { const long long tm_now = tm_clock(); tm_acc[i] += tm_now - tm_last; tm_last = tm_now; } while (0)
#define TM_FLUSH(role) do { if (lane == 0) { for (int i_ = 0; i_ < 10; ++i_) atomicAdd(&g_win_tm[role][i_], (unsigned long long)tm_acc[i_]); atomicAdd(&g_win_tm[role][10], (unsigned long long)tm_iters); } } while (0)
extern "C" int mrp_fg_debug_win_timers(unsigned long long* out) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_win_tm, sizeof(g_win_tm)) != cudaSuccess) return -1;
    unsigned long long z[36] = {0};
    return cudaMemcpyToSymbol(g_win_tm, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#else
#define TM_DECL
#define TM(i)
#define TM_FLUSH(role)
#endif
enum { PLAN_RUN = 0, PLAN_END = 1 };
struct __align__(16) WinPlan {
    long long c0;              // first row of the chunk
    uint32_t TFprev, TF;       // tick frontier before / after the chunk: the chunk's new ticks are [TFprev, TF)
    uint32_t JFprev, JF;       // trade frontier
    uint32_t slotT, slotJ;     // ring slots of the positions TFprev / JFprev
    uint32_t stg0;             // tick position of the first staged element
    int shM, shS, shT;         // element shift of the staged mids / spreads / trade bytes (16-byte rounding of the copy)
    int nr, mode;              // rows finalised by this chunk (0: the rings only advance)
};
struct __align__(16) WinStatic {
    WinPlan plan[kPlans];              // chunk q lives in slot q & 7: planned in iteration q - kDepth, produced in q, gathered in q + 1
    uint64_t full[kStages];            // mbarriers of the staging buffers
    double wt[kProdWarps][6];          // per-warp totals: x, +-x, d, d^2, EMA, counts (bit pattern)
    double wtS[kProdWarps][2];         // kSplit: the spread producers' per-warp totals d, d^2
    __align__(16) unsigned char stgT[kStages][kTradeStage];   // staged trade directions of the chunk (position JFprev - 1 onwards)
    double carryG;                     // the EMA after the chunk's last tick
    uint32_t excP1;                    // 1 + position of the last flagged tick (0 = none)
    int cnt[kConsWarps];
};
// named barriers: 1 the producers (cross-warp scan), 2 the consumers (plan), 3 "the previous chunk has been gathered":
// the consumers arrive, the producers and the trade warp wait before they overwrite ring entries
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(kWinThreads, 1)
k_windows(const DevInputs in, const DevOutputs out, const DevSegment* __restrict__ segs, const DevParams P) {
    extern __shared__ __align__(16) unsigned char win_dyn[];
    __shared__ WinStatic sh;
    const int R = P.win_ring, RJ = P.win_trade_ring, CT = P.win_chunk_ticks, CJ = P.win_chunk_trades, CTS = CT + 4;
    // every tick ring has kRingPad entries past its R: a thread stores its run linearly (no wrap test per element) and the one
    // thread in ~300 whose run crosses the end copies the overhang to the ring's head afterwards
    const int RP = R + kRingPad;
    double* __restrict__ QA = reinterpret_cast<double*>(win_dyn);
    double* __restrict__ QS = QA + RP;
    double* __restrict__ QD1 = QS + RP;
    double* __restrict__ QD2 = QD1 + RP;
    double* __restrict__ QG = QD2 + RP;
    double* __restrict__ stgM = QG + RP;                 // [kStages][CTS]
    double* __restrict__ stgS = stgM + kStages * CTS;    // [kStages][CTS]
    unsigned long long* __restrict__ QT = reinterpret_cast<unsigned long long*>(stgS + kStages * CTS);   // [RJ]
#if MRP_WIN_SLOG_SMEM
    double* __restrict__ sLog = reinterpret_cast<double*>(QT + RJ);                                 // [dir_window + 2]  c*log2(c), from the host
    uint32_t* __restrict__ QC = reinterpret_cast<uint32_t*>(sLog + P.dir_window + 2);                // [R]
#else
    const double* __restrict__ sLog = P.clog2c;          // seven look-ups per row in a 12 KB table: the L1 serves them
    uint32_t* __restrict__ QC = reinterpret_cast<uint32_t*>(QT + RJ);                                // [R]
#endif

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevSegment sg = segs[blockIdx.x];
    const int64_t ra = sg.row_begin, rb = sg.row_end, N = in.n_rows;
    const double* __restrict__ mid = in.mid + sg.tick0;
    const double* __restrict__ spread = in.spread + sg.tick0;
    const int8_t* __restrict__ tdir = in.tdir + sg.trade0;
    const uint32_t* __restrict__ kcur = in.kcur;
    const uint32_t* __restrict__ jcur = in.jcur;
    const uint32_t nTicks = (uint32_t)sg.n_ticks, nTrades = (uint32_t)sg.n_trades;
    const uint32_t Wm = (uint32_t)P.mid_window, Wd = (uint32_t)P.dir_window;
    const uint32_t Wmax = max(Wm, Wd + 1);
    const double decay = P.ewm_decay, alpha = P.ewm_alpha;

    // the prefixes start one window before the segment's first row
    const uint32_t e_first = min(__ldg(kcur + ra), nTicks), j_first = min(__ldg(jcur + ra), nTrades);
    const uint32_t baseK = sub_floor0(e_first, Wmax + 2), baseJ = sub_floor0(j_first, Wd + 2);
    double pivot = 0.0;
    if (baseK < nTicks) {
        const double s0 = __ldg(spread + baseK);
        if (fabs(s0) <= 1e150) pivot = s0;
    }
#if MRP_WIN_SLOG_SMEM
    for (int i = tid; i < P.dir_window + 2; i += kWinThreads) sLog[i] = __ldg(P.clog2c + i);   // window counts never exceed dir_window
#endif
    if (tid == 0) {
        sh.excP1 = 0u;
        sh.carryG = 0.0;
        QA[0] = QS[0] = QD1[0] = QD2[0] = QG[0] = 0.0;
        QC[0] = 0u;
        QT[0] = 0ull;
#pragma unroll
        for (int i = 0; i < kStages; ++i) mbar_init(&sh.full[i], 1);
        mbar_init_fence();
    }
    __syncthreads();

    if (warp < kProdWarps) {
        // ================================================================================ producers: ticks -> prefix rings
        // the prefixes at the tick frontier, one quantity per lane: lane 0 sum x, 1 sum +-x, 2 sum d, 3 sum d^2, 4 the EMA (re-read
        // from sh.carryG every chunk), 5 the two counters did-not-rise | moved << 32 (the other lanes shadow lane 0)
        double car_d = 0.0;
        long long car_i = 0;
        // decay^(11*2^k), decay^(11*lane), decay^352: the EMA composes over whole thread runs / warps
        double pw2[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) pw2[k] = __ldg(P.ewm_pow + (kRun << k));
        const double plane = __ldg(P.ewm_pow + kRun * lane), pwarp = __ldg(P.ewm_pow + kRun * 32);
        __syncthreads();   // the first plans are visible
        int stage = 0;
        uint32_t phase = 0u;   // staging buffer of chunk q and the parity of its mbarrier phase
        TM_DECL;
        for (int q = 0;; ++q) {
            const WinPlan& pl = sh.plan[q & (kPlans - 1)];
            if (pl.mode != PLAN_RUN) break;
            const uint32_t TFp = pl.TFprev, TF = pl.TF;
            if (lane == 4) car_d = sh.carryG;   // the EMA at the frontier (written by the owner of the previous chunk's last tick, before the chunk barrier)
            TM(0);
            mbar_wait(&sh.full[stage], phase);
            TM(1);
            const int stage_q = stage;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            if (TF > TFp) {
                const double* __restrict__ sm = stgM + stage_q * CTS + pl.shM;
                const double* __restrict__ ss = stgS + stage_q * CTS + pl.shS;
                const uint32_t k0 = TFp + (uint32_t)(kRun * tid);
                const int i0 = (int)(k0 - pl.stg0);
                const int nv = k0 < TF ? (int)min(TF - k0, (uint32_t)kRun) : 0;   // valid ticks of this thread
#if MRP_WIN_PROD_REGS
                // the run lives in registers, everything unrolled (more instruction-level parallelism, three times the code)
                const bool has_prev = k0 > pl.stg0;
                double m[kRun + 1], d[kRun];
                {
                    const int ib = nv > 0 ? i0 : 1;          // threads without a run read the buffer's head
#pragma unroll
                    for (int i = 0; i < kRun; ++i) {
                        m[i + 1] = sm[ib + i];
                        d[i] = kSplit ? 0.0 : ss[ib + i] - pivot;
                    }
                    m[0] = sm[ib - 1];
                    if (nv == 0) m[1] = 1.0;
                    if (!(nv > 0 && has_prev)) m[0] = m[1];   // the tick at the base has no predecessor: flat
                    if (__any_sync(0xffffffffu, nv < kRun)) {   // only the warp that holds the frontier has partial or empty runs
#pragma unroll
                        for (int i = 0; i < kRun; ++i)
                            if (i >= nv) { m[i + 1] = m[i]; d[i] = 0.0; }
                    }
                }
                double x[kRun];
                uint32_t umax = 0u, rsign = 0u, nneg = 0u, nmov = 0u;
#pragma unroll
                for (int j = 0; j < (kRun + 1) / 2; ++j) {
                    const double rc = rcp_pos(m[2 * j + 1]);
                    rsign |= (uint32_t)__double2hiint(rc);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = 2 * j + h;
                        if (i < kRun) {
                            const double u = h == 0 ? (m[i] - m[i + 1]) * rc : (m[i + 1] - m[i]) * rc;   // h == 0: rose <=> u < 0
                            const double r = log1p_small(u);
                            const double xx = r * r;
                            const uint32_t uh = (uint32_t)__double2hiint(u), um = uh & 0x7fffffffu;
                            umax = max(umax, um);
                            const uint32_t mv = um != 0u ? 1u : 0u;
                            const uint32_t up = h == 0 ? (uh >> 31) : (mv & ~(uh >> 31));
                            nneg += 1u - up;
                            nmov += mv;
                            x[i] = __hiloint2double((int)((uint32_t)__double2hiint(xx) | ((1u - up) << 31)), __double2loint(xx));
                        }
                    }
                }
                nneg -= (uint32_t)(kRun - nv);   // the padding ticks of a partial run are flat: they counted as "did not rise"
                double va = 0.0, vs = 0.0, v1 = 0.0, v2 = 0.0, vg = 0.0;
#pragma unroll
                for (int i = 0; i < kRun; ++i) {
                    const double ax = fabs(x[i]);
                    va += ax;
                    vs += x[i];
                    v1 += d[i];
                    v2 = fma(d[i], d[i], v2);
                    vg = fma(decay, vg, alpha * ax);
                }
                TM(2);
                const bool bad = !(umax < 0x3f500000u && (rsign >> 31) == 0u && v2 < 1e300);
                if (bad) {
                    if (nv > 0) atomicMax(&sh.excP1, k0 + (uint32_t)nv);
#pragma unroll
                    for (int i = 0; i < kRun; ++i) { x[i] = 0.0; d[i] = 0.0; }
                    nneg = nmov = 0u;
                    va = vs = v1 = v2 = vg = 0.0;
                }
#else
                // This thread's ticks M_{k0..k0+nv-1} sit at xm[0..nv-1]; the tick before them belongs to the neighbour (or to the
                // previous chunk).  The arithmetic runs as a SHORT ROLLED loop that overwrites each midprice with its
                // x = +-log(M_k / M_{k-1})^2 in place (the ring pass after the scan reads it back): the smallest code of the
                // variants measured (the three roles' hot loops fit the SM's instruction cache with room to spare), not the fastest.
                double* __restrict__ xm = const_cast<double*>(sm) + (nv > 0 ? i0 : 1);
                const double* __restrict__ sp = ss + (nv > 0 ? i0 : 1);
                double mprev = xm[-1];
                if (!(k0 > pl.stg0)) mprev = xm[0];        // the tick at the base has no predecessor: flat
                bar_sync_n(1, kProd);                      // every predecessor has been read before anybody overwrites a midprice
                double va = 0.0, vs = 0.0, v1 = 0.0, v2 = 0.0, vg = 0.0;
                uint32_t umax = 0u, rsign = 0u, nneg = 0u, nmov = 0u;
#if MRP_WIN_WHATIF & 1
#pragma unroll 1
                for (int i = 0; i < nv; ++i) { const double u = xm[i] - mprev; mprev = xm[i]; xm[i] = u * u; va += u * u; }
#else
                // one reciprocal per two returns (that of the tick between them): log(M_k / M_{k-1}) = -log1p((M_{k-1} - M_k) / M_k),
                // log(M_{k+1} / M_k) = log1p((M_{k+1} - M_k) / M_k); x = +-r^2, the sign set unless the midprice rose (-0.0 for a flat tick)
#pragma unroll kMathUnroll
                for (int i = 0; i < kRun + 1; i += 2) {
                    const bool one = i < nv, two = i + 1 < nv;
                    const double ma = one ? xm[i] : 1.0;
                    const double mb = two ? xm[i + 1] : ma;
                    const double rc = rcp_pos(ma);
                    rsign |= (uint32_t)__double2hiint(rc);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double u = h == 0 ? ((one ? mprev : ma) - ma) * rc : (mb - ma) * rc;   // h == 0: rose <=> u < 0
                        const double r = log1p_small(u);
                        const double xx = r * r;
                        const uint32_t uh = (uint32_t)__double2hiint(u), um = uh & 0x7fffffffu;
                        umax = max(umax, um);
                        const uint32_t mv = um != 0u ? 1u : 0u;   // a real move is a normal number: its high word alone tells it from zero
                        const uint32_t up = h == 0 ? (uh >> 31) : (mv & ~(uh >> 31));
                        const bool live = h == 0 ? one : two;
                        const double x = __hiloint2double((int)((uint32_t)__double2hiint(xx) | ((1u - up) << 31)), __double2loint(xx));
                        if (live) {
                            const double dd = sp[i + h] - pivot;
                            xm[i + h] = x;
                            nneg += 1u - up;
                            nmov += mv;
                            va += xx;
                            vs += x;
                            v1 += dd;
                            v2 = fma(dd, dd, v2);
                            vg = fma(decay, vg, alpha * xx);
                        }
                    }
                    mprev = mb;
                }
#endif
                TM(2);
                // flagged: a midprice that is not a positive normal number (its reciprocal is negative, or the returns round it are huge / NaN),
                // a move beyond the series range, a spread that is not finite or absurd (its square leaves the range)
                const bool bad = !(umax < 0x3f500000u && (rsign >> 31) == 0u && v2 < 1e300);
                if (bad) {
                    if (nv > 0) atomicMax(&sh.excP1, k0 + (uint32_t)nv);
                    nneg = nmov = 0u;
                    va = vs = v1 = v2 = vg = 0.0;
                }
#endif
                uint32_t vc = nneg | (nmov << 16);
                // warp scans of the thread totals (inclusive), the EMA as a scan of y -> decay^11 y + b
#pragma unroll
                for (int k = 0; k < ((MRP_WIN_WHATIF & 4) ? 0 : 5); ++k) {
                    const int dd = 1 << k;
                    const double ua = shfl_up_d(va, dd), us = shfl_up_d(vs, dd), ug = shfl_up_d(vg, dd);
                    const double u1 = kSplit ? 0.0 : shfl_up_d(v1, dd), u2 = kSplit ? 0.0 : shfl_up_d(v2, dd);
                    const uint32_t uc = __shfl_up_sync(0xffffffffu, vc, dd);
                    if (lane >= dd) {
                        va += ua; vs += us;
                        if (!kSplit) { v1 += u1; v2 += u2; }
                        vg = fma(pw2[k], ug, vg);
                        vc += uc;
                    }
                }
                if (lane == 31) {
                    sh.wt[warp][0] = va; sh.wt[warp][1] = vs; sh.wt[warp][2] = v1; sh.wt[warp][3] = v2; sh.wt[warp][4] = vg;
                    sh.wt[warp][5] = __longlong_as_double((long long)(((unsigned long long)(vc >> 16) << 32) | (vc & 0xffffu)));
                }
                // exclusive within the warp (by shuffle, not by subtraction)
                double xa = shfl_up_d(va, 1), xs = shfl_up_d(vs, 1), xg = shfl_up_d(vg, 1);
                double x1 = kSplit ? 0.0 : shfl_up_d(v1, 1), x2 = kSplit ? 0.0 : shfl_up_d(v2, 1);
                uint32_t xc = __shfl_up_sync(0xffffffffu, vc, 1);
                if (lane == 0) { xa = xs = x1 = x2 = xg = 0.0; xc = 0u; }
                TM(3);
                bar_sync_n(1, kProd);
                TM(4);
                // across the warps: lane l (< 6) walks quantity l over the warp totals; the four sums and the EMA share one
                // fused multiply-add (factor 1 or decay^352), the counters an integer add
                double mine_d = 0.0, acc_d = car_d;
                long long mine_i = 0, acc_i = car_i;
                {
                    const int ql = lane < 6 ? lane : 0;
                    const double mul = ql == 4 ? pwarp : 1.0;
#pragma unroll
                    for (int w = 0; w < kProdWarps; ++w) {
                        if (w == warp) { mine_d = acc_d; mine_i = acc_i; }
                        const double t = sh.wt[w][ql];
                        acc_d = fma(mul, acc_d, t);
                        acc_i += __double_as_longlong(t);
                    }
                }
                double gA = __shfl_sync(0xffffffffu, mine_d, 0) + xa, gS = __shfl_sync(0xffffffffu, mine_d, 1) + xs;
                double g1 = kSplit ? 0.0 : __shfl_sync(0xffffffffu, mine_d, 2) + x1, g2 = kSplit ? 0.0 : __shfl_sync(0xffffffffu, mine_d, 3) + x2;
                double gG = fma(plane, __shfl_sync(0xffffffffu, mine_d, 4), xg);   // the EMA entering this thread's first tick
                const unsigned long long mC = (unsigned long long)__shfl_sync(0xffffffffu, mine_i, 5);
                uint32_t gNeg = (uint32_t)mC + (xc & 0xffffu), gMov = (uint32_t)(mC >> 32) + (xc >> 16);
                car_d = acc_d;   // each lane keeps its own quantity's prefix for the next chunk: no broadcast
                car_i = acc_i;
                // the consumers have lifted the previous chunk's values out of the rings: its far-end entries may be overwritten
                TM(5);
                bar_sync_n(3, kWinThreads);
                TM(6);
                // into the rings: the prefix AFTER tick k lives at position k + 1.  The running sums continue from the offsets.
                uint32_t slot = pl.slotT + (uint32_t)(kRun * tid) + 1u;
                if (slot >= (uint32_t)R) slot -= (uint32_t)R;
                // linear stores from the thread's first slot (the rings are padded by one run); the overhang is copied below
                double* __restrict__ pA = QA + slot;
                double* __restrict__ pS = QS + slot;
                double* __restrict__ p1 = QD1 + slot;
                double* __restrict__ p2 = QD2 + slot;
                double* __restrict__ pG = QG + slot;
                uint32_t* __restrict__ pC = QC + slot;
#if MRP_WIN_PROD_REGS
#pragma unroll
#else
#pragma unroll kStoreUnroll
#endif
                for (int i = 0; i < ((MRP_WIN_WHATIF & 8) ? 1 : kRun); ++i) {
                    if (i < nv) {
#if MRP_WIN_PROD_REGS
                        const double xv = x[i], dv = d[i];
#else
                        const double xv = bad ? 0.0 : xm[i], dv = bad ? 0.0 : sp[i] - pivot;
#endif
                        const double ax = fabs(xv);
                        gA += ax;
                        gS += xv;
                        if (!kSplit) {
                            g1 += dv;
                            g2 = fma(dv, dv, g2);
                        }
                        gG = fma(decay, gG, alpha * ax);
                        gNeg += (uint32_t)__double2hiint(xv) >> 31;
                        gMov += ax != 0.0 ? 1u : 0u;
                        pA[i] = gA; pS[i] = gS; pG[i] = gG;
                        if (!kSplit) { p1[i] = g1; p2[i] = g2; }
                        pC[i] = (gNeg & 0xffffu) | (gMov << 16);   // two wrapping 16-bit counters
                    }
                }
                if (slot + (uint32_t)nv > (uint32_t)R) {   // one thread in ~300: the run crossed the end of the rings
#pragma unroll 1
                    for (uint32_t j = (uint32_t)R - slot; j < (uint32_t)nv; ++j) {
                        const uint32_t w = slot + j - (uint32_t)R;
                        QA[w] = pA[j]; QS[w] = pS[j]; QG[w] = pG[j]; QC[w] = pC[j];
                        if (!kSplit) { QD1[w] = p1[j]; QD2[w] = p2[j]; }
                    }
                }
                if (nv > 0 && k0 + (uint32_t)nv == TF) sh.carryG = gG;   // the EMA after the chunk's last tick
            } else {
                bar_sync_n(3, kWinThreads);
            }
            TM(7);
            __syncthreads();   // the rings hold chunk q; plan q + 1 is visible
            TM(8);
#ifdef MRP_WIN_TIMERS
            ++tm_iters;
#endif
        }
        if (warp == 0) TM_FLUSH(0);
    } else if (kSplit && tid < kFirstCons) {
        // ================================================================================ spread producers: spreads -> QD1, QD2
        // The same ticks per thread as the midprice producers, two quantities instead of six: the longest warp of a chunk walks
        // ~800 instead of ~1040 instructions.  A thread that sees an absurd / non-finite spread flags its run like the midprice
        // producers do (the two roles zero their own contributions independently: windows that contain the run are evaluated directly).
        const int stid = tid - kProd, swarp = warp - kProdWarps;
        double cD1 = 0.0, cD2 = 0.0;
        __syncthreads();   // the first plans are visible
        int stage = 0;
        uint32_t phase = 0u;
        for (int q = 0;; ++q) {
            const WinPlan& pl = sh.plan[q & (kPlans - 1)];
            if (pl.mode != PLAN_RUN) break;
            const uint32_t TFp = pl.TFprev, TF = pl.TF;
            mbar_wait(&sh.full[stage], phase);
            const int stage_q = stage;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            if (TF > TFp) {
                const double* __restrict__ ss = stgS + stage_q * CTS + pl.shS;
                const uint32_t k0 = TFp + (uint32_t)(kRun * stid);
                const int i0 = (int)(k0 - pl.stg0);
                const int nv = k0 < TF ? (int)min(TF - k0, (uint32_t)kRun) : 0;
                const int ib = nv > 0 ? i0 : 1;
                double d[kRun], v1 = 0.0, v2 = 0.0;
#pragma unroll
                for (int i = 0; i < kRun; ++i) d[i] = ss[ib + i] - pivot;
#pragma unroll
                for (int i = 0; i < kRun; ++i) {
                    if (i >= nv) d[i] = 0.0;
                    v1 += d[i];
                    v2 = fma(d[i], d[i], v2);
                }
                if (!(v2 < 1e300)) {
                    if (nv > 0) atomicMax(&sh.excP1, k0 + (uint32_t)nv);
#pragma unroll
                    for (int i = 0; i < kRun; ++i) d[i] = 0.0;
                    v1 = v2 = 0.0;
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int dd = 1 << k;
                    const double u1 = shfl_up_d(v1, dd), u2 = shfl_up_d(v2, dd);
                    if (lane >= dd) { v1 += u1; v2 += u2; }
                }
                if (lane == 31) { sh.wtS[swarp][0] = v1; sh.wtS[swarp][1] = v2; }
                double g1 = shfl_up_d(v1, 1), g2 = shfl_up_d(v2, 1);
                if (lane == 0) g1 = g2 = 0.0;
                bar_sync_n(4, kSpread);
                {   // across the warps, in the midprice producers' order: carry, then warp 0, 1, ...
                    double a1 = cD1, a2 = cD2, m1 = 0.0, m2 = 0.0;
#pragma unroll
                    for (int w = 0; w < kProdWarps; ++w) {
                        if (w == swarp) { m1 = a1; m2 = a2; }
                        a1 = fma(1.0, a1, sh.wtS[w][0]);
                        a2 = fma(1.0, a2, sh.wtS[w][1]);
                    }
                    cD1 = a1;
                    cD2 = a2;
                    g1 += m1;
                    g2 += m2;
                }
                bar_sync_n(3, kWinThreads);
                uint32_t slot = pl.slotT + (uint32_t)(kRun * stid) + 1u;
                if (slot >= (uint32_t)R) slot -= (uint32_t)R;
                double* __restrict__ p1 = QD1 + slot;
                double* __restrict__ p2 = QD2 + slot;
#pragma unroll
                for (int i = 0; i < kRun; ++i) {
                    if (i < nv) {
                        g1 += d[i];
                        g2 = fma(d[i], d[i], g2);
                        p1[i] = g1;
                        p2[i] = g2;
                    }
                }
                if (slot + (uint32_t)nv > (uint32_t)R) {
#pragma unroll 1
                    for (uint32_t j = (uint32_t)R - slot; j < (uint32_t)nv; ++j) {
                        QD1[slot + j - (uint32_t)R] = p1[j];
                        QD2[slot + j - (uint32_t)R] = p2[j];
                    }
                }
            } else {
                bar_sync_n(3, kWinThreads);
            }
            __syncthreads();   // the rings hold chunk q
        }
    } else if (tid < kFirstCons + kCons) {
        // ================================================================================ consumers: plan, gather, finalise
        const int ci = tid - kFirstCons, cw = (tid - kFirstCons) >> 5;
        DirectParams dpm;
        dpm.mid_window = Wm;
        dpm.dir_window = Wd;
        dpm.spread_divisor = P.spread_divisor;
        dpm.ewm_alpha = P.ewm_alpha;
        dpm.ewm_decay = P.ewm_decay;
        uint32_t cur_e = 0, cur_j = 0;   // this thread's candidate row of the chunk just planned
        int plan_nr = 0;
        int pstage = 0;                  // staging buffer of the next chunk to plan
        // plan the chunk that follows (TFq, JFq) and starts at row c0n, from this thread's candidate cursors; request its ticks / trades
        auto make_plan = [&](const int q_new, const long long c0n, const uint32_t TFq, const uint32_t JFq, const uint32_t slotT,
                             const uint32_t slotJ, const uint32_t pe, const uint32_t pj) {
            const long long row = c0n + ci;
            const bool inr = row < rb;
            const uint32_t ce = inr ? min(max(pe, TFq), nTicks) : 0xffffffffu, cj = inr ? min(max(pj, JFq), nTrades) : 0xffffffffu;
            const bool fits = inr && ce - TFq <= (uint32_t)CT && cj - JFq <= (uint32_t)CJ;   // monotone in ci
            const unsigned bal = __ballot_sync(0xffffffffu, fits);
            if (lane == 0) sh.cnt[cw] = __popc(bal);
            bar_sync_n(2, kCons);
            const int nr = sh.cnt[0] + sh.cnt[1] + sh.cnt[2] + sh.cnt[3];
            plan_nr = nr;
            if (ci == (nr > 0 ? nr - 1 : 0)) {
                WinPlan& np = sh.plan[q_new & (kPlans - 1)];
                const int sbuf = pstage;
                const bool end = c0n >= rb;
                uint32_t TFn = TFq, JFn = JFq;
                if (!end) {
                    TFn = nr > 0 ? ce : min(ce, TFq + (uint32_t)CT);
                    JFn = nr > 0 ? cj : min(cj, JFq + (uint32_t)CJ);
                }
                np.c0 = c0n;
                np.TFprev = TFq;
                np.TF = TFn;
                np.JFprev = JFq;
                np.JF = JFn;
                np.slotT = slotT;
                np.slotJ = slotJ;
                np.nr = nr;
                np.mode = end ? PLAN_END : PLAN_RUN;
                const uint32_t s0 = TFq > baseK ? TFq - 1u : baseK;
                np.stg0 = s0;
                np.shM = np.shS = np.shT = 0;
                uint64_t* bar = &sh.full[sbuf];
                // what the copy engine stages for the chunk: mids and spreads from the tick before the frontier, trade directions from
                // the trade before the frontier; 16-byte rounding on both ends, the element shifts are published with the plan
                uintptr_t qm = 0, qs = 0, qt = 0;
                uint32_t bm = 0, bs = 0, bt = 0;
                if (!end && TFn > TFq) {
                    const uintptr_t am = reinterpret_cast<uintptr_t>(mid + s0), as = reinterpret_cast<uintptr_t>(spread + s0);
                    const uintptr_t bytes = (uintptr_t)(TFn - s0) * 8u;
                    qm = am & ~(uintptr_t)15;
                    qs = as & ~(uintptr_t)15;
                    bm = (uint32_t)(((am + bytes + 15) & ~(uintptr_t)15) - qm);
                    bs = (uint32_t)(((as + bytes + 15) & ~(uintptr_t)15) - qs);
                    np.shM = (int)((am - qm) >> 3);
                    np.shS = (int)((as - qs) >> 3);
                }
                if (!end && JFn > JFq) {
                    const uint32_t t0 = JFq > baseJ ? JFq - 1u : baseJ;
                    const uintptr_t at = reinterpret_cast<uintptr_t>(tdir + t0);
                    qt = at & ~(uintptr_t)15;
                    bt = (uint32_t)(((at + (JFn - t0) + 15) & ~(uintptr_t)15) - qt);
                    np.shT = (int)(at - qt);
                }
                if (bm + bt != 0u) {
                    fence_proxy_async_smem();   // the readers of this buffer (a chunk barrier ago) precede the engine's writes
                    mbar_expect_tx(bar, bm + bs + bt);
                    if (bm) {
                        bulk_g2s(stgM + sbuf * CTS, reinterpret_cast<const void*>(qm), bm, bar);
                        bulk_g2s(stgS + sbuf * CTS, reinterpret_cast<const void*>(qs), bs, bar);
                    }
                    if (bt) bulk_g2s(sh.stgT[sbuf], reinterpret_cast<const void*>(qt), bt, bar);
                } else {
                    mbar_arrive(bar);           // nothing to stage: the phase completes at once
                }
            }
            cur_e = ce;
            cur_j = cj;
            if (++pstage == kStages) pstage = 0;
            // pull the ticks of the chunk after that one towards the L2
            const uint32_t pf = TFq + (uint32_t)CT + (uint32_t)ci * 16u;
            if (ci * 16 < CT && pf < nTicks) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(mid + pf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(spread + pf));
            }
        };
        // gather the rows of a produced chunk out of the rings and finalise them
        auto gather = [&](const WinPlan& pg, const uint32_t e, const uint32_t je, RowGather& gat) {
            const uint32_t TF = pg.TF, JF = pg.JF;
            uint32_t slotTe = pg.slotT + (TF - pg.TFprev), slotJe = pg.slotJ + (JF - pg.JFprev);   // slots of the chunk's frontiers
            if (slotTe >= (uint32_t)R) slotTe -= (uint32_t)R;
            if (slotJe >= (uint32_t)RJ) slotJe -= (uint32_t)RJ;
            gat.row = -1;
            if (ci < pg.nr) {
                auto slot_of = [&](uint32_t p) {   // ring slot of tick position p <= TF
                    int sl = (int)slotTe - (int)(TF - p);
                    return sl < 0 ? sl + R : sl;
                };
                const uint32_t s = sub_floor0(e, Wm), d0 = sub_floor0(e, Wd);
                const int se = slot_of(e), ss0 = slot_of(s), sd = slot_of(d0);
                const int cnt = (int)(e - s) - 1;
                const double Ae = QA[se], Se = QS[se], D1e = QD1[se], D2e = QD2[se];
                const uint32_t Ce = QC[se], Cd = QC[sd];
                gat.Ge = QG[se];
                gat.D1 = D1e - QD1[ss0];
                gat.D2 = D2e - QD2[ss0];
                gat.SA = gat.SS = gat.Gs1 = gat.xf = 0.0;
                gat.negs = gat.movm = 0;
                if (cnt > 0) {
                    int s1 = ss0 + 1, s2;
                    if (s1 == R) s1 = 0;
                    s2 = s1 + 1;
                    if (s2 == R) s2 = 0;
                    const double A1 = QA[s1];
                    const uint32_t C1 = QC[s1];
                    gat.SA = Ae - A1;
                    gat.SS = Se - QS[s1];
                    gat.Gs1 = QG[s1];
                    gat.xf = QA[s2] - A1;
                    gat.negs = (int)((Ce - C1) & 0xffffu);
                    gat.movm = (int)(((Ce >> 16) - (C1 >> 16)) & 0xffffu);
                }
                gat.negd = (int)((Ce - Cd) & 0xffffu);
                gat.movd = (int)(((Ce >> 16) - (Cd >> 16)) & 0xffffu);
                // trades: [jb, je), reversal pairs (i-1, i) for i in (jb, je)
                const uint32_t jb = sub_floor0(je, Wd);
                auto tslot = [&](uint32_t p) {
                    int sl = (int)slotJe - (int)(JF - p);
                    return sl < 0 ? sl + RJ : sl;
                };
                const int tb = tslot(jb);
                const unsigned long long Te = QT[tslot(je)], Tb = QT[tb];
                gat.pos = (int)((uint32_t)(Te - Tb) & 0xffffu);
                gat.neg = (int)((uint32_t)((Te >> 16) - (Tb >> 16)) & 0xffffu);
                gat.isum = (int)(int16_t)(uint16_t)((Te >> 48) - (Tb >> 48));
                gat.rev = 0;
                if (je - jb >= 2u) {
                    const unsigned long long T1 = QT[tb + 1 == RJ ? 0 : tb + 1];
                    gat.rev = (int)((uint32_t)((Te >> 32) - (T1 >> 32)) & 0xffffu);
                }
                gat.e = e;
                gat.je = je;
                gat.row = pg.c0 + ci;
                const uint32_t exc = sh.excP1;
                gat.flagged = exc != 0u && e <= exc + Wmax + 1u;   // the window (or the entry just before it) may hold a flagged tick
            }
        };
        auto finish = [&](const RowGather& gat) {
            if ((MRP_WIN_WHATIF & 2) && gat.row >= 0) {
                write_tick_row(out, N, gat.row, gat.Ge + gat.Gs1 + gat.xf, gat.SA, gat.SS, gat.D1 + gat.D2, (double)(gat.negs + gat.movm + gat.negd + gat.movd), 0, 0, 0, 0, 0);
                out.raw[(int64_t)MRP_F_REVERSAL_RATE * N + gat.row] = (double)(gat.rev + gat.isum);
                out.raw[(int64_t)MRP_F_AGGRESSOR_BIAS * N + gat.row] = (double)gat.pos;
                out.raw[(int64_t)MRP_F_SHANNON_ENTROPY * N + gat.row] = (double)gat.neg;
            } else
            if (gat.row >= 0) {
                if (gat.flagged) direct_tick_row(mid, spread, gat.e, dpm, out.raw, out.counts, N, gat.row);
                else finish_tick_row(out, N, P, sLog, pivot, gat);
                const uint32_t jb = sub_floor0(gat.je, Wd);
                write_trade_row(out, N, gat.row, sLog, (int)(gat.je - jb), gat.pos, gat.neg, gat.rev, gat.isum);
            }
        };
        uint32_t pre_e = 0, pre_j = 0;
        if (ra + ci < rb) {
            pre_e = __ldg(kcur + ra + ci);
            pre_j = __ldg(jcur + ra + ci);
        }
        // this thread's row cursors of the chunks in flight: eq[i] belongs to chunk q - 1 + i in iteration q (planned kDepth
        // iterations before it is produced, gathered one iteration after)
        uint32_t eq[kDepth + 2], jq[kDepth + 2];
#pragma unroll
        for (int i = 0; i < kDepth + 2; ++i) eq[i] = jq[i] = 0u;
        // the plan that follows plan `pl` (frontiers, ring slots, first row), then this thread's candidate cursors for the one after
        auto plan_after = [&](const int q_new, const WinPlan& pl) {
            const uint32_t TF = pl.TF, JF = pl.JF;
            uint32_t slotTe = pl.slotT + (TF - pl.TFprev), slotJe = pl.slotJ + (JF - pl.JFprev);
            if (slotTe >= (uint32_t)R) slotTe -= (uint32_t)R;
            if (slotJe >= (uint32_t)RJ) slotJe -= (uint32_t)RJ;
            const long long c0n = pl.c0 + pl.nr;
            make_plan(q_new, c0n, TF, JF, slotTe, slotJe, pre_e, pre_j);
            const long long c2 = c0n + plan_nr;
            pre_e = pre_j = 0;
            if (c2 + ci < rb) {
                pre_e = __ldg(kcur + c2 + ci);
                pre_j = __ldg(jcur + c2 + ci);
            }
        };
        make_plan(0, ra, baseK, baseJ, 0u, 0u, pre_e, pre_j);
        eq[1] = cur_e;
        jq[1] = cur_j;
        {   // candidates of chunk 1
            const long long c1 = ra + plan_nr;
            pre_e = pre_j = 0;
            if (c1 + ci < rb) {
                pre_e = __ldg(kcur + c1 + ci);
                pre_j = __ldg(jcur + c1 + ci);
            }
        }
#pragma unroll
        for (int p2 = 1; p2 < kDepth; ++p2) {   // plans 1 .. kDepth - 1: each needs its predecessor (written by one thread) to be visible
            bar_sync_n(2, kCons);
            plan_after(p2, sh.plan[p2 - 1]);
            eq[p2 + 1] = cur_e;
            jq[p2 + 1] = cur_j;
        }
        __syncthreads();   // the first kDepth plans are visible to everybody
        int q = 0;
        TM_DECL;
#if MRP_WIN_ORDER == 2
        // The iteration opens with the finalisation of the rows gathered in the PREVIOUS iteration — arithmetic and global stores out of
        // registers — because the shared-memory pipe is still draining the producers' ring stores of the chunk before (every load
        // issued right after the chunk barrier waits 700-1000 cycles behind them); the plan and the gather follow when the pipe is free.
        RowGather held;
        held.row = -1;
        for (;; ++q) {
            const WinPlan& pl = sh.plan[q & (kPlans - 1)];
            if (pl.mode != PLAN_RUN) break;
            TM(0);
            finish(held);
            TM(2);
            plan_after(q + kDepth, sh.plan[(q + kDepth - 1) & (kPlans - 1)]);
            eq[kDepth + 1] = cur_e;
            jq[kDepth + 1] = cur_j;
            TM(5);
            held.row = -1;
            if (q > 0) gather(sh.plan[(q - 1) & (kPlans - 1)], eq[0], jq[0], held);
            TM(1);
            bar_arrive_n(3, kWinThreads);
#pragma unroll
            for (int i = 0; i < kDepth + 1; ++i) { eq[i] = eq[i + 1]; jq[i] = jq[i + 1]; }
            TM(3);
            __syncthreads();   // the rings hold chunk q; plan q + kDepth is visible
            TM(4);
#ifdef MRP_WIN_TIMERS
            ++tm_iters;
#endif
        }
        if (cw == 0) TM_FLUSH(1);
        finish(held);
        if (q > 0) {   // the last chunk
            RowGather gat;
            gather(sh.plan[(q - 1) & (kPlans - 1)], eq[0], jq[0], gat);
            finish(gat);
        }
#else
        for (;; ++q) {
            const WinPlan& pl = sh.plan[q & (kPlans - 1)];
            if (pl.mode != PLAN_RUN) break;
            TM(0);
            // chunk q - 1 is in the rings (the producers are busy with chunk q's arithmetic and will not store before barrier 3): lift
            // its window values into registers FIRST and release the rings — everything the other roles do after barrier 3 (ring
            // stores) then overlaps the plan and the finalisation
            RowGather gat;
            gat.row = -1;
#if MRP_WIN_ORDER == 1
            plan_after(q + kDepth, sh.plan[(q + kDepth - 1) & (kPlans - 1)]);
            eq[kDepth + 1] = cur_e;
            jq[kDepth + 1] = cur_j;
            TM(5);
#endif
            if (q > 0) gather(sh.plan[(q - 1) & (kPlans - 1)], eq[0], jq[0], gat);
            TM(1);
            bar_arrive_n(3, kWinThreads);
            finish(gat);
            TM(2);
#if MRP_WIN_ORDER == 0
            // the plan of chunk q + kDepth LAST: its ticks still have more than kDepth - 1 iterations to arrive, and the cursor loads
            // it ends with (candidates of the plan after it, out of HBM) have the chunk barrier and the next gather to land — no
            // arithmetic of this role waits behind them on a scoreboard
            plan_after(q + kDepth, sh.plan[(q + kDepth - 1) & (kPlans - 1)]);
            eq[kDepth + 1] = cur_e;
            jq[kDepth + 1] = cur_j;
#endif
#pragma unroll
            for (int i = 0; i < kDepth + 1; ++i) { eq[i] = eq[i + 1]; jq[i] = jq[i + 1]; }
            TM(3);
            __syncthreads();   // the rings hold chunk q; plan q + kDepth is visible
            TM(4);
#ifdef MRP_WIN_TIMERS
            ++tm_iters;
#endif
        }
        if (cw == 0) TM_FLUSH(1);
        if (q > 0) {   // the last chunk
            RowGather gat;
            gather(sh.plan[(q - 1) & (kPlans - 1)], eq[0], jq[0], gat);
            finish(gat);
        }
#endif
    } else {
        // ================================================================================ the trade warp: trades -> QT
        unsigned long long cJ = 0ull;    // the counters at the trade frontier
        __syncthreads();   // the first plans are visible
        int stage = 0;
        uint32_t phase = 0u;
        TM_DECL;
        for (int q = 0;; ++q) {
            const WinPlan& pl = sh.plan[q & (kPlans - 1)];
            if (pl.mode != PLAN_RUN) break;
            TM(0);
            const uint32_t JFp = pl.JFprev, JF = pl.JF;
            uint32_t slot0 = pl.slotJ;
            if (JF > JFp) mbar_wait(&sh.full[stage], phase);   // the chunk's trade directions have landed
            const uint32_t t0 = JFp > baseJ ? JFp - 1u : baseJ;
            const signed char* __restrict__ sb = reinterpret_cast<const signed char*>(sh.stgT[stage]) + pl.shT;   // sb[j - t0] = direction of trade j
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            bool released = false;   // barrier 3 (the previous chunk has been gathered) is due before the first store into the ring
            for (uint32_t jp = JFp; jp < JF; jp += 32u * kTradeRun) {
                const uint32_t j0 = jp + (uint32_t)(lane * kTradeRun);
                const int nt = j0 < JF ? (int)min(JF - j0, (uint32_t)kTradeRun) : 0;   // this lane's trades of the pass
                const signed char* __restrict__ mine = sb + (j0 - t0);                // mine[i] = direction of trade j0 + i
                const int dfirst = (j0 > baseJ && nt > 0) ? (int)mine[-1] : 0;        // the trade at the base (trade 0) closes no pair
                // fields: pos | neg << 16 | reversal << 32 | dir << 48 (feature_processor.cpp:235, :242-245, :255-259); within a pass
                // no field but the top one can carry.  Two short rolled loops (lane total, then the prefixes) instead of three unrolled
                // ones: this role's instructions share the SM's instruction cache with the other two
                auto fields = [](const int d, const int dprev) {
                    return (unsigned long long)(d > 0 ? 1 : 0) + ((unsigned long long)(d < 0 ? 1 : 0) << 16) +
                           ((unsigned long long)((d != 0 && d == -dprev) ? 1 : 0) << 32) + ((unsigned long long)(long long)d << 48);
                };
                unsigned long long t = 0ull;
                {
                    int dprev = dfirst;
#pragma unroll 1
                    for (int i = 0; i < nt; ++i) {
                        const int d = (int)mine[i];
                        t += fields(d, dprev);
                        dprev = d;
                    }
                }
                unsigned long long v = t;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const unsigned long long uv = shfl_up_u64(v, 1 << k);
                    if (lane >= (1 << k)) v += uv;
                }
                unsigned long long ex = shfl_up_u64(v, 1);
                if (lane == 0) ex = 0ull;
                const unsigned long long base = add4x16(cJ, ex);
                if (!released) {
                    TM(1);
                    bar_sync_n(3, kWinThreads);
                    released = true;
                    TM(2);
                }
                uint32_t slot = slot0 + (uint32_t)(lane * kTradeRun) + 1u;
                if (slot >= (uint32_t)RJ) slot -= (uint32_t)RJ;
                {
                    int dprev = dfirst;
                    unsigned long long c = 0ull;
#pragma unroll 1
                    for (int i = 0; i < nt; ++i) {
                        const int d = (int)mine[i];
                        c += fields(d, dprev);
                        dprev = d;
                        QT[slot] = add4x16(base, c);
                        if (++slot == (uint32_t)RJ) slot = 0u;
                    }
                }
                cJ = add4x16(cJ, __shfl_sync(0xffffffffu, v, 31));
                slot0 += 32u * kTradeRun;
                if (slot0 >= (uint32_t)RJ) slot0 -= (uint32_t)RJ;
            }
            if (!released) bar_sync_n(3, kWinThreads);
            TM(3);
            __syncthreads();   // the rings hold chunk q
            TM(4);
#ifdef MRP_WIN_TIMERS
            ++tm_iters;
#endif
        }
        TM_FLUSH(2);
    }
}

// Ring geometry for given windows: the tick ring holds one window + one chunk, likewise the trade ring; everything must fit the
// 227 KB a CTA can opt into.  Returns false when the windows are too long for the on-chip rings.
static size_t win_smem(int ring, int chunk_ticks, int trade_ring, int dir_window) {
    return (size_t)(ring + kRingPad) * 44 + (size_t)2 * kStages * (chunk_ticks + 4) * 8 + (size_t)trade_ring * 8 +
           (MRP_WIN_SLOG_SMEM ? (size_t)(dir_window + 2) * 8 : 0);
}
bool win_geometry(int mid_window, int dir_window, int* ring, int* chunk_ticks, int* trade_ring, int* chunk_trades, size_t* smem) {
    const int wmax = mid_window > dir_window + 1 ? mid_window : dir_window + 1;
    const size_t budget = 232448 - sizeof(WinStatic) - 1024;
    const int cj = kChunkTrades;
    const int rj = ((dir_window + 2 + cj + 8 + 1) / 2) * 2;
    for (int ct = kProd * kRun; ct >= 4 * kRun; ct -= 2 * kRun) {   // even (16-byte staging rows), whole thread runs
        const int r = ((wmax + 2 + ct + 8 + 1) / 2) * 2;
        if (win_smem(r, ct, rj, dir_window) <= budget) {
            *ring = r;
            *chunk_ticks = ct;
            *trade_ring = rj;
            *chunk_trades = cj;
            *smem = win_smem(r, ct, rj, dir_window);
            return true;
        }
    }
    return false;
}

cudaError_t launch_windows(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg, const DevParams& P,
                           cudaStream_t st) {
    const size_t smem = win_smem(P.win_ring, P.win_chunk_ticks, P.win_trade_ring, P.dir_window);
    static int ready_for[64];   // the opt-in is per device (idempotent; a race only repeats the call)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (ready_for[dev & 63] < (int)smem) {
        e = cudaFuncSetAttribute(k_windows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ready_for[dev & 63] = (int)smem;
    }
    k_windows<<<nseg, kWinThreads, smem, st>>>(in, out, segs, P);
    return cudaGetLastError();
}

}  // namespace mrp
