// kernels_window.cu — the cursor-addressed rolling-window features (tick axis + trade axis), one kernel.
//
//   tick axis : ewm_volatility, realized_variance, directional_volatility, spread_volatility,
//               tick_direction_entropy      (ref: feature_processor.cpp:48-107, :215-229; feature_engine.cpp:130-150)
//   trade axis: reversal_rate, aggressor_bias, shannon_entropy
//                                           (ref: feature_processor.cpp:231-247, :254-269; feature_engine.cpp:87-93)
//
// The reference re-walks a 1800-tick (1500-trade) deque for every snapshot.  Here every window quantity is
// a SLIDING sum: going from snapshot t-1 to t the window gains the ticks [k_{t-1}, k_t) and loses the ticks
// [k_{t-1}-W, k_t-W), so
//        sum_t = sum_{t-1} + A_t - B_t
// and a chunk of 128 consecutive snapshots needs one cooperative pass over its NEW ticks (one reciprocal and
// a 6-term log1p series per tick instead of two logs), two short per-snapshot loops over a shared-memory ring
// (the added and the evicted ticks) and one 128-wide scan per quantity.  Each tick is read from HBM once.
//
// Per-tick ring entries (8 B each):
//   y_k = +r_k^2 when the return r_k = log(M_k / M_{k-1}) is > 0, -r_k^2 otherwise (-0.0 for a flat tick), so that
//         sum|y| = sum r^2, sum y = (up part) - (down part), sign/zero tests give the tick direction;
//   d_k = spread_k - pivot (pivot: a spread value near the window, removes the cancellation against the level).
// The restarted exponentially weighted variance (feature_processor.cpp:64-79) over a window with first return f is
//        w = sum_{i>f} alpha (1-alpha)^(e-1-i) x_i + (1-alpha)^(cnt-1) x_f
// i.e. an exponentially weighted window sum E plus a correction for the first term; E slides like the other sums
// once it is rescaled by (1-alpha)^-(position in chunk)  (F = E * decay^-(k - k_chunk)).
//
// Ticks the fast formulation cannot represent — a non-positive / non-finite midprice, a 50 ms move above 2^-10
// relative (series range), a non-finite spread — are flagged; every snapshot whose window contains a flagged tick
// is evaluated by direct_tick_row(), a plain transcription of the reference loops with true logs, and the sliding
// state is rebuilt from a clean halo once the flagged tick has left all windows.  A snapshot that brings more new
// ticks than the ring can hold is handled the same way.  There is no CPU path.
#include "mrp_device.cuh"

namespace mrp {

#ifndef MRP_WIN_MINB
#define MRP_WIN_MINB 3
#endif
constexpr int kWinThreads = 256;
constexpr int kWinRows = 128;            // snapshots per chunk: threads 0..127 add side, 128..255 evict side + trades
// the log1p series covers relative moves below 2^-10 per tick (high word 0x3f500000)

struct __align__(16) WinShared {
    double A[5][kWinRows];      // per-row deltas, then (after the scan) per-row window values
    int AI[2][kWinRows];        // up | down << 16 tick counts: [0] returns of the mid window, [1] direction window
    int TD[2][kWinRows];        // trades: [0] pos | neg << 16, [1] reversals | dir-sum << 16
    uint32_t LM[kWinRows];      // 1 + position of the last tick that moved the midprice (running max)
    uint32_t E[kWinRows + 1];   // clamped tick cursors: E[0] = previous cursor, E[i+1] = row c0+i
    uint32_t J[kWinRows + 1];   // clamped trade cursors
    uint32_t excP1;             // 1 + position of the last flagged tick (0 = none)
};

// 1/a to full precision for a normal, positive a (same refinement as the compiler's division fast path,
// without its special-case branch: callers flag non-positive / non-finite inputs themselves).
__device__ __forceinline__ double rcp_pos(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}

// log(1+u) for |u| <= 2^-10: truncation below 2^-60 relative.
__device__ __forceinline__ double log1p_small(double u) {
    double q = -1.0 / 6.0;
    q = fma(q, u, 1.0 / 5.0);
    q = fma(q, u, -1.0 / 4.0);
    q = fma(q, u, 1.0 / 3.0);
    q = fma(q, u, -0.5);
    q = fma(q, u, 1.0);
    return q * u;
}

__device__ __forceinline__ uint32_t sub_floor0(uint32_t a, uint32_t w) { return a > w ? a - w : 0u; }

// In-place inclusive scan of 128 doubles by one warp (lane owns 4 consecutive rows) + running carry.
__device__ __forceinline__ void scan128(double* __restrict__ a, double& carry, int lane) {
    double2 p = *reinterpret_cast<double2*>(a + lane * 4), q = *reinterpret_cast<double2*>(a + lane * 4 + 2);
    p.y += p.x;
    q.x += p.y;
    q.y += q.x;
    double s = q.y;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = shfl_up_d(s, d);
        if (lane >= d) s += t;
    }
    const double off = carry + (s - q.y);
    p.x += off; p.y += off; q.x += off; q.y += off;
    *reinterpret_cast<double2*>(a + lane * 4) = p;
    *reinterpret_cast<double2*>(a + lane * 4 + 2) = q;
    carry = __shfl_sync(0xffffffffu, q.y, 31);
}
__device__ __forceinline__ void scan128(int* __restrict__ a, int& carry, int lane) {
    int4 v = *reinterpret_cast<int4*>(a + lane * 4);
    v.y += v.x;
    v.z += v.y;
    v.w += v.z;
    int s = v.w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
    }
    const int off = carry + (s - v.w);
    v.x += off; v.y += off; v.z += off; v.w += off;
    *reinterpret_cast<int4*>(a + lane * 4) = v;
    carry = __shfl_sync(0xffffffffu, v.w, 31);
}

// ---- direct evaluation of one snapshot's tick-axis features: the reference loops, in the reference order
//      (feature_processor.cpp:48-107, :215-229).  Used for windows that contain a flagged tick.
struct DirectParams {
    uint32_t mid_window, dir_window;
    int spread_divisor;
    double ewm_alpha, ewm_decay;
};

__device__ __forceinline__ void write_tick_row(const DevOutputs& out, int64_t N, int64_t r, double ewm, double rv, double dv, double sv,
                                               double tent, int n_valid, int n_up, int t_up, int t_dn, int t_tot) {
    out.raw[(int64_t)MRP_F_EWM_VOLATILITY * N + r] = ewm;
    out.raw[(int64_t)MRP_F_REALIZED_VARIANCE * N + r] = rv;
    out.raw[(int64_t)MRP_F_DIRECTIONAL_VOLATILITY * N + r] = dv;
    out.raw[(int64_t)MRP_F_SPREAD_VOLATILITY * N + r] = sv;
    out.raw[(int64_t)MRP_F_TICK_DIRECTION_ENTROPY * N + r] = tent;
    if (out.counts) {
        out.counts[(int64_t)MRP_C_RET_VALID * N + r] = n_valid;
        out.counts[(int64_t)MRP_C_RET_UP * N + r] = n_up;
        out.counts[(int64_t)MRP_C_TICK_UP * N + r] = t_up;
        out.counts[(int64_t)MRP_C_TICK_DOWN * N + r] = t_dn;
        out.counts[(int64_t)MRP_C_TICK_TOTAL * N + r] = t_tot;
    }
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

// trade-axis features of one snapshot from its window counts (feature_processor.cpp:231-247, :254-269)
// a / n for a count n inside the host's 1/n table: reciprocal, product, ONE correction step (the quotient's remainder is
// exact in an fma) — the correctly rounded quotient for the small integers that meet here, for 4 instructions and a load
__device__ __forceinline__ double div_by_count(double a, int n, const double* __restrict__ inv_cnt) {
    const double b = (double)n, r = __ldg(inv_cnt + n);
    const double q = a * r;
    return fma(fma(-q, b, a), r, q);
}

// inv_cnt == nullptr: plain divisions (the direct path)
__device__ __forceinline__ void write_trade_row(const DevOutputs& out, int64_t N, int64_t r, const double* __restrict__ clog2c,
                                                const double* __restrict__ inv_cnt, int nd, int pos, int neg, int rev, int isum) {
    const int tot = pos + neg;                                  // :261-269
    double rr, ab, ent;
    if (inv_cnt) {
        rr = nd > 1 ? div_by_count((double)rev, nd, inv_cnt) : 0.0;                                   // :239 — divides by size, not size-1
        ab = nd > 0 ? div_by_count((double)isum, nd, inv_cnt) : __longlong_as_double(0x7ff8000000000000ll);  // :246 — 0/0 = NaN with no trades
        // binary entropy = (T log2 T - pos log2 pos - neg log2 neg) / T with c*log2(c) tabulated by the host
        ent = tot > 0 ? div_by_count((__ldg(clog2c + tot) - __ldg(clog2c + pos)) - __ldg(clog2c + neg), tot, inv_cnt) : 0.0;
    } else {
        rr = nd > 1 ? (double)rev / (double)nd : 0.0;
        ab = (double)isum / (double)nd;
        ent = tot > 0 ? ((__ldg(clog2c + tot) - __ldg(clog2c + pos)) - __ldg(clog2c + neg)) / (double)tot : 0.0;
    }
    out.raw[(int64_t)MRP_F_REVERSAL_RATE * N + r] = rr;
    out.raw[(int64_t)MRP_F_AGGRESSOR_BIAS * N + r] = ab;
    out.raw[(int64_t)MRP_F_SHANNON_ENTROPY * N + r] = ent;
    if (out.counts) {
        out.counts[(int64_t)MRP_C_TRADE_POS * N + r] = pos;
        out.counts[(int64_t)MRP_C_TRADE_NEG * N + r] = neg;
        out.counts[(int64_t)MRP_C_TRADE_REVERSALS * N + r] = rev;
        out.counts[(int64_t)MRP_C_TRADE_TOTAL * N + r] = nd;
        out.counts[(int64_t)MRP_C_TRADE_DIR_SUM * N + r] = isum;
    }
}

__device__ __noinline__ void direct_trade_row(const int8_t* __restrict__ tdir, uint32_t je, uint32_t dir_window,
                                              const double* __restrict__ clog2c, double* __restrict__ raw, int32_t* __restrict__ counts,
                                              int64_t N, int64_t r) {
    const uint32_t jb = sub_floor0(je, dir_window);
    int pos = 0, neg = 0, rev = 0, isum = 0;
    for (uint32_t j = jb; j < je; ++j) {
        const int d = tdir[j];
        pos += d > 0;
        neg += d < 0;
        isum += d;
        if (j > jb && d != 0 && d == -(int)tdir[j - 1]) ++rev;
    }
    DevOutputs o{};
    o.raw = raw;
    o.counts = counts;
    write_trade_row(o, N, r, clog2c, nullptr, (int)(je - jb), pos, neg, rev, isum);
}

// explicit 32-bit shared-window accessors for the ring (keeps address formation out of the inner loops)
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
template <int OFF>
__device__ __forceinline__ double lds_f64_off(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint32_t lds_hi32(uint32_t a) {  // high word of the double at a
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(v) : "r"(a));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64_off(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}

// y = +-r^2 of the return log(M / Mprev) = sgn * log1p(sgn * v) written with ONE reciprocal per pair of ticks:
// v = (M - Mprev) * rcp with rcp = 1/Mprev (sgn = +1) or 1/M (sgn = -1, log(M/Mprev) = -log1p((Mprev - M)/M)).
// t = Mprev - M; the sign bit of y is set unless the midprice rose (t < 0).
__device__ __forceinline__ double signed_sq_return(double t, double v_signed) {
    const double r = log1p_small(v_signed);
    const double x = r * r;
    const int hi = __double2hiint(x) | (~__double2hiint(t) & 0x80000000);
    return __hiloint2double(hi, __double2loint(x));
}
__device__ __forceinline__ bool small_mag(double v, uint32_t hi_limit) {  // |v| < the double whose high word is hi_limit; NaN -> false
    return ((uint32_t)__double2hiint(v) & 0x7fffffffu) < hi_limit;
}

// visit the ring slots of the positions [k0, k1) in order (at most one wrap): body(shared address, position)
template <int RING, typename F>
__device__ __forceinline__ void for_ring(uint32_t ring_base, uint32_t k0, uint32_t k1, F&& body) {
    while (k0 < k1) {
        const uint32_t slot = k0 & (RING - 1);
        const uint32_t n = min(k1 - k0, (uint32_t)RING - slot);
        uint32_t a = ring_base + slot * 8u;
        const uint32_t aend = a + n * 8u;
#pragma unroll 1
        for (; a < aend; a += 8u) body(a);
        k0 += n;
    }
}

__device__ __forceinline__ void scan128_max(uint32_t* __restrict__ a, uint32_t& carry, int lane) {
    uint4 v = *reinterpret_cast<uint4*>(a + lane * 4);
    v.y = max(v.y, v.x);
    v.z = max(v.z, v.y);
    v.w = max(v.w, v.z);
    uint32_t s = v.w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s = max(s, t);
    }
    uint32_t off = __shfl_up_sync(0xffffffffu, s, 1);
    off = lane == 0 ? carry : max(carry, off);
    v.x = max(v.x, off); v.y = max(v.y, off); v.z = max(v.z, off); v.w = max(v.w, off);
    *reinterpret_cast<uint4*>(a + lane * 4) = v;
    carry = __shfl_sync(0xffffffffu, v.w, 31);
}

// One snapshot's five tick-axis features from its window values.  The sliding sums cover the ticks [s, e) of the mid window
// (SA = sum |y|, SS = sum y, D1 / D2 = sums of d and d^2, Eall = exponentially weighted sum, negs = sign-set count) and [d0, e) of
// the direction window (dirw = sign-set | moved << 16); ys / yf are the ring entries of ticks s and s + 1; lm = 1 + position
// of the last tick that moved the midprice.
__device__ __forceinline__ void finish_tick_row(const DevOutputs& out, const int64_t N, const int64_t row, const DevParams& P,
                                                const uint32_t e, const double pivot, const double SAall, const double SSall,
                                                const double D1, const double D2, const double Eall, const int negs, const int dirw,
                                                const uint32_t lm, const double ys, const double yf) {
    double ewm = 0.0, rv = 0.0, dv = 0.0, sv = 0.0, tent = 0.0;
    int cnt = 0, nUp = 0, t_up = 0, t_dn = 0, t_tot = 0;
    if (e > 0u) {
        const uint32_t s = sub_floor0(e, (uint32_t)P.mid_window);
        cnt = (int)(e - s) - 1;
        if (cnt > 0) {
            // the window's returns are (s, e): take tick s out of the sums again
            const double SA = SAall - fabs(ys), SS = SSall - ys;
            nUp = cnt - (negs - (int)((uint32_t)__double2hiint(ys) >> 31));
            // a window of flat ticks only is exactly zero in the reference; do not let rounding residue of the sliding sums
            // through the square roots
            const bool moved = lm > s + 1u;
            const double tot = (SA > 0.0 && moved) ? SA : 0.0;
            rv = tot * __ldg(P.inv_cnt + cnt);  // feature_processor.cpp:61 (tabulated 1/count)
            double w = 0.0;
            if (tot != 0.0) {
                // :64-79, see the file header: E over (s, e) plus the restart correction of the first return
                w = Eall + __ldg(P.ewm_pow + cnt) * (fabs(yf) - P.ewm_alpha * fabs(ys));
                if (!(w > 0.0)) w = 0.0;
            }
            ewm = sqrt(w);
            double sumU = 0.5 * (SA + SS), sumD = 0.5 * (SA - SS);
            if (!(sumU > 0.0) || !moved) sumU = 0.0;
            if (!(sumD > 0.0) || !moved) sumD = 0.0;
            const double au = sumU * __ldg(P.inv_cnt + nUp);          // :95-98 (inv_cnt[0] = 0: empty side)
            const double ad = sumD * __ldg(P.inv_cnt + (cnt - nUp));
            const double diff = au - ad;
            dv = sqrt(fabs(diff)) * ((diff >= 0) ? 1.0 : -1.0);
        }
        {   // :101-107 with s_k = pivot + d_k: mean = (sum s)/1500, var = sum (s_k - mean)^2 / 1500
            const double n = (double)(e - s);
            const double mean = (D1 + n * pivot) * P.inv_spread_divisor;
            const double delta = mean - pivot;
            double ss = D2 - 2.0 * delta * D1 + n * delta * delta;
            if (!(ss > 0.0)) ss = 0.0;
            sv = sqrt(ss * P.inv_spread_divisor);
        }
        const uint32_t d0 = sub_floor0(e, (uint32_t)P.dir_window);
        t_tot = (int)(e - d0);
        const int negd = dirw & 0xffff, movd = (dirw >> 16) & 0xffff;
        t_up = t_tot - negd;            // sign clear <=> the midprice rose
        t_dn = negd + movd - t_tot;     // sign set and moved
        const int zero = t_tot - movd;
        // :222-229: -sum p log2 p = (T log2 T - sum c log2 c) / T, c*log2(c) tabulated by the host
        tent = ((__ldg(P.clog2c + t_tot) - __ldg(P.clog2c + t_up)) - __ldg(P.clog2c + t_dn) - __ldg(P.clog2c + zero)) *
               __ldg(P.inv_cnt + t_tot);
        if (cnt < 0) cnt = 0;
    }
    write_tick_row(out, N, row, ewm, rv, dv, sv, tent, cnt, nUp, t_up, t_dn, t_tot);
}

// ================================================================================================
// Pipelined chunk loop for clean data: warps 4-7 (producers) fetch the cursors, run the tick phase and the trade deltas
// of chunk q+1 while warps 0-3 (consumers) run the row loops, the scans and the finalisation of chunk q.  The global-load
// latency of the tick phase is thereby hidden behind shared-memory / ALU work instead of stalling the whole CTA at a
// barrier.  Two chunks of new ticks must fit the ring next to one window, so a pipelined chunk holds at most 112 rows.
// Anything irregular — a flagged tick, a snapshot with more new ticks than fit — makes the producers publish BAIL; the
// kernel then continues from that row with the sequential chunk loop below, which owns all the special cases.
// Named barriers: 1 producers, 2 consumers, 3/4 "descriptor b is full", 5/6 "descriptor b is free again".
// ================================================================================================
constexpr int kPipeRows = 112;
enum { PIPE_RUN = 0, PIPE_END = 1, PIPE_BAIL = 2 };

struct __align__(16) PipeDesc {
    int TD[2][kWinRows];        // trade deltas of the chunk's rows, scanned in place by the consumers
    uint32_t E[kWinRows + 1];   // clamped tick cursors, E[0] = the previous chunk's last
    uint32_t J[kWinRows];       // clamped trade cursor of each row
    long long c0;               // first row of the chunk
    uint32_t Elast;
    int nr, mode;
};
struct __align__(16) PipeShared {
    double A[5][kWinRows];
    int AI[2][kWinRows];
    uint32_t LM[kWinRows];
    PipeDesc d[2];
    uint32_t exc;
    int cnt[4];
};

__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// returns (through ps.d[..]) the first row it did not process; called by all 256 threads with a freshly restarted state
template <int RING>
__device__ __forceinline__ int64_t pipelined_chunks(PipeShared& ps, const uint32_t ring_base, const DevInputs& in, const DevOutputs& out,
                                                    const DevSegment& sg, const DevParams& P, int64_t c0, const int64_t out_from,
                                                    uint32_t Eprev, uint32_t Jprev, const uint32_t baseK, const uint32_t baseJ,
                                                    const double pivot) {
    constexpr uint32_t mask = RING - 1;
    constexpr int kD = RING * 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rb = sg.row_end, N = in.n_rows;
    const double* __restrict__ mid = in.mid + sg.tick0;
    const double* __restrict__ spread = in.spread + sg.tick0;
    const int8_t* __restrict__ tdir = in.tdir + sg.trade0;
    const uint32_t Wm = (uint32_t)P.mid_window, Wd = (uint32_t)P.dir_window;
    const uint32_t Wmax = max(Wm, Wd + 1);
    const uint32_t span_pipe = ((uint32_t)RING - Wmax - 8u) / 2u;
    const double decay = P.ewm_decay;
    if (tid == 0) ps.exc = 0u;
    __syncthreads();

    if (warp >= 4) {
        // ------------------------------------------------------------------------------------ producers
        const int pt = tid - kWinRows, pw = warp - 4;
        int q = 0, published = 0;
        // the cursors of a round are requested one round ahead
        uint32_t nce = 0xffffffffu, ncj = 0xffffffffu;
        if (c0 + pt < rb && pt < kPipeRows) {
            nce = __ldg(in.kcur + c0 + pt);
            ncj = __ldg(in.jcur + c0 + pt);
        }
        for (;; ++q) {
            const int b = q & 1;
            PipeDesc& D = ps.d[b];
            if (q >= 2) bar_sync_n(5 + b, kWinThreads);   // the consumers have released this descriptor (chunk q-2)
            int mode = PIPE_RUN, nr = 0;
            uint32_t Elast = Eprev;
            if (c0 >= rb) mode = PIPE_END;
            else {
                const int64_t row = c0 + pt;
                uint32_t ce = 0xffffffffu, cj = 0xffffffffu;
                if (row < rb && pt < kPipeRows) {
                    ce = max(nce, baseK);
                    cj = max(ncj, baseJ);
                }
                const bool fits = row < rb && pt < kPipeRows && ce - Eprev <= span_pipe;   // monotone in pt
                const unsigned bal = __ballot_sync(0xffffffffu, fits);
                if (lane == 0) ps.cnt[pw] = __popc(bal);
                D.E[pt + 1] = ce;
                D.J[pt] = cj;
                if (pt == 0) D.E[0] = Eprev;
                bar_sync_n(1, kWinRows);
                nr = ps.cnt[0] + ps.cnt[1] + ps.cnt[2] + ps.cnt[3];
                if (nr == 0) mode = PIPE_BAIL;
                else {
                    Elast = D.E[nr];
                    {   // next round's cursors: in flight during this round's work
                        const int64_t nrow = c0 + nr + pt;
                        nce = ncj = 0xffffffffu;
                        if (nrow < rb && pt < kPipeRows) {
                            nce = __ldg(in.kcur + nrow);
                            ncj = __ldg(in.jcur + nrow);
                        }
                    }
                    // trade window boundaries of this thread's row: requested before the tick phase, used after it
                    const uint32_t j0 = pt > 0 ? D.J[pt - 1] : Jprev, j1 = cj;
                    uint32_t q0 = 0, q1 = 0;
                    int dprev = 0, dq = 0;
                    if (pt < nr) {
                        q0 = max(baseJ, sub_floor0(j0, Wd));
                        q1 = max(baseJ, sub_floor0(j1, Wd));
                        if (j0 > baseJ) dprev = (int)tdir[j0 - 1];   // the trade at the base (trade 0) closes no pair
                        if (q1 > q0) dq = (int)tdir[q0];
                    }
                    // tick phase: new ticks -> ring, two per thread (see the sequential loop for the derivation)
                    uint32_t k = (Eprev & ~1u) + 2u * pt;
                    double Ma = 1.0, Mp = 1.0, Mb = 1.0, Sa = pivot, Sb = pivot;
                    auto fetch = [&](uint32_t kk, double& fMa, double& fMp, double& fMb, double& fSa, double& fSb) {
                        if (kk < Elast) {
                            // (defaults must not copy fMa: that would wait for its load before the others are even issued)
                            fMa = __ldg(mid + kk);
                            fSa = fSb = pivot;
                            if (kk != baseK) fMp = __ldg(mid + kk - 1);
                            if (kk >= Eprev) fSa = __ldg(spread + kk);
                            if (kk + 1u < Elast) {
                                fMb = __ldg(mid + kk + 1);
                                fSb = __ldg(spread + kk + 1);
                            }
                        }
                    };
                    // two pairs ahead: the loads of pairs k+256 and k+512 are in flight while pair k is worked on.  The three
                    // register sets rotate by unrolling, not by copying (a copy would wait for the load it copies).
                    auto work = [&](const uint32_t kk, double wMa, double wMp, double wMb, const double wSa, const double wSb) {
                        const bool va = kk >= Eprev, vb = kk + 1u < Elast;
                        if (kk == baseK) wMp = wMa;
                        if (!vb) wMb = wMa;
                        const double ra = rcp_pos(wMa);
                        const double ta = wMp - wMa, tb = wMa - wMb;
                        const double va_ = ta * ra, vb_ = -tb * ra;
                        const double ya = signed_sq_return(ta, va_), yb = signed_sq_return(tb, vb_);
                        const double da = wSa - pivot, db = wSb - pivot;
                        const bool ok = (uint32_t)(__double2hiint(wMa) - 0x00100000) < 0x7fe00000u && small_mag(va_, 0x3f500000u) &&
                                        small_mag(vb_, 0x3f500000u) && small_mag(da, 0x5f138d35u) && small_mag(db, 0x5f138d35u);
                        const uint32_t a = ring_base + (kk & mask) * 8u;
                        if (va) { sts_f64_off<0>(a, ya); sts_f64_off<kD>(a, da); }
                        if (vb) { sts_f64_off<8>(a, yb); sts_f64_off<kD + 8>(a, db); }
                        if (!ok) ps.exc = 1u;
                    };
                    constexpr uint32_t S = 2u * kWinRows;
                    double Ma1 = 1.0, Mp1 = 1.0, Mb1 = 1.0, Sa1 = pivot, Sb1 = pivot;
                    double Ma2 = 1.0, Mp2 = 1.0, Mb2 = 1.0, Sa2 = pivot, Sb2 = pivot;
                    fetch(k, Ma, Mp, Mb, Sa, Sb);
                    fetch(k + S, Ma1, Mp1, Mb1, Sa1, Sb1);
                    while (k < Elast) {
                        fetch(k + 2u * S, Ma2, Mp2, Mb2, Sa2, Sb2);
                        work(k, Ma, Mp, Mb, Sa, Sb);
                        if (k + S >= Elast) break;
                        fetch(k + 3u * S, Ma, Mp, Mb, Sa, Sb);
                        work(k + S, Ma1, Mp1, Mb1, Sa1, Sb1);
                        if (k + 2u * S >= Elast) break;
                        fetch(k + 4u * S, Ma1, Mp1, Mb1, Sa1, Sb1);
                        work(k + 2u * S, Ma2, Mp2, Mb2, Sa2, Sb2);
                        k += 3u * S;
                    }
                    // trade deltas of this chunk's rows, read straight from global memory
                    int w0 = 0, w1 = 0;
                    if (pt < nr) {
                        for (uint32_t j = j0; j < j1; ++j) {
                            const int d = tdir[j];
                            w0 += (d > 0 ? 1 : 0) + (d < 0 ? 65536 : 0);
                            w1 += ((d != 0 && d == -dprev) ? 1 : 0) + (d << 16);   // feature_processor.cpp:235
                            dprev = d;
                        }
                        int d = dq;
                        for (uint32_t j = q0; j < q1; ++j) {
                            const int dn = tdir[j + 1];
                            w0 -= (d > 0 ? 1 : 0) + (d < 0 ? 65536 : 0);
                            w1 -= ((dn != 0 && dn == -d) ? 1 : 0) + (d << 16);   // the pair (j, j+1) leaves with trade j
                            d = dn;
                        }
                    }
                    D.TD[0][pt] = w0;
                    D.TD[1][pt] = w1;
                    bar_sync_n(1, kWinRows);
                    if (ps.exc != 0u) mode = PIPE_BAIL;   // a flagged tick: the sequential loop takes over from this chunk
                }
            }
            if (pt == 0) {
                D.c0 = c0;
                D.Elast = Elast;
                D.nr = nr;
                D.mode = mode;
            }
            __threadfence_block();
            bar_arrive_n(3 + b, kWinThreads);
            if (mode != PIPE_RUN) break;
            ++published;
            const uint32_t jn = D.J[nr - 1];
            bar_sync_n(1, kWinRows);      // everybody has read what it needs from D before the next round may overwrite its twin
            Eprev = Elast;
            Jprev = jn;
            c0 += nr;
        }
        // drain: the "free again" arrival of the last processed chunk has no later round to consume it
        if (published >= 1) bar_sync_n(5 + ((q - 1) & 1), kWinThreads);
    } else {
        // ------------------------------------------------------------------------------------ consumers
        const int ct = tid;
        // scan ownership: warp 0: A0, A4, TD0 | warp 1: A1, AI0, TD1 | warp 2: A2, AI1 | warp 3: A3, LM
        double carry_a = 0.0, carry_b = 0.0;
        int carry_i = 0, carry_t = 0;
        uint32_t carry_lm = 0;
        for (int q = 0;; ++q) {
            const int b = q & 1;
            PipeDesc& D = ps.d[b];
            bar_sync_n(3 + b, kWinThreads);
            if (D.mode != PIPE_RUN) break;
            const int nr = D.nr;
            const int64_t row = D.c0 + ct;
            const uint32_t E0 = D.E[0], Elast = D.Elast;
            uint32_t e0 = 0, e1 = 0;
            if (ct < nr) { e0 = D.E[ct]; e1 = D.E[ct + 1]; }
            {   // ticks that arrive: [e0, e1)
                double sAll = 0.0, sSgn = 0.0, hh = 0.0, d1 = 0.0, d2 = 0.0;
                uint32_t neg = 0, mov = 0, lm = 0;
                uint32_t k0 = e0;
                while (k0 < e1) {
                    const uint32_t slot = k0 & mask;
                    const uint32_t n = min(e1 - k0, (uint32_t)RING - slot);
                    const uint32_t a0 = ring_base + slot * 8u, aend = a0 + n * 8u;
                    uint32_t am = 0;
#pragma unroll 1
                    for (uint32_t a = a0; a < aend; a += 8u) {
                        const double y = lds_f64(a), d = lds_f64_off<kD>(a);
                        sAll += fabs(y);
                        sSgn += y;
                        hh = fma(hh, decay, fabs(y));
                        neg += (uint32_t)__double2hiint(y) >> 31;
                        if (((uint32_t)__double2hiint(y) << 1) != 0u) { ++mov; am = a; }
                        d1 += d;
                        d2 = fma(d, d, d2);
                    }
                    if (am != 0u) lm = k0 + ((am - a0) >> 3) + 1u;
                    k0 += n;
                }
                // ticks that leave the direction window: [q0, q1)
                const uint32_t q0 = max(baseK, sub_floor0(e0, Wd)), q1 = max(baseK, sub_floor0(e1, Wd));
                uint32_t qneg = 0, qmov = 0;
                for_ring<RING>(ring_base, q0, q1, [&](uint32_t a) {
                    const uint32_t hi = lds_hi32(a);
                    qneg += hi >> 31;
                    if ((hi << 1) != 0u) ++qmov;
                });
                ps.A[0][ct] = sAll;
                ps.A[1][ct] = sSgn;
                ps.A[2][ct] = d1;
                ps.A[3][ct] = d2;
                ps.A[4][ct] = (ct < nr) ? P.ewm_alpha * hh * __ldg(P.ewm_ipow + (e1 - E0)) : 0.0;
                ps.AI[0][ct] = (int)neg;
                ps.AI[1][ct] = (int)(neg - qneg) + ((int)(mov - qmov) << 16);
                ps.LM[ct] = lm;
            }
            {   // ticks that leave the mid window: [s0, s1) — same thread, same slots: no exchange needed
                const uint32_t s0 = max(baseK, sub_floor0(e0, Wm)), s1 = max(baseK, sub_floor0(e1, Wm));
                double bAll = 0.0, bSgn = 0.0, b1 = 0.0, b2 = 0.0, hE = 0.0;
                uint32_t neg = 0;
                for_ring<RING>(ring_base, s0, s1, [&](uint32_t a) {
                    const double y = lds_f64(a), d = lds_f64_off<kD>(a);
                    bAll += fabs(y);
                    bSgn += y;
                    hE = fma(hE, decay, fabs(y));
                    neg += (uint32_t)__double2hiint(y) >> 31;
                    b1 += d;
                    b2 = fma(d, d, b2);
                });
                if (s1 > s0) {
                    ps.A[0][ct] -= bAll;
                    ps.A[1][ct] -= bSgn;
                    ps.A[2][ct] -= b1;
                    ps.A[3][ct] -= b2;
                    ps.A[4][ct] -= P.ewm_alpha * hE * __ldg(P.ewm_pow + (e1 - s1)) * __ldg(P.ewm_ipow + (e1 - E0));
                    ps.AI[0][ct] -= (int)neg;
                }
            }
            bar_sync_n(2, kWinRows);
            if (warp == 0) {
                scan128(ps.A[0], carry_a, lane);
                scan128(ps.A[4], carry_b, lane);
                carry_b *= __ldg(P.ewm_pow + (Elast - E0));   // F is relative to the chunk's first tick
                scan128(D.TD[0], carry_t, lane);
            } else if (warp == 1) {
                scan128(ps.A[1], carry_a, lane);
                scan128(ps.AI[0], carry_i, lane);
                scan128(D.TD[1], carry_t, lane);
            } else if (warp == 2) {
                scan128(ps.A[2], carry_a, lane);
                scan128(ps.AI[1], carry_i, lane);
            } else {
                scan128(ps.A[3], carry_a, lane);
                scan128_max(ps.LM, carry_lm, lane);
            }
            bar_sync_n(2, kWinRows);
            if (ct < nr && row >= out_from) {
                const uint32_t e = e1;
                finish_tick_row(out, N, row, P, e, pivot, ps.A[0][ct], ps.A[1][ct], ps.A[2][ct], ps.A[3][ct],
                                ps.A[4][ct] * __ldg(P.ewm_pow + (e - E0)), ps.AI[0][ct], ps.AI[1][ct], ps.LM[ct],
                                lds_f64(ring_base + (sub_floor0(e, Wm) & mask) * 8u),
                                lds_f64(ring_base + ((sub_floor0(e, Wm) + 1u) & mask) * 8u));
                const uint32_t je = D.J[ct];
                const uint32_t jb = sub_floor0(je, Wd);
                const int w0 = D.TD[0][ct], w1 = D.TD[1][ct];
                const int pos = w0 & 0xffff, neg = (w0 >> 16) & 0xffff, rev = w1 & 0xffff;
                const int isum = (w1 - rev) >> 16;
                write_trade_row(out, N, row, P.clog2c, P.inv_cnt, (int)(je - jb), pos, neg, rev, isum);
            }
            bar_sync_n(2, kWinRows);      // the A arrays and this descriptor are free again
            bar_arrive_n(5 + b, kWinThreads);
        }
    }
    __syncthreads();
    // exactly one of the two descriptors carries the terminating mode; it names the first unprocessed row
    const int64_t resume = ps.d[0].mode != PIPE_RUN ? (int64_t)ps.d[0].c0 : (int64_t)ps.d[1].c0;
    __syncthreads();
    return resume;
}

template <int RING>
__global__ void __launch_bounds__(kWinThreads, MRP_WIN_MINB)
k_windows(const DevInputs in, const DevOutputs out, const DevSegment* __restrict__ segs, const DevParams P) {
    extern __shared__ __align__(16) double ring[];  // y: ring[0, RING)   d: ring[RING, 2 RING)
    __shared__ __align__(16) unsigned char shraw[sizeof(PipeShared) > sizeof(WinShared) ? sizeof(PipeShared) : sizeof(WinShared)];
    WinShared& sh = *reinterpret_cast<WinShared*>(shraw);
    PipeShared& ps = *reinterpret_cast<PipeShared*>(shraw);   // the pipelined and the sequential loop never overlap in time
    bool try_pipeline = P.win_pipeline != 0;
    constexpr uint32_t mask = RING - 1;
    constexpr int kD = RING * 8;  // byte offset of the d ring
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = tid & (kWinRows - 1);
    const bool addside = tid < kWinRows;
    const DevSegment sg = segs[blockIdx.x];
    const int64_t ra = sg.row_begin, rb = sg.row_end, sr0 = sg.stream_row0, N = in.n_rows;
    const double* __restrict__ mid = in.mid + sg.tick0;
    const double* __restrict__ spread = in.spread + sg.tick0;
    const int8_t* __restrict__ tdir = in.tdir + sg.trade0;
    const uint32_t* __restrict__ kcur = in.kcur;
    const uint32_t* __restrict__ jcur = in.jcur;
    const uint32_t Wm = (uint32_t)P.mid_window, Wd = (uint32_t)P.dir_window;
    const uint32_t Wmax = max(Wm, Wd + 1);
    const uint32_t span_max = (uint32_t)P.win_span_max;
    const double decay = P.ewm_decay;
    DirectParams dp;
    dp.mid_window = Wm;
    dp.dir_window = Wd;
    dp.spread_divisor = P.spread_divisor;
    dp.ewm_alpha = P.ewm_alpha;
    dp.ewm_decay = P.ewm_decay;

    // block-uniform sliding state
    int64_t c0 = ra, out_from = ra, restart_at = ra;  // next row, first row to write, row the state is rebuilt for
    bool need_restart = true;
    uint32_t Eprev = 0, Jprev = 0, baseK = 0, baseJ = 0;
    double pivot = 0.0;
    double carry = 0.0;              // the scan carry of the quantity this warp owns (warps 0..4)
    int carry_i0 = 0, carry_i1 = 0;  // warp 5: tick counts, warp 6: trade counts
    uint32_t carry_lm = 0;           // warp 7: 1 + position of the last tick that moved the midprice
    uint32_t pre_e = 0, pre_j = 0;   // cursors of the NEXT chunk's row, requested one chunk ahead
    bool pre_valid = false;
    if (tid == 0) sh.excP1 = 0u;
    __syncthreads();

    while (c0 < rb) {
        // ------------------------------------------------------------------ (re)start from a clean halo
        if (need_restart) {
            const int64_t rs = restart_at;
            int64_t h = sr0;
            baseK = 0;
            baseJ = 0;
            if (rs > sr0) {
                const uint32_t e = kcur[rs], j = jcur[rs];
                baseK = sub_floor0(e, Wmax + 1) & ~1u;  // even: tick pairs never straddle the base
                baseJ = sub_floor0(j, Wd + 1);
                // first row whose cursors lie beyond the bases (cursors are non-decreasing): 256-ary search
                int64_t lo = sr0, hi = rs;
                while (lo < hi) {
                    const int64_t step = (hi - lo + kWinThreads - 1) / kWinThreads;
                    const int64_t p = lo + (int64_t)tid * step;
                    const bool below = p < hi && !(kcur[p] > baseK || jcur[p] > baseJ);
                    const int c = __syncthreads_count(below);
                    if (c == 0) hi = lo;
                    else {
                        const int64_t nlo = lo + (int64_t)(c - 1) * step + 1;
                        const int64_t nhi = lo + (int64_t)c * step;
                        lo = nlo;
                        hi = nhi < hi ? nhi : hi;
                    }
                }
                h = lo;
            }
            c0 = h;
            out_from = rs;
            Eprev = baseK;
            Jprev = baseJ;
            pivot = 0.0;
            if ((int64_t)baseK < sg.n_ticks) {
                const double s0 = __ldg(spread + baseK);
                if (fabs(s0) <= 1e150) pivot = s0;
            }
            carry = 0.0;
            carry_i0 = carry_i1 = 0;
            carry_lm = 0;
            need_restart = false;
            pre_valid = false;
            if (try_pipeline) {
                // clean data runs through the producer / consumer pipeline; it hands back the first row it left alone
                try_pipeline = false;
                const int64_t resume = pipelined_chunks<RING>(ps, ring_base, in, out, sg, P, c0, out_from, Eprev, Jprev, baseK, baseJ, pivot);
                if (tid == 0) sh.excP1 = 0u;
                __syncthreads();
                c0 = resume;
                if (c0 >= rb) break;
                out_from = c0 > out_from ? c0 : out_from;
                restart_at = out_from;
                need_restart = true;
                continue;
            }
        }

        // ------------------------------------------------------------------ cursors of up to 128 rows
        const int64_t row = c0 + i;
        uint32_t ce = 0xffffffffu, cj = 0xffffffffu;
        if (row < rb) {
            ce = max(pre_valid ? pre_e : __ldg(kcur + row), baseK);
            cj = max(pre_valid ? pre_j : __ldg(jcur + row), baseJ);
        }
        if (addside) {
            sh.E[i + 1] = ce;
            sh.J[i + 1] = cj;
            if (i == 0) { sh.E[0] = Eprev; sh.J[0] = Jprev; }
        }
        // rows whose new ticks still fit the ring (monotone in i)
        int nr = __syncthreads_count(addside && row < rb && ce - Eprev <= span_max);
        if (nr == 0) {
            // one snapshot with more new ticks than the ring holds: evaluate it directly, then start over behind it
            if (c0 >= out_from) {
                if (tid == 0) direct_tick_row(mid, spread, __ldg(kcur + c0), dp, out.raw, out.counts, N, c0);
                else if (tid == kWinRows) direct_trade_row(tdir, __ldg(jcur + c0), Wd, P.clog2c, out.raw, out.counts, N, c0);
            }
            restart_at = c0 + 1;
            out_from = c0 + 1;
            c0 = c0 + 1;
            need_restart = true;
            __syncthreads();
            continue;
        }
        uint32_t Elast = sh.E[nr];

        // ------------------------------------------------------------------ tick phase: new ticks -> ring, two per thread
        for (uint32_t k = (Eprev & ~1u) + 2u * tid; k < Elast; k += 2u * kWinThreads) {
            // ticks a = k (new unless k < Eprev) and b = k + 1 (new unless past the chunk); k is even
            const bool va = k >= Eprev, vb = k + 1u < Elast;
            // all five loads are issued before any value is touched (a default such as `Mp = Ma` would wait for Ma's load first)
            const double Ma = __ldg(mid + k);
            double Mp = 1.0, Mb = 1.0, Sa = pivot, Sb = pivot;
            if (k != baseK) Mp = __ldg(mid + k - 1);
            if (va) Sa = __ldg(spread + k);
            if (vb) {
                Mb = __ldg(mid + k + 1);
                Sb = __ldg(spread + k + 1);
            }
            if (k == baseK) Mp = Ma;   // the tick at the base (tick 0) has no return: flat
            if (!vb) Mb = Ma;          // defaults give flat, unflagged ticks
            const double ra = rcp_pos(Ma);
            const double ta = Mp - Ma, tb = Ma - Mb;
            const double va_ = ta * ra, vb_ = -tb * ra;        // (Mp-Ma)/Ma and (Mb-Ma)/Ma
            const double ya = signed_sq_return(ta, va_), yb = signed_sq_return(tb, vb_);
            const double da = Sa - pivot, db = Sb - pivot;
            // flagged: a midprice that is not a positive normal number, a move beyond the series range, a spread that is
            // not finite (or absurd).  One flag per pair, attributed to tick b (conservative by one position).
            const bool ok = (uint32_t)(__double2hiint(Ma) - 0x00100000) < 0x7fe00000u && small_mag(va_, 0x3f500000u) &&
                            small_mag(vb_, 0x3f500000u) && small_mag(da, 0x5f138d35u) && small_mag(db, 0x5f138d35u);
            const uint32_t a = ring_base + (k & mask) * 8u;
            if (va) { sts_f64_off<0>(a, ya); sts_f64_off<kD>(a, da); }
            if (vb) { sts_f64_off<8>(a, yb); sts_f64_off<kD + 8>(a, db); }
            if (!ok) atomicMax(&sh.excP1, k + 2u);
        }
        __syncthreads();
        const uint32_t excP1 = sh.excP1;
        if (excP1 != 0u && excP1 - 1u >= baseK) {
            // The sliding state has absorbed (or is about to absorb) a flagged tick.  Rows whose window may still
            // contain it are evaluated directly; the first row past it restarts the state from a clean halo.
            const int nflag = __syncthreads_count(addside && i < nr && (int64_t)ce - (int64_t)Wmax - 2 <= (int64_t)(excP1 - 1u));
            if (nflag == 0) {
                restart_at = c0;
                need_restart = true;
                continue;
            }
            nr = nflag;
            Elast = sh.E[nr];
        }
        // request the next chunk's cursors now; they are consumed at the top of the next iteration
        {
            const int64_t nrow = c0 + nr + i;
            pre_valid = true;
            if (nrow < rb) {
                pre_e = __ldg(kcur + nrow);
                pre_j = __ldg(jcur + nrow);
            }
            // and pull the next chunk's ticks towards the L2
            const int64_t pf = (int64_t)Elast + (int64_t)tid * 16;
            if (tid < 96 && pf < sg.n_ticks) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(mid + pf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(spread + pf));
            }
        }

        // ------------------------------------------------------------------ row phase: per-row deltas
        uint32_t e0 = 0, e1 = 0;
        if (i < nr) { e0 = sh.E[i]; e1 = sh.E[i + 1]; }
        const uint32_t E0 = sh.E[0];
        double bAll = 0.0, bSgn = 0.0, b1 = 0.0, b2 = 0.0, bF = 0.0;
        int bNeg = 0;
        if (addside) {
            // ticks that arrive: [e0, e1)
            double sAll = 0.0, sSgn = 0.0, hh = 0.0, d1 = 0.0, d2 = 0.0;
            uint32_t neg = 0, mov = 0, lm = 0;
            {
                uint32_t k0 = e0;
                while (k0 < e1) {
                    const uint32_t slot = k0 & mask;
                    const uint32_t n = min(e1 - k0, (uint32_t)RING - slot);
                    const uint32_t a0 = ring_base + slot * 8u, aend = a0 + n * 8u;
                    uint32_t am = 0;
#pragma unroll 1
                    for (uint32_t a = a0; a < aend; a += 8u) {
                        const double y = lds_f64(a), d = lds_f64_off<kD>(a);
                        sAll += fabs(y);
                        sSgn += y;
                        hh = fma(hh, decay, fabs(y));
                        neg += (uint32_t)__double2hiint(y) >> 31;
                        if (((uint32_t)__double2hiint(y) << 1) != 0u) { ++mov; am = a; }  // r^2 of a move is normal: high word suffices
                        d1 += d;
                        d2 = fma(d, d, d2);
                    }
                    if (am != 0u) lm = k0 + ((am - a0) >> 3) + 1u;
                    k0 += n;
                }
            }
            // ticks that leave the direction window: [q0, q1)
            const uint32_t q0 = max(baseK, sub_floor0(e0, Wd)), q1 = max(baseK, sub_floor0(e1, Wd));
            uint32_t qneg = 0, qmov = 0;
            for_ring<RING>(ring_base, q0, q1, [&](uint32_t a) {
                const uint32_t hi = lds_hi32(a);
                qneg += hi >> 31;
                // r^2 of a real move is far above 2^-1022: the high word alone tells zero from non-zero
                if ((hi << 1) != 0u) ++qmov;
            });
            sh.A[0][i] = sAll;
            sh.A[1][i] = sSgn;
            sh.A[2][i] = d1;
            sh.A[3][i] = d2;
            sh.A[4][i] = (i < nr) ? P.ewm_alpha * hh * __ldg(P.ewm_ipow + (e1 - E0)) : 0.0;
            sh.AI[0][i] = (int)neg;
            sh.AI[1][i] = (int)(neg - qneg) + ((int)(mov - qmov) << 16);
            sh.LM[i] = lm;
        } else {
            // trades first (independent of the ring): deltas of this row's trade window, read straight from global
            int w0 = 0, w1 = 0;
            if (i < nr) {
                const uint32_t j0 = sh.J[i], j1 = sh.J[i + 1];
                int dprev = j0 > baseJ ? (int)tdir[j0 - 1] : 0;  // the trade at the base (trade 0) closes no pair
                for (uint32_t j = j0; j < j1; ++j) {
                    const int d = tdir[j];
                    w0 += (d > 0 ? 1 : 0) + (d < 0 ? 65536 : 0);
                    w1 += ((d != 0 && d == -dprev) ? 1 : 0) + (d << 16);  // feature_processor.cpp:235
                    dprev = d;
                }
                const uint32_t q0 = max(baseJ, sub_floor0(j0, Wd)), q1 = max(baseJ, sub_floor0(j1, Wd));
                if (q1 > q0) {
                    int d = tdir[q0];
                    for (uint32_t j = q0; j < q1; ++j) {
                        const int dn = tdir[j + 1];
                        w0 -= (d > 0 ? 1 : 0) + (d < 0 ? 65536 : 0);
                        w1 -= ((dn != 0 && dn == -d) ? 1 : 0) + (d << 16);  // the pair (j, j+1) leaves with trade j
                        d = dn;
                    }
                }
            }
            sh.TD[0][i] = w0;
            sh.TD[1][i] = w1;
            // ticks that leave the mid window: [s0, s1)
            const uint32_t s0 = max(baseK, sub_floor0(e0, Wm)), s1 = max(baseK, sub_floor0(e1, Wm));
            double hE = 0.0;
            uint32_t neg = 0;
            for_ring<RING>(ring_base, s0, s1, [&](uint32_t a) {
                const double y = lds_f64(a), d = lds_f64_off<kD>(a);
                bAll += fabs(y);
                bSgn += y;
                hE = fma(hE, decay, fabs(y));
                neg += (uint32_t)__double2hiint(y) >> 31;
                b1 += d;
                b2 = fma(d, d, b2);
            });
            bNeg = (int)neg;
            if (s1 > s0) bF = P.ewm_alpha * hE * __ldg(P.ewm_pow + (e1 - s1)) * __ldg(P.ewm_ipow + (e1 - E0));
        }
        __syncthreads();
        if (!addside) {
            sh.A[0][i] -= bAll;
            sh.A[1][i] -= bSgn;
            sh.A[2][i] -= b1;
            sh.A[3][i] -= b2;
            sh.A[4][i] -= bF;
            sh.AI[0][i] -= bNeg;
        }
        __syncthreads();

        // ------------------------------------------------------------------ one warp per quantity: 128-wide scans
        if (warp < 5) {
            scan128(sh.A[warp], carry, lane);
            if (warp == 4) carry *= __ldg(P.ewm_pow + (Elast - E0));  // F is relative to the chunk's first tick
        } else if (warp == 5) {
            scan128(sh.AI[0], carry_i0, lane);
            scan128(sh.AI[1], carry_i1, lane);
        } else if (warp == 6) {
            scan128(sh.TD[0], carry_i0, lane);
            scan128(sh.TD[1], carry_i1, lane);
        } else {
            scan128_max(sh.LM, carry_lm, lane);
        }
        __syncthreads();

        // ------------------------------------------------------------------ finalise and write rows >= out_from
        if (i < nr && row >= out_from) {
            if (addside) {
                const uint32_t e = e1;  // == the row's own cursor (clamping only affects halo rows)
                const bool flagged = excP1 != 0u && excP1 - 1u >= baseK && (int64_t)e - (int64_t)Wmax - 2 <= (int64_t)(excP1 - 1u);
                if (flagged) {
                    direct_tick_row(mid, spread, e, dp, out.raw, out.counts, N, row);
                } else {
                    finish_tick_row(out, N, row, P, e, pivot, sh.A[0][i], sh.A[1][i], sh.A[2][i], sh.A[3][i],
                                    sh.A[4][i] * __ldg(P.ewm_pow + (e - E0)), sh.AI[0][i], sh.AI[1][i], sh.LM[i],
                                    lds_f64(ring_base + (sub_floor0(e, Wm) & mask) * 8u),
                                    lds_f64(ring_base + ((sub_floor0(e, Wm) + 1u) & mask) * 8u));
                }
            } else {
                const uint32_t je = sh.J[i + 1];
                const uint32_t jb = sub_floor0(je, Wd);
                const int w0 = sh.TD[0][i], w1 = sh.TD[1][i];
                const int pos = w0 & 0xffff, neg = (w0 >> 16) & 0xffff, rev = w1 & 0xffff;
                const int isum = (w1 - rev) >> 16;
                write_trade_row(out, N, row, P.clog2c, P.inv_cnt, (int)(je - jb), pos, neg, rev, isum);
            }
        }
        Eprev = Elast;
        Jprev = sh.J[nr];
        c0 += nr;
        __syncthreads();
    }
}

size_t win_smem_bytes(int ring) { return (size_t)ring * 16u; }

cudaError_t launch_windows(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg, const DevParams& P,
                           cudaStream_t st) {
    const size_t smem = win_smem_bytes(P.win_ring);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<nseg, kWinThreads, smem, st>>>(in, out, segs, P);
        return cudaGetLastError();
    };
    if (P.win_ring == 4096) return go(k_windows<4096>);
    if (P.win_ring == 8192) return go(k_windows<8192>);
    return cudaErrorInvalidValue;
}

}  // namespace mrp
