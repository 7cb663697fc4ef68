// kernels_row.cu — snapshot-axis (book) pass.
//
//   k_book   : midprice, log_spread, log_return, ofi, signed_volume_pressure, order_arrival_rate, depth_imbalance,
//              market_depth, lob_slope, price_gap, liquidity_stress, and the int8 depth-change directions
//              (ref: feature_processor.cpp:32-43, 111-153, 157-212, 271-317; order_book.cpp:208-239)
#include "mrp_device.cuh"

namespace mrp {

// ------------------------------------------------------------------------------------------------
// book pass: one thread per snapshot row, 256 rows per chunk, chunks walked in time order so that the
// two EMAs (ofi: 0.2/0.8, liquidity stress: 0.1/0.9) are carried as block scans of y -> a*y + b.
// A segment that starts mid-stream first walks `ema_warm_rows` halo rows (only what feeds the EMAs
// and the lag-21 midprice); 0.9^halo is below 1e-17, far under the fp64 noise of the features.
// ------------------------------------------------------------------------------------------------
// log(1 + u) by its Taylor series for |u| <= 2^-7 (truncation below 1e-16 relative).  The reference takes
// differences of two rounded logs of nearby prices (feature_processor.cpp:33,38,166,171), whose own rounding
// noise (1 ulp of log 5000 ~ 1.8e-15) is ~4e-11 of a one-tick distance; the series gives the exact quantity for a
// handful of fused multiply-adds instead of two ~90-instruction logs.  Callers fall back to the two-log form
// whenever the argument is out of range (empty side, price 0, NaN, a deep level far from the touch).
constexpr double kSeriesMax = 0x1p-7;
__device__ __forceinline__ double log1p_series(double u) {
    double q = -1.0 / 8.0;
    q = fma(q, u, 1.0 / 7.0);
    q = fma(q, u, -1.0 / 6.0);
    q = fma(q, u, 1.0 / 5.0);
    q = fma(q, u, -1.0 / 4.0);
    q = fma(q, u, 1.0 / 3.0);
    q = fma(q, u, -0.5);
    q = fma(q, u, 1.0);
    return q * u;
}
// The row's size-weighted log distances in the reference's own form (feature_processor.cpp:164-175,186-200), for rows
// where some level lies outside the series range (rare: empty side, NaN, a level > 0.78 % from the midprice).
__device__ __noinline__ void lob_sums_direct(const double* __restrict__ bid_px, const double* __restrict__ ask_px,
                                             const int32_t* __restrict__ bid_sz, const int32_t* __restrict__ ask_sz, int64_t N, int64_t r,
                                             int L, double midp, double* bid_lw, double* ask_lw) {
    const double log_mid = log(midp);
    double bl = 0.0, al = 0.0;
    for (int i = 0; i < L; ++i) {
        const double bp = bid_px[(int64_t)i * N + r], ap = ask_px[(int64_t)i * N + r];
        if (bp > 0) bl += fabs(log_mid - log(bp)) * bid_sz[(int64_t)i * N + r];
        if (ap > 0) al += fabs(log(ap) - log_mid) * ask_sz[(int64_t)i * N + r];
    }
    *bid_lw = bl;
    *ask_lw = al;
}

__device__ __forceinline__ double ema_scan(double b, const double* __restrict__ pw, double carry_in, double* sWarp,
                                           int lane, int warp, double& carry_out) {
    double v = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = shfl_up_d(v, d);
        if (lane >= d) v = __ldg(pw + d) * t + v;
    }
    if (lane == 31) sWarp[warp] = v;
    __syncthreads();
    const double p32 = __ldg(pw + 32);
    double cw = carry_in, co = carry_in;
#pragma unroll
    for (int u = 0; u < kWarps; ++u) {
        if (u == warp) cw = co;
        co = p32 * co + sWarp[u];
    }
    carry_out = co;
    return __ldg(pw + lane + 1) * cw + v;
}

constexpr int kStressCache = 4;  // levels 1..4 keep their last exp(-distance*decay) per thread

// LG = levels per load batch; it divides the depth L, so the level loop carries no bounds predicates.
template <int LG>
__global__ void __launch_bounds__(kThreads, 2)
k_book(const DevInputs in, const DevOutputs out, const DevSegment* __restrict__ segs, const DevParams P) {
    __shared__ double sInv[2][kThreads], sMid[2][kThreads];
    __shared__ double sWarpLiq[kWarps], sCarLiq;
    __shared__ double sWOfi[kWarps], sWSt[kWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevSegment sg = segs[blockIdx.x];
    const int64_t ra = sg.row_begin, rb = sg.row_end, sr0 = sg.stream_row0;
    const int64_t N = in.n_rows;
    const int L = P.L;
    int64_t wa = ra - P.ema_warm_rows;
    if (wa < sr0) wa = sr0;
    const int back = P.return_lag + 1;  // feature_normalizer.hpp:19-24
    const bool want_dirs = out.bdir != nullptr;

    double yOfi = 0.0, ySt = 0.0;
    // exp(-distance*decay) of the near-touch levels: on a regular tick grid the arguments repeat row after row
    double cx[kStressCache], ce[kStressCache];
#pragma unroll
    for (int j = 0; j < kStressCache; ++j) { cx[j] = 0.0; ce[j] = 1.0; }
    if (tid == 0) sCarLiq = 0.0;
    int buf = 0;
    __syncthreads();

    for (int64_t c0 = wa; c0 < rb; c0 += kThreads, buf ^= 1) {
        const int64_t t = c0 + tid;
        const bool act = t < rb;
        const bool full = act && t >= ra;       // rows that are written (the rest only warm the EMAs)
        const bool hasprev = act && t > sr0;    // a previous snapshot exists in this stream
        const int64_t tc = act ? t : rb - 1;    // clamp loads of idle lanes
        const bool own_prev = lane == 0 && hasprev;  // lane 0 fetches the previous row itself; the others ask their neighbour

        // row scalars first: one batch of independent loads
        const unsigned long long ts = in.ts[tc];
        const double buy = in.buy[tc], sell = in.sell[tc];
        const int adds = in.adds[tc];
        unsigned long long pts = own_prev ? in.ts[tc - 1] : 0ull;

        double bb = 0.0, ba = 0.0, midp = 0.0, inv_mid = 0.0;
        double raw_ofi = 0.0, liq_b = 0.0, liq_a = 0.0;
        double bid_depth = 0.0, ask_depth = 0.0, bid_lw = 0.0, ask_lw = 0.0;
        bool oor = false;  // some level outside the log1p series range
        double p0b = 0, p1b = 0, p0a = 0, p1a = 0;
        int s0b = 0, s1b = 0, s0a = 0, s1a = 0;
        const double* __restrict__ pbp = in.bid_px + tc;
        const double* __restrict__ pap = in.ask_px + tc;
        const int32_t* __restrict__ pbs = in.bid_sz + tc;
        const int32_t* __restrict__ pas = in.ask_sz + tc;
        int8_t* __restrict__ obd = out.bdir + tc;
        int8_t* __restrict__ oad = out.adir + tc;
        for (int i0 = 0; i0 < L; i0 += LG) {
            // all loads of a batch of levels are issued before any is used
            double bp[LG], ap[LG];
            int bs[LG], as[LG], qbs[LG], qas[LG];
#pragma unroll
            for (int g = 0; g < LG; ++g) {
                const int64_t o = (int64_t)g * N;
                bp[g] = pbp[o];
                ap[g] = pap[o];
                bs[g] = pbs[o];
                as[g] = pas[o];
                qbs[g] = own_prev ? pbs[o - 1] : 0;
                qas[g] = own_prev ? pas[o - 1] : 0;
            }
            if (i0 == 0) {
                bb = bp[0];  // level 0 == best bid / ask (feature_engine.cpp:22-28)
                ba = ap[0];
                midp = (ba + bb) / 2;  // feature_processor.cpp:34,161
                inv_mid = 1.0 / midp;
            }
#pragma unroll
            for (int g = 0; g < LG; ++g) {
                const int i = i0 + g;
                // first snapshot compares with an empty book (order_book.cpp:115,168)
                const int nbs = __shfl_up_sync(0xffffffffu, bs[g], 1), nas = __shfl_up_sync(0xffffffffu, as[g], 1);
                const int qb = hasprev ? (lane == 0 ? qbs[g] : nbs) : 0;
                const int qa = hasprev ? (lane == 0 ? qas[g] : nas) : 0;
                const int bd = (bs[g] > qb) - (bs[g] < qb);  // order_book.cpp:208-239
                const int ad = (as[g] > qa) - (as[g] < qa);
                if (want_dirs && full) {
                    obd[(int64_t)g * N] = (int8_t)bd;
                    oad[(int64_t)g * N] = (int8_t)ad;
                }
                const double w = i < kOfiWeights ? P.ofi_w[i < kOfiWeights ? i : 0] : __ldg(P.ofi_weight + i);  // exp(-i*0.5), :121
                const double bid_contrib = (double)(bd * bs[g]) * w;   // (int8 * int) * double, as the reference
                const double ask_contrib = (double)(ad * as[g]) * w;
                raw_ofi += bid_contrib - ask_contrib;      // :125
                const double fbs = (double)bs[g], fas = (double)as[g];
                if (i < P.stress_levels && (i == 0 || i > kStressCache || i0 != 0)) {   // :282-299 (levels 1..kStressCache of
                    const bool ub = bp[g] > 0.0 && fbs >= P.stress_min_size;              //  the first batch: cached, below)
                    const bool ua = ap[g] > 0.0 && fas >= P.stress_min_size;
                    if (i == 0) {  // distance 0 exactly: weight 1
                        if (ub) liq_b += fbs;
                        if (ua) liq_a += fas;
                    } else {
                        const double xb = -(bb - bp[g]) * P.stress_distance_decay, xa = -(ap[g] - ba) * P.stress_distance_decay;
                        if (ub) liq_b += exp(xb) * fbs;
                        if (ua) liq_a += exp(xa) * fas;
                    }
                }
                if (i == 0) { p0b = bp[g]; p0a = ap[g]; s0b = bs[g]; s0a = as[g]; }
                if (i == 1) { p1b = bp[g]; p1a = ap[g]; s1b = bs[g]; s1a = as[g]; }
                {   // :164-175, :189-200 — branch-free: out-of-range arguments only raise the row's flag
                    const double ub_ = (bp[g] - midp) * inv_mid, ua_ = (ap[g] - midp) * inv_mid;
                    const bool hb = bp[g] > 0, ha = ap[g] > 0;
                    oor |= (hb && !(fabs(ub_) <= kSeriesMax)) || (ha && !(fabs(ua_) <= kSeriesMax));
                    const double lb = fabs(log1p_series(ub_)), la = fabs(log1p_series(ua_));
                    bid_lw += hb ? lb * fbs : 0.0;   // |log mid - log bp| * size
                    ask_lw += ha ? la * fas : 0.0;   // |log ap - log mid| * size
                    bid_depth += hb ? fbs : 0.0;
                    ask_depth += ha ? fas : 0.0;
                }
            }
            if (i0 == 0) {
                // near-touch liquidity weights exp(-distance*decay) of levels 1..kStressCache (:282-299): on a regular tick
                // grid the arguments repeat row after row, so each thread remembers its last bid-side argument per level
                // and the ask side shares it when the book is symmetric
#pragma unroll
                for (int j = 0; j < kStressCache; ++j) {
                    constexpr int kTop = LG - 1;
                    const int g = j + 1 < kTop ? j + 1 : kTop;   // == j + 1 whenever the level exists in this batch
                    if (j + 1 < LG && j + 1 < P.stress_levels) {
                        const double fbs = (double)bs[g], fas = (double)as[g];
                        const bool ub = bp[g] > 0.0 && fbs >= P.stress_min_size;
                        const bool ua = ap[g] > 0.0 && fas >= P.stress_min_size;
                        const double xb = -(bb - bp[g]) * P.stress_distance_decay, xa = -(ap[g] - ba) * P.stress_distance_decay;
                        if (ub) {
                            if (xb != cx[j]) { cx[j] = xb; ce[j] = exp(xb); }
                            liq_b += ce[j] * fbs;
                        }
                        if (ua) liq_a += ((xa == cx[j]) ? ce[j] : exp(xa)) * fas;
                    }
                }
            }
            pbp += (int64_t)LG * N;
            pap += (int64_t)LG * N;
            pbs += (int64_t)LG * N;
            pas += (int64_t)LG * N;
            obd += (int64_t)LG * N;
            oad += (int64_t)LG * N;
        }
        if (oor && full) lob_sums_direct(in.bid_px, in.ask_px, in.bid_sz, in.ask_sz, N, tc, L, midp, &bid_lw, &ask_lw);
        const double liq = liq_b + liq_a;  // :301

        // neighbours: previous row's liquidity, timestamp; lag-21 midprice through shared memory
        sInv[buf][tid] = inv_mid;
        sMid[buf][tid] = midp;
        if (lane == 31) sWarpLiq[warp] = liq;
        {
            const unsigned long long nts = __shfl_up_sync(0xffffffffu, ts, 1);
            if (lane != 0) pts = nts;
            if (!hasprev) pts = 0ull;
        }
        __syncthreads();
        double pliq = shfl_up_d(liq, 1);
        if (lane == 0) pliq = warp == 0 ? sCarLiq : sWarpLiq[warp - 1];
        if (!hasprev) pliq = 0.0;  // cache_.prev_liquidity starts at 0 (feature_processor.hpp:27)
        // lag-(return_lag+1) midprice: read NOW (two barriers separate this read from the next chunk's writes)
        double old_mid, old_inv;
        {
            const int j = tid - back;
            old_mid = j >= 0 ? sMid[buf][j] : sMid[buf ^ 1][kThreads + j];
            old_inv = j >= 0 ? sInv[buf][j] : sInv[buf ^ 1][kThreads + j];
        }

        const double total_volume = buy + sell;                                               // :129
        const double x_ofi = (total_volume > P.ofi_min_total_vol) ? raw_ofi / total_volume : 0.0;  // :130
        double raw_stress = 0.0;                                                              // :304-308
        if (pliq > 0.0) raw_stress = -((liq - pliq) / pliq);

        double coOfi, coSt;
        const double ofi = ema_scan(act ? P.ofi_alpha * x_ofi : 0.0, P.ofi_pow, yOfi, sWOfi, lane, warp, coOfi);  // :133
        const double stress = ema_scan(act ? P.stress_alpha * raw_stress : 0.0, P.stress_pow, ySt, sWSt, lane, warp, coSt);  // :311-313
        yOfi = coOfi;
        ySt = coSt;
        if (tid == kThreads - 1) sCarLiq = liq;

        if (full) {
            // log_spread = log(ask0) - log(bid0), :33
            double log_spread;
            {
                const double u = (ba - bb) / bb;
                if (bb > 0.0 && fabs(u) <= kSeriesMax) log_spread = log1p_series(u);
                else log_spread = log(ba) - log(bb);
            }
            // log_return, :37-42: the normaliser deque holds min(t, W) earlier rows when row t asks for index `lag`
            const int64_t tl = t - sr0;
            const int64_t have = tl < P.norm_window ? tl : P.norm_window;
            double log_ret = 0.0;
            if ((int64_t)P.return_lag < have && old_mid > 0.0) {
                const double u = (midp - old_mid) * old_inv;
                log_ret = fabs(u) <= kSeriesMax ? log1p_series(u) : log(midp) - log(old_mid);
            }
            double rate = 0.0;  // :143-151
            if (pts > 0ull) {
                const unsigned long long delta_ns = ts - pts;
                const double seconds = (double)delta_ns * 1e-9;
                rate = (seconds > 0.0) ? (double)adds / seconds : 0.0;
            }
            const double depth = bid_depth + ask_depth;
            const double bid_slope = (bid_depth > 0.0) ? bid_lw / bid_depth : 0.0;  // :202-204
            const double ask_slope = (ask_depth > 0.0) ? ask_lw / ask_depth : 0.0;
            const double gb = (p0b * s0b - p1b * s1b) / (double)(s0b + s1b);        // :207-211
            const double ga = (p0a * s0a - p1a * s1a) / (double)(s0a + s1a);
            double* raw = out.raw + t;
            raw[(int64_t)MRP_F_MIDPRICE * N] = midp;
            raw[(int64_t)MRP_F_LOG_SPREAD * N] = log_spread;
            raw[(int64_t)MRP_F_LOG_RETURN * N] = log_ret;
            raw[(int64_t)MRP_F_OFI * N] = ofi;
            raw[(int64_t)MRP_F_SIGNED_VOLUME_PRESSURE * N] = (total_volume > 0.0) ? (buy - sell) / total_volume : 0.0;  // :137-140
            raw[(int64_t)MRP_F_ORDER_ARRIVAL_RATE * N] = rate;
            raw[(int64_t)MRP_F_DEPTH_IMBALANCE * N] = (depth > 0.0) ? (bid_depth - ask_depth) / depth : 0.0;  // :181-183
            raw[(int64_t)MRP_F_MARKET_DEPTH * N] = depth;                           // :178
            raw[(int64_t)MRP_F_LOB_SLOPE * N] = bid_slope + ask_slope;
            raw[(int64_t)MRP_F_PRICE_GAP * N] = gb + ga;
            raw[(int64_t)MRP_F_LIQUIDITY_STRESS * N] = stress;
        }
    }
}

cudaError_t launch_book(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg,
                        const DevParams& P, cudaStream_t st) {
    const int L = P.L;
    if (L % 5 == 0) k_book<5><<<nseg, kThreads, 0, st>>>(in, out, segs, P);
    else if (L % 4 == 0) k_book<4><<<nseg, kThreads, 0, st>>>(in, out, segs, P);
    else if (L % 2 == 0) k_book<2><<<nseg, kThreads, 0, st>>>(in, out, segs, P);
    else k_book<1><<<nseg, kThreads, 0, st>>>(in, out, segs, P);
    return cudaGetLastError();
}

}  // namespace mrp
