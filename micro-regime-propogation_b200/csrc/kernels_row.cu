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
// log(p / m) for a price p near the midprice m: the reference takes |log m - log p| of two rounded logs
// (feature_processor.cpp:166,171), whose rounding noise (1 ulp of log m ~ 1.8e-15) is ~3e-11 of a one-tick
// distance.  The Taylor series of log1p((p-m)/m) gives the same quantity to ~3e-16 relative for the cost of a
// dozen fused multiply-adds instead of a full log; beyond |u| = 2^-6 it falls back to the two-log form.
__device__ __forceinline__ double log_ratio(double p, double m, double inv_m, double log_m) {
    const double u = (p - m) * inv_m;
    if (fabs(u) <= 0.015625) {
        double q = 1.0 / 11.0;
        q = fma(q, u, -1.0 / 10.0);
        q = fma(q, u, 1.0 / 9.0);
        q = fma(q, u, -1.0 / 8.0);
        q = fma(q, u, 1.0 / 7.0);
        q = fma(q, u, -1.0 / 6.0);
        q = fma(q, u, 1.0 / 5.0);
        q = fma(q, u, -1.0 / 4.0);
        q = fma(q, u, 1.0 / 3.0);
        q = fma(q, u, -0.5);
        q = fma(q, u, 1.0);
        return q * u;
    }
    return log(p) - log_m;
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

constexpr int kLevelGroup = 5;

__global__ void __launch_bounds__(kThreads, 2)
k_book(const DevInputs in, const DevOutputs out, const DevSegment* __restrict__ segs, const DevParams P) {
    __shared__ double sLM[2][kThreads], sMid[2][kThreads];
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

    double yOfi = 0.0, ySt = 0.0;
    if (tid == 0) sCarLiq = 0.0;
    int buf = 0;
    __syncthreads();

    for (int64_t c0 = wa; c0 < rb; c0 += kThreads, buf ^= 1) {
        const int64_t t = c0 + tid;
        const bool act = t < rb;
        const bool full = act && t >= ra;       // rows that are written (the rest only warm the EMAs)
        const bool hasprev = act && t > sr0;    // a previous snapshot exists in this stream
        const int64_t tc = act ? t : rb - 1;    // clamp loads of idle lanes

        // row scalars first: one batch of independent loads
        const unsigned long long ts = in.ts[tc];
        const double buy = in.buy[tc], sell = in.sell[tc];
        const int adds = in.adds[tc];
        unsigned long long pts = (lane == 0 && hasprev) ? in.ts[tc - 1] : 0ull;

        double bb = 0.0, ba = 0.0, midp = 0.0, log_mid = 0.0, inv_mid = 0.0;
        double raw_ofi = 0.0, liq_b = 0.0, liq_a = 0.0;
        double bid_depth = 0.0, ask_depth = 0.0, bid_lw = 0.0, ask_lw = 0.0;
        double p0b = 0, p1b = 0, p0a = 0, p1a = 0;
        int s0b = 0, s1b = 0, s0a = 0, s1a = 0;
        for (int i0 = 0; i0 < L; i0 += kLevelGroup) {
            // all loads of a group of levels are issued before any is used (the level loop is otherwise
            // a chain of load -> use stalls: 55 % of the kernel's cycles in the first profile)
            double bp[kLevelGroup], ap[kLevelGroup];
            int bs[kLevelGroup], as[kLevelGroup], pbs[kLevelGroup], pas[kLevelGroup];
#pragma unroll
            for (int g = 0; g < kLevelGroup; ++g) {
                const bool lv = i0 + g < L;
                const int64_t o = (int64_t)(lv ? i0 + g : 0) * N + tc;
                bp[g] = lv ? in.bid_px[o] : 0.0;
                ap[g] = lv ? in.ask_px[o] : 0.0;
                bs[g] = lv ? in.bid_sz[o] : 0;
                as[g] = lv ? in.ask_sz[o] : 0;
                // lane 0 fetches the previous row itself; the other lanes get it from their neighbour
                pbs[g] = (lane == 0 && hasprev && lv) ? in.bid_sz[o - 1] : 0;
                pas[g] = (lane == 0 && hasprev && lv) ? in.ask_sz[o - 1] : 0;
            }
            if (i0 == 0) {
                bb = bp[0];  // level 0 == best bid / ask (feature_engine.cpp:22-28)
                ba = ap[0];
                midp = (ba + bb) / 2;  // feature_processor.cpp:34,161
                log_mid = log(midp);
                inv_mid = 1.0 / midp;
            }
#pragma unroll
            for (int g = 0; g < kLevelGroup; ++g) {
                const int i = i0 + g;
                const int nbs = __shfl_up_sync(0xffffffffu, bs[g], 1), nas = __shfl_up_sync(0xffffffffu, as[g], 1);
                if (i < L) {
                    // first snapshot compares with an empty book (order_book.cpp:115,168)
                    const int qb = hasprev ? (lane == 0 ? pbs[g] : nbs) : 0;
                    const int qa = hasprev ? (lane == 0 ? pas[g] : nas) : 0;
                    const int bd = (bs[g] > qb) - (bs[g] < qb);  // order_book.cpp:208-239
                    const int ad = (as[g] > qa) - (as[g] < qa);
                    if (full && out.bdir) {
                        const int64_t o = (int64_t)i * N + tc;
                        out.bdir[o] = (int8_t)bd;
                        out.adir[o] = (int8_t)ad;
                    }
                    const double w = __ldg(P.ofi_weight + i);  // exp(-i*0.5), feature_processor.cpp:121
                    const double bid_contrib = (double)(bd * bs[g]) * w;
                    const double ask_contrib = (double)(ad * as[g]) * w;
                    raw_ofi += bid_contrib - ask_contrib;      // :125
                    if (i < P.stress_levels) {                 // :282-299
                        // exp(-dist*decay): level 0 has dist == 0 exactly (weight 1); on a contiguous book both
                        // sides share the argument, so one exp serves both
                        const double xb = -(bb - bp[g]) * P.stress_distance_decay, xa = -(ap[g] - ba) * P.stress_distance_decay;
                        const double eb = (i == 0) ? 1.0 : exp(xb);
                        const double ea = (i == 0) ? 1.0 : ((xa == xb) ? eb : exp(xa));
                        if (bp[g] > 0.0 && (double)bs[g] >= P.stress_min_size) liq_b += eb * bs[g];
                        if (ap[g] > 0.0 && (double)as[g] >= P.stress_min_size) liq_a += ea * as[g];
                    }
                    if (i == 0) { p0b = bp[g]; p0a = ap[g]; s0b = bs[g]; s0a = as[g]; }
                    if (i == 1) { p1b = bp[g]; p1a = ap[g]; s1b = bs[g]; s1a = as[g]; }
                    if (full) {                                // :164-175, :189-200
                        if (bp[g] > 0) {
                            bid_lw += fabs(log_ratio(bp[g], midp, inv_mid, log_mid)) * bs[g];  // |log mid - log bp|
                            bid_depth += bs[g];
                        }
                        if (ap[g] > 0) {
                            ask_lw += fabs(log_ratio(ap[g], midp, inv_mid, log_mid)) * as[g];  // |log ap - log mid|
                            ask_depth += as[g];
                        }
                    }
                }
            }
        }
        const double log_bb = log(bb), log_ba = log(ba);
        const double liq = liq_b + liq_a;  // :301

        // neighbours: previous row's liquidity, timestamp; lag-21 midprice through shared memory
        sLM[buf][tid] = log_mid;
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
        double old_mid, old_lm;
        {
            const int j = tid - back;
            old_mid = j >= 0 ? sMid[buf][j] : sMid[buf ^ 1][kThreads + j];
            old_lm = j >= 0 ? sLM[buf][j] : sLM[buf ^ 1][kThreads + j];
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
            // log_return, :37-42: the normaliser deque holds min(t, W) earlier rows when row t asks for index `lag`
            const int64_t tl = t - sr0;
            const int64_t have = tl < P.norm_window ? tl : P.norm_window;
            double log_ret = 0.0;
            if ((int64_t)P.return_lag < have && old_mid > 0.0) log_ret = log_mid - old_lm;
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
            raw[(int64_t)MRP_F_LOG_SPREAD * N] = log_ba - log_bb;                   // :33
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
    k_book<<<nseg, kThreads, 0, st>>>(in, out, segs, P);
    return cudaGetLastError();
}

}  // namespace mrp
