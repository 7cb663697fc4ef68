// kernels_row.cu — snapshot-axis (book) pass.
//
//   k_book   : midprice, log_spread, log_return, ofi, signed_volume_pressure, order_arrival_rate, depth_imbalance,
//              market_depth, lob_slope, price_gap, liquidity_stress, and the int8 depth-change directions
//              (ref: feature_processor.cpp:32-43, 111-153, 157-212, 271-317; order_book.cpp:208-239)
#include "mrp_device.cuh"

namespace mrp {

// ------------------------------------------------------------------------------------------------
// book pass: one thread per snapshot row, kThreads (384) rows per chunk, chunks walked in time order so that the
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

// Library exponentials / logarithms sit on rare paths (a level that moved, a row outside the series range) but were inlined a
// dozen times into the level loop: 1.2 k of its 2 k instructions, 350 of the tail's 1 k.  Out of line the chunk loop is 1.6 k
// instructions instead of 3.1 k and fits the SM's 32 KB instruction cache (profiles/r2b_windows_experiments.md has the method).
__device__ __noinline__ double exp_rare(double x) { return exp(x); }
__device__ __noinline__ double log_diff_rare(double a, double b) { return log(a) - log(b); }

// The EMA y_t = decay * y_{t-1} + b_t over a chunk, as a scan of affine maps.  pw2 = decay^(1,2,4,8,16,32)
// (kernel-parameter constants: operands, not loads); pw[n] = decay^n per lane.  Two halves around ONE block barrier so
// that several EMAs share it: ema_warp (in-warp scan, publishes the warp's total), then ema_block (carries across warps).
__device__ __forceinline__ double ema_warp(double b, const double (&pw2)[6], double* sWarp, int lane, int warp) {
    double v = b;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int d = 1 << k;
        const double t = shfl_up_d(v, d);
        if (lane >= d) v = pw2[k] * t + v;
    }
    if (lane == 31) sWarp[warp] = v;
    return v;
}
// A thread folds only the warps BEFORE its own (a warp-uniform trip count: 5.5 steps on average instead of all 12 for every thread,
// which was 8 % of the kernel's instructions); the EMA that leaves the chunk is simply the last row's value — the last thread
// publishes it (decay^32 * carry-into-the-last-warp + the last warp's total: the very expression the full fold ended with).
__device__ __forceinline__ double ema_block(double v, const double (&pw2)[6], double pl, double carry_in, const double* sWarp, int warp) {
    const double p32 = pw2[5];
    double cw = carry_in;
#pragma unroll 1
    for (int u = 0; u < warp; ++u) cw = p32 * cw + sWarp[u];
    return pl * cw + v;
}

#ifndef MRP_BOOK_LG10
#define MRP_BOOK_LG10 0
#endif
#ifndef MRP_BOOK_NS
#define MRP_BOOK_NS 2
#endif
#ifndef MRP_BOOK_MINB
#define MRP_BOOK_MINB 1
#endif
constexpr int kStressCache = 4;  // levels 1..4 keep their last exp(-distance*decay) per thread

// ---- bulk asynchronous copies (TMA, UBLKCP) with mbarrier completion.  One lane per (array, level) of a batch requests the
// chunk's run of that level (kThreads rows) as ONE copy; nobody spends issue slots or registers on per-row loads.  Source
// addresses are rounded down / up to the 16 bytes the copy engine wants and every thread adds the resulting element
// shift (published in the stage header) to its slot index, so any row count and any 8 / 4-byte aligned base works.
constexpr int kPxSlot = kThreads + 4;    // doubles: a chunk's rows + the <= 1 element alignment shift, padded to 16 bytes
constexpr int kSzSlot = kThreads + 12;   // ints: 4 pad + previous row + a chunk's rows + the <= 3 element shift, padded
constexpr int kSzPad = 4;                // the copy lands 16 bytes into the slot, so "previous row" of row 0 never indexes below it

// one landing stage: LG levels of a chunk
template <int LG>
struct __align__(16) BookStage {
    double bp[LG][kPxSlot], ap[LG][kPxSlot];
    int bs[LG][kSzSlot], as[LG][kSzSlot];
    int4 shift[LG];   // slot index of the chunk's first row: x bid_px, y ask_px, z bid_sz, w ask_sz
};

// LG = levels per staged batch (it divides the depth L, so the level loop carries no bounds predicates), NS = stages.
// T = threads (= rows per chunk).  DIRECT: no staging at all — every thread loads its own row's levels from global memory.  That
// form is slower on its own (few warps, exposed load latency) but needs 6 KB of shared memory instead of 94 and a third of the
// registers per CTA, so two of its 96-thread CTAs fit on an SM BESIDE k_windows' one CTA (which holds 215 KB of shared memory
// but leaves 60 % of the issue slots idle): mrp_fg_run gives it a share of the segments to work off in k_windows' shadow.
template <int LG, int NS, int T, bool DIRECT>
__global__ void __launch_bounds__(T, DIRECT ? 1 : MRP_BOOK_MINB)
k_book(const DevInputs in, const DevOutputs out, const DevSegment* __restrict__ segs, const DevParams P) {
    constexpr int kThreads = T, kWarps = T / 32;   // shadow the namespace constants: everything below is per instantiation
    static_assert(DIRECT || 4 * LG <= 32 * kWarps, "one lane per (array, level) run of a batch");
    extern __shared__ __align__(16) unsigned char book_dyn[];
    BookStage<LG>* __restrict__ stg = reinterpret_cast<BookStage<LG>*>(book_dyn);
    double* __restrict__ sOfiW = reinterpret_cast<double*>(book_dyn + (DIRECT ? 0 : sizeof(BookStage<LG>) * NS));  // [L] depth weights
    __shared__ double sInv[2][kThreads], sMid[2][kThreads];
    __shared__ double sWarpLiq[kWarps], sCarLiq;
    __shared__ double sWOfi[kWarps], sWSt[kWarps];
    __shared__ double sYOfi[2], sYSt[2];   // the two EMAs after the previous chunk's last row (double-buffered: readers and the writer of a chunk sit behind the same barrier)
    __shared__ __align__(8) uint64_t sFull[NS], sEmpty[NS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevSegment sg = segs[blockIdx.x];
    const int64_t ra = sg.row_begin, rb = sg.row_end, sr0 = sg.stream_row0;
    const int64_t N = in.n_rows;
    const int L = P.L;
    const int nb = L / LG;
    int64_t wa = ra - P.ema_warm_rows;
    if (wa < sr0) wa = sr0;
    const int back = P.return_lag + 1;  // feature_normalizer.hpp:19-24
    const bool want_dirs = out.bdir != nullptr;

    // exp(-distance*decay) of the near-touch levels: on a regular tick grid the arguments repeat row after row
    double cx[kStressCache], ce[kStressCache];
#pragma unroll
    for (int j = 0; j < kStressCache; ++j) { cx[j] = 0.0; ce[j] = 1.0; }
    if (tid == 0) {
        sCarLiq = 0.0;
        sYOfi[0] = sYOfi[1] = sYSt[0] = sYSt[1] = 0.0;
        if constexpr (!DIRECT) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                mbar_init(&sFull[s], 4 * LG);     // one arrival (with its byte count) per requesting lane
                mbar_init(&sEmpty[s], kWarps);    // one arrival per warp that has finished reading the stage
            }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
    }
    for (int i = tid; i < L; i += kThreads) sOfiW[i] = P.ofi_weight[i];
    int buf = 0;
    __syncthreads();

    // ---- staging pipeline: batch b = (chunk, level group) lands in stage b % NS; batch b + NS - 1 is requested when batch b
    // starts, after every warp has released the stage's previous occupant.  The batch's 4*LG runs are spread over the
    // warps: lane j of warp w owns run w + kWarps j (array a, level g of the group) and keeps its invariants in registers.
    int64_t ic = wa;   // chunk of the next batch to request
    int ib = 0;        // its level group
    int nreq = 0;      // batches requested so far
    const int my_run = warp + kWarps * lane;
    const bool has_run = my_run < 4 * LG;
    const int run_a = has_run ? my_run / LG : 0, run_g = has_run ? my_run - run_a * LG : 0;
    const bool run_px = run_a < 2;
    const int run_esh = run_px ? 3 : 2;
    const uintptr_t run_base = reinterpret_cast<uintptr_t>(run_a == 0 ? (const void*)in.bid_px : run_a == 1 ? (const void*)in.ask_px
                                                          : run_a == 2 ? (const void*)in.bid_sz : (const void*)in.ask_sz);
    // byte offset of the run's landing slot and of its shift word inside a stage
    const uint32_t run_dst = run_a == 0 ? (uint32_t)(offsetof(BookStage<LG>, bp) + run_g * kPxSlot * 8)
                           : run_a == 1 ? (uint32_t)(offsetof(BookStage<LG>, ap) + run_g * kPxSlot * 8)
                           : run_a == 2 ? (uint32_t)(offsetof(BookStage<LG>, bs) + (run_g * kSzSlot + kSzPad) * 4)
                                        : (uint32_t)(offsetof(BookStage<LG>, as) + (run_g * kSzSlot + kSzPad) * 4);
    const uint32_t run_shw = (uint32_t)(offsetof(BookStage<LG>, shift) + run_g * 16 + run_a * 4);
    int64_t run_e = (int64_t)run_g * N + wa;                      // element index of (level, chunk's first row) of the next request
    const int64_t step_in = (int64_t)LG * N;                      // next level group, same chunk
    const int64_t step_out = kThreads - (int64_t)(nb - 1) * LG * N;   // first level group of the next chunk
    auto request = [&]() {
        if constexpr (DIRECT) return;
        if (ic < rb) {
            const int s = nreq % NS;
            if (nreq >= NS) mbar_wait(&sEmpty[s], (uint32_t)((nreq / NS - 1) & 1));
            if (has_run) {
                const int prev = (!run_px && ic > sr0) ? 1 : 0;   // sizes also bring the row before the chunk
                const int64_t left = rb - ic;
                const int nrows = left < kThreads ? (int)left : kThreads;
                const uintptr_t p0 = run_base + ((uintptr_t)(run_e - prev) << run_esh);
                const uintptr_t p1 = p0 + ((uintptr_t)(nrows + prev) << run_esh);
                const uintptr_t q0 = p0 & ~(uintptr_t)15, q1 = (p1 + 15) & ~(uintptr_t)15;
                const uint32_t bytes = (uint32_t)(q1 - q0);
                const int sh = (int)((p0 - q0) >> run_esh) + prev + (run_px ? 0 : kSzPad);
                unsigned char* sbase = reinterpret_cast<unsigned char*>(&stg[s]);
                *reinterpret_cast<int*>(sbase + run_shw) = sh;
                mbar_expect_tx(&sFull[s], bytes);
                bulk_g2s(sbase + run_dst, reinterpret_cast<const void*>(q0), bytes, &sFull[s]);
            }
            ++nreq;
            if (++ib == nb) { ib = 0; ic += kThreads; run_e += step_out; }
            else run_e += step_in;
        }
    };
#pragma unroll
    for (int s = 0; s < NS - 1; ++s) request();
    int nuse = 0;      // batches consumed so far
    const double plOfi = __ldg(P.ofi_pow + lane + 1), plSt = __ldg(P.stress_pow + lane + 1);   // decay^(lane+1)

    for (int64_t c0 = wa; c0 < rb; c0 += kThreads, buf ^= 1) {
        const int64_t t = c0 + tid;
        const bool act = t < rb;
        const bool full = act && t >= ra;       // rows that are written (the rest only warm the EMAs)
        const bool hasprev = act && t > sr0;    // a previous snapshot exists in this stream
        const int64_t tc = act ? t : rb - 1;    // clamp loads of idle lanes
        const bool own_prev = lane == 0 && hasprev;  // lane 0 fetches the previous row itself; the others ask their neighbour

        // row scalars: independent loads, consumed after the level loop
        const unsigned long long ts = in.ts[tc];
        const double buy = in.buy[tc], sell = in.sell[tc];
        const int adds = in.adds[tc];
        unsigned long long pts = own_prev ? in.ts[tc - 1] : 0ull;

        double bb = 0.0, ba = 0.0, midp = 0.0, inv_mid = 0.0;
        double raw_ofi = 0.0, liq_b = 0.0, liq_a = 0.0;
        double bid_depth = 0.0, ask_depth = 0.0, bid_lw = 0.0, ask_lw = 0.0;
        bool oor = false;  // some level outside the log1p series range
        double spread0 = 0.0;        // log ask0 - log bid0 from the level-0 series
        bool spread0_ok = false;
        double p0b = 0, p1b = 0, p0a = 0, p1a = 0;
        int s0b = 0, s1b = 0, s0a = 0, s1a = 0;
        int8_t* __restrict__ obd = out.bdir + tc;
        int8_t* __restrict__ oad = out.adir + tc;
        for (int i0 = 0; i0 < L; i0 += LG) {
            request();   // keeps NS-1 batches in flight
            const int sidx = DIRECT ? 0 : nuse % NS;
            if constexpr (!DIRECT) mbar_wait(&sFull[sidx], (uint32_t)((nuse / NS) & 1));
            ++nuse;
            const BookStage<LG>& S = stg[sidx];
            // DIRECT: this row's LG levels straight from global memory, all 6 * LG loads issued before the first use
            double dbp[DIRECT ? LG : 1], dap[DIRECT ? LG : 1];
            int dbs[DIRECT ? LG : 1], das[DIRECT ? LG : 1], dqb[DIRECT ? LG : 1], dqa[DIRECT ? LG : 1];
            if constexpr (DIRECT) {
                const int64_t e = (int64_t)i0 * N + tc;
#pragma unroll
                for (int g = 0; g < LG; ++g) {
                    dbp[g] = __ldg(in.bid_px + e + (int64_t)g * N);
                    dap[g] = __ldg(in.ask_px + e + (int64_t)g * N);
                    dbs[g] = __ldg(in.bid_sz + e + (int64_t)g * N);
                    das[g] = __ldg(in.ask_sz + e + (int64_t)g * N);
                    dqb[g] = hasprev ? __ldg(in.bid_sz + e + (int64_t)g * N - 1) : 0;
                    dqa[g] = hasprev ? __ldg(in.ask_sz + e + (int64_t)g * N - 1) : 0;
                }
            }
            if (i0 == 0) {
                if constexpr (DIRECT) {
                    bb = dbp[0];
                    ba = dap[0];
                } else {
                    const int4 sh = S.shift[0];
                    bb = S.bp[0][sh.x + tid];  // level 0 == best bid / ask (feature_engine.cpp:22-28)
                    ba = S.ap[0][sh.y + tid];
                }
                midp = (ba + bb) / 2;  // feature_processor.cpp:34,161
                inv_mid = 1.0 / midp;
            }
#pragma unroll
            for (int g = 0; g < LG; ++g) {
                const int i = i0 + g;
                double bp, ap;
                int bs, as, qb = 0, qa = 0;   // first snapshot compares with an empty book (order_book.cpp:115,168)
                if constexpr (DIRECT) {
                    bp = dbp[g]; ap = dap[g]; bs = dbs[g]; as = das[g]; qb = dqb[g]; qa = dqa[g];
                } else {
                    const int4 sh = S.shift[g];
                    bp = S.bp[g][sh.x + tid]; ap = S.ap[g][sh.y + tid];
                    bs = S.bs[g][sh.z + tid]; as = S.as[g][sh.w + tid];
                    if (hasprev) {
                        qb = S.bs[g][sh.z + tid - 1];
                        qa = S.as[g][sh.w + tid - 1];
                    }
                }
                const int bd = (bs > qb) - (bs < qb);  // order_book.cpp:208-239
                const int ad = (as > qa) - (as < qa);
                if (want_dirs && full) {
                    obd[(int64_t)g * N] = (int8_t)bd;
                    oad[(int64_t)g * N] = (int8_t)ad;
                }
                const double w = sOfiW[i];  // exp(-i*0.5), :121
                const double bid_contrib = (double)(bd * bs) * w;   // (int8 * int) * double, as the reference
                const double ask_contrib = (double)(ad * as) * w;
                raw_ofi += bid_contrib - ask_contrib;      // :125
                if (i == 0) { p0b = bp; p0a = ap; s0b = bs; s0a = as; }
                if (i == 1) { p1b = bp; p1a = ap; s1b = bs; s1a = as; }
                {   // :164-175, :189-200 — branch-free.  A side without a price (0, negative, NaN) takes part with the
                    // midprice as its price and size 0: distance 0, weight 0.  Out-of-range arguments only raise the row's flag.
                    const bool hb = bp > 0, ha = ap > 0;
                    const double ub_ = ((hb ? bp : midp) - midp) * inv_mid, ua_ = ((ha ? ap : midp) - midp) * inv_mid;
                    oor |= !(fabs(ub_) <= kSeriesMax) || !(fabs(ua_) <= kSeriesMax);
                    const double sb_ = log1p_series(ub_), sa_ = log1p_series(ua_);
                    const double fbe = (double)(hb ? bs : 0), fae = (double)(ha ? as : 0);
                    bid_lw = fma(fabs(sb_), fbe, bid_lw);   // |log mid - log bp| * size
                    ask_lw = fma(fabs(sa_), fae, ask_lw);   // |log ap - log mid| * size
                    bid_depth += fbe;
                    ask_depth += fae;
                    if (i == 0) {   // log ask0 - log bid0 (:33) = log(ask0 / mid) - log(bid0 / mid)
                        spread0 = sa_ - sb_;
                        spread0_ok = hb && ha;
                    }
                    if (i < P.stress_levels) {   // :282-299 — the levels nearest the touch, from the values already in registers
                        const bool ub = hb && fbe >= P.stress_min_size;   // (a side without a price has fbe == 0 and hb == false)
                        const bool ua = ha && fae >= P.stress_min_size;
                        if (i == 0) {  // distance 0 exactly: weight 1
                            liq_b += ub ? fbe : 0.0;
                            liq_a += ua ? fae : 0.0;
                        } else {
                            const double xb = -(bb - bp) * P.stress_distance_decay, xa = -(ap - ba) * P.stress_distance_decay;
                            if (g >= 1 && g <= kStressCache && i0 == 0) {
                                // weights exp(-distance*decay): on a regular tick grid the arguments repeat row after row, so each
                                // thread remembers its last bid-side argument per level; the ask side shares it when the book is symmetric
                                const int j = (g >= 1 && g <= kStressCache) ? g - 1 : 0;
                                if ((ub && xb != cx[j]) || (ua && xa != cx[j])) {   // rare: something moved
                                    if (ub) {
                                        if (xb != cx[j]) { cx[j] = xb; ce[j] = exp_rare(xb); }
                                        liq_b += ce[j] * fbe;
                                    }
                                    if (ua) liq_a += ((xa == cx[j]) ? ce[j] : exp_rare(xa)) * fae;
                                } else {
                                    liq_b += ub ? ce[j] * fbe : 0.0;
                                    liq_a += ua ? ce[j] * fae : 0.0;
                                }
                            } else {
                                if (ub) liq_b += exp_rare(xb) * fbe;
                                if (ua) liq_a += exp_rare(xa) * fae;
                            }
                        }
                    }
                }
            }
            obd += (int64_t)LG * N;
            oad += (int64_t)LG * N;
            if constexpr (!DIRECT) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sEmpty[sidx]);   // this warp is done with the stage
            }
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

        const double vOfi = ema_warp(act ? P.ofi_alpha * x_ofi : 0.0, P.ofi_pw2, sWOfi, lane, warp);            // :133
        const double vSt = ema_warp(act ? P.stress_alpha * raw_stress : 0.0, P.stress_pw2, sWSt, lane, warp);  // :311-313
        __syncthreads();
        // slot `buf` was written by the previous chunk's last thread (two block barriers ago); this chunk's last thread writes the other
        const double ofi = ema_block(vOfi, P.ofi_pw2, plOfi, sYOfi[buf], sWOfi, warp);
        const double stress = ema_block(vSt, P.stress_pw2, plSt, sYSt[buf], sWSt, warp);
        if (tid == kThreads - 1) {
            sCarLiq = liq;
            sYOfi[buf ^ 1] = ofi;
            sYSt[buf ^ 1] = stress;
        }

        if (full) {
            // log_spread = log(ask0) - log(bid0), :33
            const double log_spread = (spread0_ok && !oor) ? spread0 : log_diff_rare(ba, bb);
            // log_return, :37-42: the normaliser deque holds min(t, W) earlier rows when row t asks for index `lag`
            const int64_t tl = t - sr0;
            const int64_t have = tl < P.norm_window ? tl : P.norm_window;
            double log_ret = 0.0;
            if ((int64_t)P.return_lag < have && old_mid > 0.0) {
                const double u = (midp - old_mid) * old_inv;
                log_ret = fabs(u) <= kSeriesMax ? log1p_series(u) : log_diff_rare(midp, old_mid);
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

template <int LG, int NS>
static cudaError_t launch_book_t(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg,
                                 const DevParams& P, cudaStream_t st) {
    constexpr int kStageBytes = (int)sizeof(BookStage<LG>) * NS;
    const int bytes = kStageBytes + P.L * (int)sizeof(double);  // + the depth weights
    // the opt-in is per device: remembered per device ordinal (idempotent; a race only repeats the call)
    static int ready_for[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int& ready = ready_for[dev & 63];
    if (ready < bytes) {
        e = cudaFuncSetAttribute(k_book<LG, NS, kThreads, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        ready = bytes;
    }
    k_book<LG, NS, kThreads, false><<<nseg, kThreads, bytes, st>>>(in, out, segs, P);
    return cudaGetLastError();
}

cudaError_t launch_book(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg,
                        const DevParams& P, cudaStream_t st) {
    if (nseg <= 0) return cudaSuccess;
    const int L = P.L;
#if MRP_BOOK_LG10
    if (L % 10 == 0) return launch_book_t<10, 2>(in, out, segs, nseg, P, st);
#endif
    if (L % 5 == 0) return launch_book_t<5, MRP_BOOK_NS>(in, out, segs, nseg, P, st);
    if (L % 4 == 0) return launch_book_t<4, 2>(in, out, segs, nseg, P, st);
    if (L % 2 == 0) return launch_book_t<2, 3>(in, out, segs, nseg, P, st);
    return launch_book_t<1, 4>(in, out, segs, nseg, P, st);
}

// The staging-free form for the segments that are worked off beside k_windows (see the kernel's header).  Its dynamic shared
// memory request is padded so that exactly kDirectPerSm of its CTAs fit into what k_windows leaves of an SM.
#ifndef MRP_BOOK_DIRECT_THREADS
#define MRP_BOOK_DIRECT_THREADS 96
#endif
constexpr int kDirectThreads = MRP_BOOK_DIRECT_THREADS;
bool book_direct_supported(const DevParams& P) { return P.L % 5 == 0 && P.return_lag + 1 <= kDirectThreads; }
cudaError_t launch_book_direct(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg, const DevParams& P,
                               int smem_pad, cudaStream_t st) {
    if (nseg <= 0) return cudaSuccess;
    const int bytes = P.L * (int)sizeof(double) + smem_pad;
    static bool carve_set[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!carve_set[dev & 63]) {   // the same L1 / shared-memory split as k_windows, or the two could not share an SM
        e = cudaFuncSetAttribute(k_book<5, 1, kDirectThreads, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        carve_set[dev & 63] = true;
    }
    k_book<5, 1, kDirectThreads, true><<<nseg, kDirectThreads, bytes, st>>>(in, out, segs, P);
    return cudaGetLastError();
}

}  // namespace mrp
