// kernels_tick.cu — tick-axis passes (the 50 ms midprice / spread streams).
//
//   k_tick_returns : ewm_volatility, realized_variance, directional_volatility, tick_direction_entropy
//                    (ref: feature_processor.cpp:48-98 and :215-229; tick directions feature_engine.cpp:134-145)
//   k_tick_spread  : spread_volatility  (ref: feature_processor.cpp:100-107)
//
// One CTA streams through one row segment.  Per chunk of 1024 ticks each thread takes 4 consecutive
// ticks (one log each), a block scan turns the per-tick terms into running prefix values, and the
// values are parked in a shared-memory ring that always holds the last (window + chunk) ticks.  A
// snapshot row whose cursor falls inside the chunk is then finalised from ring look-ups at its window
// end (cursor-1) and window start (cursor-window): every window sum of the reference's O(1800) loops is
// a difference of two ring entries.
#include "mrp_device.cuh"

namespace mrp {

// ------------------------------------------------------------------------------------------------
// returns pass
// ------------------------------------------------------------------------------------------------
struct __align__(16) RetEntry {
    double pu;  // running sum of r^2 over valid returns with r > 0      (feature_processor.cpp:87-89)
    double pd;  // running sum of r^2 over valid returns with !(r > 0)   (feature_processor.cpp:90-92)
    double g;   // G_k = decay*G_{k-1} + alpha*r_k^2 on valid returns (global EW recurrence, see finalise)
    unsigned long long cnt;  // five chunk-local inclusive counts, 11 bits each
};
constexpr int F_UP = 0, F_VALID = 11, F_BAD = 22, F_TUP = 33, F_TDN = 44;

struct RetPoint {
    double pu, pd, g;
    uint32_t up, valid, bad, tup, tdn;  // segment-running counts
};

// 32-byte entries, 16 bytes of padding after every 4 (a thread stores 4 consecutive entries with
// 128-bit stores: lane stride 144 B keeps each quarter-warp on 32 distinct banks).
__device__ __forceinline__ unsigned ret_byte(int slot) { return (unsigned)slot * 32u + ((unsigned)slot >> 2) * 16u; }

__global__ void __launch_bounds__(kThreads, 2)
k_tick_returns(const DevInputs in, const DevOutputs out, const DevSegment* __restrict__ segs, const DevParams P) {
    extern __shared__ __align__(16) unsigned char ring[];
    __shared__ double sWarpM[kWarps], sWarpLm[kWarps];
    __shared__ double sCarM, sCarLm;
    __shared__ double sPU[kWarps], sPD[kWarps], sA[kWarps], sB[kWarps];
    __shared__ unsigned long long sC[kWarps];
    __shared__ uint32_t sBase[kMaxRingChunks][5];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevSegment sg = segs[blockIdx.x];
    const int64_t ra = sg.row_begin, rb = sg.row_end;
    const double* __restrict__ mid = in.mid + sg.tick0;
    const uint32_t* __restrict__ kcur = in.kcur;
    const int RING = P.ret_ring;
    const int NCHB = (RING + kChunk - 1) / kChunk + 1;
    const int64_t kend = kcur[rb - 1];
    const int wmax = max(P.mid_window, P.dir_window + 1);
    int64_t base = (int64_t)kcur[ra] - wmax - 2;
    if (base < 0) base = 0;
    int64_t nchunks = (kend - base + kChunk - 1) / kChunk;
    if (nchunks < 1) nchunks = 1;

    double cPU = 0.0, cPD = 0.0, cG = 0.0;
    uint32_t cUp = 0, cValid = 0, cBad = 0, cTup = 0, cTdn = 0;
    if (tid == 0) {
        double m0 = base > 0 ? mid[base - 1] : 0.0;
        sCarM = m0;
        sCarLm = base > 0 ? log(m0) : 0.0;
    }
    int64_t r_next = ra;
    int cs_slot = 0, cmod = 0;

    double Mn[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        int64_t kk = base + tid * kPerThread + i;
        Mn[i] = kk < kend ? __ldg(mid + kk) : 1.0;
    }

    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t cs = base + c * kChunk;
        // cursor of the first row this thread may finalise: requested now, consumed after the scan
        int64_t e_pref = r_next + tid < rb ? (int64_t)kcur[r_next + tid] : (int64_t)0x7fffffffffffffll;
        double M[kPerThread], lm[kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) M[i] = Mn[i];
        if (c + 1 < nchunks) {  // register prefetch of the next chunk hides the HBM latency behind this chunk's math
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                int64_t kk = cs + kChunk + tid * kPerThread + i;
                Mn[i] = kk < kend ? __ldg(mid + kk) : 1.0;
            }
        }
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) lm[i] = log(M[i]);
        if (lane == 31) {
            sWarpM[warp] = M[kPerThread - 1];
            sWarpLm[warp] = lm[kPerThread - 1];
        }
        if (tid == 0) {
            sBase[cmod][0] = cUp;
            sBase[cmod][1] = cValid;
            sBase[cmod][2] = cBad;
            sBase[cmod][3] = cTup;
            sBase[cmod][4] = cTdn;
        }
        __syncthreads();
        double Mp = shfl_up_d(M[kPerThread - 1], 1), lmp = shfl_up_d(lm[kPerThread - 1], 1);
        if (lane == 0) {
            Mp = warp == 0 ? sCarM : sWarpM[warp - 1];
            lmp = warp == 0 ? sCarLm : sWarpLm[warp - 1];
        }

        // ---- thread-local inclusive prefixes over its 4 ticks
        double lpu[kPerThread], lpd[kPerThread], lg[kPerThread], lap[kPerThread];
        unsigned long long lc[kPerThread];
        double pu = 0.0, pd = 0.0, g = 0.0, ap = 1.0;
        unsigned long long cn = 0ull;
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const int64_t kk = cs + tid * kPerThread + i;
            const bool hasp = kk > 0;  // the very first tick of a stream has no predecessor (feature_engine.cpp:134-141)
            const bool valid = hasp && (M[i] != 0.0) && (Mp != 0.0);  // feature_processor.cpp:56
            const double r = lm[i] - lmp;                             // :57
            double x = r * r;
            const bool bad = valid && !finite_d(x);
            const bool up = valid && (r > 0.0);                       // :87
            if (!valid || bad) x = 0.0;
            pu += up ? x : 0.0;
            pd += (valid && !up) ? x : 0.0;
            if (valid) {
                g = P.ewm_decay * g + P.ewm_alpha * x;
                ap *= P.ewm_decay;
            }
            const int d = hasp ? ((M[i] > Mp) - (M[i] < Mp)) : 0;     // feature_engine.cpp:138
            cn += (up ? (1ull << F_UP) : 0ull) + (valid ? (1ull << F_VALID) : 0ull) + (bad ? (1ull << F_BAD) : 0ull) +
                  (d > 0 ? (1ull << F_TUP) : 0ull) + (d < 0 ? (1ull << F_TDN) : 0ull);
            lpu[i] = pu; lpd[i] = pd; lg[i] = g; lap[i] = ap; lc[i] = cn;
            Mp = M[i];
            lmp = lm[i];
        }

        // ---- block scan of the thread aggregates: sums, the affine map y -> ap*y + g, packed counts
        double ipu = pu, ipd = pd, iA = ap, iB = g;
        unsigned long long ic = cn;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double tpu = shfl_up_d(ipu, d), tpd = shfl_up_d(ipd, d), tA = shfl_up_d(iA, d), tB = shfl_up_d(iB, d);
            const unsigned long long tc = shfl_up_u64(ic, d);
            if (lane >= d) {
                ipu += tpu;
                ipd += tpd;
                iB = iA * tB + iB;
                iA = iA * tA;
                ic += tc;
            }
        }
        if (lane == 31) { sPU[warp] = ipu; sPD[warp] = ipd; sA[warp] = iA; sB[warp] = iB; sC[warp] = ic; }
        __syncthreads();
        double epu = shfl_up_d(ipu, 1), epd = shfl_up_d(ipd, 1), eA = shfl_up_d(iA, 1), eB = shfl_up_d(iB, 1);
        unsigned long long ec = shfl_up_u64(ic, 1);
        if (lane == 0) { epu = 0.0; epd = 0.0; eA = 1.0; eB = 0.0; ec = 0ull; }
        double wpu = 0.0, wpd = 0.0, wA = 1.0, wB = 0.0, tpu = 0.0, tpd = 0.0, tA = 1.0, tB = 0.0;
        unsigned long long wc = 0ull, tc = 0ull;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            if (w == warp) { wpu = tpu; wpd = tpd; wA = tA; wB = tB; wc = tc; }
            tpu += sPU[w];
            tpd += sPD[w];
            tB = sA[w] * tB + sB[w];
            tA = tA * sA[w];
            tc += sC[w];
        }
        const double PUin = cPU + (wpu + epu), PDin = cPD + (wpd + epd);
        const double Gin = eA * (wA * cG + wB) + eB;
        const unsigned long long Cin = wc + ec;
        if (tid == kThreads - 1) { sCarM = M[kPerThread - 1]; sCarLm = lm[kPerThread - 1]; }

        {
            int slot = cs_slot + tid * kPerThread;
            if (slot >= RING) slot -= RING;
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                RetEntry en;
                en.pu = PUin + lpu[i];
                en.pd = PDin + lpd[i];
                en.g = lap[i] * Gin + lg[i];
                en.cnt = Cin + lc[i];
                *reinterpret_cast<RetEntry*>(ring + ret_byte(slot + i)) = en;
            }
        }
        cPU += tpu;
        cPD += tpd;
        cG = tA * cG + tB;
        cUp += (uint32_t)((tc >> F_UP) & 0x7ffu);
        cValid += (uint32_t)((tc >> F_VALID) & 0x7ffu);
        cBad += (uint32_t)((tc >> F_BAD) & 0x7ffu);
        cTup += (uint32_t)((tc >> F_TUP) & 0x7ffu);
        cTdn += (uint32_t)((tc >> F_TDN) & 0x7ffu);
        __syncthreads();

        // ---- finalise every row whose cursor lies in (cs, cs+chunk]  (rows with cursor 0 ride along in chunk 0)
        auto get = [&](int64_t k) -> RetPoint {
            const int off = (int)(k - cs);
            int slot = cs_slot + off;
            if (slot < 0) slot += RING;
            else if (slot >= RING) slot -= RING;
            const RetEntry en = *reinterpret_cast<const RetEntry*>(ring + ret_byte(slot));
            int ci = cmod + (off >> kChunkShift);
            if (ci < 0) ci += NCHB;
            RetPoint p;
            p.pu = en.pu; p.pd = en.pd; p.g = en.g;
            p.up = sBase[ci][0] + (uint32_t)((en.cnt >> F_UP) & 0x7ffu);
            p.valid = sBase[ci][1] + (uint32_t)((en.cnt >> F_VALID) & 0x7ffu);
            p.bad = sBase[ci][2] + (uint32_t)((en.cnt >> F_BAD) & 0x7ffu);
            p.tup = sBase[ci][3] + (uint32_t)((en.cnt >> F_TUP) & 0x7ffu);
            p.tdn = sBase[ci][4] + (uint32_t)((en.cnt >> F_TDN) & 0x7ffu);
            return p;
        };
        const int64_t ce = (c + 1 == nchunks) ? (int64_t)0x7fffffffffffll : cs + kChunk;
        for (bool first = true;; first = false) {
            const int64_t r = r_next + tid;
            const int64_t e = first ? e_pref : (r < rb ? (int64_t)kcur[r] : (int64_t)0x7fffffffffffffll);
            const bool q = e <= ce;
            if (q) {
                double rv = 0.0, ewm = 0.0, dv = 0.0, tent = 0.0;
                uint32_t n_valid = 0, n_up = 0, t_up = 0, t_dn = 0, t_tot = 0;
                if (e > 0) {
                    const int64_t hi = e - 1;
                    int64_t s = e - P.mid_window;
                    if (s < 0) s = 0;
                    const RetPoint H = get(hi), S = get(s);
                    n_valid = H.valid - S.valid;
                    n_up = H.up - S.up;
                    const uint32_t nbad = H.bad - S.bad;
                    const double sumU = H.pu - S.pu, sumD = H.pd - S.pd;
                    if (n_valid > 0) {
                        const double tot = sumU + sumD;
                        rv = tot * __ldg(P.inv_cnt + n_valid);  // feature_processor.cpp:61 (tabulated 1/count)
                        // EW variance restarted at the window's first valid return f (:64-79).  With the global
                        // recurrence G (same decay, same inputs) the restarted state obeys
                        //   w_k - G_k = decay^(#valid in (f,k]) * (w_f - G_f),  w_f = r_f^2,
                        // and because decay + alpha = 1:  w_end = G_end + decay^cnt * (r_f^2 - G_s).
                        double w = 0.0;
                        if (tot != 0.0) {
                            int64_t f = s + 1;
                            RetPoint Fm = S, F = get(f);
                            while (F.valid == Fm.valid) { Fm = F; ++f; F = get(f); }
                            const double xf = (F.pu - Fm.pu) + (F.pd - Fm.pd);
                            w = H.g + __ldg(P.ewm_pow + n_valid) * (xf - S.g);
                            if (!(w > 0.0)) w = 0.0;
                        }
                        ewm = sqrt(w);
                        const uint32_t n_dn = n_valid - n_up;
                        const double au = sumU * __ldg(P.inv_cnt + n_up);   // :95-98 (inv_cnt[0] = 0 covers the empty side)
                        const double ad = sumD * __ldg(P.inv_cnt + n_dn);
                        const double diff = au - ad;
                        dv = sqrt(fabs(diff)) * ((diff >= 0) ? 1.0 : -1.0);
                        if (nbad) { rv = ewm = dv = __longlong_as_double(0x7ff8000000000000ll); }
                    }
                    int64_t d0 = e - P.dir_window;  // tick-direction deque, feature_engine.cpp:143-145
                    if (d0 < 0) d0 = 0;
                    t_up = H.tup;
                    t_dn = H.tdn;
                    if (d0 > 0) {
                        const RetPoint D = get(d0 - 1);
                        t_up -= D.tup;
                        t_dn -= D.tdn;
                    }
                    t_tot = (uint32_t)(e - d0);
                    const uint32_t zero = t_tot - t_up - t_dn;
                    // feature_processor.cpp:222-229:  -sum p_i log2 p_i with p_i = c_i/T
                    //   = (T log2 T - sum c_i log2 c_i) / T, with c log2 c tabulated by the host (exact 0 for one bucket)
                    tent = ((__ldg(P.clog2c + t_tot) - __ldg(P.clog2c + t_up)) - __ldg(P.clog2c + t_dn) - __ldg(P.clog2c + zero)) /
                           (double)t_tot;
                }
                out.raw[(int64_t)MRP_F_EWM_VOLATILITY * in.n_rows + r] = ewm;
                out.raw[(int64_t)MRP_F_REALIZED_VARIANCE * in.n_rows + r] = rv;
                out.raw[(int64_t)MRP_F_DIRECTIONAL_VOLATILITY * in.n_rows + r] = dv;
                out.raw[(int64_t)MRP_F_TICK_DIRECTION_ENTROPY * in.n_rows + r] = tent;
                if (out.counts) {
                    out.counts[(int64_t)MRP_C_RET_VALID * in.n_rows + r] = (int32_t)n_valid;
                    out.counts[(int64_t)MRP_C_RET_UP * in.n_rows + r] = (int32_t)n_up;
                    out.counts[(int64_t)MRP_C_TICK_UP * in.n_rows + r] = (int32_t)t_up;
                    out.counts[(int64_t)MRP_C_TICK_DOWN * in.n_rows + r] = (int32_t)t_dn;
                    out.counts[(int64_t)MRP_C_TICK_TOTAL * in.n_rows + r] = (int32_t)t_tot;
                }
            }
            const int nq = __syncthreads_count(q);
            r_next += nq;
            if (nq < kThreads) break;
        }
        cs_slot += kChunk;
        if (cs_slot >= RING) cs_slot -= RING;
        cmod = (cmod + 1 == NCHB) ? 0 : cmod + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// spread pass
// ------------------------------------------------------------------------------------------------
// ring entry: running sums of d = s - pivot and d^2 (pivot = the segment's first finite spread, so the
// window variance is evaluated without cancellation against the spread LEVEL), and the running count of
// non-finite spreads (GetSpread() is NaN while a side is empty, order_book.cpp:248-253).
struct SprEntry {
    double d1, d2;
    uint32_t bad, pad;
};
// 24-byte entries + 8 bytes per 4: a thread's 4 entries start 104 B apart -> conflict-free 64-bit accesses
__device__ __forceinline__ unsigned spr_byte(int slot) { return (unsigned)slot * 24u + ((unsigned)slot >> 2) * 8u; }

__global__ void __launch_bounds__(kThreads, 3)
k_tick_spread(const DevInputs in, const DevOutputs out, const DevSegment* __restrict__ segs, const DevParams P) {
    extern __shared__ __align__(16) unsigned char ring[];
    __shared__ double sD1[kWarps], sD2[kWarps];
    __shared__ uint32_t sBad[kWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevSegment sg = segs[blockIdx.x];
    const int64_t ra = sg.row_begin, rb = sg.row_end;
    const double* __restrict__ spread = in.spread + sg.tick0;
    const uint32_t* __restrict__ kcur = in.kcur;
    const int RING = P.spr_ring;
    const int64_t kend = kcur[rb - 1];
    int64_t base = (int64_t)kcur[ra] - P.mid_window - 2;
    if (base < 0) base = 0;
    int64_t nchunks = (kend - base + kChunk - 1) / kChunk;
    if (nchunks < 1) nchunks = 1;

    double pivot = base < kend ? __ldg(spread + base) : 0.0;
    if (!finite_d(pivot)) pivot = 0.0;
    double cD1 = 0.0, cD2 = 0.0;
    uint32_t cBad = 0;
    int64_t r_next = ra;
    int cs_slot = 0;

    double Sn[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        int64_t kk = base + tid * kPerThread + i;
        Sn[i] = kk < kend ? __ldg(spread + kk) : pivot;
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t cs = base + c * kChunk;
        // cursor of the first row this thread may finalise: requested now, consumed after the scan
        int64_t e_pref = r_next + tid < rb ? (int64_t)kcur[r_next + tid] : (int64_t)0x7fffffffffffffll;
        double l1[kPerThread], l2[kPerThread];
        uint32_t lb[kPerThread];
        double a1 = 0.0, a2 = 0.0;
        uint32_t ab = 0;
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            double d = Sn[i] - pivot;
            const bool bad = !finite_d(d);
            if (bad) d = 0.0;
            a1 += d;
            a2 += d * d;
            ab += bad ? 1u : 0u;
            l1[i] = a1; l2[i] = a2; lb[i] = ab;
        }
        if (c + 1 < nchunks) {
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                int64_t kk = cs + kChunk + tid * kPerThread + i;
                Sn[i] = kk < kend ? __ldg(spread + kk) : pivot;
            }
        }
        double i1 = a1, i2 = a2;
        uint32_t ib = ab;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double t1 = shfl_up_d(i1, d), t2 = shfl_up_d(i2, d);
            const uint32_t tb = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { i1 += t1; i2 += t2; ib += tb; }
        }
        if (lane == 31) { sD1[warp] = i1; sD2[warp] = i2; sBad[warp] = ib; }
        __syncthreads();
        double e1 = shfl_up_d(i1, 1), e2 = shfl_up_d(i2, 1);
        uint32_t eb = __shfl_up_sync(0xffffffffu, ib, 1);
        if (lane == 0) { e1 = 0.0; e2 = 0.0; eb = 0; }
        double w1 = 0.0, w2 = 0.0, t1 = 0.0, t2 = 0.0;
        uint32_t wb = 0, tb = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            if (w == warp) { w1 = t1; w2 = t2; wb = tb; }
            t1 += sD1[w];
            t2 += sD2[w];
            tb += sBad[w];
        }
        const double in1 = cD1 + (w1 + e1), in2 = cD2 + (w2 + e2);
        const uint32_t inb = cBad + wb + eb;
        {
            int slot = cs_slot + tid * kPerThread;
            if (slot >= RING) slot -= RING;
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                SprEntry* en = reinterpret_cast<SprEntry*>(ring + spr_byte(slot + i));
                en->d1 = in1 + l1[i];
                en->d2 = in2 + l2[i];
                en->bad = inb + lb[i];
            }
        }
        cD1 += t1;
        cD2 += t2;
        cBad += tb;
        __syncthreads();

        auto get = [&](int64_t k) -> SprEntry {
            int slot = cs_slot + (int)(k - cs);
            if (slot < 0) slot += RING;
            else if (slot >= RING) slot -= RING;
            return *reinterpret_cast<const SprEntry*>(ring + spr_byte(slot));
        };
        const int64_t ce = (c + 1 == nchunks) ? (int64_t)0x7fffffffffffll : cs + kChunk;
        for (bool first = true;; first = false) {
            const int64_t r = r_next + tid;
            const int64_t e = first ? e_pref : (r < rb ? (int64_t)kcur[r] : (int64_t)0x7fffffffffffffll);
            const bool q = e <= ce;
            if (q) {
                double sv = 0.0;
                if (e > 0) {
                    int64_t s = e - P.mid_window;
                    if (s < 0) s = 0;
                    const SprEntry H = get(e - 1);
                    double D1 = H.d1, D2 = H.d2;
                    uint32_t nbad = H.bad;
                    if (s > base) {  // s == base only when base == 0 (nothing precedes the window)
                        const SprEntry S = get(s - 1);
                        D1 -= S.d1;
                        D2 -= S.d2;
                        nbad -= S.bad;
                    }
                    // feature_processor.cpp:101-107 with s_i = pivot + d_i:
                    //   mean = (sum s)/1500,  var = sum (s_i - mean)^2 / 1500   (the deque holds n <= 1800 values)
                    const double n = (double)(e - s);
                    const double div = (double)P.spread_divisor;
                    const double mean = (D1 + n * pivot) / div;
                    const double delta = mean - pivot;
                    double ss = D2 - 2.0 * delta * D1 + n * delta * delta;
                    if (!(ss > 0.0)) ss = 0.0;
                    sv = sqrt(ss / div);
                    if (nbad) sv = __longlong_as_double(0x7ff8000000000000ll);
                }
                out.raw[(int64_t)MRP_F_SPREAD_VOLATILITY * in.n_rows + r] = sv;
            }
            const int nq = __syncthreads_count(q);
            r_next += nq;
            if (nq < kThreads) break;
        }
        cs_slot += kChunk;
        if (cs_slot >= RING) cs_slot -= RING;
    }
}

// host-side launch helpers ---------------------------------------------------------------------
size_t ret_ring_bytes(int ring) { return (size_t)ring * 36u; }
size_t spr_ring_bytes(int ring) { return (size_t)ring * 26u; }

cudaError_t launch_tick_returns(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg,
                                const DevParams& P, cudaStream_t st) {
    const size_t smem = ret_ring_bytes(P.ret_ring);
    cudaError_t e = cudaFuncSetAttribute(k_tick_returns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_tick_returns<<<nseg, kThreads, smem, st>>>(in, out, segs, P);
    return cudaGetLastError();
}

cudaError_t launch_tick_spread(const DevInputs& in, const DevOutputs& out, const DevSegment* segs, int nseg,
                               const DevParams& P, cudaStream_t st) {
    const size_t smem = spr_ring_bytes(P.spr_ring);
    cudaError_t e = cudaFuncSetAttribute(k_tick_spread, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_tick_spread<<<nseg, kThreads, smem, st>>>(in, out, segs, P);
    return cudaGetLastError();
}

}  // namespace mrp
