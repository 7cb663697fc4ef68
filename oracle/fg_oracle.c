/*
 * fg_oracle.c — plain-C restatement of microregime::FeatureProcessor + FeatureNormalizer.
 *
 * TEST INFRASTRUCTURE ONLY (see fg_oracle.h).  Every block cites the reference lines it restates
 * (paths relative to the reference root).  The per-row window loops are the reference's own
 * O(window) loops in the reference's own order, so that floating-point rounding is identical;
 * nothing here uses prefix sums or closed forms.
 *
 * Parity: pinned against the compiled reference (oracle/_ref, tests/test_oracle_vs_reference.py).
 */
#include "fg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n;       /* rows in this stream                                   */
    int64_t stride;  /* level / column stride of the row-axis arrays (= total rows of the batch) */
    const uint64_t* ts;
    const double *bid_px, *ask_px;
    const int32_t *bid_sz, *ask_sz;
    const double *buy, *sell;
    const int32_t* adds;
    const uint32_t *kcur, *jcur;
    const double *mid, *spread; /* this stream's ticks  */
    int64_t K;
    const int8_t* tdir;         /* this stream's trades */
    int64_t T;
    double *raw, *norm;         /* column c at [c*stride + r] */
    int8_t *bdir, *adir;
    int32_t* counts;
} stream_t;

#define RAW(c, r) s->raw[(int64_t)(c) * s->stride + (r)]
#define NORM(c, r) s->norm[(int64_t)(c) * s->stride + (r)]
#define CNT(c, r) s->counts[(int64_t)(c) * s->stride + (r)]
#define BP(i, r) s->bid_px[(int64_t)(i) * s->stride + (r)]
#define AP(i, r) s->ask_px[(int64_t)(i) * s->stride + (r)]
#define BS(i, r) s->bid_sz[(int64_t)(i) * s->stride + (r)]
#define AS(i, r) s->ask_sz[(int64_t)(i) * s->stride + (r)]

static inline int sgn_i(int32_t now, int32_t before) { return (now > before) - (now < before); }

/* Everything of one row that does not depend on another row's OUTPUT.
 * x_ofi / liq receive the per-row inputs of the two EMAs. */
static void row_features(const mrp_fg_params* p, const stream_t* s, int64_t t, double* x_ofi, double* liq) {
    const int L = p->depth_levels;
    const double best_bid = BP(0, t); /* feature_engine.cpp:22-28: best_* = level 0 (0.0 when the side is empty) */
    const double best_ask = AP(0, t);

    /* ---- ProcessPriceAndSpread, feature_processor.cpp:32-35 (log_return needs row t-21: done later) */
    RAW(MRP_F_LOG_SPREAD, t) = log(best_ask) - log(best_bid);
    const double midprice = (best_ask + best_bid) / 2;
    RAW(MRP_F_MIDPRICE, t) = midprice;

    /* ---- the rolling mid / spread deques at this snapshot: feature_engine.cpp:130-150 (cap 1800) */
    const int64_t e = s->kcur[t];
    const int64_t b = e > p->mid_window ? e - p->mid_window : 0;
    const double* M = s->mid + b;
    const double* S = s->spread + b;
    const int64_t m = e - b; /* midprices.size() */

    /* ---- ProcessVolatility, feature_processor.cpp:48-108 */
    {
        double realized_variance = 0.0; /* :53-61 */
        int count = 0;
        for (int64_t i = 1; i < m; ++i) {
            if (M[i - 1] == 0.0 || M[i] == 0.0) continue;
            double ret = log(M[i]) - log(M[i - 1]);
            realized_variance += ret * ret;
            count++;
        }
        RAW(MRP_F_REALIZED_VARIANCE, t) = (count > 0) ? realized_variance / count : 0.0;

        const double alpha = 2.0 / (p->ewm_span + 1); /* :64 */
        double ewm_var = 0.0;
        int ewm_count = 0;
        for (int64_t i = 1; i < m; ++i) { /* :67-78, restarted for every snapshot */
            if (M[i - 1] == 0.0 || M[i] == 0.0) continue;
            double ret = log(M[i]) - log(M[i - 1]);
            if (ewm_count == 0) {
                ewm_var = ret * ret;
            } else {
                ewm_var = (1 - alpha) * ewm_var + alpha * (ret * ret);
            }
            ewm_count++;
        }
        RAW(MRP_F_EWM_VOLATILITY, t) = (ewm_count > 0) ? sqrt(ewm_var) : 0.0; /* :79 */

        double up_var = 0.0, down_var = 0.0; /* :82-98; zero returns count as "down" */
        int up_count = 0, down_count = 0;
        for (int64_t i = 1; i < m; ++i) {
            if (M[i - 1] == 0.0 || M[i] == 0.0) continue;
            double ret = log(M[i]) - log(M[i - 1]);
            if (ret > 0) {
                up_var += ret * ret;
                up_count++;
            } else {
                down_var += ret * ret;
                down_count++;
            }
        }
        double avg_up_var = (up_count > 0) ? up_var / up_count : 0.0;
        double avg_down_var = (down_count > 0) ? down_var / down_count : 0.0;
        double diff = avg_up_var - avg_down_var;
        RAW(MRP_F_DIRECTIONAL_VOLATILITY, t) = sqrt(fabs(diff)) * ((diff >= 0) ? 1.0 : -1.0);

        double acc = 0.0; /* :101-107; the divisor is ROLLING_WINDOW (1500), not the deque length */
        for (int64_t i = 0; i < m; ++i) acc += S[i];
        double mean_spread = acc / p->spread_divisor;
        double spread_var = 0.0;
        for (int64_t i = 0; i < m; ++i) spread_var += (S[i] - mean_spread) * (S[i] - mean_spread);
        spread_var /= p->spread_divisor;
        RAW(MRP_F_SPREAD_VOLATILITY, t) = sqrt(spread_var);

        if (s->counts) {
            CNT(MRP_C_RET_VALID, t) = count;
            CNT(MRP_C_RET_UP, t) = up_count;
        }
    }

    /* ---- depth-change directions, order_book.cpp:149-158,208-239: both branches of compute_delta
     *      reduce to sign(new_size - old_size); the book before the first snapshot is all zeros. */
    double raw_ofi = 0.0;
    for (int i = 0; i < L; ++i) {
        int32_t pb = t > 0 ? BS(i, t - 1) : 0, pa = t > 0 ? AS(i, t - 1) : 0;
        int8_t bd = (int8_t)sgn_i(BS(i, t), pb), ad = (int8_t)sgn_i(AS(i, t), pa);
        if (s->bdir) s->bdir[(int64_t)i * s->stride + t] = bd;
        if (s->adir) s->adir[(int64_t)i * s->stride + t] = ad;
        /* ---- ProcessOrderFlow, feature_processor.cpp:120-126 */
        double weight = exp(-i * p->ofi_depth_decay);
        double bid_contrib = bd * BS(i, t) * weight; /* (int8 * int) * double, as in the reference */
        double ask_contrib = ad * AS(i, t) * weight;
        raw_ofi += bid_contrib - ask_contrib;
    }
    const double total_volume = s->buy[t] + s->sell[t]; /* :129-130 */
    *x_ofi = (total_volume > p->ofi_min_total_vol) ? raw_ofi / total_volume : 0.0;
    const double net_signed_volume = s->buy[t] - s->sell[t]; /* :137-140 */
    RAW(MRP_F_SIGNED_VOLUME_PRESSURE, t) = (total_volume > 0.0) ? net_signed_volume / total_volume : 0.0;

    /* order arrival rate, :143-152: cache_.last_arrival_time_ns is the previous row's timestamp_ns */
    {
        uint64_t last = t > 0 ? s->ts[t - 1] : 0;
        double rate = 0.0;
        if (last > 0) {
            uint64_t delta_ns = s->ts[t] - last;
            double seconds = (double)delta_ns * 1e-9;
            rate = (seconds > 0.0) ? (double)s->adds[t] / seconds : 0.0;
        }
        RAW(MRP_F_ORDER_ARRIVAL_RATE, t) = rate;
    }

    /* ---- ProcessLiquidity, feature_processor.cpp:157-212 */
    {
        double bid_depth = 0.0, ask_depth = 0.0;
        double mid = (best_ask + best_bid) / 2;
        double log_mid = log(mid);
        for (int i = 0; i < L; ++i) { /* :164-175 */
            if (BP(i, t) > 0) bid_depth += BS(i, t);
            if (AP(i, t) > 0) ask_depth += AS(i, t);
        }
        RAW(MRP_F_MARKET_DEPTH, t) = bid_depth + ask_depth; /* :178 */
        RAW(MRP_F_DEPTH_IMBALANCE, t) =
            (bid_depth + ask_depth > 0.0) ? (bid_depth - ask_depth) / (bid_depth + ask_depth) : 0.0; /* :181-183 */

        double bid_log_weighted = 0.0, ask_log_weighted = 0.0, bid_sum = 0.0, ask_sum = 0.0; /* :186-204 */
        for (int i = 0; i < L; ++i) {
            if (BP(i, t) > 0) {
                double dist = fabs(log_mid - log(BP(i, t)));
                bid_log_weighted += dist * BS(i, t);
                bid_sum += BS(i, t);
            }
            if (AP(i, t) > 0) {
                double dist = fabs(log(AP(i, t)) - log_mid);
                ask_log_weighted += dist * AS(i, t);
                ask_sum += AS(i, t);
            }
        }
        double bid_slope = (bid_sum > 0.0) ? bid_log_weighted / bid_sum : 0.0;
        double ask_slope = (ask_sum > 0.0) ? ask_log_weighted / ask_sum : 0.0;
        RAW(MRP_F_LOB_SLOPE, t) = bid_slope + ask_slope;

        /* :207-211 — integer denominators, no zero guard; needs L >= 2 */
        double gb = (BP(0, t) * BS(0, t) - BP(1, t) * BS(1, t)) / (BS(0, t) + BS(1, t));
        double ga = (AP(0, t) * AS(0, t) - AP(1, t) * AS(1, t)) / (AS(0, t) + AS(1, t));
        RAW(MRP_F_PRICE_GAP, t) = gb + ga;
    }

    /* ---- ProcessMicrostructureTransitions, feature_processor.cpp:215-248 */
    {
        /* rolling_tick_directions: dir of tick k = sign(M_k - M_{k-1}), first tick 0, cap 1500
         * (feature_engine.cpp:134-145) */
        int up = 0, down = 0, zero = 0;
        int64_t d0 = e > p->dir_window ? e - p->dir_window : 0;
        for (int64_t k = d0; k < e; ++k) {
            int dir = 0;
            if (k > 0) {
                double cur = s->mid[k], prev = s->mid[k - 1];
                dir = (cur > prev) ? 1 : ((cur < prev) ? -1 : 0);
            }
            if (dir > 0) ++up;
            else if (dir < 0) ++down;
            else ++zero;
        }
        double total = up + down + zero;
        double p_up = up / total, p_down = down / total, p_zero = zero / total;
        double tick_entropy = 0.0;
        if (p_up > 0) tick_entropy -= p_up * log2(p_up);
        if (p_down > 0) tick_entropy -= p_down * log2(p_down);
        if (p_zero > 0) tick_entropy -= p_zero * log2(p_zero);
        RAW(MRP_F_TICK_DIRECTION_ENTROPY, t) = tick_entropy;

        /* rolling_trade_directions: last <= 1500 trade dirs (feature_engine.cpp:88-93) */
        const int64_t je = s->jcur[t];
        const int64_t jb = je > p->dir_window ? je - p->dir_window : 0;
        const int8_t* D = s->tdir + jb;
        const int64_t nd = je - jb;
        int reversals = 0; /* :233-239 — divides by size, not size-1 */
        for (int64_t i = 1; i < nd; ++i)
            if (D[i] != 0 && D[i] == -D[i - 1]) ++reversals;
        RAW(MRP_F_REVERSAL_RATE, t) = (nd > 1) ? (double)reversals / (double)nd : 0.0;
        int isum = 0; /* :242-247 — std::accumulate with an int init; 0/0 = NaN with no trades */
        for (int64_t i = 0; i < nd; ++i) isum += D[i];
        double dsum = isum;
        RAW(MRP_F_AGGRESSOR_BIAS, t) = dsum / (double)nd;

        /* ---- ProcessEngineeredFeatures, feature_processor.cpp:254-269 */
        int pos = 0, neg = 0;
        for (int64_t i = 0; i < nd; ++i) {
            if (D[i] > 0) ++pos;
            else if (D[i] < 0) ++neg;
        }
        int tot = pos + neg;
        double p_pos = tot > 0 ? (double)pos / tot : 0.0;
        double p_neg = tot > 0 ? (double)neg / tot : 0.0;
        double entropy = 0.0;
        if (p_pos > 0) entropy -= p_pos * log2(p_pos);
        if (p_neg > 0) entropy -= p_neg * log2(p_neg);
        RAW(MRP_F_SHANNON_ENTROPY, t) = entropy;

        if (s->counts) {
            CNT(MRP_C_TICK_UP, t) = up;
            CNT(MRP_C_TICK_DOWN, t) = down;
            CNT(MRP_C_TICK_TOTAL, t) = up + down + zero;
            CNT(MRP_C_TRADE_POS, t) = pos;
            CNT(MRP_C_TRADE_NEG, t) = neg;
            CNT(MRP_C_TRADE_REVERSALS, t) = reversals;
            CNT(MRP_C_TRADE_TOTAL, t) = (int32_t)nd;
            CNT(MRP_C_TRADE_DIR_SUM, t) = isum;
        }
    }

    /* ---- weighted near-touch liquidity, feature_processor.cpp:274-301 */
    {
        double bid_weighted = 0.0, ask_weighted = 0.0;
        for (int i = 0; i < p->stress_levels; ++i) {
            double bid_price = BP(i, t), ask_price = AP(i, t);
            int bid_size = BS(i, t), ask_size = AS(i, t);
            if (bid_price > 0.0 && bid_size >= p->stress_min_size) {
                double dist = best_bid - bid_price;
                double weight = exp(-dist * p->stress_distance_decay);
                bid_weighted += weight * bid_size;
            }
            if (ask_price > 0.0 && ask_size >= p->stress_min_size) {
                double dist = ask_price - best_ask;
                double weight = exp(-dist * p->stress_distance_decay);
                ask_weighted += weight * ask_size;
            }
        }
        *liq = bid_weighted + ask_weighted;
    }
}

static void run_stream(const mrp_fg_params* p, const stream_t* s, int nthreads) {
    const int64_t n = s->n;
    if (n == 0) return;
    double* x_ofi = (double*)malloc(sizeof(double) * (size_t)n);
    double* liq = (double*)malloc(sizeof(double) * (size_t)n);
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t t = 0; t < n; ++t) row_features(p, s, t, &x_ofi[t], &liq[t]);

    /* ---- the three row-to-row dependencies (FeatureProcessor::Cache, feature_processor.hpp:25-41) */
    double prev_ofi = 0.0, prev_liq = 0.0, prev_stress = 0.0;
    const int64_t back = (int64_t)p->return_lag + 1; /* feature_normalizer.hpp:19-24: window[size-index-1] BEFORE row t is added */
    for (int64_t t = 0; t < n; ++t) {
        /* log_return, feature_processor.cpp:37-42 */
        double old_mid = 0.0;
        int64_t have = t < p->norm_window ? t : p->norm_window; /* normaliser deque size before AddFeatureSet */
        if ((int64_t)p->return_lag < have) old_mid = RAW(MRP_F_MIDPRICE, t - back);
        RAW(MRP_F_LOG_RETURN, t) = (old_mid > 0.0) ? log(RAW(MRP_F_MIDPRICE, t)) - log(old_mid) : 0.0;
        /* ofi EMA, feature_processor.cpp:133-134 */
        double ofi = p->ofi_alpha * x_ofi[t] + (1.0 - p->ofi_alpha) * prev_ofi;
        RAW(MRP_F_OFI, t) = ofi;
        prev_ofi = ofi;
        /* liquidity stress EMA, feature_processor.cpp:303-317 */
        double raw_stress = 0.0;
        if (prev_liq > 0.0) {
            double change = (liq[t] - prev_liq) / prev_liq;
            raw_stress = -change;
        }
        double st = p->stress_alpha * raw_stress + (1.0 - p->stress_alpha) * prev_stress;
        RAW(MRP_F_LIQUIDITY_STRESS, t) = st;
        prev_liq = liq[t];
        prev_stress = st;
    }
    free(x_ofi);
    free(liq);

    /* ---- FeatureNormalizer, feature_normalizer.cpp:32-83: add the row, evict the oldest beyond
     *      WINDOW_SIZE, then z-score with the running sums.  Columns are independent, so the
     *      kFeatures iteration order (:10-30) does not matter. */
    const int64_t W = p->norm_window;
#pragma omp parallel for schedule(static, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int c = 0; c < MRP_FG_NUM_FEATURES; ++c) {
        double sum = 0.0, sum2 = 0.0;
        for (int64_t t = 0; t < n; ++t) {
            double val = RAW(c, t);
            sum += 1.0 * val; /* :39-40, sign = +1 */
            sum2 += 1.0 * val * val;
            int64_t size = t + 1;
            if (size > W) { /* :48-51 */
                double old = RAW(c, t - W);
                sum += -1.0 * old;
                sum2 += -1.0 * old * old;
                size = W;
            }
            double mean = sum / (double)size; /* :62-68 */
            double variance = (sum2 / (double)size) - (mean * mean);
            double stddev = (variance > 0.0) ? sqrt(variance) : 1.0;
            NORM(c, t) = (val - mean) / stddev;
            if (c == 0 && s->counts) CNT(MRP_C_NORM_N, t) = (int32_t)size;
        }
    }
}

int mrp_oracle_run(const mrp_fg_params* p, const mrp_fg_soa_view* v, double* raw, double* norm,
                   int8_t* bid_dir, int8_t* ask_dir, int32_t* counts, int nthreads) {
    if (!p || !v || !raw || !norm || v->n_streams < 1 || v->memory != MRP_MEM_HOST) return 1;
    if (p->depth_levels < 2 || p->stress_levels > p->depth_levels) return 1;
    const int S = v->n_streams;
    const int64_t N = v->row_offsets[S];
    for (int i = 0; i < S; ++i) {
        stream_t s;
        const int64_t r0 = v->row_offsets[i], k0 = v->tick_offsets[i], j0 = v->trade_offsets[i];
        s.n = v->row_offsets[i + 1] - r0;
        s.stride = N;
        s.ts = v->ts_ns + r0;
        s.bid_px = v->bid_px + r0;
        s.ask_px = v->ask_px + r0;
        s.bid_sz = v->bid_sz + r0;
        s.ask_sz = v->ask_sz + r0;
        s.buy = v->buy_vol + r0;
        s.sell = v->sell_vol + r0;
        s.adds = v->adds + r0;
        s.kcur = v->tick_cursor + r0;
        s.jcur = v->trade_cursor + r0;
        s.mid = v->mid + k0;
        s.spread = v->spread + k0;
        s.K = v->tick_offsets[i + 1] - k0;
        s.tdir = v->trade_dir + j0;
        s.T = v->trade_offsets[i + 1] - j0;
        s.raw = raw + r0;
        s.norm = norm + r0;
        s.bdir = bid_dir ? bid_dir + r0 : 0;
        s.adir = ask_dir ? ask_dir + r0 : 0;
        s.counts = counts ? counts + r0 : 0;
        for (int64_t t = 0; t < s.n; ++t)
            if ((int64_t)s.kcur[t] > s.K || (int64_t)s.jcur[t] > s.T) return 1;
        run_stream(p, &s, nthreads);
    }
    return 0;
}

int mrp_oracle_default_params(mrp_fg_params* p) {
    if (!p) return 1;
    memset(p, 0, sizeof(*p));
    p->depth_levels = 10;   /* common_constants.hpp:8 */
    p->mid_window = 1800;   /* feature_engine.cpp:146 */
    p->dir_window = 1500;   /* common_constants.hpp:9 */
    p->spread_divisor = 1500;
    p->ewm_span = 1500;
    p->norm_window = 30000; /* common_constants.hpp:10 */
    p->return_lag = 20;     /* feature_processor.cpp:37 with common_constants.hpp:11 */
    p->stress_levels = 5;
    p->ofi_depth_decay = 0.5;
    p->ofi_alpha = 0.2;
    p->ofi_min_total_vol = 1e-6;
    p->stress_min_size = 5.0;
    p->stress_distance_decay = 10.0;
    p->stress_alpha = 0.1;
    return 0;
}

/* ---- vol-of-vol extension: PARITY UNPINNED (no reference implementation exists; restates include/mrp_fg.h's definition).
 * Every window is summed afresh in long double, oldest row first: no sliding state, nothing shared with the CUDA algorithm. */
static void rolling_stat(const double* x, int64_t n, int window, int zscore, double* out) {
    for (int64_t t = 0; t < n; ++t) {
        const int64_t cnt = t + 1 < window ? t + 1 : window;
        long double s = 0.0L, s2 = 0.0L;
        for (int64_t i = t - cnt + 1; i <= t; ++i) {
            s += (long double)x[i];
            s2 += (long double)x[i] * (long double)x[i];
        }
        const double mean = (double)(s / (long double)cnt);
        /* variance from the exactly summed moments, rounded once: the quantity the device approximates */
        const long double m = s / (long double)cnt;
        const double var = (double)(s2 / (long double)cnt - m * m);
        if (zscore) out[t] = (x[t] - mean) / (var > 0.0 ? sqrt(var) : 1.0);
        else out[t] = var > 0.0 ? sqrt(var) : (var == var ? 0.0 : var);
    }
}

int mrp_oracle_vol_of_vol(const double* ewm, const int64_t* row_offsets, int n_streams, const int32_t* horizons, int n_horizons,
                          int norm_window, double* raw, double* norm) {
    if (!ewm || !row_offsets || !horizons || !raw || !norm || n_streams < 1 || n_horizons < 1 || norm_window < 1) return 1;
    const int64_t N = row_offsets[n_streams];
    for (int k = 0; k < n_horizons; ++k) {
        if (horizons[k] < 1) return 1;
        for (int s = 0; s < n_streams; ++s) {
            const int64_t a = row_offsets[s], n = row_offsets[s + 1] - a;
            rolling_stat(ewm + a, n, horizons[k], 0, raw + (int64_t)k * N + a);
            rolling_stat(raw + (int64_t)k * N + a, n, norm_window, 1, norm + (int64_t)k * N + a);
        }
    }
    return 0;
}
