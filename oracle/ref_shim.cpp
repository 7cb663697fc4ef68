// ref_shim.cpp — drives the UNMODIFIED reference classes over a structure-of-arrays batch.
//
// TEST INFRASTRUCTURE ONLY (see fg_oracle.h).  This file contains no feature arithmetic: it
// includes the reference's own headers from /root/reference and is linked with the reference's
// own translation units (feature_processor.cpp, feature_normalizer.cpp, feature_engine.cpp,
// order_book.cpp — compiled where they lie, see oracle/Makefile) into oracle/_ref/libmrp_ref.so.
// It only converts between the staged arrays of include/mrp_fg.h and the reference's call
// sequence (dual_feature_pipeline.cpp:70-116):
//     FeatureEngine::UpdateMidpriceAndSpread   per 50 ms tick
//     FeatureEngine::update_trade              per trade
//     FeatureEngine::update_events('A')        per add
//     FeatureEngine::generate_snapshot         per snapshot  (use_book = 1: the reference's real
//                                              OrderBookManager is rebuilt to the row's top-L book,
//                                              so GetL3Snapshot / GetDepthChange / compute_delta run)
//     FeatureProcessor::GetRawFeatureSet + GetProcessedFeatureSet
#include <cstdint>
#include <vector>
#include <cmath>
#include <iostream>
#include <optional>
#include <string>
#include <thread>
#include <atomic>

#include "feature_processor.hpp"
#include "feature_engine.hpp"
#include "order_book.hpp"
#include "common_constants.hpp"

#include "../include/mrp_fg.h"

namespace {

using microregime::FeatureProcessor;
using microregime::FeatureSet;

inline void store(const FeatureSet& f, double* out, int64_t stride, int64_t r) {
    const double v[MRP_FG_NUM_FEATURES] = {
        f.midprice, f.log_spread, f.log_return, f.ewm_volatility, f.realized_variance,
        f.directional_volatility, f.spread_volatility, f.ofi, f.signed_volume_pressure,
        f.order_arrival_rate, f.depth_imbalance, f.market_depth, f.lob_slope, f.price_gap,
        f.tick_direction_entropy, f.reversal_rate, f.aggressor_bias, f.shannon_entropy,
        f.liquidity_stress};
    for (int c = 0; c < MRP_FG_NUM_FEATURES; ++c) out[(int64_t)c * stride + r] = v[c];
}

void run_stream(const mrp_fg_soa_view* v, int s, const double* trade_size, int use_book, double* raw,
                double* norm, int8_t* bid_dir, int8_t* ask_dir, double* snap_buy, double* snap_sell) {
    constexpr int L = (int)microregime::DEPTH_LEVELS;
    const int64_t N = v->row_offsets[v->n_streams];
    const int64_t r0 = v->row_offsets[s], r1 = v->row_offsets[s + 1];
    const int64_t k0 = v->tick_offsets[s], j0 = v->trade_offsets[s];

    OrderBookManager book;
    FeatureEngine engine(book, "SYN");
    FeatureProcessor processor;

    uint64_t next_id = 1;
    std::vector<uint64_t> live;
    int64_t k = 0, j = 0;
    int32_t prev_b[L] = {0}, prev_a[L] = {0};

    for (int64_t r = r0; r < r1; ++r) {
        for (; k < (int64_t)v->tick_cursor[r]; ++k)
            engine.UpdateMidpriceAndSpread(v->mid[k0 + k], v->spread[k0 + k]);
        for (; j < (int64_t)v->trade_cursor[r]; ++j)
            engine.update_trade(0.0, trade_size ? trade_size[j0 + j] : 1.0, v->trade_dir[j0 + j]);
        engine.most_recent_timestamp_ns = v->ts_ns[r];

        FeatureInputSnapshot snap;
        if (use_book) {
            for (int a = 0; a < v->adds[r]; ++a) engine.update_events('A');
            for (uint64_t id : live) book.ApplyCancel(id, 0);
            live.clear();
            for (int i = 0; i < L; ++i) {
                if (v->bid_px[(int64_t)i * N + r] > 0) {
                    book.ApplyAdd(next_id, v->bid_px[(int64_t)i * N + r], v->bid_sz[(int64_t)i * N + r], BookSide::Bid);
                    live.push_back(next_id++);
                }
                if (v->ask_px[(int64_t)i * N + r] > 0) {
                    book.ApplyAdd(next_id, v->ask_px[(int64_t)i * N + r], v->ask_sz[(int64_t)i * N + r], BookSide::Ask);
                    live.push_back(next_id++);
                }
            }
            snap = engine.generate_snapshot();
        } else {
            L3Snapshot l3;
            for (int i = 0; i < L; ++i) {
                l3.bid[i] = {v->bid_px[(int64_t)i * N + r], v->bid_sz[(int64_t)i * N + r]};
                l3.ask[i] = {v->ask_px[(int64_t)i * N + r], v->ask_sz[(int64_t)i * N + r]};
            }
            snap = engine.generate_snapshot_from_l3(l3, v->ts_ns[r]);
            snap.rolling_buy_volume = v->buy_vol[r];
            snap.rolling_sell_volume = v->sell_vol[r];
            snap.adds_since_last_snapshot = v->adds[r];
            for (int i = 0; i < L; ++i) { // staged-array rule: sign of the level's size change
                int32_t b = l3.bid[i].size, a = l3.ask[i].size;
                snap.bid_depth_change_direction[i] = (int8_t)((b > prev_b[i]) - (b < prev_b[i]));
                snap.ask_depth_change_direction[i] = (int8_t)((a > prev_a[i]) - (a < prev_a[i]));
                prev_b[i] = b;
                prev_a[i] = a;
            }
            snap.instrument = "SYN";
        }
        if (snap_buy) snap_buy[r] = snap.rolling_buy_volume;
        if (snap_sell) snap_sell[r] = snap.rolling_sell_volume;
        for (int i = 0; i < L; ++i) {
            if (bid_dir) bid_dir[(int64_t)i * N + r] = snap.bid_depth_change_direction[i];
            if (ask_dir) ask_dir[(int64_t)i * N + r] = snap.ask_depth_change_direction[i];
        }
        FeatureSet fr = processor.GetRawFeatureSet(snap);
        FeatureSet fn = processor.GetProcessedFeatureSet(fr);
        store(fr, raw, N, r);
        store(fn, norm, N, r);
    }
}

}  // namespace

extern "C" {

int mrp_ref_depth_levels(void) { return (int)microregime::DEPTH_LEVELS; }

// Returns 0 on success.  `trade_size` may be NULL unless use_book != 0.
int mrp_ref_run(const mrp_fg_soa_view* v, const double* trade_size, int use_book, double* raw, double* norm,
                int8_t* bid_dir, int8_t* ask_dir, double* snap_buy, double* snap_sell, int nthreads) {
    if (!v || !raw || !norm || v->memory != MRP_MEM_HOST || v->n_streams < 1) return 1;
    if (use_book && !trade_size) return 1;
    // The reference prints from inside the hot path (feature_normalizer.cpp:64-66, order_book.cpp:161-167);
    // silence std::cout for the duration.
    std::streambuf* saved = std::cout.rdbuf(nullptr);
    const int S = v->n_streams;
    int nt = nthreads < 1 ? 1 : (nthreads > S ? S : nthreads);
    int rc = 0;
    try {
        if (nt == 1) {
            for (int s = 0; s < S; ++s) run_stream(v, s, trade_size, use_book, raw, norm, bid_dir, ask_dir, snap_buy, snap_sell);
        } else {
            std::atomic<int> next{0};
            std::atomic<int> failed{0};
            std::vector<std::thread> pool;
            for (int w = 0; w < nt; ++w)
                pool.emplace_back([&] {
                    try {
                        for (int s; (s = next.fetch_add(1)) < S;)
                            run_stream(v, s, trade_size, use_book, raw, norm, bid_dir, ask_dir, snap_buy, snap_sell);
                    } catch (...) {
                        failed = 1;
                    }
                });
            for (auto& t : pool) t.join();
            rc = failed ? 2 : 0;
        }
    } catch (...) {
        rc = 2;
    }
    std::cout.rdbuf(saved);
    std::cout.clear();
    return rc;
}

}  // extern "C"
