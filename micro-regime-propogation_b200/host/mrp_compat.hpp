// mrp_compat.hpp — the reference's boundary types for the feature_generation stage.
//
// When the adaptor is built inside the reference tree (its include directories on the -I path) the reference's own
// headers are used, so FeatureSet / FeatureInputSnapshot / DataReciever ARE the reference's types:
//     feature_generation/include/feature_set.hpp:10-44
//     data_ingestion/include/feature_snapshot.hpp:12-44
//     feature_generation/include/data_reciever.hpp:7-16
// Built stand-alone (this repository, the GPU box: the reference is not present) equivalent declarations with the same
// member names are provided so that the adaptor, its tools and its tests compile unchanged.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#if defined(MRP_USE_REFERENCE_HEADERS) || (__has_include("feature_set.hpp") && __has_include("feature_snapshot.hpp") && __has_include("data_reciever.hpp"))
#include <cmath>
#include <iostream>
#include <optional>
#include "feature_set.hpp"
#include "feature_snapshot.hpp"
#include "data_reciever.hpp"
#define MRP_HAVE_REFERENCE_HEADERS 1
namespace mrp_host {
constexpr std::size_t kDepthLevels = microregime::DEPTH_LEVELS;
}
#else
#ifndef MRP_DEPTH_LEVELS
#define MRP_DEPTH_LEVELS 10
#endif
namespace mrp_host {
constexpr std::size_t kDepthLevels = MRP_DEPTH_LEVELS;
}
namespace microregime {

// Output row: declaration order == CSV column order.
struct FeatureSet {
    std::uint64_t timestamp_ns = 0;
    std::string instrument;
    double midprice = 0, log_spread = 0, log_return = 0;
    double ewm_volatility = 0, realized_variance = 0, directional_volatility = 0, spread_volatility = 0;
    double ofi = 0, signed_volume_pressure = 0, order_arrival_rate = 0;
    double depth_imbalance = 0, market_depth = 0, lob_slope = 0, price_gap = 0;
    double tick_direction_entropy = 0, reversal_rate = 0, aggressor_bias = 0;
    double shannon_entropy = 0, liquidity_stress = 0;
};

// Sink the pipeline pushes (raw, normalised) rows into.
class DataReciever {
public:
    virtual void ingest_feature_set(const std::string& symbol, std::uint64_t timestamp_ns, const FeatureSet& raw_features,
                                    const FeatureSet& normalized) = 0;
    virtual ~DataReciever() = default;
};

}  // namespace microregime

// Input row produced by FeatureEngine::generate_snapshot.
struct FeatureInputSnapshot {
    std::uint64_t timestamp_ns = 0;
    std::string instrument;
    double best_bid_price = 0, best_ask_price = 0;
    std::array<double, mrp_host::kDepthLevels> bid_prices{}, ask_prices{};
    std::array<int, mrp_host::kDepthLevels> bid_sizes{}, ask_sizes{};
    double rolling_buy_volume = 0, rolling_sell_volume = 0;
    int adds_since_last_snapshot = 0;
    const std::deque<double>* rolling_midprices = nullptr;
    const std::deque<double>* rolling_spreads = nullptr;
    const std::deque<std::int8_t>* rolling_tick_directions = nullptr;
    const std::deque<std::int8_t>* rolling_trade_directions = nullptr;
    std::array<std::int8_t, mrp_host::kDepthLevels> bid_depth_change_direction{}, ask_depth_change_direction{};
    std::uint32_t reserved = 0;
};
#endif

namespace mrp_host {

// The 19 feature members of FeatureSet in declaration (== CSV, == mrp_fg_feature) order.
inline constexpr double microregime::FeatureSet::*kFeatureMembers[19] = {
    &microregime::FeatureSet::midprice, &microregime::FeatureSet::log_spread, &microregime::FeatureSet::log_return,
    &microregime::FeatureSet::ewm_volatility, &microregime::FeatureSet::realized_variance,
    &microregime::FeatureSet::directional_volatility, &microregime::FeatureSet::spread_volatility, &microregime::FeatureSet::ofi,
    &microregime::FeatureSet::signed_volume_pressure, &microregime::FeatureSet::order_arrival_rate,
    &microregime::FeatureSet::depth_imbalance, &microregime::FeatureSet::market_depth, &microregime::FeatureSet::lob_slope,
    &microregime::FeatureSet::price_gap, &microregime::FeatureSet::tick_direction_entropy, &microregime::FeatureSet::reversal_rate,
    &microregime::FeatureSet::aggressor_bias, &microregime::FeatureSet::shannon_entropy, &microregime::FeatureSet::liquidity_stress};

inline constexpr const char* kFeatureNames[19] = {
    "midprice", "log_spread", "log_return", "ewm_volatility", "realized_variance", "directional_volatility", "spread_volatility",
    "ofi", "signed_volume_pressure", "order_arrival_rate", "depth_imbalance", "market_depth", "lob_slope", "price_gap",
    "tick_direction_entropy", "reversal_rate", "aggressor_bias", "shannon_entropy", "liquidity_stress"};

}  // namespace mrp_host
