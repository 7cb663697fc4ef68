// ref_csv_shim.cpp — feeds rows to the UNMODIFIED reference CsvWriter (TEST INFRASTRUCTURE ONLY, see fg_oracle.h).
//
// The reference defines CsvWriter inside features_to_csv.cpp, next to main(); this translation unit includes that
// file where it lies (with main renamed) so that the reference's own class formats the golden CSV files the
// product's CsvWriter is compared with byte for byte.  No formatting logic lives here.
#include <cstdint>
#include <vector>
#include <cmath>
#include <iostream>
#include <optional>
#include <stdexcept>

#define main mrp_ref_features_to_csv_main
#include "features_to_csv.cpp"   // -I$(REF)/feature_generation/src/data
#undef main

#include "../include/mrp_fg.h"

// link-time substitute for the databento-backed reader (data_ingestion/src/data/dbn_reader.cpp): never constructed here
struct DbnMboReader::Impl {};
DbnMboReader::DbnMboReader(const std::string&, const std::string& instrument) : instrument_(instrument) {
    throw std::runtime_error("DbnMboReader: no databento in the checker build");
}
DbnMboReader::~DbnMboReader() = default;
bool DbnMboReader::has_next() const { return false; }
MarketEvent DbnMboReader::next_event() { throw std::runtime_error("DbnMboReader: no databento in the checker build"); }

extern "C" int mrp_ref_write_csv(const char* base_filename, uint64_t interval_ns, const char* date, const char* instrument, int64_t n,
                                 const uint64_t* grid_ts, const uint64_t* row_ts, const double* raw, const double* norm) {
    std::streambuf* saved = std::cout.rdbuf(nullptr);  // the constructor prints the directory it creates
    int rc = 0;
    try {
        microregime::CsvWriter w(base_filename, interval_ns, date);
        for (int64_t r = 0; r < n; ++r) {
            microregime::FeatureSet f[2];
            for (int k = 0; k < 2; ++k) {
                const double* c = k == 0 ? raw : norm;
                microregime::FeatureSet& s = f[k];
                s.timestamp_ns = row_ts[r];
                s.instrument = instrument;
                s.midprice = c[0 * n + r]; s.log_spread = c[1 * n + r]; s.log_return = c[2 * n + r];
                s.ewm_volatility = c[3 * n + r]; s.realized_variance = c[4 * n + r]; s.directional_volatility = c[5 * n + r];
                s.spread_volatility = c[6 * n + r]; s.ofi = c[7 * n + r]; s.signed_volume_pressure = c[8 * n + r];
                s.order_arrival_rate = c[9 * n + r]; s.depth_imbalance = c[10 * n + r]; s.market_depth = c[11 * n + r];
                s.lob_slope = c[12 * n + r]; s.price_gap = c[13 * n + r]; s.tick_direction_entropy = c[14 * n + r];
                s.reversal_rate = c[15 * n + r]; s.aggressor_bias = c[16 * n + r]; s.shannon_entropy = c[17 * n + r];
                s.liquidity_stress = c[18 * n + r];
            }
            w.ingest_feature_set(instrument, grid_ts[r], f[0], f[1]);
        }
    } catch (...) {
        rc = 2;
    }
    std::cout.rdbuf(saved);
    std::cout.clear();
    return rc;
}
