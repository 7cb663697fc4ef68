// gpu_feature_processor.hpp — C++20 host adaptor over the C ABI (include/mrp_fg.h).
//
// Stands in for microregime::FeatureProcessor (feature_generation/include/feature_processor.hpp:9-46) in the
// pipeline loop (feature_generation/src/core/dual_feature_pipeline.cpp:103-117).  The reference computes one row per
// call; a GPU launch per row is meaningless, so this adaptor STAGES what the reference's calls would have consumed
// and computes a whole instrument-day at flush():
//
//   reference call                                         adaptor call
//   FeatureEngine::UpdateMidpriceAndSpread(mid, spread)    on_tick(mid, spread)            (feature_engine.cpp:130-150)
//   FeatureEngine::update_trade(price, size, dir)          on_trade(dir)                   (feature_engine.cpp:87-93)
//   FeatureProcessor::GetRawFeatureSet(snapshot) +         stage(snapshot, grid_time_ns)   (feature_processor.cpp:11-29)
//     GetProcessedFeatureSet(raw) + sink.ingest_feature_set
//   (end of the day)                                       flush(sink)  -> sink.ingest_feature_set(symbol, grid, raw, norm) per row
//
// Errors of the C ABI surface as std::runtime_error (the reference's main maps std::exception to exit code 1,
// features_to_csv.cpp:141-147).  There is no CPU fallback: without a B200 the constructor throws.
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mrp_fg.h"
#include "mrp_compat.hpp"

namespace microregime {

class GpuFeatureProcessor {
public:
    // `params` = nullptr: the reference constants (mrp_fg_default_params) with depth_levels = DEPTH_LEVELS.
    explicit GpuFeatureProcessor(std::string symbol, int device = 0, const mrp_fg_params* params = nullptr);
    ~GpuFeatureProcessor();
    GpuFeatureProcessor(const GpuFeatureProcessor&) = delete;
    GpuFeatureProcessor& operator=(const GpuFeatureProcessor&) = delete;

    // hooks: call where the pipeline calls the FeatureEngine methods of the same instrument
    void on_tick(double midprice, double spread);
    void on_trade(std::int8_t direction);

    // Replaces GetRawFeatureSet + GetProcessedFeatureSet + the sink call for one snapshot: copies what it needs
    // (the snapshot's deque pointers are borrowed and mutate on the next event, feature_snapshot.hpp:33-36).
    void stage(const FeatureInputSnapshot& snapshot, std::uint64_t grid_timestamp_ns);

    std::size_t staged_rows() const { return ts_.size(); }

    // ---- the reference's own two signatures (feature_processor.hpp:14-15): ONE row per call.
    // GetRawFeatureSet stages the snapshot (the on_tick / on_trade hooks must have been fed, as for stage()) and computes
    // this row NOW, by a GPU pass over the trailing rows that can still influence it (norm_window + EMA warm-up + 32
    // rows, with their ticks and trades).  GetProcessedFeatureSet returns the z-scores of the row the last
    // GetRawFeatureSet computed, as the reference's call order has it (dual_feature_pipeline.cpp:107-108).
    // A launch, an upload of the window and a read-back per row: milliseconds.  For code and tests written against the
    // reference interface; throughput comes from stage() + flush().  Rows stay staged, so flush() still delivers the day.
    FeatureSet GetRawFeatureSet(const FeatureInputSnapshot& snapshot);
    FeatureSet GetProcessedFeatureSet(const FeatureSet& raw_feature_set);

    // Run the GPU pass over everything staged and deliver the rows, in order, to the sink; clears the stage.
    void flush(DataReciever& sink);

    // Same, but returns the column-major results instead of calling a sink: raw/norm [19][n].
    void flush_columns(std::vector<double>& raw, std::vector<double>& norm, std::vector<std::uint64_t>& grid_ts);

    // Deliver two instruments' rows interleaved exactly as the reference does (base row t, then future row t;
    // dual_feature_pipeline.cpp:106-117).
    static void flush_pair(GpuFeatureProcessor& base, GpuFeatureProcessor& future, DataReciever& base_sink, DataReciever& future_sink);

    const std::string& symbol() const { return symbol_; }

    // The trade hook of INTEGRATION.md §3: OrderEngine::process_event only sees the FeatureEngine of the event's instrument
    // (order_engine.cpp:34,107), so the pipeline registers which staging object belongs to which engine once, and the one
    // line added next to `feature_engine->update_trade(price, size, direction)` forwards the direction.
    static void attach(const void* feature_engine, GpuFeatureProcessor* processor);
    static void detach(const GpuFeatureProcessor* processor);   // called by the destructor

private:
    void run(std::vector<double>& raw, std::vector<double>& norm);
    void clear();
    [[noreturn]] void fail(const char* what) const;

    std::string symbol_;
    mrp_fg_ctx* ctx_ = nullptr;
    int levels_ = 0;
    mrp_fg_params params_{};
    double last_norm_[MRP_FG_NUM_FEATURES] = {};   // z-scores of the row computed by the last GetRawFeatureSet
    bool last_valid_ = false;
    // snapshot axis (level-major book arrays are transposed at flush)
    std::vector<std::uint64_t> ts_, grid_;
    std::vector<double> bid_px_, ask_px_;   // row-major staging: [n][L]
    std::vector<std::int32_t> bid_sz_, ask_sz_;
    std::vector<double> buy_, sell_;
    std::vector<std::int32_t> adds_;
    std::vector<std::uint32_t> kcur_, jcur_;
    std::vector<std::string> instrument_;
    // tick / trade axes
    std::vector<double> mid_, spread_;
    std::vector<std::int8_t> tdir_;
};

// next to feature_engine->update_trade(...): forwards to the GpuFeatureProcessor attached to that engine (no-op when none is)
void mrp_trade_hook(const void* feature_engine, std::int8_t direction);

}  // namespace microregime
