// host_capi.cpp — small extern "C" surface over the C++ adaptor so that the Python tests and the validation harness
// can drive it through ctypes.  Not part of the drop-in boundary (that is include/mrp_fg.h + the C++ classes).
#include <cstring>
#include <string>

#include "binary_sink.hpp"
#include "csv_writer.hpp"
#include "gpu_feature_processor.hpp"

namespace {
thread_local std::string g_err;
}

extern "C" {

const char* mrp_host_last_error(void) { return g_err.c_str(); }

// Write n rows through CsvWriter.  bulk = 0: one ingest_feature_set call per row (the reference's path);
// bulk = 1: the threaded column path.  raw / norm: column-major [19][n].
int mrp_host_write_csv(const char* project_root, const char* base_filename, uint64_t interval_ns, const char* date, const char* instrument,
                       int64_t n, const uint64_t* grid_ts, const uint64_t* row_ts, const double* raw, const double* norm, int bulk) {
    try {
        microregime::CsvWriter w(base_filename, interval_ns, date, project_root);
        if (bulk) {
            w.write_columns(instrument, grid_ts, raw, norm, n, 0);
        } else {
            microregime::FeatureSet a, b;
            for (int64_t r = 0; r < n; ++r) {
                a.timestamp_ns = b.timestamp_ns = row_ts ? row_ts[r] : grid_ts[r];
                a.instrument = b.instrument = instrument;
                for (int c = 0; c < 19; ++c) {
                    a.*(mrp_host::kFeatureMembers[c]) = raw[(size_t)c * n + r];
                    b.*(mrp_host::kFeatureMembers[c]) = norm[(size_t)c * n + r];
                }
                w.ingest_feature_set(instrument, grid_ts[r], a, b);
            }
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// The same rows through BinarySink (NPY files in the reference's directory).  bulk = 0: one ingest_feature_set call per row.
int mrp_host_write_binary(const char* project_root, const char* base_filename, uint64_t interval_ns, const char* date, const char* instrument,
                          int64_t n, const uint64_t* grid_ts, const uint64_t* row_ts, const double* raw, const double* norm, int bulk) {
    try {
        microregime::BinarySink w(base_filename, interval_ns, date, project_root);
        if (bulk) {
            w.write_columns(instrument, grid_ts, raw, norm, n);
        } else {
            microregime::FeatureSet a, b;
            for (int64_t r = 0; r < n; ++r) {
                a.timestamp_ns = b.timestamp_ns = row_ts ? row_ts[r] : grid_ts[r];
                a.instrument = b.instrument = instrument;
                for (int c = 0; c < 19; ++c) {
                    a.*(mrp_host::kFeatureMembers[c]) = raw[(size_t)c * n + r];
                    b.*(mrp_host::kFeatureMembers[c]) = norm[(size_t)c * n + r];
                }
                w.ingest_feature_set(instrument, grid_ts[r], a, b);
            }
            w.close();
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// Drive GpuFeatureProcessor exactly as a pipeline would (ticks, trades, snapshots in event order) over a staged
// host batch of ONE stream with L == DEPTH_LEVELS, and write the reference-layout CSVs.  Returns 0, or 1 with
// mrp_host_last_error() set (e.g. no GPU: the adaptor has no CPU fallback).
int mrp_host_features_to_csv(const mrp_fg_soa_view* v, const char* project_root, const char* base_filename, const char* date,
                             const char* instrument, uint64_t first_grid_ns, uint64_t interval_ns, int device, int bulk) {
    try {
        if (!v || v->n_streams != 1 || v->memory != MRP_MEM_HOST) throw std::runtime_error("expects a single-stream HOST view");
        const int64_t n = v->row_offsets[1];
        constexpr int L = (int)mrp_host::kDepthLevels;
        microregime::GpuFeatureProcessor proc(instrument, device);
        int64_t k = 0, j = 0;
        FeatureInputSnapshot s;
        s.instrument = instrument;
        for (int64_t r = 0; r < n; ++r) {
            for (; k < (int64_t)v->tick_cursor[r]; ++k) proc.on_tick(v->mid[k], v->spread[k]);
            for (; j < (int64_t)v->trade_cursor[r]; ++j) proc.on_trade(v->trade_dir[j]);
            s.timestamp_ns = v->ts_ns[r];
            for (int i = 0; i < L; ++i) {
                s.bid_prices[i] = v->bid_px[(size_t)i * n + r];
                s.ask_prices[i] = v->ask_px[(size_t)i * n + r];
                s.bid_sizes[i] = v->bid_sz[(size_t)i * n + r];
                s.ask_sizes[i] = v->ask_sz[(size_t)i * n + r];
            }
            s.best_bid_price = s.bid_prices[0];
            s.best_ask_price = s.ask_prices[0];
            s.rolling_buy_volume = v->buy_vol[r];
            s.rolling_sell_volume = v->sell_vol[r];
            s.adds_since_last_snapshot = v->adds[r];
            proc.stage(s, first_grid_ns + (uint64_t)r * interval_ns);
        }
        microregime::CsvWriter w(base_filename, interval_ns, date, project_root);
        if (bulk) {
            std::vector<double> raw, norm;
            std::vector<uint64_t> grid;
            proc.flush_columns(raw, norm, grid);
            w.write_columns(instrument, grid.data(), raw.data(), norm.data(), n, 0);
        } else {
            proc.flush(w);
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// The reference's per-row call sequence (dual_feature_pipeline.cpp:106-108) over a staged host batch of ONE stream:
// rows before `first_row` are only staged, every later row goes through GetRawFeatureSet + GetProcessedFeatureSet.
// raw_out / norm_out: [19][n - first_row].  With params == NULL the reference constants are used.
int mrp_host_rows_by_reference_calls(const mrp_fg_soa_view* v, const mrp_fg_params* params, int64_t first_row, double* raw_out,
                                     double* norm_out, int device) {
    try {
        if (!v || v->n_streams != 1 || v->memory != MRP_MEM_HOST) throw std::runtime_error("expects a single-stream HOST view");
        const int64_t n = v->row_offsets[1];
        if (first_row < 0 || first_row > n) throw std::runtime_error("first_row outside the stream");
        const int64_t m = n - first_row;
        constexpr int L = (int)mrp_host::kDepthLevels;
        microregime::GpuFeatureProcessor proc("SYN", device, params);
        int64_t k = 0, j = 0;
        FeatureInputSnapshot s;
        s.instrument = "SYN";
        for (int64_t r = 0; r < n; ++r) {
            for (; k < (int64_t)v->tick_cursor[r]; ++k) proc.on_tick(v->mid[k], v->spread[k]);
            for (; j < (int64_t)v->trade_cursor[r]; ++j) proc.on_trade(v->trade_dir[j]);
            s.timestamp_ns = v->ts_ns[r];
            for (int i = 0; i < L; ++i) {
                s.bid_prices[i] = v->bid_px[(size_t)i * n + r];
                s.ask_prices[i] = v->ask_px[(size_t)i * n + r];
                s.bid_sizes[i] = v->bid_sz[(size_t)i * n + r];
                s.ask_sizes[i] = v->ask_sz[(size_t)i * n + r];
            }
            s.best_bid_price = s.bid_prices[0];
            s.best_ask_price = s.ask_prices[0];
            s.rolling_buy_volume = v->buy_vol[r];
            s.rolling_sell_volume = v->sell_vol[r];
            s.adds_since_last_snapshot = v->adds[r];
            if (r < first_row) {
                proc.stage(s, s.timestamp_ns);
                continue;
            }
            const microregime::FeatureSet raw = proc.GetRawFeatureSet(s);
            const microregime::FeatureSet norm = proc.GetProcessedFeatureSet(raw);
            if (raw.timestamp_ns != s.timestamp_ns || norm.timestamp_ns != s.timestamp_ns || norm.instrument != s.instrument)
                throw std::runtime_error("timestamp / instrument not carried through");
            for (int c = 0; c < MRP_FG_NUM_FEATURES; ++c) {
                raw_out[(size_t)c * m + (r - first_row)] = raw.*(mrp_host::kFeatureMembers[c]);
                norm_out[(size_t)c * m + (r - first_row)] = norm.*(mrp_host::kFeatureMembers[c]);
            }
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int mrp_host_depth_levels(void) { return (int)mrp_host::kDepthLevels; }

}  // extern "C"
