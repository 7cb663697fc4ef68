// binary_sink.cpp — see binary_sink.hpp.
#include "binary_sink.hpp"

#include <cstdio>
#include <cstring>
#include <filesystem>
#include <stdexcept>

#include "csv_writer.hpp"

namespace microregime {

namespace {
const char* const kNames[19] = {"midprice", "log_spread", "log_return", "ewm_volatility", "realized_variance", "directional_volatility",
                                "spread_volatility", "ofi", "signed_volume_pressure", "order_arrival_rate", "depth_imbalance", "market_depth",
                                "lob_slope", "price_gap", "tick_direction_entropy", "reversal_rate", "aggressor_bias", "shannon_entropy",
                                "liquidity_stress"};

// NPY 1.0: magic, version, little-endian header length, a Python dict literal padded with spaces to a 64-byte boundary, '\n'
void write_npy(const std::string& path, const char* descr, const std::string& shape, const void* data, std::size_t bytes) {
    std::string dict = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (" + shape + "), }";
    std::size_t total = 10 + dict.size() + 1;
    const std::size_t pad = (64 - total % 64) % 64;
    dict.append(pad, ' ');
    dict.push_back('\n');
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("BinarySink: cannot open " + path);
    const unsigned char head[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const std::uint16_t hl = (std::uint16_t)dict.size();
    const unsigned char len[2] = {(unsigned char)(hl & 0xff), (unsigned char)(hl >> 8)};
    bool ok = std::fwrite(head, 1, 8, f) == 8 && std::fwrite(len, 1, 2, f) == 2 && std::fwrite(dict.data(), 1, dict.size(), f) == dict.size();
    if (ok && bytes) ok = std::fwrite(data, 1, bytes, f) == bytes;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("BinarySink: short write to " + path);
}
}  // namespace

BinarySink::BinarySink(const std::string& base_filename, std::uint64_t snapshot_interval_ns, const std::string& date,
                       const std::string& project_root, int window_size, int rolling_window)
    : stem_(base_filename) {
    namespace fs = std::filesystem;
    // the reference's directory rule (features_to_csv.cpp:18-30), shared with CsvWriter
    const fs::path dir = fs::path(project_root) / "data" / CsvWriter::directory_name(snapshot_interval_ns, window_size, rolling_window) / date;
    fs::create_directories(dir);
    dir_ = dir.string();
}

BinarySink::~BinarySink() {
    try {
        close();
    } catch (...) {   // a destructor must not throw; close() explicitly to see errors
    }
}

void BinarySink::ingest_feature_set(const std::string& symbol, std::uint64_t timestamp_ns, const FeatureSet& raw_features,
                                    const FeatureSet& normalized) {
    if (instrument_.empty()) instrument_ = symbol;
    ts_.push_back(timestamp_ns);
    for (int c = 0; c < 19; ++c) {
        raw_rows_.push_back(raw_features.*(mrp_host::kFeatureMembers[c]));
        norm_rows_.push_back(normalized.*(mrp_host::kFeatureMembers[c]));
    }
}

void BinarySink::write_columns(const std::string& instrument, const std::uint64_t* grid_ts, const double* raw, const double* norm, std::int64_t n) {
    const std::string shape = "19, " + std::to_string(n);
    write_npy(dir_ + "/" + stem_ + "_raw.npy", "<f8", shape, raw, (std::size_t)n * 19 * 8);
    write_npy(dir_ + "/" + stem_ + "_norm.npy", "<f8", shape, norm, (std::size_t)n * 19 * 8);
    write_npy(dir_ + "/" + stem_ + "_timestamp_ns.npy", "<u8", std::to_string(n) + ",", grid_ts, (std::size_t)n * 8);
    std::string js = "{\"instrument\": \"" + instrument + "\", \"rows\": " + std::to_string(n) + ", \"layout\": \"[feature][row], float64\", \"columns\": [";
    for (int c = 0; c < 19; ++c) js += std::string(c ? ", \"" : "\"") + kNames[c] + "\"";
    js += "]}\n";
    std::FILE* f = std::fopen((dir_ + "/" + stem_ + "_columns.json").c_str(), "wb");
    if (!f || std::fwrite(js.data(), 1, js.size(), f) != js.size() || std::fclose(f) != 0) throw std::runtime_error("BinarySink: cannot write the column index");
    written_ = true;
}

void BinarySink::close() {
    if (written_) return;
    const std::int64_t n = (std::int64_t)ts_.size();
    std::vector<double> raw((std::size_t)n * 19), norm((std::size_t)n * 19);   // the per-row path: transpose to [19][n]
    for (std::int64_t r = 0; r < n; ++r)
        for (int c = 0; c < 19; ++c) {
            raw[(std::size_t)c * n + r] = raw_rows_[(std::size_t)r * 19 + c];
            norm[(std::size_t)c * n + r] = norm_rows_[(std::size_t)r * 19 + c];
        }
    write_columns(instrument_, ts_.data(), raw.data(), norm.data(), n);
}

}  // namespace microregime
