#include "csv_writer.hpp"

#include <memory>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <thread>

namespace microregime {

std::string CsvWriter::directory_name(std::uint64_t snapshot_interval_ns, int window_size, int rolling_window) {
    // std::to_string(double) prints six decimals: 0.500000 (features_to_csv.cpp:20)
    const std::string seconds = std::to_string(static_cast<double>(snapshot_interval_ns) / 1000000000);
    return "output_Snapshot" + seconds + "_Window" + std::to_string(window_size) + "_Events" + std::to_string(rolling_window);
}

const char* CsvWriter::header() {
    return "timestamp_ns,instrument,midprice,log_spread,log_return,ewm_volatility,realized_variance,directional_volatility,"
           "spread_volatility,ofi,signed_volume_pressure,order_arrival_rate,depth_imbalance,market_depth,lob_slope,price_gap,"
           "tick_direction_entropy,reversal_rate,aggressor_bias,shannon_entropy,liquidity_stress\n";
}

CsvWriter::CsvWriter(const std::string& base_filename, std::uint64_t snapshot_interval_ns, const std::string& date,
                     const std::string& project_root, int window_size, int rolling_window) {
    const std::filesystem::path dir =
        std::filesystem::path(project_root) / "data" / directory_name(snapshot_interval_ns, window_size, rolling_window) / date;
    std::filesystem::create_directories(dir);
    dir_ = dir.string();
    raw_csv_.open(dir / (base_filename + "_raw.csv"), std::ios::out);
    norm_csv_.open(dir / (base_filename + "_norm.csv"), std::ios::out);
    if (!raw_csv_.is_open() || !norm_csv_.is_open()) throw std::runtime_error("CsvWriter: cannot open output files in " + dir_);
    raw_csv_ << header();
    norm_csv_ << header();
}

CsvWriter::~CsvWriter() = default;

int CsvWriter::format_value(char* out, double v) {
    // operator<<(double) with std::scientific and precision 15 == printf("%.15e"); non-finite values print as
    // nan / -nan / inf / -inf (glibc)
    if (std::isnan(v)) {
        const char* s = std::signbit(v) ? "-nan" : "nan";
        const int n = (int)std::strlen(s);
        std::memcpy(out, s, n);
        return n;
    }
    if (std::isinf(v)) {
        const char* s = v < 0 ? "-inf" : "inf";
        const int n = (int)std::strlen(s);
        std::memcpy(out, s, n);
        return n;
    }
    const auto r = std::to_chars(out, out + 32, v, std::chars_format::scientific, 15);
    return (int)(r.ptr - out);
}

namespace {

std::size_t format_row(char* out, std::uint64_t ts, const std::string& instrument, const double* cols, std::int64_t n, std::int64_t r) {
    char* p = out;
    p = std::to_chars(p, p + 24, ts).ptr;
    *p++ = ',';
    std::memcpy(p, instrument.data(), instrument.size());
    p += instrument.size();
    for (int c = 0; c < 19; ++c) {
        *p++ = ',';
        p += CsvWriter::format_value(p, cols[(std::size_t)c * n + r]);
    }
    *p++ = '\n';
    return (std::size_t)(p - out);
}

}  // namespace

void CsvWriter::ingest_feature_set(const std::string&, std::uint64_t timestamp_ns, const FeatureSet& raw_features, const FeatureSet& normalized) {
    char buf[1024];
    double v[19];
    for (const FeatureSet* f : {&raw_features, &normalized}) {
        for (int c = 0; c < 19; ++c) v[c] = f->*(mrp_host::kFeatureMembers[c]);
        std::string inst = f->instrument.size() < 256 ? f->instrument : f->instrument.substr(0, 255);
        const std::size_t len = format_row(buf, timestamp_ns, inst, v, 1, 0);
        (f == &raw_features ? raw_csv_ : norm_csv_).write(buf, (std::streamsize)len);
    }
}

void CsvWriter::write_columns(const std::string& instrument, const std::uint64_t* grid_ts, const double* raw, const double* norm,
                              std::int64_t n, int threads) {
    if (n <= 0) return;
    constexpr std::int64_t block = 1 << 13;   // rows a thread formats per round: 4.5 MB of text, stays near its cache
    int nt = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    nt = (int)std::min<std::int64_t>(nt, (n + block - 1) / block);
    const std::size_t row_cap = 24 + instrument.size() + 19 * 26 + 2;
    // two sets of per-thread buffers: one is being written to the file while the other is being formatted
    std::unique_ptr<char[]> bufs[2];
    std::vector<std::size_t> lens[2];
    for (int k = 0; k < 2; ++k) {
        bufs[k].reset(new char[(std::size_t)nt * block * row_cap]);   // uninitialised on purpose
        lens[k].assign((std::size_t)nt, 0);
    }
    std::thread writer;
    int k = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const double* cols = pass == 0 ? raw : norm;
        std::ofstream& out = pass == 0 ? raw_csv_ : norm_csv_;
        for (std::int64_t b0 = 0; b0 < n; b0 += block * nt, k ^= 1) {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; ++t) {
                const std::int64_t lo = std::min(n, b0 + (std::int64_t)t * block), hi = std::min(n, lo + block);
                lens[k][(std::size_t)t] = 0;
                if (lo >= hi) continue;
                pool.emplace_back([&, t, lo, hi, k] {
                    char* buf = bufs[k].get() + (std::size_t)t * block * row_cap;
                    std::size_t len = 0;
                    for (std::int64_t r = lo; r < hi; ++r) len += format_row(buf + len, grid_ts[r], instrument, cols, n, r);
                    lens[k][(std::size_t)t] = len;
                });
            }
            for (auto& th : pool) th.join();
            if (writer.joinable()) writer.join();   // the previous round's text is on its way: file order is round order
            writer = std::thread([&, k, nt, block, row_cap, o = &out] {
                for (int t = 0; t < nt; ++t)
                    if (lens[k][(std::size_t)t]) o->write(bufs[k].get() + (std::size_t)t * block * row_cap, (std::streamsize)lens[k][(std::size_t)t]);
            });
        }
    }
    if (writer.joinable()) writer.join();
}

}  // namespace microregime
