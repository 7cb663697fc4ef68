// csv_writer.hpp — the reference's CSV sink, byte for byte.
//
// Mirrors microregime::CsvWriter (feature_generation/src/data/features_to_csv.cpp:16-111):
//   directory  <root>/data/output_Snapshot<std::to_string(double seconds)>_Window<WINDOW_SIZE>_Events<ROLLING_WINDOW>/<date>/
//   files      <base_filename>_raw.csv, <base_filename>_norm.csv                                     (:24-30)
//   header     timestamp_ns,instrument,midprice,...,liquidity_stress                                 (:55-81)
//   rows       integer grid timestamp, instrument, then 19 values as std::scientific with 15 digits  (:83-110)
// plus a bulk path that formats whole column blocks on several threads (the per-row virtual call is the
// end-to-end bottleneck once the features come off a GPU); both paths produce identical bytes.
#pragma once

#include <cstdint>
#include <fstream>
#include <string>
#include <vector>

#include "mrp_compat.hpp"

namespace microregime {

class CsvWriter : public DataReciever {
public:
    CsvWriter(const std::string& base_filename, std::uint64_t snapshot_interval_ns, const std::string& date,
              const std::string& project_root = ".", int window_size = 30000, int rolling_window = 1500);
    ~CsvWriter() override;

    void ingest_feature_set(const std::string& symbol, std::uint64_t timestamp_ns, const FeatureSet& raw_features,
                            const FeatureSet& normalized) override;

    // n rows at once: raw / norm are column-major [19][n]; one instrument string for all rows.
    void write_columns(const std::string& instrument, const std::uint64_t* grid_ts, const double* raw, const double* norm,
                       std::int64_t n, int threads = 0);

    const std::string& directory() const { return dir_; }
    static std::string directory_name(std::uint64_t snapshot_interval_ns, int window_size, int rolling_window);
    static const char* header();
    // One value as the reference's stream prints it (e.g. 5.646500000000000e+03, nan, -inf); returns the length.
    static int format_value(char* out, double v);

private:
    std::string dir_;
    std::ofstream raw_csv_, norm_csv_;
};

}  // namespace microregime
