// binary_sink.hpp — a column-major binary sink beside CsvWriter.
//
// The reference's only sink is CSV (features_to_csv.cpp:16-111): ~400 bytes of text per row and instrument, which the Python
// stage then parses back (regime_classifier/python/classifier.py:15-18 reads <base>_norm.csv with pandas).  Once the features
// come off a GPU that text is ~75x the cost of computing them.  BinarySink keeps the reference's directory and file stem and
// writes, per instrument,
//     <base_filename>_raw.npy, <base_filename>_norm.npy     float64 [19][n]  (C order: row c = feature c in CSV column order)
//     <base_filename>_timestamp_ns.npy                      uint64  [n]      (the grid time the CSV's first column carries)
//     <base_filename>_columns.json                          {"instrument": ..., "columns": [...19 names...], "rows": n}
// in NumPy's NPY 1.0 format, so the consumer maps them without parsing: np.load(path, mmap_mode="r").
// It is a DataReciever (same virtual as CsvWriter: rows are buffered and written at close) and has the bulk column path the GPU
// adaptor uses (one write per matrix).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "mrp_compat.hpp"

namespace microregime {

class BinarySink : public DataReciever {
public:
    BinarySink(const std::string& base_filename, std::uint64_t snapshot_interval_ns, const std::string& date,
               const std::string& project_root = ".", int window_size = 30000, int rolling_window = 1500);
    ~BinarySink() override;   // writes what ingest_feature_set buffered (unless write_columns / close already did)

    void ingest_feature_set(const std::string& symbol, std::uint64_t timestamp_ns, const FeatureSet& raw_features,
                            const FeatureSet& normalized) override;

    // n rows at once: raw / norm column-major [19][n] exactly as the C ABI delivers them.
    void write_columns(const std::string& instrument, const std::uint64_t* grid_ts, const double* raw, const double* norm, std::int64_t n);

    void close();
    const std::string& directory() const { return dir_; }

private:
    std::string dir_, stem_, instrument_;
    std::vector<std::uint64_t> ts_;
    std::vector<double> raw_rows_, norm_rows_;   // row-major staging of the per-row path
    bool written_ = false;
};

}  // namespace microregime
