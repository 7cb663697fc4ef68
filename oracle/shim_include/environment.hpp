// Stand-in for the reference's git-ignored environment.hpp (used at feature_generation/src/data/features_to_csv.cpp:11,24):
// the checker decides where the reference's CsvWriter writes.
#pragma once
#include <cstdlib>
#include <filesystem>
namespace microregime {
inline std::filesystem::path get_project_root() {
    const char* e = std::getenv("MRP_REF_PROJECT_ROOT");
    return e ? std::filesystem::path(e) : std::filesystem::current_path();
}
}  // namespace microregime
