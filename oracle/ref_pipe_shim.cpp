// ref_pipe_shim.cpp — the reference's own features_to_csv program behind a C entry point (TEST INFRASTRUCTURE ONLY, see fg_oracle.h).
//
// Compiled twice by oracle/Makefile:
//   _ref/libmrp_refpipe.so   the UNMODIFIED reference: features_to_csv.cpp (main renamed) + dual_feature_pipeline.cpp + event_parser.cpp +
//                            order_engine.cpp + feature_engine.cpp + order_book.cpp + feature_processor.cpp + feature_normalizer.cpp,
//                            where they lie under /root/reference, with the synthetic MBO reader (mbo_synth_reader.cpp);
//   _ref/libmrp_gpupipe.so   the same sources with the INTEGRATION.md §3 patch applied to temporary copies (sed) and the product's
//                            C++20 adaptor (host/gpu_feature_processor.cpp, compiled against the reference's real headers) -> C ABI -> CUDA.
// Both write the reference's four CSV files under $MRP_REF_PROJECT_ROOT; tests/test_gpu_pipeline_config0.py diffs them.
#include <cstdint>
#include <vector>
#include <cmath>
#include <iostream>
#include <optional>
#include <stdexcept>

#define main mrp_ref_features_to_csv_main
#include "features_to_csv.cpp"   // -I$(REF)/feature_generation/src/data
#undef main

extern "C" int mrp_pipe_run(const char* date, const char* base_asset, const char* future, uint64_t interval_ns) {
    std::streambuf* saved = std::cout.rdbuf(nullptr);   // the pipeline and the normaliser print from inside the loop
    int rc = 0;
    try {
        microregime::run_feature_extraction(date, base_asset, future, interval_ns);
    } catch (const std::exception& e) {
        std::cerr << "mrp_pipe_run: " << e.what() << std::endl;
        rc = 1;   // what the reference's main returns (features_to_csv.cpp:141-147)
    } catch (...) {
        rc = 2;
    }
    std::cout.rdbuf(saved);
    std::cout.clear();
    return rc;
}
