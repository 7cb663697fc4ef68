/*
 * fg_oracle.h — CPU restatement of the reference feature_generation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load
 * or call it, and only as the checker.  The product (micro-regime-propogation_b200/) never
 * links or imports it and has no CPU fallback.
 *
 * Parity status: PINNED against the reference's own code.  The reference ships no golden
 * vectors or known-answer tests for this path, so the pin is the reference
 * source itself, compiled unmodified from /root/reference into oracle/_ref/ (see
 * oracle/Makefile, oracle/ref_shim.cpp) and compared with this restatement in
 * tests/test_oracle_vs_reference.py; outputs of that compiled reference on seeded inputs are
 * committed under tests/golden/ (made by tests/golden/make_golden.py).
 */
#ifndef MRP_FG_ORACLE_H
#define MRP_FG_ORACLE_H

#include "../include/mrp_fg.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Run the restated FeatureProcessor + FeatureNormalizer over every stream of a HOST view.
 *   raw, norm        double [19][n_rows]   (required)
 *   bid_dir, ask_dir int8   [L][n_rows]    (may be NULL)
 *   counts           int32  [MRP_FG_NUM_COUNTS][n_rows] (may be NULL)
 * nthreads > 1 parallelises the per-row window loops with OpenMP (results identical: every
 * row's loops run in the reference order; the EMA / normaliser chains stay sequential).
 * Returns 0, or 1 on bad arguments. */
int mrp_oracle_run(const mrp_fg_params* p, const mrp_fg_soa_view* host_view, double* raw,
                   double* norm, int8_t* bid_dir, int8_t* ask_dir, int32_t* counts, int nthreads);

/* vol-of-vol EXTENSION — PARITY UNPINNED.  The reference has no such feature, so there is nothing to
 * pin this function against: it restates the DEFINITION in include/mrp_fg.h (mrp_fg_vol_of_vol) with direct O(window) loops
 * in extended precision, and is the only checker the CUDA extension has.
 *   ewm        double [n_rows]   the ewm_volatility column (raw feature MRP_F_EWM_VOLATILITY)
 *   row_offsets int64 [n_streams + 1]
 *   raw, norm  double [n_horizons][n_rows]
 * Returns 0, or 1 on bad arguments. */
int mrp_oracle_vol_of_vol(const double* ewm, const int64_t* row_offsets, int n_streams, const int32_t* horizons, int n_horizons,
                          int norm_window, double* raw, double* norm);

#ifdef __cplusplus
}
#endif
#endif
