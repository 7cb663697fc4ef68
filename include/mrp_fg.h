/*
 * mrp_fg.h — C ABI of the B200 feature-generation engine.
 *
 * This is the drop-in boundary for ONE stage of JackDamato/micro-regime-propogation:
 * the per-snapshot feature engine (`microregime::FeatureProcessor` + `FeatureNormalizer`).
 * Nothing like this ABI exists in the reference (it is a single C++ process); each entry
 * point below names the reference interface it stands in for, as
 *   ref: <file>:<line>   (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C, no exceptions across the boundary: every call returns an int status
 *     (MRP_FG_OK == 0) and `mrp_fg_last_error()` gives a thread-local message
 *     (ref: the reference signals errors with C++ exceptions that `main` maps to exit
 *      code 1, feature_generation/src/data/features_to_csv.cpp:141-147);
 *   - one context per (host thread, GPU); contexts share no state
 *     (ref: one FeatureProcessor per instrument, feature_generation/include/dual_feature_pipeline.hpp:46-47);
 *   - all kernels are enqueued on the caller's CUDA stream (passed as void*, a cudaStream_t);
 *   - there is NO CPU fallback: if the CUDA device or the kernels are unavailable the
 *     call fails with MRP_FG_ERR_CUDA.
 *
 * Data model (structure of arrays; "stream" = one instrument-day = one FeatureProcessor instance)
 *   A batch holds S independent streams stored back to back on three axes:
 *     snapshot axis  N rows    one row per FeatureInputSnapshot
 *                              (ref: data_ingestion/include/feature_snapshot.hpp:12-44)
 *     tick axis      K ticks   the 50 ms (midprice, spread) samples pushed by
 *                              FeatureEngine::UpdateMidpriceAndSpread
 *                              (ref: data_ingestion/src/core/feature_engine.cpp:130-150)
 *     trade axis     T trades  the trade directions pushed by FeatureEngine::update_trade
 *                              (ref: data_ingestion/src/core/feature_engine.cpp:87-93)
 *   Row t of a stream carries two cursors: tick_cursor[t] / trade_cursor[t] = how many ticks /
 *   trades of THIS stream had been pushed when the snapshot was taken.  The reference's
 *   rolling deques are then exactly
 *     rolling_midprices / rolling_spreads = ticks [max(0,k-mid_window), k)        (cap 1800)
 *     rolling_tick_directions             = tick dirs [max(0,k-dir_window), k)    (cap 1500)
 *     rolling_trade_directions            = trades [max(0,j-dir_window), j)       (cap 1500)
 *   Per-level book arrays are level-major: element (level i, row r) at [i * n_rows + r].
 *   Limits: a stream holds at most 2^32 - 65536 ticks and trades (uint32 cursors; longer streams are split by time),
 *   mid_window and dir_window <= about 4000 (on-chip ring), depth_levels in [2, 1024].
 *   The per-level depth-change directions (ref: data_ingestion/src/core/order_book.cpp:208-239,
 *   both branches reduce to sign(new_size - old_size)) are derived on the device from the
 *   sizes of rows t and t-1 (row 0 compares with an all-zero book, order_book.cpp:115,168).
 */
#ifndef MRP_FG_H
#define MRP_FG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRP_FG_ABI_VERSION 8   /* 3: + mrp_fg_download_rows; 4: + mrp_fg_vol_of_vol, mrp_fg_download_vol_of_vol;
                                  5: + mrp_fg_norm_moments, mrp_fg_mean_scale, mrp_fg_standardize;
                                  6: + mrp_fg_halo_rows(_for), mrp_fg_moments_rows, mrp_fg_norm_moments_rows;
                                  7: + mrp_fg_halo_query, mrp_fg_halo_export, mrp_fg_halo_import, mrp_fg_build_state,
                                     mrp_fg_poison_state, mrp_fg_poison_apply;
                                  8: + mrp_fg_plan_shape, mrp_fg_plan_host_chunks; mrp_fg_process_host takes raw == NULL or norm == NULL */

/* status codes */
#define MRP_FG_OK 0
#define MRP_FG_ERR_INVALID 1 /* bad argument / inconsistent view            */
#define MRP_FG_ERR_CUDA 2    /* CUDA runtime or kernel failure, or no GPU   */
#define MRP_FG_ERR_STATE 3   /* call order (run before bind, ...)           */
#define MRP_FG_ERR_NOMEM 4

/* Output column order == FeatureSet declaration order == CSV column order
 * (ref: feature_generation/include/feature_set.hpp:15-43,
 *       feature_generation/src/data/features_to_csv.cpp:62-80). */
enum mrp_fg_feature {
    MRP_F_MIDPRICE = 0,
    MRP_F_LOG_SPREAD,
    MRP_F_LOG_RETURN,
    MRP_F_EWM_VOLATILITY,
    MRP_F_REALIZED_VARIANCE,
    MRP_F_DIRECTIONAL_VOLATILITY,
    MRP_F_SPREAD_VOLATILITY,
    MRP_F_OFI,
    MRP_F_SIGNED_VOLUME_PRESSURE,
    MRP_F_ORDER_ARRIVAL_RATE,
    MRP_F_DEPTH_IMBALANCE,
    MRP_F_MARKET_DEPTH,
    MRP_F_LOB_SLOPE,
    MRP_F_PRICE_GAP,
    MRP_F_TICK_DIRECTION_ENTROPY,
    MRP_F_REVERSAL_RATE,
    MRP_F_AGGRESSOR_BIAS,
    MRP_F_SHANNON_ENTROPY,
    MRP_F_LIQUIDITY_STRESS,
    MRP_FG_NUM_FEATURES /* 19 */
};

/* Optional integer window-count columns (bit-exact against the oracle). */
enum mrp_fg_count {
    MRP_C_RET_VALID = 0, /* valid return pairs in the mid window   (feature_processor.cpp:55-60)  */
    MRP_C_RET_UP,        /* of those, ret > 0                       (feature_processor.cpp:87-89)  */
    MRP_C_TICK_UP,       /* tick dirs > 0 in the dir window         (feature_processor.cpp:217-221)*/
    MRP_C_TICK_DOWN,
    MRP_C_TICK_TOTAL,
    MRP_C_TRADE_POS,     /* trade dirs > 0                          (feature_processor.cpp:255-259)*/
    MRP_C_TRADE_NEG,
    MRP_C_TRADE_REVERSALS, /*                                       (feature_processor.cpp:233-238)*/
    MRP_C_TRADE_TOTAL,
    MRP_C_TRADE_DIR_SUM,   /*                                       (feature_processor.cpp:242-245)*/
    MRP_C_NORM_N,          /* normaliser window size n              (feature_normalizer.cpp:71)    */
    MRP_FG_NUM_COUNTS /* 11 */
};

/* Every compile-time constant of the reference path, as a run-time field.
 * Defaults (mrp_fg_default_params) are the reference values. */
typedef struct mrp_fg_params {
    int32_t depth_levels;    /* DEPTH_LEVELS = 10       ref: data_ingestion/include/common_constants.hpp:8  */
    int32_t mid_window;      /* 1800 mids/spreads kept  ref: data_ingestion/src/core/feature_engine.cpp:146 */
    int32_t dir_window;      /* ROLLING_WINDOW = 1500   ref: common_constants.hpp:9, feature_engine.cpp:90,143 */
    int32_t spread_divisor;  /* ROLLING_WINDOW = 1500   ref: feature_processor.cpp:101,106 (not the deque length) */
    int32_t ewm_span;        /* ROLLING_WINDOW: alpha = 2/(span+1)   ref: feature_processor.cpp:64 */
    int32_t norm_window;     /* WINDOW_SIZE = 30000     ref: common_constants.hpp:10, feature_normalizer.cpp:48 */
    int32_t return_lag;      /* 10e9 / SNAPSHOT_INTERVAL_NS = 20 -> reads the row 21 back
                                ref: feature_processor.cpp:37, feature_normalizer.hpp:19-24 */
    int32_t stress_levels;   /* LEVELS_TO_USE = 5       ref: feature_processor.cpp:274 */
    double ofi_depth_decay;  /* 0.5                     ref: feature_processor.cpp:114 */
    double ofi_alpha;        /* 0.2                     ref: feature_processor.cpp:115 */
    double ofi_min_total_vol;/* 1e-6                    ref: feature_processor.cpp:116 */
    double stress_min_size;  /* 5.0                     ref: feature_processor.cpp:275 */
    double stress_distance_decay; /* 10.0               ref: feature_processor.cpp:276 */
    double stress_alpha;     /* 0.1                     ref: feature_processor.cpp:277 */
    int32_t norm_mode;       /* MRP_NORM_COMPENSATED (default) or MRP_NORM_FAITHFUL */
    int32_t emit_dirs;       /* also write the int8 per-level depth-change directions */
    int32_t emit_counts;     /* also write the int32 window-count columns             */
    int32_t reserved;
} mrp_fg_params;

/* norm_mode */
#define MRP_NORM_COMPENSATED 0 /* exact windowed sums (block partials + in-tile scans); more accurate than the reference */
#define MRP_NORM_FAITHFUL 1    /* re-enacts the reference's add-then-evict running sums in its own
                                  rounding order (feature_normalizer.cpp:37-51): serial per column, slow, for validation */

/* A batch of S streams, structure of arrays.  `memory` says where the pointers live. */
#define MRP_MEM_DEVICE 0
#define MRP_MEM_HOST 1

typedef struct mrp_fg_soa_view {
    int32_t n_streams;           /* S >= 1 */
    int32_t memory;              /* MRP_MEM_DEVICE or MRP_MEM_HOST for the array pointers below */
    /* offsets are ALWAYS host arrays of S+1 entries; stream s owns [off[s], off[s+1]) */
    const int64_t* row_offsets;
    const int64_t* tick_offsets;
    const int64_t* trade_offsets;
    /* snapshot axis, n_rows = row_offsets[S] */
    const uint64_t* ts_ns;        /* [n_rows]     FeatureInputSnapshot::timestamp_ns              */
    const double* bid_px;         /* [L][n_rows]  bid_prices  (level 0 == best_bid_price)         */
    const double* ask_px;         /* [L][n_rows]  ask_prices  (level 0 == best_ask_price)         */
    const int32_t* bid_sz;        /* [L][n_rows]  bid_sizes                                       */
    const int32_t* ask_sz;        /* [L][n_rows]  ask_sizes                                       */
    const double* buy_vol;        /* [n_rows]     rolling_buy_volume       } NULL: derived on the device   */
    const double* sell_vol;       /* [n_rows]     rolling_sell_volume      } by mrp_fg_build_state         */
    const int32_t* adds;          /* [n_rows]     adds_since_last_snapshot }                               */
    const uint32_t* tick_cursor;  /* [n_rows]     ticks of this stream pushed before the snapshot */
    const uint32_t* trade_cursor; /* [n_rows]     trades of this stream pushed before the snapshot*/
    /* tick axis, n_ticks = tick_offsets[S] */
    const double* mid;            /* [n_ticks]    OrderBookManager::GetMidPrice() samples */
    const double* spread;         /* [n_ticks]    OrderBookManager::GetSpread() samples   */
    /* trade axis, n_trades = trade_offsets[S] */
    const int8_t* trade_dir;      /* [n_trades]   +1 / -1 (order_engine.cpp:106)          */
} mrp_fg_soa_view;

typedef struct mrp_fg_ctx mrp_fg_ctx;

/* Fill `p` with the reference constants (common_constants.hpp:8-11, feature_processor.cpp:64,114-116,274-277). */
int mrp_fg_default_params(mrp_fg_params* p);

/* Replaces: constructing a FeatureProcessor (feature_processor.hpp:11-12).  `device` = CUDA ordinal. */
int mrp_fg_create(const mrp_fg_params* p, int device, mrp_fg_ctx** out);
int mrp_fg_destroy(mrp_fg_ctx* ctx);

/* Bind caller-owned DEVICE arrays (view->memory must be MRP_MEM_DEVICE); nothing is copied except
 * the three small offset arrays.  Replaces: handing FeatureInputSnapshot rows to
 * FeatureProcessor::GetRawFeatureSet (feature_processor.cpp:11), for all rows at once.
 * Alignment: natural alignment of the element type is enough (8 bytes for prices, 4 for sizes); the book arrays are
 * fetched in 16-byte units, so the copy engine may read the rest of the aligned 16-byte block that holds an array's
 * first or last element (never a byte outside the memory page those elements live in). */
int mrp_fg_bind_inputs(mrp_fg_ctx* ctx, const mrp_fg_soa_view* view);

/* Copy a HOST view (view->memory == MRP_MEM_HOST; pinned memory recommended) into context-owned
 * device buffers, asynchronously on `stream`, and bind them.  Buffers are reused across calls
 * when the shape allows. */
int mrp_fg_upload(mrp_fg_ctx* ctx, const mrp_fg_soa_view* view, void* stream);

/* Host-resident data in, host-resident features out (what a pipeline holding pinned host arrays calls): every stream
 * of the HOST view is cut into time chunks of at most `chunk_rows` snapshots (0 = default, 2 Mi rows; the first and last chunks of a call are shorter) that are uploaded,
 * computed and downloaded on two alternating device buffer sets / CUDA streams, so PCIe runs in both directions while
 * the SMs compute.  A chunk is preceded by a halo (norm_window + EMA warm-up rows, one window of ticks / trades)
 * re-read from the host; results equal the single-shot upload / run / download sequence to within the EMA warm-up
 * (< 1e-17 of the column scale), including the normaliser's "poisoned for ever" columns, whose state is carried across chunks.  raw, norm: HOST double [19][n_rows] (pinned memory recommended); ONE of them may be
 * NULL — that matrix is not copied back (a consumer of the `_norm.csv` half only: half the device-to-host bytes).  Synchronous.
 * Replaces, for every row: FeatureProcessor::GetRawFeatureSet + GetProcessedFeatureSet (feature_processor.cpp:11-29). */
int mrp_fg_process_host(mrp_fg_ctx* ctx, const mrp_fg_soa_view* view, double* raw, double* norm, int64_t chunk_rows);

/* Enqueue the whole feature pass (raw features, then rolling z-scores) for every bound stream.
 * Replaces, for every row: FeatureProcessor::GetRawFeatureSet + GetProcessedFeatureSet
 * (feature_processor.cpp:11-29, feature_normalizer.cpp:32-83).  Asynchronous. */
int mrp_fg_run(mrp_fg_ctx* ctx, void* stream);

/* Device pointers to the context-owned outputs (valid until the next bind/upload/destroy).
 *   raw, norm : double [19][n_rows]     column c, row r at [c * n_rows + r]
 *   bid_dir, ask_dir : int8 [L][n_rows] (NULL unless params.emit_dirs)
 *   counts : int32 [MRP_FG_NUM_COUNTS][n_rows] (NULL unless params.emit_counts)
 * Any out pointer may be NULL. */
int mrp_fg_outputs(mrp_fg_ctx* ctx, const double** raw, const double** norm,
                   const int8_t** bid_dir, const int8_t** ask_dir, const int32_t** counts);

/* Copy outputs to HOST buffers (same layouts), asynchronously on `stream`; NULL = skip. */
int mrp_fg_download(mrp_fg_ctx* ctx, double* raw, double* norm, int8_t* bid_dir, int8_t* ask_dir,
                    int32_t* counts, void* stream);

/* Copy the rows [row_begin, row_end) of the RAW and z-scored features to HOST buffers laid out [19][row_end - row_begin]
 * (what a time shard keeps after dropping its halo, or the newest row of a trailing window); NULL = skip. */
int mrp_fg_download_rows(mrp_fg_ctx* ctx, int64_t row_begin, int64_t row_end, double* raw, double* norm, void* stream);

/* The rolling state FeatureEngine keeps between snapshots, derived on the device from per-event arrays instead of staged per row
 * (ref: data_ingestion/src/core/feature_engine.cpp:87-150).  Call after bind / upload and before mrp_fg_run; the view may then
 * leave buy_vol / sell_vol / adds NULL (28 bytes per row less on the wire).
 *   trade_size  [n_trades] int32, aligned with trade_dir: the sizes FeatureEngine::update_trade received (order_engine.cpp:104).
 *               rolling_buy_volume / rolling_sell_volume of a snapshot become the exact integer sums over the last dir_window
 *               trades before its trade cursor, with the reference's eviction rule (feature_engine.cpp:95-110: a leaving trade is
 *               subtracted from buy when its direction is > 0, from sell otherwise).  NULL = keep the bound arrays.
 *   add_cursor  [n_rows] uint32: 'A' events of the stream before the snapshot; adds_since_last_snapshot = the difference of
 *               consecutive cursors (feature_engine.cpp:74-75,119-122).  NULL = keep the bound array.
 *   memory      MRP_MEM_DEVICE or MRP_MEM_HOST for the two pointers.
 * (The tick directions of feature_engine.cpp:134-141 are always derived on the device from the midprice ticks.) */
int mrp_fg_build_state(mrp_fg_ctx* ctx, const int32_t* trade_size, const uint32_t* add_cursor, int memory, void* stream);

/* ONE STREAM SPLIT BY TIME ACROSS GPUs (BASELINE.json configs[4]).  Every state of the path has finite memory, so a shard that
 * is preceded by mrp_fg_halo_rows() rows of its stream (and one window of ticks / trades in front of those) reproduces the
 * whole-stream result on its own rows.  The halo is the TAIL of the left neighbour's shard; these two calls move it as ONE
 * contiguous device buffer, which the caller ships with whatever it has (cudaMemcpyPeerAsync, ncclSend / ncclRecv, MPI):
 *
 *   mrp_fg_halo_query   what the export will hold (info->bytes sizes the caller's message buffer).  The geometry is read from the
 *                       device once per bind (one small synchronous copy).
 *   mrp_fg_halo_export  packs the last mrp_fg_halo_rows() rows of the bound stream (ONE stream), the ticks from
 *                       (tick cursor of the first of those rows) - max(mid_window, dir_window + 1) - 2 and the trades from
 *                       (its trade cursor) - dir_window - 2 into dst_device (dst_capacity bytes; NULL = a context-owned
 *                       buffer, returned through *packed_device); cursors are rebased to the packed ticks / trades.
 *                       Asynchronous on `stream` once the geometry is known.
 *   mrp_fg_halo_import  writes a packed halo into the FRONT of the bound stream: its first info->n_rows rows, first
 *                       info->n_ticks ticks and first info->n_trades trades are overwritten (the bound arrays must be
 *                       writable, and the caller lays out its shard with that headroom: own rows from row info->n_rows on,
 *                       own ticks from tick info->n_ticks on, own cursors = info->n_ticks / n_trades + counts within the shard).
 *                       After mrp_fg_run the rows [info->n_rows, n_rows) are the shard's own (mrp_fg_download_rows,
 *                       mrp_fg_moments_rows).
 * Replaces: nothing in the reference (a single process walks the whole stream, dual_feature_pipeline.cpp:67-120). */
typedef struct mrp_fg_halo_info {
    int64_t n_rows;    /* halo rows                                   */
    int64_t n_ticks;   /* ticks the halo rows (and the shard's first rows) can see */
    int64_t n_trades;  /* trades likewise                             */
    int64_t bytes;     /* size of the packed buffer                   */
} mrp_fg_halo_info;
int mrp_fg_halo_query(mrp_fg_ctx* ctx, mrp_fg_halo_info* info, void* stream);
int mrp_fg_halo_export(mrp_fg_ctx* ctx, void* dst_device, int64_t dst_capacity, const void** packed_device, mrp_fg_halo_info* info, void* stream);
int mrp_fg_halo_import(mrp_fg_ctx* ctx, const void* packed_device, const mrp_fg_halo_info* info, void* stream);

/* A NaN (or an infinity, once evicted) poisons the reference's running sums FOR EVER (feature_normalizer.cpp:37-51): every later
 * z-score of that column is NaN.  Within one bound stream mrp_fg_run reproduces this; across time shards the state must travel:
 *   mrp_fg_poison_state  after mrp_fg_run: state[s*57 + c*3 + {0,1,2}] = first NaN / +inf / -inf row of column c of stream s in
 *                        ABSOLUTE stream rows (row_base[s] = absolute row of the bound stream's first row, NULL = 0), INT64_MAX = none.
 *   mrp_fg_poison_apply  rewrites the bound rows' z-scores for what happened BEFORE the shard: state_before = the element-wise
 *                        minimum of the states of all earlier shards (an all-gather of 57 integers per rank).  Synchronises.
 * (mrp_fg_process_host carries the state from chunk to chunk by itself.) */
int mrp_fg_poison_state(mrp_fg_ctx* ctx, const int64_t* row_base, int64_t* state, void* stream);
int mrp_fg_poison_apply(mrp_fg_ctx* ctx, const int64_t* state_before, const int64_t* row_base, void* stream);

/* EXTENSION WITHOUT A REFERENCE COUNTERPART (BASELINE.json configs[3]: "vol-of-vol over 3 horizons"; the reference has no
 * such feature, SURVEY.md 0.1 D2 — PARITY UNPINNED: its only checker is the restatement of this definition in oracle/).
 * After mrp_fg_run, for each of n_horizons (1..8) row horizons h:
 *     vov_h(t) = population standard deviation of the ewm_volatility column over the rows t-n+1 .. t of the same stream,
 *                n = min(t + 1, h); 0 when the variance is not positive
 *     z_h(t)   = rolling z-score of vov_h over norm_window rows, arithmetic of feature_normalizer.cpp:61-69
 * A non-finite value makes the rows whose window contains it non-finite, and up to one tile (the shared divisor below) of rows
 * after them (the sliding sums of that tile carry it).  The horizons and norm_window must share
 * a divisor >= 64.  raw / norm receive DEVICE pointers to double [n_horizons][n_rows] owned by the context (either may be
 * NULL); the 19 reference columns and the CSV layout are untouched. */
int mrp_fg_vol_of_vol(mrp_fg_ctx* ctx, const int32_t* horizons, int32_t n_horizons, const double** raw, const double** norm, void* stream);
/* Copy what the last mrp_fg_vol_of_vol produced to HOST buffers, double [n_horizons][n_rows] each; NULL = skip. */
int mrp_fg_download_vol_of_vol(mrp_fg_ctx* ctx, double* raw, double* norm, void* stream);

/* Per-column moments of the RAW features: out[c*3 + {0,1,2}] = {non-NaN cells, sum x, sum x^2} (double [19*3], host) over all
 * bound rows, or over the rows [row_begin, row_end) only — a time shard passes its OWN rows, so that the halo rows it
 * recomputed (which its left neighbour also counts) stay out.  These are what a caller allreduces across GPUs to get
 * global (StandardScaler-style) statistics (ref: regime_classifier/python/data_checks.py:79-80).  Synchronises `stream`. */
int mrp_fg_moments(mrp_fg_ctx* ctx, double* out, void* stream);
int mrp_fg_moments_rows(mrp_fg_ctx* ctx, int64_t row_begin, int64_t row_end, double* out, void* stream);

/* Rows of history that can still influence a row's features: norm_window + the EMA warm-up + a margin.  A time shard (or a
 * chunk of a host-resident stream) that is preceded by this many rows of its stream — plus one window of ticks / trades in
 * front of them — reproduces the whole-stream result on its own rows.  The one definition every caller uses. */
int mrp_fg_halo_rows(const mrp_fg_ctx* ctx);
int mrp_fg_halo_rows_for(const mrp_fg_params* p);   /* the same from parameters alone (NULL = defaults); needs no GPU */

/* How a batch would be cut — host arithmetic, needs no GPU.  Rows per segment of the two segment kernels (chosen per batch against
 * the device's multiprocessor count: waves x (segment + halo) is minimised; sm_count <= 0 = 148, a B200) and rows per normaliser
 * tile (a divisor of norm_window) for `n_streams` streams delimited by row_offsets[n_streams + 1].  Results are bit-identical for
 * a stream alone and inside a batch only when both are cut the same way; these are the two lengths that decide it.  Either out
 * pointer may be NULL.  (No reference counterpart: the reference walks one row at a time.) */
int mrp_fg_plan_shape(const mrp_fg_params* p, int32_t n_streams, const int64_t* row_offsets, int32_t sm_count, int64_t* segment_rows,
                      int32_t* tile_rows);
/* The chunk lengths mrp_fg_process_host uses for one stream of `rows` snapshots (chunk_rows as in that call; first_stream /
 * last_stream: whether the stream opens / closes the call — the first chunks of a call are chunk/8, /4, /2 and the last ones halve
 * what is left, because the first upload and the last download overlap nothing).  Writes at most `cap` lengths, returns their
 * number (-1: bad argument).  Host arithmetic. */
int mrp_fg_plan_host_chunks(int64_t rows, int64_t chunk_rows, int first_stream, int last_stream, int64_t* chunks, int cap);

/* The step the reference's Python stage takes next (sklearn StandardScaler over the z-scored columns of ALL rows of ALL dates:
 * regime_classifier/python/classifier.py:35-37, data_checks.py:78-80), as three calls so that several GPUs / processes can combine:
 *   mrp_fg_norm_moments   out[c*3 + {0,1,2}] = {number of non-NaN cells, sum z, sum z^2} of column c of the z-scored features
 *                         (double [19*3], host; NaN cells are missing values as for the scaler; deterministic; synchronises `stream`)
 *   (callers ADD the moment arrays of all contexts / ranks: one all-reduce of 57 doubles)
 *   mrp_fg_mean_scale     mean_scale[c*2 + {0,1}] = {mean, population standard deviation} from combined moments; a constant
 *                         column keeps scale 1 (sklearn's rule).  Pure host arithmetic, no context needed.
 *   mrp_fg_standardize    (z - mean) / scale for every bound row into a context-owned device array double [19][n_rows]
 *                         (*out_device, may be NULL) and / or a HOST buffer of the same layout (out_host, may be NULL). */
int mrp_fg_norm_moments(mrp_fg_ctx* ctx, double* out, void* stream);
int mrp_fg_norm_moments_rows(mrp_fg_ctx* ctx, int64_t row_begin, int64_t row_end, double* out, void* stream);   /* own rows of a time shard */
int mrp_fg_mean_scale(const double* moments, double* mean_scale);
int mrp_fg_standardize(mrp_fg_ctx* ctx, const double* moments, double* out_host, const double** out_device, void* stream);

/* Block until everything enqueued on `stream` by this context has finished; surfaces kernel faults. */
int mrp_fg_sync(mrp_fg_ctx* ctx, void* stream);

/* Introspection for benchmarks. */
int mrp_fg_num_kernels(void);                        /* kernels launched by one mrp_fg_run            */
const char* mrp_fg_kernel_name(int i);
/* With timing enabled, mrp_fg_run brackets every kernel with CUDA events on the launching stream. */
int mrp_fg_enable_timing(mrp_fg_ctx* ctx, int on);
int mrp_fg_kernel_times_ms(mrp_fg_ctx* ctx, float* ms /* [mrp_fg_num_kernels()] */);
int64_t mrp_fg_launch_count(const mrp_fg_ctx* ctx); /* kernels launched so far by this context */

const char* mrp_fg_last_error(void);
int mrp_fg_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MRP_FG_H */
