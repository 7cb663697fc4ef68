// api.cu — implementation of the C ABI declared in include/mrp_fg.h.
//
// Host-side plumbing only: argument checking, device buffers, the segment / tile tables, launch order
// and timing.  No feature arithmetic lives here and there is no CPU fallback: every entry point that
// needs the GPU fails with MRP_FG_ERR_CUDA when the device or a kernel is unavailable.
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <numeric>
#include <vector>

#include "mrp_device.cuh"

namespace mrp {
cudaError_t launch_windows(const DevInputs&, const DevOutputs&, const DevSegment*, int, const DevParams&, cudaStream_t);
cudaError_t launch_book(const DevInputs&, const DevOutputs&, const DevSegment*, int, const DevParams&, cudaStream_t);
cudaError_t launch_book_direct(const DevInputs&, const DevOutputs&, const DevSegment*, int, const DevParams&, int, cudaStream_t);
bool book_direct_supported(const DevParams&);
cudaError_t launch_normalize(const DevOutputs&, int64_t, const DevTile*, int, double*, int*, void*, const int64_t*, int, long long*, const long long*,
                             const DevParams&, cudaStream_t);
cudaError_t launch_norm_poison(const DevOutputs&, int64_t, void*, const int64_t*, int, long long*, const long long*, const DevParams&, cudaStream_t);
size_t norm_poison_bytes(int);
cudaError_t launch_normalize_faithful(const DevOutputs&, int64_t, const int64_t*, int, const DevParams&, cudaStream_t);
int col_moment_blocks(int64_t N);
cudaError_t launch_col_moments(const double* x, int64_t col_stride, int64_t N, double* part, cudaStream_t st);
cudaError_t launch_standardize(const double* x, double* out, int64_t N, const double* mean_scale_dev, cudaStream_t st);
size_t state_prefix_bytes(int64_t T);
size_t state_partial_bytes(int64_t T);
cudaError_t launch_build_state(const int8_t* dir, const int32_t* size, int64_t T, const uint32_t* jcur, const uint32_t* add_cursor, int64_t n,
                               int window, unsigned long long* part, unsigned long long* P, double* buy, double* sell, int32_t* adds, cudaStream_t st);
cudaError_t launch_rolling(const double* x, int64_t x_stride, double* out, int64_t out_stride, const void* tiles, int ntiles, double* tot,
                           const int32_t* windows_dev, int njobs, int tile_rows, int mode, cudaStream_t st);
bool win_geometry(int mid_window, int dir_window, int* ring, int* chunk_ticks, int* trade_ring, int* chunk_trades, size_t* smem);
}  // namespace mrp

using namespace mrp;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return MRP_FG_ERR_CUDA;
}
#define CU(call)                                             \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

constexpr int kNumKernels = 3;
const char* kKernelNames[kNumKernels] = {"windows", "book", "normalize"};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap && p) return MRP_FG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            g_err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
            return e == cudaErrorMemoryAllocation ? MRP_FG_ERR_NOMEM : MRP_FG_ERR_CUDA;
        }
        cap = bytes;
        return MRP_FG_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

// pinned host staging: the source of an asynchronous copy must stay valid (and unpaged) until the copy has run
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap && p) return MRP_FG_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) {
            g_err = std::string("cudaMallocHost(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
            return MRP_FG_ERR_NOMEM;
        }
        cap = bytes;
        return MRP_FG_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct mrp_fg_ctx {
    mrp_fg_params p{};
    int device = 0;
    int sm_count = 148;   // multiprocessors of the device (segment lengths are chosen per batch against it)
    bool bound = false;
    int n_streams = 0;
    int64_t n_rows = 0, n_ticks = 0, n_trades = 0;
    std::vector<int64_t> row_off, tick_off, trade_off;
    DevInputs in{};
    DevOutputs out{};
    DevParams dp{};
    // owned copies of the inputs (mrp_fg_upload)
    DevBuf b_ts, b_bpx, b_apx, b_bsz, b_asz, b_buy, b_sell, b_adds, b_kcur, b_jcur, b_mid, b_spread, b_tdir;
    // outputs
    DevBuf b_raw, b_norm, b_bdir, b_adir, b_counts;
    // tables
    DevBuf b_segs, b_tiles, b_tile_tot, b_flags, b_tables, b_row_off, b_poison;
    int nseg = 0, ntiles = 0;
    // the segment / tile / row-offset tables are built on the host at bind time and copied on the stream of the next run
    PinBuf h_tables;
    size_t h_seg_bytes = 0, h_tile_bytes = 0, h_off_bytes = 0;
    bool tables_dirty = false;
    cudaEvent_t tables_ready = nullptr;   // recorded after the copy: runs on other streams wait for it
    // vol-of-vol extension (mrp_fg_vol_of_vol)
    DevBuf b_vov_raw, b_vov_norm, b_vov_tiles, b_vov_tot, b_vov_win;
    // global standardisation (mrp_fg_norm_moments / mrp_fg_standardize)
    DevBuf b_mom_part, b_scaled, b_mean_scale;
    int vov_n = 0;
    uint64_t shape_gen = 0;            // bumped whenever the bound shape changes
    uint64_t vov_gen = ~0ull;          // shape_gen the cached vol-of-vol tile list was built for
    std::vector<int32_t> vov_h;        // ... and its horizons
    int vov_nt = 0;
    int64_t vov_tile = 0;
    // timing
    bool timing = false, have_times = false;
    cudaEvent_t ev[kNumKernels + 1] = {};
    bool ev_made = false;
    int64_t launches = 0;
    // mrp_fg_process_host: two child contexts (device buffer sets), each with its own non-blocking stream
    // mrp_fg_process_host: kLanes device buffer sets / streams.  Two are enough: three were measured (475.8 ms against 477.0 ms per
    // 50 M rows — a lane's next upload waits for its previous download, but the host's per-chunk preparation is not what the
    // upload engine waits for), so the third set of device buffers is not spent.
    static constexpr int kLanes = 2;
    mrp_fg_ctx* lane[kLanes] = {};
    cudaStream_t lane_stream[kLanes] = {};
    cudaEvent_t lane_done[kLanes] = {};
    uint32_t* lane_cursors[kLanes] = {};   // pinned: rebased tick / trade cursors of the chunk in flight
    size_t lane_cursor_cap[kLanes] = {};
    long long* pin_carry = nullptr;        // pinned: [57] "none" (resets the carry at a stream's start) + [kLanes] the lanes' slice_row0
    cudaEvent_t run_chain = nullptr;       // the chunks' kernels run in stream order across the two lanes (the carry is a chain)
    // non-finite state of the normaliser carried across time chunks / shards (k_norm_poison): [n_streams][19][3] absolute rows + the
    // absolute row of each bound stream's first row; null = the bound streams start at their stream's first row
    DevBuf b_carry, b_slice_row0;
    long long* carry = nullptr;              // device; may point into another context's buffer (mrp_fg_process_host lanes)
    const long long* slice_row0 = nullptr;   // device
    // rolling state derived on the device (mrp_fg_build_state)
    DevBuf b_st_buy, b_st_sell, b_st_adds, b_st_size, b_st_addcur, b_st_part, b_st_prefix;
    // time-shard halo (mrp_fg_halo_export): packed tail of the bound stream
    DevBuf b_halo;
    mrp_fg_halo_info halo_info{};
    int64_t halo_t0 = 0, halo_j0 = 0;       // first tick / trade of the packed tail, in the bound stream's indices
    uint64_t halo_gen = ~0ull;              // shape_gen the geometry was read for
    // the window and the book kernels write disjoint columns: the book pass may run on a second stream beside the window pass
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_aux[2] = {nullptr, nullptr};
    int two_streams = 0;
    // Co-scheduling (mrp_fg_run): k_windows keeps one CTA per SM busy at ~40 % of the issue slots; a share of the book pass's
    // segments is worked off beside it by the staging-free k_book form (one 160-thread CTA fits into what k_windows leaves of an
    // SM).  k_windows runs on an internal high-priority stream so that its next CTA, not a second staging-free CTA, takes the
    // slot a finished k_windows CTA frees.
    double corun_frac = 0.0;              // share of the segments given to the staging-free form (0 = off)
    cudaStream_t hi_stream = nullptr;     // k_windows when co-scheduling
    cudaEvent_t ev_win = nullptr;
};

namespace {

int check_params(const mrp_fg_params& p) {
    if (p.depth_levels < 2 || p.depth_levels > 1024) return fail(MRP_FG_ERR_INVALID, "depth_levels must be in [2,1024] (price_gap reads levels 0 and 1)");
    if (p.stress_levels < 0 || p.stress_levels > p.depth_levels) return fail(MRP_FG_ERR_INVALID, "stress_levels must be in [0,depth_levels]");
    if (p.mid_window < 1 || p.dir_window < 1 || p.spread_divisor < 1 || p.ewm_span < 0 || p.norm_window < 1)
        return fail(MRP_FG_ERR_INVALID, "windows must be positive");
    if (p.return_lag < 0 || p.return_lag + 1 > kThreads) return fail(MRP_FG_ERR_INVALID, "return_lag must be below the book kernel's chunk length (384 rows)");
    if (p.norm_mode != MRP_NORM_COMPENSATED && p.norm_mode != MRP_NORM_FAITHFUL) return fail(MRP_FG_ERR_INVALID, "unknown norm_mode");
    return MRP_FG_OK;
}

// Normaliser tile: a divisor of the window W (so that a window is a whole number of tiles).  A warp sweeps 128 rows
// per step, 4 per lane: multiples of 4 keep the 256-bit accesses, little waste in the last step.  Larger tiles
// amortise the per-tile look-back (measured: 4.8 ms at 1000 rows, 4.4 ms at 2000-3000 on 50 M rows), but the batch must
// still offer a round of tiles for 148 SMs x 2 CTAs, so the choice depends on the row count (round 1 asked for 3000 tiles, which
// picked 500-row tiles for 32 x 46 600 rows: 0.244 ms against 0.176 ms with 2000-row tiles; 200 k rows still prefer 500: 0.058 ms).
int pick_tile_rows(int W, int64_t n_rows) {
    if (const char* e = std::getenv("MRP_FG_NORM_TILE")) {  // experiments
        const int t = std::atoi(e);
        if (t >= 4 && W % t == 0) return t;
    }
    if (W <= 256) return W;
    int pick = 0;
    double best = -1.0;
    for (int t = 128; t <= 3072 && t <= W; t += 4)
        if (W % t == 0) {
            const double eff = (double)t / (double)(((t + 127) / 128) * 128);   // last-step waste
            const double fill = std::min(1.0, (double)(n_rows / t) / 300.0);    // enough tiles for the machine: one round of 148 SMs x 2 CTAs
                                                                                 // (measured at 32 x 46 600 rows: 500-row tiles 0.244 ms, 1000-2000 rows 0.176-0.181)
            const double size = std::min(1.0, 0.6 + 0.4 * (double)t / 2000.0);  // look-back amortisation
            const double s = eff * fill * size;
            if (s > best + 1e-12) { best = s; pick = t; }
        }
    if (pick) return pick;
    for (int t = std::min(W, 4096); t >= 16; --t)   // windows without a convenient divisor
        if (W % t == 0) return t;
    return 0;
}

// Segment length of the two segment kernels for a batch: see setup_shape.
int64_t pick_segment_rows(const int64_t* row_off, int S, int sm_count, int warm_rows) {
    const int sms = std::max(1, sm_count);
    int64_t seg_rows = 16384;
    double best = 1e300;
    for (int64_t cand = 4 * kThreads; cand <= 16384; cand += kThreads) {
        int64_t nseg = 0;
        for (int s2 = 0; s2 < S; ++s2) nseg += (row_off[s2 + 1] - row_off[s2] + cand - 1) / cand;
        const int64_t waves = (nseg + sms - 1) / sms;
        const double cost = (double)waves * (double)(cand + warm_rows);
        if (cost <= best) { best = cost; seg_rows = cand; }   // ties: the longer segment (fewer halos)
    }
    return seg_rows;
}

constexpr int64_t kDefaultHostChunk = 2 << 20;   // measured at 50 M rows: 1 / 2 / 3 / 4 Mi rows -> 489 / 481 / 482 / 491 ms
// Length of the chunk of mrp_fg_process_host that starts `rem` rows before the end of its stream: see mrp_fg_process_host.
int64_t next_host_chunk(int64_t rem, int64_t chunk_rows, int& head, bool last_stream) {
    const int64_t min_chunk = std::max<int64_t>(1, chunk_rows >> 3);
    int64_t take = std::min(rem, chunk_rows);
    if (head < 3 && rem > 2 * chunk_rows) take = std::max<int64_t>(1, chunk_rows >> (3 - head++));
    else if (last_stream && rem <= 2 * chunk_rows && rem > min_chunk) take = (rem + 1) / 2;
    return take;
}

// Rows after which both row-axis EMAs have forgotten their start (decay^rows < 1e-17 for the slower one; also covers the
// return lag), rounded up to whole 128-row blocks.  -1: the alphas are out of range.
int ema_warm_rows_for(const mrp_fg_params& p) {
    const double slow = std::max(1.0 - p.ofi_alpha, 1.0 - p.stress_alpha);
    int warm = 0;
    if (slow > 0.0 && slow < 1.0) warm = (int)std::ceil(std::log(1e-17) / std::log(slow));
    else if (slow >= 1.0) return -1;
    warm = std::max(warm, p.return_lag + 1);
    return ((warm + 127) / 128) * 128;   // whole 128-row blocks (384 rows for the reference's alphas), whatever the book kernel's chunk length
}

// Build device parameters, ring geometry and the host-computed constant tables.
int setup_params(mrp_fg_ctx* c) {
    const mrp_fg_params& p = c->p;
    DevParams& d = c->dp;
    d = DevParams{};
    d.L = p.depth_levels;
    d.mid_window = p.mid_window;
    d.dir_window = p.dir_window;
    d.spread_divisor = p.spread_divisor;
    d.norm_window = p.norm_window;
    d.return_lag = p.return_lag;
    d.stress_levels = p.stress_levels;
    d.emit_counts = p.emit_counts;
    d.ewm_alpha = 2.0 / (p.ewm_span + 1);  // feature_processor.cpp:64
    d.ewm_decay = 1 - d.ewm_alpha;         // :74
    d.inv_spread_divisor = 1.0 / p.spread_divisor;
    d.ofi_alpha = p.ofi_alpha;
    d.ofi_decay = 1.0 - p.ofi_alpha;       // :133
    d.ofi_min_total_vol = p.ofi_min_total_vol;
    d.stress_alpha = p.stress_alpha;
    d.stress_decay = 1.0 - p.stress_alpha; // :313
    d.stress_min_size = p.stress_min_size;
    d.stress_distance_decay = p.stress_distance_decay;

    // window kernel geometry: the prefix rings hold one window plus the new ticks / trades of one chunk
    if (p.mid_window > 32767 || p.dir_window > 32767) return fail(MRP_FG_ERR_INVALID, "mid_window / dir_window must be <= 32767 (16-bit window counters)");
    {
        size_t smem = 0;
        if (!win_geometry(p.mid_window, p.dir_window, &d.win_ring, &d.win_chunk_ticks, &d.win_trade_ring, &d.win_chunk_trades, &smem))
            return fail(MRP_FG_ERR_INVALID, "mid_window / dir_window too large for the on-chip prefix rings (limit about 4000 ticks)");
    }
    if (!(d.ewm_decay > 0.0 && d.ewm_decay < 1.0)) return fail(MRP_FG_ERR_INVALID, "ewm_span must be >= 2 (decay = 1 - 2/(span+1) in (0,1))");

    // EMA halo of the book pass
    d.ema_warm_rows = ema_warm_rows_for(p);
    if (d.ema_warm_rows < 0) return fail(MRP_FG_ERR_INVALID, "EMA alphas must be in (0,1]");

    if (pick_tile_rows(p.norm_window, 1 << 30) <= 0) return fail(MRP_FG_ERR_INVALID, "norm_window above 4096 needs a divisor in [16,4096]");

    // constant tables, evaluated by the host libm so that they are bit-identical to the reference's factors
    std::vector<double> tab;
    const size_t o_w = 0;
    for (int i = 0; i < p.depth_levels; ++i) tab.push_back(std::exp(-i * p.ofi_depth_decay));  // feature_processor.cpp:121
    const size_t o_e = tab.size();
    for (int n = 0; n <= d.win_ring; ++n) tab.push_back((double)powl((long double)d.ewm_decay, (long double)n));
    const size_t o_ei = tab.size();
    for (int n = 0; n <= 512; ++n) tab.push_back((double)powl((long double)d.ewm_decay, -(long double)n));
    const size_t o_o = tab.size();
    for (int n = 0; n <= 256; ++n) tab.push_back((double)powl((long double)d.ofi_decay, (long double)n));
    for (int k = 0; k < 6; ++k) {
        d.ofi_pw2[k] = (double)powl((long double)d.ofi_decay, (long double)(1 << k));
        d.stress_pw2[k] = (double)powl((long double)d.stress_decay, (long double)(1 << k));
    }
    const size_t o_s = tab.size();
    for (int n = 0; n <= 256; ++n) tab.push_back((double)powl((long double)d.stress_decay, (long double)n));
    const size_t o_i = tab.size();
    for (int n = 0; n <= p.norm_window; ++n) tab.push_back(n > 0 ? 1.0 / (double)n : 0.0);
    for (int n = 0; n < 256; ++n) tab.push_back(0.0);  // idle lanes of the normaliser's last partial block read past 1/W
    const size_t o_h = tab.size();
    const int hmax = std::max(p.dir_window, p.mid_window) + 1;
    for (int n = 0; n <= hmax; ++n) tab.push_back(n > 0 ? (double)((long double)n * log2l((long double)n)) : 0.0);
    const size_t o_c = tab.size();
    for (int n = 0; n <= hmax; ++n) tab.push_back(n > 0 ? 1.0 / (double)n : 0.0);
    int rc = c->b_tables.ensure(tab.size() * sizeof(double));
    if (rc) return rc;
    CU(cudaMemcpy(c->b_tables.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());   // pageable source on the legacy stream: landed before any non-blocking stream reads it (create time only)
    const double* t = static_cast<const double*>(c->b_tables.p);
    d.ofi_weight = t + o_w;
    d.ewm_pow = t + o_e;
    d.ewm_ipow = t + o_ei;
    d.ofi_pow = t + o_o;
    d.stress_pow = t + o_s;
    d.inv_cnt_norm = t + o_i;
    d.clog2c = t + o_h;
    d.inv_cnt = t + o_c;
    // the same table entries for full windows, as kernel parameters
    d.inv_ret_full = tab[o_c + std::max(p.mid_window - 1, 0)];
    d.pow_ret_full = tab[o_e + std::max(p.mid_window - 1, 0)];
    d.inv_dir_full = tab[o_c + p.dir_window];
    d.clog2c_dir_full = tab[o_h + p.dir_window];
    return MRP_FG_OK;
}

int check_view(const mrp_fg_ctx* c, const mrp_fg_soa_view* v, int want_memory) {
    (void)c;
    if (!v) return fail(MRP_FG_ERR_INVALID, "view is NULL");
    if (v->memory != want_memory) return fail(MRP_FG_ERR_INVALID, want_memory == MRP_MEM_DEVICE ? "mrp_fg_bind_inputs needs a DEVICE view" : "mrp_fg_upload needs a HOST view");
    if (v->n_streams < 1) return fail(MRP_FG_ERR_INVALID, "n_streams must be >= 1");
    if (!v->row_offsets || !v->tick_offsets || !v->trade_offsets) return fail(MRP_FG_ERR_INVALID, "offset arrays are required");
    const int S = v->n_streams;
    if (v->row_offsets[0] != 0 || v->tick_offsets[0] != 0 || v->trade_offsets[0] != 0) return fail(MRP_FG_ERR_INVALID, "offsets must start at 0");
    for (int s = 0; s < S; ++s) {
        if (v->row_offsets[s + 1] < v->row_offsets[s] || v->tick_offsets[s + 1] < v->tick_offsets[s] || v->trade_offsets[s + 1] < v->trade_offsets[s])
            return fail(MRP_FG_ERR_INVALID, "offsets must be non-decreasing");
        // cursors are uint32 and the kernels step positions by up to 2048 before comparing: keep a margin below 2^32
        if (v->tick_offsets[s + 1] - v->tick_offsets[s] > 0xffff0000ll || v->trade_offsets[s + 1] - v->trade_offsets[s] > 0xffff0000ll)
            return fail(MRP_FG_ERR_INVALID, "a stream may hold at most 2^32-65536 ticks / trades (cursors are uint32); split longer streams by time (timeshard)");
    }
    const int64_t N = v->row_offsets[S];
    if (N > 0 && (!v->ts_ns || !v->bid_px || !v->ask_px || !v->bid_sz || !v->ask_sz || !v->tick_cursor || !v->trade_cursor))
        return fail(MRP_FG_ERR_INVALID, "a snapshot-axis array is NULL");
    // buy_vol / sell_vol / adds may be NULL: mrp_fg_build_state derives them from per-event arrays before mrp_fg_run
    if (v->tick_offsets[S] > 0 && (!v->mid || !v->spread)) return fail(MRP_FG_ERR_INVALID, "mid / spread are NULL");
    if (v->trade_offsets[S] > 0 && !v->trade_dir) return fail(MRP_FG_ERR_INVALID, "trade_dir is NULL");
    return MRP_FG_OK;
}

// Shape-dependent state: outputs, segment table, tile table.
int setup_shape(mrp_fg_ctx* c, const mrp_fg_soa_view* v) {
    const int S = v->n_streams;
    c->n_streams = S;
    c->carry = nullptr;          // a new binding starts at its streams' first rows unless the caller says otherwise
    c->slice_row0 = nullptr;
    c->row_off.assign(v->row_offsets, v->row_offsets + S + 1);
    c->tick_off.assign(v->tick_offsets, v->tick_offsets + S + 1);
    c->trade_off.assign(v->trade_offsets, v->trade_offsets + S + 1);
    c->n_rows = c->row_off[S];
    c->n_ticks = c->tick_off[S];
    c->n_trades = c->trade_off[S];
    const int64_t N = c->n_rows;
    const int L = c->p.depth_levels;

    int rc;
    if ((rc = c->b_raw.ensure((size_t)N * MRP_FG_NUM_FEATURES * sizeof(double)))) return rc;
    if ((rc = c->b_norm.ensure((size_t)N * MRP_FG_NUM_FEATURES * sizeof(double)))) return rc;
    c->out.raw = static_cast<double*>(c->b_raw.p);
    c->out.norm = static_cast<double*>(c->b_norm.p);
    c->out.bdir = c->out.adir = nullptr;
    c->out.counts = nullptr;
    if (c->p.emit_dirs) {
        if ((rc = c->b_bdir.ensure((size_t)N * L))) return rc;
        if ((rc = c->b_adir.ensure((size_t)N * L))) return rc;
        c->out.bdir = static_cast<int8_t*>(c->b_bdir.p);
        c->out.adir = static_cast<int8_t*>(c->b_adir.p);
    }
    if (c->p.emit_counts) {
        if ((rc = c->b_counts.ensure((size_t)N * MRP_FG_NUM_COUNTS * sizeof(int32_t)))) return rc;
        c->out.counts = static_cast<int32_t*>(c->b_counts.p);   // every count column of every row is written by the kernels
    }

    // row segments: enough CTAs to fill 148 SMs several times over, long enough to amortise the window halo
    // Both segment kernels run one CTA per SM, so a launch costs waves x (segment + halo): the halo (one EMA warm-up of the book pass,
    // one window of ticks) is paid per segment, an unfilled last wave per launch.  Pick the whole-chunk length that minimises that
    // product for THIS batch (streams are cut separately).  50 M rows: 16 128-16 384 rows as before; 32 x 46 600 rows: 5 632 rows,
    // 288 segments = two waves, instead of 2 304 rows in five.
    int64_t seg_rows = pick_segment_rows(c->row_off.data(), S, c->sm_count, c->dp.ema_warm_rows);
    if (const char* e = std::getenv("MRP_FG_SEGMENT_ROWS")) {
        long v2 = std::atol(e);
        if (v2 >= 1) seg_rows = v2;
    }
    std::vector<DevSegment> segs;
    std::vector<DevTile> tiles;
    c->dp.tile_rows = pick_tile_rows(c->p.norm_window, N);   // depends on the batch size
    c->dp.tiles_back = c->p.norm_window / c->dp.tile_rows;
    const int TR = c->dp.tile_rows;
    for (int s = 0; s < S; ++s) {
        const int64_t r0 = c->row_off[s], r1 = c->row_off[s + 1];
        for (int64_t a = r0; a < r1; a += seg_rows) {
            DevSegment g{};
            g.row_begin = a;
            g.row_end = std::min(a + seg_rows, r1);
            g.stream_row0 = r0;
            g.stream_row1 = r1;
            g.tick0 = c->tick_off[s];
            g.n_ticks = c->tick_off[s + 1] - c->tick_off[s];
            g.trade0 = c->trade_off[s];
            g.n_trades = c->trade_off[s + 1] - c->trade_off[s];
            segs.push_back(g);
        }
        const int first = (int)tiles.size();
        int idx = 0;
        for (int64_t a = r0; a < r1; a += TR, ++idx) {
            DevTile t{};
            t.row = a;
            t.stream_row0 = r0;
            t.n = (int32_t)std::min<int64_t>(TR, r1 - a);
            t.index = idx;
            t.first_id = first;
            t.stream = s;
            tiles.push_back(t);
        }
    }
    ++c->shape_gen;
    c->nseg = (int)segs.size();
    c->ntiles = (int)tiles.size();
    if ((rc = c->b_segs.ensure(segs.size() * sizeof(DevSegment)))) return rc;
    if ((rc = c->b_tiles.ensure(tiles.size() * sizeof(DevTile)))) return rc;
    if ((rc = c->b_tile_tot.ensure(tiles.size() * 2 * MRP_FG_NUM_FEATURES * sizeof(double)))) return rc;
    if ((rc = c->b_flags.ensure((tiles.size() + 1) * sizeof(int)))) return rc;
    if ((rc = c->b_row_off.ensure((S + 1) * sizeof(int64_t)))) return rc;
    if ((rc = c->b_poison.ensure(norm_poison_bytes(S)))) return rc;
    // the tables go to pinned staging; upload_tables() copies them on the stream of the next run (no legacy-stream traffic)
    c->h_seg_bytes = segs.size() * sizeof(DevSegment);
    c->h_tile_bytes = tiles.size() * sizeof(DevTile);
    c->h_off_bytes = (size_t)(S + 1) * sizeof(int64_t);
    if (c->tables_ready) CU(cudaEventSynchronize(c->tables_ready));   // a previous copy may still be reading the staging
    if ((rc = c->h_tables.ensure(c->h_seg_bytes + c->h_tile_bytes + c->h_off_bytes))) return rc;
    unsigned char* h = static_cast<unsigned char*>(c->h_tables.p);
    if (c->h_seg_bytes) std::memcpy(h, segs.data(), c->h_seg_bytes);
    if (c->h_tile_bytes) std::memcpy(h + c->h_seg_bytes, tiles.data(), c->h_tile_bytes);
    std::memcpy(h + c->h_seg_bytes + c->h_tile_bytes, c->row_off.data(), c->h_off_bytes);
    c->tables_dirty = true;
    return MRP_FG_OK;
}

// Copy the tables of the current shape on `st` (first run after a bind) or make `st` wait for the copy that did.
int upload_tables(mrp_fg_ctx* c, cudaStream_t st) {
    if (!c->tables_ready) CU(cudaEventCreateWithFlags(&c->tables_ready, cudaEventDisableTiming));
    if (c->tables_dirty) {
        const unsigned char* h = static_cast<const unsigned char*>(c->h_tables.p);
        if (c->h_seg_bytes) CU(cudaMemcpyAsync(c->b_segs.p, h, c->h_seg_bytes, cudaMemcpyHostToDevice, st));
        if (c->h_tile_bytes) CU(cudaMemcpyAsync(c->b_tiles.p, h + c->h_seg_bytes, c->h_tile_bytes, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->b_row_off.p, h + c->h_seg_bytes + c->h_tile_bytes, c->h_off_bytes, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->tables_ready, st));
        c->tables_dirty = false;
    } else {
        CU(cudaStreamWaitEvent(st, c->tables_ready, 0));
    }
    return MRP_FG_OK;
}

}  // namespace

extern "C" {

int mrp_fg_abi_version(void) { return MRP_FG_ABI_VERSION; }
const char* mrp_fg_last_error(void) { return g_err.c_str(); }

int mrp_fg_default_params(mrp_fg_params* p) {
    if (!p) return fail(MRP_FG_ERR_INVALID, "params is NULL");
    std::memset(p, 0, sizeof(*p));
    p->depth_levels = 10;
    p->mid_window = 1800;
    p->dir_window = 1500;
    p->spread_divisor = 1500;
    p->ewm_span = 1500;
    p->norm_window = 30000;
    p->return_lag = 20;
    p->stress_levels = 5;
    p->ofi_depth_decay = 0.5;
    p->ofi_alpha = 0.2;
    p->ofi_min_total_vol = 1e-6;
    p->stress_min_size = 5.0;
    p->stress_distance_decay = 10.0;
    p->stress_alpha = 0.1;
    p->norm_mode = MRP_NORM_COMPENSATED;
    return MRP_FG_OK;
}

int mrp_fg_create(const mrp_fg_params* p, int device, mrp_fg_ctx** out) {
    if (!out) return fail(MRP_FG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    mrp_fg_params pp;
    if (p) pp = *p;
    else mrp_fg_default_params(&pp);
    int rc = check_params(pp);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(MRP_FG_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                         " (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(MRP_FG_ERR_INVALID, "device ordinal out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(MRP_FG_ERR_CUDA, std::string("device '") + prop.name + "' is not sm_100-class; the kernels are built for sm_100a only");
    mrp_fg_ctx* c = new mrp_fg_ctx();
    c->p = pp;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (const char* e = std::getenv("MRP_FG_TWO_STREAMS")) c->two_streams = std::atoi(e) != 0;
    if (const char* e = std::getenv("MRP_FG_CORUN")) c->corun_frac = std::min(0.9, std::max(0.0, std::atof(e)));
    rc = setup_params(c);
    if (rc) {
        delete c;
        return rc;
    }
    *out = c;
    return MRP_FG_OK;
}

int mrp_fg_destroy(mrp_fg_ctx* c) {
    if (!c) return MRP_FG_OK;
    cudaSetDevice(c->device);
    for (int i = 0; i < mrp_fg_ctx::kLanes; ++i) {
        if (c->lane[i]) mrp_fg_destroy(c->lane[i]);
        if (c->lane_stream[i]) cudaStreamDestroy(c->lane_stream[i]);
        if (c->lane_done[i]) cudaEventDestroy(c->lane_done[i]);
        if (c->lane_cursors[i]) cudaFreeHost(c->lane_cursors[i]);
    }
    if (c->pin_carry) cudaFreeHost(c->pin_carry);
    if (c->run_chain) cudaEventDestroy(c->run_chain);
    DevBuf* bufs[] = {&c->b_ts, &c->b_bpx, &c->b_apx, &c->b_bsz, &c->b_asz, &c->b_buy, &c->b_sell, &c->b_adds, &c->b_kcur,
                      &c->b_jcur, &c->b_mid, &c->b_spread, &c->b_tdir, &c->b_raw, &c->b_norm, &c->b_bdir, &c->b_adir,
                      &c->b_counts, &c->b_segs, &c->b_tiles, &c->b_tile_tot, &c->b_flags, &c->b_tables, &c->b_row_off, &c->b_poison,
                      &c->b_vov_raw, &c->b_vov_norm, &c->b_vov_tiles, &c->b_vov_tot, &c->b_vov_win, &c->b_mom_part, &c->b_scaled,
                      &c->b_mean_scale, &c->b_halo, &c->b_st_buy, &c->b_st_sell, &c->b_st_adds, &c->b_st_size, &c->b_st_addcur,
                      &c->b_st_part, &c->b_st_prefix, &c->b_carry, &c->b_slice_row0};
    for (DevBuf* b : bufs) b->release();
    c->h_tables.release();
    if (c->tables_ready) cudaEventDestroy(c->tables_ready);
    if (c->ev_made)
        for (auto& ev : c->ev) cudaEventDestroy(ev);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->hi_stream) cudaStreamDestroy(c->hi_stream);
    if (c->ev_win) cudaEventDestroy(c->ev_win);
    for (cudaEvent_t ev : {c->ev_fork, c->ev_join, c->ev_aux[0], c->ev_aux[1]})
        if (ev) cudaEventDestroy(ev);
    delete c;
    return MRP_FG_OK;
}

int mrp_fg_bind_inputs(mrp_fg_ctx* c, const mrp_fg_soa_view* v) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    int rc = check_view(c, v, MRP_MEM_DEVICE);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    c->bound = false;
    if ((rc = setup_shape(c, v))) return rc;
    DevInputs& in = c->in;
    in.n_rows = c->n_rows;
    in.ts = v->ts_ns;
    in.bid_px = v->bid_px;
    in.ask_px = v->ask_px;
    in.bid_sz = v->bid_sz;
    in.ask_sz = v->ask_sz;
    in.buy = v->buy_vol;
    in.sell = v->sell_vol;
    in.adds = v->adds;
    in.kcur = v->tick_cursor;
    in.jcur = v->trade_cursor;
    in.mid = v->mid;
    in.spread = v->spread;
    in.tdir = v->trade_dir;
    c->bound = true;
    return MRP_FG_OK;
}

}  // extern "C"

namespace {

// `rows` slices of `width` bytes between two pitched layouts.  The runtime rejects pitches above cudaDeviceProp::memPitch
// (2^31 - 1 bytes: a [L][n_rows] host array of more than 268 M rows), so such layouts take one plain copy per slice.
cudaError_t copy_pitched(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaMemcpyKind kind,
                         cudaStream_t st) {
    constexpr size_t kMaxPitch = 0x7fffffffull;
    if (dpitch <= kMaxPitch && spitch <= kMaxPitch) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
    for (size_t r = 0; r < rows; ++r) {
        cudaError_t e = cudaMemcpyAsync(static_cast<unsigned char*>(dst) + r * dpitch, static_cast<const unsigned char*>(src) + r * spitch,
                                        width, kind, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Copy a HOST view into the context's device buffers and bind them.  `level_stride` = elements between consecutive
// levels of the level-major host arrays (== n_rows for a self-contained view; larger when the view is a row slice of a
// bigger batch, mrp_fg_process_host).
int upload_impl(mrp_fg_ctx* c, const mrp_fg_soa_view* v, int64_t level_stride, cudaStream_t st, bool validate) {
    int rc;
    CU(cudaSetDevice(c->device));
    c->bound = false;
    if (validate) {
        // the cursors are on the host here: validate what the kernels rely on
        const int S = v->n_streams;
        for (int s = 0; s < S; ++s) {
            const int64_t K = v->tick_offsets[s + 1] - v->tick_offsets[s], T = v->trade_offsets[s + 1] - v->trade_offsets[s];
            uint32_t pk = 0, pj = 0;
            for (int64_t r = v->row_offsets[s]; r < v->row_offsets[s + 1]; ++r) {
                const uint32_t k = v->tick_cursor[r], j = v->trade_cursor[r];
                if (k < pk || j < pj) return fail(MRP_FG_ERR_INVALID, "cursors must be non-decreasing within a stream");
                if ((int64_t)k > K || (int64_t)j > T) return fail(MRP_FG_ERR_INVALID, "a cursor points past the end of its stream's tick / trade arrays");
                pk = k;
                pj = j;
            }
        }
    }
    if ((rc = setup_shape(c, v))) return rc;
    const int64_t N = c->n_rows, K = c->n_ticks, T = c->n_trades;
    const int L = c->p.depth_levels;
    struct Copy { DevBuf* b; const void* src; size_t elem; int64_t count; bool levels; };
    const Copy copies[] = {
        {&c->b_ts, v->ts_ns, 8, N, false},        {&c->b_bpx, v->bid_px, 8, N, true},     {&c->b_apx, v->ask_px, 8, N, true},
        {&c->b_bsz, v->bid_sz, 4, N, true},       {&c->b_asz, v->ask_sz, 4, N, true},     {&c->b_buy, v->buy_vol, 8, N, false},
        {&c->b_sell, v->sell_vol, 8, N, false},   {&c->b_adds, v->adds, 4, N, false},     {&c->b_kcur, v->tick_cursor, 4, N, false},
        {&c->b_jcur, v->trade_cursor, 4, N, false}, {&c->b_mid, v->mid, 8, K, false},     {&c->b_spread, v->spread, 8, K, false},
        {&c->b_tdir, v->trade_dir, 1, T, false},
    };
    for (const Copy& cp : copies) {
        const size_t bytes = (size_t)cp.count * cp.elem * (cp.levels ? (size_t)L : 1u);
        if (!cp.src && bytes) continue;   // an omitted rolling-state array (mrp_fg_build_state supplies it)
        if ((rc = cp.b->ensure(bytes))) return rc;
        if (!bytes) continue;
        if (cp.levels && level_stride != N)
            CU(copy_pitched(cp.b->p, (size_t)N * cp.elem, cp.src, (size_t)level_stride * cp.elem, (size_t)N * cp.elem, (size_t)L,
                            cudaMemcpyHostToDevice, st));
        else
            CU(cudaMemcpyAsync(cp.b->p, cp.src, bytes, cudaMemcpyHostToDevice, st));
    }
    DevInputs& in = c->in;
    in.n_rows = N;
    in.ts = static_cast<const uint64_t*>(c->b_ts.p);
    in.bid_px = static_cast<const double*>(c->b_bpx.p);
    in.ask_px = static_cast<const double*>(c->b_apx.p);
    in.bid_sz = static_cast<const int32_t*>(c->b_bsz.p);
    in.ask_sz = static_cast<const int32_t*>(c->b_asz.p);
    in.buy = v->buy_vol ? static_cast<const double*>(c->b_buy.p) : nullptr;
    in.sell = v->sell_vol ? static_cast<const double*>(c->b_sell.p) : nullptr;
    in.adds = v->adds ? static_cast<const int32_t*>(c->b_adds.p) : nullptr;
    in.kcur = static_cast<const uint32_t*>(c->b_kcur.p);
    in.jcur = static_cast<const uint32_t*>(c->b_jcur.p);
    in.mid = static_cast<const double*>(c->b_mid.p);
    in.spread = static_cast<const double*>(c->b_spread.p);
    in.tdir = static_cast<const int8_t*>(c->b_tdir.p);
    c->bound = true;
    return MRP_FG_OK;
}

}  // namespace

extern "C" {

int mrp_fg_upload(mrp_fg_ctx* c, const mrp_fg_soa_view* v, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    int rc = check_view(c, v, MRP_MEM_HOST);
    if (rc) return rc;
    return upload_impl(c, v, v->row_offsets[v->n_streams], static_cast<cudaStream_t>(stream), true);
}

// Host-resident data in, host-resident features out, with PCIe busy in both directions while the SMs compute: every
// stream is cut into time chunks; chunk i is uploaded (with a halo of norm_window + EMA warm-up rows and one window of
// ticks / trades in front, re-read from the host: < 1 % extra traffic at the default chunk size), computed and downloaded on
// one of two lanes (device buffer set + non-blocking stream) while chunk i+1 travels on the other.  Every state of the
// path has finite memory, so a chunk's own rows are exact (EMAs agree to < 1e-17 of their scale).
int mrp_fg_process_host(mrp_fg_ctx* c, const mrp_fg_soa_view* v, double* raw, double* norm, int64_t chunk_rows) {
    if (!c || (!raw && !norm)) return fail(MRP_FG_ERR_INVALID, "NULL argument (one of raw / norm may be NULL, not both)");
    int rc = check_view(c, v, MRP_MEM_HOST);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const int S = v->n_streams, L = c->p.depth_levels;
    const int64_t NT = v->row_offsets[S];
    if (NT == 0) return MRP_FG_OK;
    if (!v->buy_vol || !v->sell_vol || !v->adds)
        return fail(MRP_FG_ERR_INVALID, "mrp_fg_process_host needs buy_vol / sell_vol / adds in the view (mrp_fg_build_state works on a bound batch)");
    const int64_t H = mrp_fg_halo_rows(c);
    if (chunk_rows <= 0) chunk_rows = kDefaultHostChunk;
    const int64_t wmax = std::max(c->p.mid_window, c->p.dir_window + 1);
    constexpr int kLanes = mrp_fg_ctx::kLanes;
    for (int i = 0; i < kLanes; ++i) {
        if (!c->lane[i]) {
            mrp_fg_params p = c->p;
            p.emit_dirs = p.emit_counts = 0;
            p.norm_mode = c->p.norm_mode;
            if ((rc = mrp_fg_create(&p, c->device, &c->lane[i]))) return rc;
            CU(cudaStreamCreateWithFlags(&c->lane_stream[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&c->lane_done[i], cudaEventDisableTiming));
        }
    }
    // a NaN (or an evicted infinity) poisons the reference's running sums FOR EVER (feature_normalizer.cpp:37-51): the state
    // travels from chunk to chunk in a device table; the chunks' kernels are chained across the two lanes by one event
    constexpr int kCarry = MRP_FG_NUM_FEATURES * 3;
    if (!c->pin_carry) {
        CU(cudaMallocHost(reinterpret_cast<void**>(&c->pin_carry), sizeof(long long) * (kCarry + kLanes)));
        for (int i = 0; i < kCarry; ++i) c->pin_carry[i] = 0x7fffffffffffffffll;
        CU(cudaEventCreateWithFlags(&c->run_chain, cudaEventDisableTiming));
    }
    if ((rc = c->b_carry.ensure(sizeof(long long) * kCarry))) return rc;
    for (int i = 0; i < kLanes; ++i)
        if ((rc = c->lane[i]->b_slice_row0.ensure(sizeof(long long)))) return rc;
    bool chained = false;
    bool busy[kLanes] = {};
    int li = 0;
    // whatever way this function is left, nothing may still be reading the caller's inputs or writing the caller's outputs
    struct Drain {
        mrp_fg_ctx* c;
        ~Drain() {
            for (int i = 0; i < mrp_fg_ctx::kLanes; ++i)
                if (c->lane_stream[i]) cudaStreamSynchronize(c->lane_stream[i]);
        }
    } drain{c};
    for (int s = 0; s < S; ++s) {
        const int64_t r0 = v->row_offsets[s], r1 = v->row_offsets[s + 1];
        const int64_t k0 = v->tick_offsets[s], K = v->tick_offsets[s + 1] - k0;
        const int64_t j0 = v->trade_offsets[s], T = v->trade_offsets[s + 1] - j0;
        // Chunk schedule: the first chunk's upload and the last chunk's download overlap nothing, so the call starts with
        // chunk/8, chunk/4, chunk/2 and ends by halving what is left down to chunk/8 (measured at 50 M rows, 2 Mi-row chunks:
        // the exposed head and tail shrink from ~40 ms to ~10 ms of a 480 ms call)
        int head = s == 0 ? 0 : 3;
        int64_t b = r0;
        for (int64_t a = r0; a < r1; a = b, li = (li + 1) % kLanes) {
            b = a + next_host_chunk(r1 - a, chunk_rows, head, s == S - 1);
            const int64_t e0 = std::max(r0, a - H);           // first halo row
            const int64_t n = b - e0;
            int64_t tb = 0, jb = 0;                            // tick / trade bases: the slice's "stream start"
            if (e0 > r0) {
                tb = std::max<int64_t>(0, (int64_t)v->tick_cursor[e0] - wmax - 2);
                jb = std::max<int64_t>(0, (int64_t)v->trade_cursor[e0] - c->p.dir_window - 2);
            }
            const int64_t t1 = v->tick_cursor[b - 1], j1 = v->trade_cursor[b - 1];
            if (t1 > K || j1 > T || tb > t1 || jb > j1) return fail(MRP_FG_ERR_INVALID, "a cursor points past the end of its stream's tick / trade arrays");
            if (busy[li]) CU(cudaEventSynchronize(c->lane_done[li]));   // the lane's buffers (device and pinned) are free again
            if (c->lane_cursor_cap[li] < (size_t)n * 2) {
                if (c->lane_cursors[li]) cudaFreeHost(c->lane_cursors[li]);
                c->lane_cursors[li] = nullptr;
                c->lane_cursor_cap[li] = 0;
                CU(cudaMallocHost(reinterpret_cast<void**>(&c->lane_cursors[li]), sizeof(uint32_t) * (size_t)n * 2));
                c->lane_cursor_cap[li] = (size_t)n * 2;
            }
            uint32_t* kc = c->lane_cursors[li];
            uint32_t* jc = kc + n;
            {
                uint32_t pk = 0, pj = 0;
                for (int64_t i = 0; i < n; ++i) {
                    const int64_t k = (int64_t)v->tick_cursor[e0 + i] - tb, j = (int64_t)v->trade_cursor[e0 + i] - jb;
                    if (k < (int64_t)pk || j < (int64_t)pj) return fail(MRP_FG_ERR_INVALID, "cursors must be non-decreasing within a stream");
                    kc[i] = pk = (uint32_t)k;
                    jc[i] = pj = (uint32_t)j;
                }
            }
            const int64_t ro[2] = {0, n}, to[2] = {0, t1 - tb}, tro[2] = {0, j1 - jb};
            mrp_fg_soa_view sub{};
            sub.n_streams = 1;
            sub.memory = MRP_MEM_HOST;
            sub.row_offsets = ro;
            sub.tick_offsets = to;
            sub.trade_offsets = tro;
            sub.ts_ns = v->ts_ns + e0;
            sub.bid_px = v->bid_px + e0;
            sub.ask_px = v->ask_px + e0;
            sub.bid_sz = v->bid_sz + e0;
            sub.ask_sz = v->ask_sz + e0;
            sub.buy_vol = v->buy_vol + e0;
            sub.sell_vol = v->sell_vol + e0;
            sub.adds = v->adds + e0;
            sub.tick_cursor = kc;
            sub.trade_cursor = jc;
            sub.mid = v->mid + k0 + tb;
            sub.spread = v->spread + k0 + tb;
            sub.trade_dir = v->trade_dir + j0 + jb;
            mrp_fg_ctx* ln = c->lane[li];
            cudaStream_t st = c->lane_stream[li];
            if ((rc = upload_impl(ln, &sub, NT, st, false))) return rc;
            if (chained) CU(cudaStreamWaitEvent(st, c->run_chain, 0));
            if (a == r0) CU(cudaMemcpyAsync(c->b_carry.p, c->pin_carry, sizeof(long long) * kCarry, cudaMemcpyHostToDevice, st));   // a new stream: nothing poisoned yet
            c->pin_carry[kCarry + li] = e0 - r0;            // the slice's first row, in stream rows (the lane's previous chunk has completed: busy[] above)
            CU(cudaMemcpyAsync(ln->b_slice_row0.p, c->pin_carry + kCarry + li, sizeof(long long), cudaMemcpyHostToDevice, st));
            ln->carry = static_cast<long long*>(c->b_carry.p);
            ln->slice_row0 = static_cast<const long long*>(ln->b_slice_row0.p);
            if ((rc = mrp_fg_run(ln, st))) return rc;
            CU(cudaEventRecord(c->run_chain, st));
            chained = true;
            const int64_t skip = a - e0, own = b - a;
            if (raw)    // a matrix the caller does not want stays on the device: half the device-to-host bytes
                CU(copy_pitched(raw + a, (size_t)NT * 8, ln->out.raw + skip, (size_t)n * 8, (size_t)own * 8, MRP_FG_NUM_FEATURES,
                                cudaMemcpyDeviceToHost, st));
            if (norm)
                CU(copy_pitched(norm + a, (size_t)NT * 8, ln->out.norm + skip, (size_t)n * 8, (size_t)own * 8, MRP_FG_NUM_FEATURES,
                                cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(c->lane_done[li], st));
            busy[li] = true;
            c->launches += kNumKernels + 1;
        }
    }
    for (int i = 0; i < kLanes; ++i)
        if (busy[i]) CU(cudaEventSynchronize(c->lane_done[i]));
    CU(cudaGetLastError());
    (void)L;
    return MRP_FG_OK;
}

int mrp_fg_run(mrp_fg_ctx* c, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "mrp_fg_run before mrp_fg_bind_inputs / mrp_fg_upload");
    if (c->n_rows == 0) return MRP_FG_OK;
    if (!c->in.buy || !c->in.sell || !c->in.adds)
        return fail(MRP_FG_ERR_STATE, "rolling_buy_volume / rolling_sell_volume / adds are not bound: pass them in the view or call mrp_fg_build_state");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const DevSegment* segs = static_cast<const DevSegment*>(c->b_segs.p);
    {
        int rc = upload_tables(c, st);
        if (rc) return rc;
    }
    const bool tm = c->timing;
    if (tm && !c->ev_made) {
        for (auto& ev : c->ev) CU(cudaEventCreate(&ev));
        c->ev_made = true;
    }
    const int n_direct = (c->corun_frac > 0.0 && !c->two_streams && book_direct_supported(c->dp) && c->nseg >= 2)
                             ? (int)(c->corun_frac * c->nseg) : 0;
    if (n_direct > 0 && !c->hi_stream) {
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CU(cudaStreamCreateWithPriority(&c->hi_stream, cudaStreamNonBlocking, greatest));
        CU(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, least));
        CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_win, cudaEventDisableTiming));
    }
    if (c->two_streams && !c->aux_stream) {
        CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        CU(cudaEventCreate(&c->ev_aux[0]));
        CU(cudaEventCreate(&c->ev_aux[1]));
    }
    if (tm) CU(cudaEventRecord(c->ev[0], st));
    if (c->two_streams) {
        // fork: the book pass on the auxiliary stream, the window pass on the caller's; join before the normaliser
        CU(cudaEventRecord(c->ev_fork, st));
        CU(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        if (tm) CU(cudaEventRecord(c->ev_aux[0], c->aux_stream));
        CU(launch_book(c->in, c->out, segs, c->nseg, c->dp, c->aux_stream));
        if (tm) CU(cudaEventRecord(c->ev_aux[1], c->aux_stream));
        CU(cudaEventRecord(c->ev_join, c->aux_stream));
        CU(launch_windows(c->in, c->out, segs, c->nseg, c->dp, st));
        if (tm) CU(cudaEventRecord(c->ev[1], st));
        CU(cudaStreamWaitEvent(st, c->ev_join, 0));
    } else if (n_direct > 0) {
        CU(cudaEventRecord(c->ev_fork, st));
        CU(cudaStreamWaitEvent(c->hi_stream, c->ev_fork, 0));
        CU(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        CU(launch_windows(c->in, c->out, segs, c->nseg, c->dp, c->hi_stream));   // first: one CTA on every SM
        CU(cudaEventRecord(c->ev_win, c->hi_stream));
        CU(launch_book_direct(c->in, c->out, segs, n_direct, c->dp, 0, c->aux_stream));   // ... then what still fits beside them
        CU(cudaEventRecord(c->ev_join, c->aux_stream));
        CU(cudaStreamWaitEvent(st, c->ev_win, 0));
        if (tm) CU(cudaEventRecord(c->ev[1], st));   // "windows" = until k_windows is done (part of the book pass ran beside it)
        CU(launch_book(c->in, c->out, segs + n_direct, c->nseg - n_direct, c->dp, st));   // "book" = what is left of it after that
        CU(cudaStreamWaitEvent(st, c->ev_join, 0));
        c->launches += 1;
    } else {
        CU(launch_windows(c->in, c->out, segs, c->nseg, c->dp, st));
        if (tm) CU(cudaEventRecord(c->ev[1], st));
        CU(launch_book(c->in, c->out, segs, c->nseg, c->dp, st));
    }
    if (tm) CU(cudaEventRecord(c->ev[2], st));
    CU(launch_normalize(c->out, c->n_rows, static_cast<const DevTile*>(c->b_tiles.p), c->ntiles,
                        static_cast<double*>(c->b_tile_tot.p), static_cast<int*>(c->b_flags.p), c->b_poison.p,
                        static_cast<const int64_t*>(c->b_row_off.p), c->n_streams, c->carry, c->slice_row0, c->dp, st));
    c->launches += kNumKernels + 1;  // + the (normally empty) non-finite fix-up pass
    if (c->p.norm_mode == MRP_NORM_FAITHFUL) {
        CU(launch_normalize_faithful(c->out, c->n_rows, static_cast<const int64_t*>(c->b_row_off.p), c->n_streams, c->dp, st));
        c->launches += 1;
    }
    if (tm) {
        CU(cudaEventRecord(c->ev[kNumKernels], st));
        c->have_times = true;
    }
    return MRP_FG_OK;
}

int mrp_fg_outputs(mrp_fg_ctx* c, const double** raw, const double** norm, const int8_t** bid_dir, const int8_t** ask_dir,
                   const int32_t** counts) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    if (raw) *raw = c->out.raw;
    if (norm) *norm = c->out.norm;
    if (bid_dir) *bid_dir = c->out.bdir;
    if (ask_dir) *ask_dir = c->out.adir;
    if (counts) *counts = c->out.counts;
    return MRP_FG_OK;
}

int mrp_fg_download(mrp_fg_ctx* c, double* raw, double* norm, int8_t* bid_dir, int8_t* ask_dir, int32_t* counts, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t N = (size_t)c->n_rows, L = (size_t)c->p.depth_levels;
    if (N == 0) return MRP_FG_OK;
    if (raw) CU(cudaMemcpyAsync(raw, c->out.raw, N * MRP_FG_NUM_FEATURES * 8, cudaMemcpyDeviceToHost, st));
    if (norm) CU(cudaMemcpyAsync(norm, c->out.norm, N * MRP_FG_NUM_FEATURES * 8, cudaMemcpyDeviceToHost, st));
    if ((bid_dir || ask_dir) && !c->out.bdir) return fail(MRP_FG_ERR_STATE, "directions were not requested (params.emit_dirs)");
    if (bid_dir) CU(cudaMemcpyAsync(bid_dir, c->out.bdir, N * L, cudaMemcpyDeviceToHost, st));
    if (ask_dir) CU(cudaMemcpyAsync(ask_dir, c->out.adir, N * L, cudaMemcpyDeviceToHost, st));
    if (counts && !c->out.counts) return fail(MRP_FG_ERR_STATE, "counts were not requested (params.emit_counts)");
    if (counts) CU(cudaMemcpyAsync(counts, c->out.counts, N * MRP_FG_NUM_COUNTS * 4, cudaMemcpyDeviceToHost, st));
    return MRP_FG_OK;
}

// ---- vol-of-vol: an extension without a reference counterpart (see include/mrp_fg.h and kernels_volvol.cu)
int mrp_fg_vol_of_vol(mrp_fg_ctx* c, const int32_t* horizons, int32_t n_horizons, const double** raw, const double** norm, void* stream) {
    if (!c || !horizons) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    if (n_horizons < 1 || n_horizons > 8) return fail(MRP_FG_ERR_INVALID, "n_horizons must be in [1,8]");
    // one tile length for every window: the greatest common divisor of the horizons and the z-score window
    int64_t g = c->p.norm_window;
    for (int k = 0; k < n_horizons; ++k) {
        if (horizons[k] < 1) return fail(MRP_FG_ERR_INVALID, "horizons must be positive");
        g = std::gcd<int64_t, int64_t>(g, horizons[k]);
    }
    if (g < 64) return fail(MRP_FG_ERR_INVALID, "the horizons and norm_window must share a divisor >= 64 (the tile length of the rolling sums)");
    int64_t tile = g;
    while (tile > 2048 && tile % 2 == 0) tile /= 2;   // any divisor of g divides every window; keep tiles short enough to spread over the SMs
    if (tile > 8192) tile = g;
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t N = c->n_rows;
    c->vov_n = n_horizons;
    if (raw) *raw = nullptr;
    if (norm) *norm = nullptr;
    if (N == 0) return MRP_FG_OK;
    int rc;
    if ((rc = c->b_vov_raw.ensure((size_t)N * n_horizons * sizeof(double)))) return rc;
    if ((rc = c->b_vov_norm.ensure((size_t)N * n_horizons * sizeof(double)))) return rc;
    const std::vector<int32_t> hv(horizons, horizons + n_horizons);
    if (c->vov_gen != c->shape_gen || c->vov_h != hv || c->vov_tile != tile) {   // the tile list depends on the shape and the windows only
        struct HostTile { int64_t row; int32_t n; int32_t index; };
        std::vector<HostTile> tiles;
        for (int s = 0; s < c->n_streams; ++s) {
            int idx = 0;
            for (int64_t a = c->row_off[s]; a < c->row_off[s + 1]; a += tile, ++idx)
                tiles.push_back({a, (int32_t)std::min<int64_t>(tile, c->row_off[s + 1] - a), idx});
        }
        const int nt_new = (int)tiles.size();
        if ((rc = c->b_vov_tiles.ensure(tiles.size() * sizeof(HostTile)))) return rc;
        if ((rc = c->b_vov_tot.ensure((size_t)nt_new * n_horizons * 2 * sizeof(double)))) return rc;
        if ((rc = c->b_vov_win.ensure(16 * sizeof(int32_t)))) return rc;
        int32_t win[16];
        for (int k = 0; k < 8; ++k) {
            win[k] = k < n_horizons ? horizons[k] : 1;
            win[8 + k] = c->p.norm_window;
        }
        // (pageable sources: the copies are staged by the runtime before the calls return)
        CU(cudaMemcpyAsync(c->b_vov_tiles.p, tiles.data(), tiles.size() * sizeof(HostTile), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->b_vov_win.p, win, sizeof(win), cudaMemcpyHostToDevice, st));
        c->vov_gen = c->shape_gen;
        c->vov_h = hv;
        c->vov_tile = tile;
        c->vov_nt = nt_new;
    }
    const int nt = c->vov_nt;
    double* vraw = static_cast<double*>(c->b_vov_raw.p);
    double* vnorm = static_cast<double*>(c->b_vov_norm.p);
    const int32_t* wdev = static_cast<const int32_t*>(c->b_vov_win.p);
    // horizon k: rolling standard deviation of the ewm_volatility column ...
    CU(launch_rolling(c->out.raw + (int64_t)MRP_F_EWM_VOLATILITY * N, 0, vraw, N, c->b_vov_tiles.p, nt, static_cast<double*>(c->b_vov_tot.p),
                      wdev, n_horizons, (int)tile, /*std*/ 0, st));
    // ... and its rolling z-score over norm_window rows
    CU(launch_rolling(vraw, N, vnorm, N, c->b_vov_tiles.p, nt, static_cast<double*>(c->b_vov_tot.p), wdev + 8, n_horizons, (int)tile, /*z*/ 1, st));
    c->launches += 4;
    if (raw) *raw = vraw;
    if (norm) *norm = vnorm;
    return MRP_FG_OK;
}

int mrp_fg_download_vol_of_vol(mrp_fg_ctx* c, double* raw, double* norm, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    if (!c->bound || c->vov_n == 0) return fail(MRP_FG_ERR_STATE, "mrp_fg_vol_of_vol has not run");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)c->n_rows * c->vov_n * sizeof(double);
    if (bytes == 0) return MRP_FG_OK;
    if (raw) CU(cudaMemcpyAsync(raw, c->b_vov_raw.p, bytes, cudaMemcpyDeviceToHost, st));
    if (norm) CU(cudaMemcpyAsync(norm, c->b_vov_norm.p, bytes, cudaMemcpyDeviceToHost, st));
    return MRP_FG_OK;
}

int mrp_fg_download_rows(mrp_fg_ctx* c, int64_t row_begin, int64_t row_end, double* raw, double* norm, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    if (row_begin < 0 || row_end < row_begin || row_end > c->n_rows) return fail(MRP_FG_ERR_INVALID, "row range outside the bound rows");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t N = (size_t)c->n_rows, n = (size_t)(row_end - row_begin);
    if (n == 0) return MRP_FG_OK;
    // one strided copy per feature matrix: 19 column slices of n rows each
    if (raw) CU(copy_pitched(raw, n * 8, c->out.raw + row_begin, N * 8, n * 8, MRP_FG_NUM_FEATURES, cudaMemcpyDeviceToHost, st));
    if (norm) CU(copy_pitched(norm, n * 8, c->out.norm + row_begin, N * 8, n * 8, MRP_FG_NUM_FEATURES, cudaMemcpyDeviceToHost, st));
    return MRP_FG_OK;
}

}  // extern "C"

namespace {
// {non-NaN count, sum, sum of squares} of the rows [row_begin, row_end) of every column of a [19][n_rows] matrix: deterministic
// block partials on the device, added in block order in extended precision on the host
int column_moments(mrp_fg_ctx* c, const double* matrix, int64_t row_begin, int64_t row_end, double* out, cudaStream_t st) {
    if (row_begin < 0 || row_end < row_begin || row_end > c->n_rows) return fail(MRP_FG_ERR_INVALID, "row range outside the bound rows");
    const int64_t n = row_end - row_begin;
    const int nb = col_moment_blocks(n);
    std::vector<double> part((size_t)nb * MRP_FG_NUM_FEATURES * 3);
    if (nb > 0) {
        int rc = c->b_mom_part.ensure(part.size() * sizeof(double));
        if (rc) return rc;
        CU(launch_col_moments(matrix + row_begin, c->n_rows, n, static_cast<double*>(c->b_mom_part.p), st));
        c->launches += 1;
        CU(cudaMemcpyAsync(part.data(), c->b_mom_part.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    for (int f = 0; f < MRP_FG_NUM_FEATURES; ++f) {
        long double cnt = 0, s = 0, s2 = 0;
        for (int b = 0; b < nb; ++b) {
            const double* p = &part[((size_t)b * MRP_FG_NUM_FEATURES + f) * 3];
            cnt += p[0];
            s += p[1];
            s2 += p[2];
        }
        out[f * 3 + 0] = (double)cnt;
        out[f * 3 + 1] = (double)s;
        out[f * 3 + 2] = (double)s2;
    }
    return MRP_FG_OK;
}
}  // namespace

extern "C" {

int mrp_fg_moments_rows(mrp_fg_ctx* c, int64_t row_begin, int64_t row_end, double* out, void* stream) {
    if (!c || !out) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    CU(cudaSetDevice(c->device));
    return column_moments(c, c->out.raw, row_begin, row_end, out, static_cast<cudaStream_t>(stream));
}

int mrp_fg_moments(mrp_fg_ctx* c, double* out, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    return mrp_fg_moments_rows(c, 0, c->n_rows, out, stream);
}

int mrp_fg_norm_moments_rows(mrp_fg_ctx* c, int64_t row_begin, int64_t row_end, double* out, void* stream) {
    if (!c || !out) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    CU(cudaSetDevice(c->device));
    return column_moments(c, c->out.norm, row_begin, row_end, out, static_cast<cudaStream_t>(stream));
}

int mrp_fg_norm_moments(mrp_fg_ctx* c, double* out, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    return mrp_fg_norm_moments_rows(c, 0, c->n_rows, out, stream);
}

int mrp_fg_halo_rows_for(const mrp_fg_params* p) {
    mrp_fg_params d;
    if (!p) {
        mrp_fg_default_params(&d);
        p = &d;
    }
    const int warm = ema_warm_rows_for(*p);
    return warm < 0 ? -1 : p->norm_window + warm + 32;
}

int mrp_fg_halo_rows(const mrp_fg_ctx* c) { return c ? mrp_fg_halo_rows_for(&c->p) : -1; }

int mrp_fg_plan_shape(const mrp_fg_params* p, int32_t n_streams, const int64_t* row_offsets, int32_t sm_count, int64_t* segment_rows,
                      int32_t* tile_rows) {
    mrp_fg_params d;
    if (!p) {
        mrp_fg_default_params(&d);
        p = &d;
    }
    if (n_streams < 1 || !row_offsets) return fail(MRP_FG_ERR_INVALID, "n_streams / row_offsets");
    for (int s = 0; s < n_streams; ++s)
        if (row_offsets[s + 1] < row_offsets[s]) return fail(MRP_FG_ERR_INVALID, "row_offsets must be non-decreasing");
    const int warm = ema_warm_rows_for(*p);
    if (warm < 0) return fail(MRP_FG_ERR_INVALID, "EMA alphas must be in (0,1]");
    long forced = 0;
    if (const char* e = std::getenv("MRP_FG_SEGMENT_ROWS")) forced = std::atol(e);
    if (segment_rows) *segment_rows = forced >= 1 ? forced : pick_segment_rows(row_offsets, n_streams, sm_count > 0 ? sm_count : 148, warm);
    if (tile_rows) {
        const int t = pick_tile_rows(p->norm_window, row_offsets[n_streams]);
        if (t <= 0) return fail(MRP_FG_ERR_INVALID, "norm_window above 4096 needs a divisor in [16,4096]");
        *tile_rows = t;
    }
    return MRP_FG_OK;
}

int mrp_fg_plan_host_chunks(int64_t rows, int64_t chunk_rows, int first_stream, int last_stream, int64_t* chunks, int cap) {
    if (rows < 0 || (cap > 0 && !chunks)) return -1;
    if (chunk_rows <= 0) chunk_rows = kDefaultHostChunk;
    int head = first_stream ? 0 : 3, n = 0;
    for (int64_t a = 0; a < rows; ++n) {
        const int64_t take = next_host_chunk(rows - a, chunk_rows, head, last_stream != 0);
        if (n < cap) chunks[n] = take;
        a += take;
    }
    return n;
}

int mrp_fg_mean_scale(const double* moments, double* mean_scale) {
    if (!moments || !mean_scale) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    for (int f = 0; f < MRP_FG_NUM_FEATURES; ++f) {
        const long double n = moments[f * 3], s = moments[f * 3 + 1], s2 = moments[f * 3 + 2];
        double mean = 0.0, scale = 1.0;
        if (n > 0) {
            const long double m = s / n;
            long double var = s2 / n - m * m;
            if (var < 0) var = 0;
            mean = (double)m;
            // sklearn's constant-feature rule (utils/extmath.py: _is_constant_feature) and zero handling: such columns are left unscaled
            const long double eps = 2.220446049250313e-16L;
            const long double bound = n * eps * var + (n * m * eps) * (n * m * eps);
            scale = (var <= bound || !(var > 0)) ? 1.0 : (double)sqrtl(var);
        }
        mean_scale[f * 2] = mean;
        mean_scale[f * 2 + 1] = scale;
    }
    return MRP_FG_OK;
}

int mrp_fg_standardize(mrp_fg_ctx* c, const double* moments, double* out_host, const double** out_device, void* stream) {
    if (!c || !moments) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t N = c->n_rows;
    if (out_device) *out_device = nullptr;
    if (N == 0) return MRP_FG_OK;
    double ms[MRP_FG_NUM_FEATURES * 2];
    int rc = mrp_fg_mean_scale(moments, ms);
    if (rc) return rc;
    if ((rc = c->b_scaled.ensure((size_t)N * MRP_FG_NUM_FEATURES * sizeof(double)))) return rc;
    if ((rc = c->b_mean_scale.ensure(sizeof(ms)))) return rc;
    CU(cudaMemcpyAsync(c->b_mean_scale.p, ms, sizeof(ms), cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns
    CU(launch_standardize(c->out.norm, static_cast<double*>(c->b_scaled.p), N, static_cast<const double*>(c->b_mean_scale.p), st));
    c->launches += 1;
    if (out_host) CU(cudaMemcpyAsync(out_host, c->b_scaled.p, (size_t)N * MRP_FG_NUM_FEATURES * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out_device) *out_device = static_cast<const double*>(c->b_scaled.p);
    return MRP_FG_OK;
}

}  // extern "C"

namespace mrp {
cudaError_t launch_rebase(const uint32_t* src, uint32_t* dst, int64_t n, uint32_t base, cudaStream_t st);
}

namespace {
// byte offsets of the arrays inside a packed halo (every array starts on a 16-byte boundary)
struct HaloLayout {
    size_t ts, bpx, apx, bsz, asz, buy, sell, adds, kcur, jcur, mid, spread, tdir, total;
};
HaloLayout halo_layout(int64_t H, int64_t K, int64_t T, int L) {
    HaloLayout o{};
    size_t at = 0;
    auto take = [&](size_t bytes) {
        const size_t a = at;
        at += (bytes + 15) & ~(size_t)15;
        return a;
    };
    o.ts = take((size_t)H * 8);
    o.bpx = take((size_t)H * L * 8);
    o.apx = take((size_t)H * L * 8);
    o.bsz = take((size_t)H * L * 4);
    o.asz = take((size_t)H * L * 4);
    o.buy = take((size_t)H * 8);
    o.sell = take((size_t)H * 8);
    o.adds = take((size_t)H * 4);
    o.kcur = take((size_t)H * 4);
    o.jcur = take((size_t)H * 4);
    o.mid = take((size_t)K * 8);
    o.spread = take((size_t)K * 8);
    o.tdir = take((size_t)T);
    o.total = at;
    return o;
}
}  // namespace

extern "C" {

}  // extern "C"

namespace {
// where the tail of the bound stream starts on the tick / trade axes: read from the device once per bind
int halo_geometry(mrp_fg_ctx* c, cudaStream_t st) {
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    if (c->n_streams != 1) return fail(MRP_FG_ERR_INVALID, "the halo calls need ONE bound stream (a time shard)");
    const int64_t N = c->n_rows, H = mrp_fg_halo_rows(c);
    if (N < H) return fail(MRP_FG_ERR_INVALID, "a time shard must hold at least mrp_fg_halo_rows() rows: the halo comes from ONE neighbour");
    if (c->halo_gen == c->shape_gen) return MRP_FG_OK;
    const int64_t a = N - H;
    uint32_t kj[2];
    CU(cudaMemcpyAsync(&kj[0], c->in.kcur + a, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&kj[1], c->in.jcur + a, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int64_t wmax = std::max(c->p.mid_window, c->p.dir_window + 1);
    c->halo_t0 = std::max<int64_t>(0, (int64_t)kj[0] - wmax - 2);
    c->halo_j0 = std::max<int64_t>(0, (int64_t)kj[1] - c->p.dir_window - 2);
    if (c->halo_t0 > c->n_ticks || c->halo_j0 > c->n_trades) return fail(MRP_FG_ERR_INVALID, "a cursor points past the end of its stream's tick / trade arrays");
    c->halo_info.n_rows = H;
    c->halo_info.n_ticks = c->n_ticks - c->halo_t0;
    c->halo_info.n_trades = c->n_trades - c->halo_j0;
    c->halo_info.bytes = (int64_t)halo_layout(H, c->halo_info.n_ticks, c->halo_info.n_trades, c->p.depth_levels).total;
    c->halo_gen = c->shape_gen;
    return MRP_FG_OK;
}
}  // namespace

extern "C" {

// the rolling state FeatureEngine keeps between snapshots, from per-event arrays (kernels_state.cu)
int mrp_fg_build_state(mrp_fg_ctx* c, const int32_t* trade_size, const uint32_t* add_cursor, int memory, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "mrp_fg_build_state after mrp_fg_bind_inputs / mrp_fg_upload");
    if (memory != MRP_MEM_DEVICE && memory != MRP_MEM_HOST) return fail(MRP_FG_ERR_INVALID, "memory must be MRP_MEM_DEVICE or MRP_MEM_HOST");
    if (!trade_size && !add_cursor) return MRP_FG_OK;
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t N = c->n_rows, T = c->n_trades;
    int rc;
    if (memory == MRP_MEM_HOST) {   // stage the per-event arrays
        if (trade_size) {
            if ((rc = c->b_st_size.ensure((size_t)T * 4))) return rc;
            if (T) CU(cudaMemcpyAsync(c->b_st_size.p, trade_size, (size_t)T * 4, cudaMemcpyHostToDevice, st));
            trade_size = static_cast<const int32_t*>(c->b_st_size.p);
        }
        if (add_cursor) {
            if ((rc = c->b_st_addcur.ensure((size_t)N * 4))) return rc;
            if (N) CU(cudaMemcpyAsync(c->b_st_addcur.p, add_cursor, (size_t)N * 4, cudaMemcpyHostToDevice, st));
            add_cursor = static_cast<const uint32_t*>(c->b_st_addcur.p);
        }
    }
    if (trade_size) {
        if ((rc = c->b_st_buy.ensure((size_t)N * 8))) return rc;
        if ((rc = c->b_st_sell.ensure((size_t)N * 8))) return rc;
    }
    if (add_cursor && (rc = c->b_st_adds.ensure((size_t)N * 4))) return rc;
    int64_t tmax = 0;
    for (int s = 0; s < c->n_streams; ++s) tmax = std::max(tmax, c->trade_off[s + 1] - c->trade_off[s]);
    if (trade_size) {
        if ((rc = c->b_st_part.ensure(state_partial_bytes(tmax)))) return rc;
        if ((rc = c->b_st_prefix.ensure(state_prefix_bytes(tmax)))) return rc;
    }
    for (int s = 0; s < c->n_streams; ++s) {   // streams one after the other on `st`: they share the prefix scratch
        const int64_t r0 = c->row_off[s], n = c->row_off[s + 1] - r0, j0 = c->trade_off[s], Ts = c->trade_off[s + 1] - j0;
        CU(launch_build_state(c->in.tdir + j0, trade_size ? trade_size + j0 : nullptr, Ts, c->in.jcur + r0, add_cursor ? add_cursor + r0 : nullptr, n,
                              c->p.dir_window, static_cast<unsigned long long*>(c->b_st_part.p), static_cast<unsigned long long*>(c->b_st_prefix.p),
                              trade_size ? static_cast<double*>(c->b_st_buy.p) + r0 : nullptr, trade_size ? static_cast<double*>(c->b_st_sell.p) + r0 : nullptr,
                              add_cursor ? static_cast<int32_t*>(c->b_st_adds.p) + r0 : nullptr, st));
        c->launches += trade_size ? 4 : 1;
    }
    if (trade_size) {
        c->in.buy = static_cast<const double*>(c->b_st_buy.p);
        c->in.sell = static_cast<const double*>(c->b_st_sell.p);
    }
    if (add_cursor) c->in.adds = static_cast<const int32_t*>(c->b_st_adds.p);
    return MRP_FG_OK;
}

// The normaliser's non-finite state across time shards (see k_norm_poison): [n_streams][19][3] first NaN / +inf / -inf in ABSOLUTE stream rows
int mrp_fg_poison_state(mrp_fg_ctx* c, const int64_t* row_base, int64_t* state, void* stream) {
    if (!c || !state) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = c->n_streams;
    std::vector<uint32_t> loc((size_t)S * MRP_FG_NUM_FEATURES * 3);
    CU(cudaMemcpyAsync(loc.data(), c->b_poison.p, loc.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int s = 0; s < S; ++s)
        for (int k = 0; k < MRP_FG_NUM_FEATURES * 3; ++k) {
            const uint32_t v = loc[(size_t)s * MRP_FG_NUM_FEATURES * 3 + k];
            state[(size_t)s * MRP_FG_NUM_FEATURES * 3 + k] = v == 0xffffffffu ? INT64_MAX : (int64_t)v + (row_base ? row_base[s] : 0);
        }
    return MRP_FG_OK;
}

int mrp_fg_poison_apply(mrp_fg_ctx* c, const int64_t* state_before, const int64_t* row_base, void* stream) {
    if (!c || !state_before) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = c->n_streams;
    const size_t nst = (size_t)S * MRP_FG_NUM_FEATURES * 3;
    int rc;
    if ((rc = c->b_carry.ensure(nst * sizeof(long long)))) return rc;
    if ((rc = c->b_slice_row0.ensure((size_t)S * sizeof(long long)))) return rc;
    std::vector<long long> base((size_t)S, 0);
    if (row_base)
        for (int s = 0; s < S; ++s) base[s] = row_base[s];
    // (pageable sources: staged by the runtime before the calls return)
    CU(cudaMemcpyAsync(c->b_carry.p, state_before, nst * sizeof(long long), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->b_slice_row0.p, base.data(), (size_t)S * sizeof(long long), cudaMemcpyHostToDevice, st));
    CU(launch_norm_poison(c->out, c->n_rows, c->b_poison.p, static_cast<const int64_t*>(c->b_row_off.p), S, static_cast<long long*>(c->b_carry.p),
                          static_cast<const long long*>(c->b_slice_row0.p), c->dp, st));
    c->launches += 1;
    CU(cudaStreamSynchronize(st));   // `base` leaves scope
    return MRP_FG_OK;
}

int mrp_fg_halo_query(mrp_fg_ctx* c, mrp_fg_halo_info* info, void* stream) {
    if (!c || !info) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    int rc = halo_geometry(c, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    *info = c->halo_info;
    return MRP_FG_OK;
}

int mrp_fg_halo_export(mrp_fg_ctx* c, void* dst_device, int64_t dst_capacity, const void** packed_device, mrp_fg_halo_info* info, void* stream) {
    if (!c || !info) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = halo_geometry(c, st);
    if (rc) return rc;
    const int64_t N = c->n_rows, H = c->halo_info.n_rows, a = N - H;
    const int L = c->p.depth_levels;
    if (dst_device) {
        if (dst_capacity < c->halo_info.bytes) return fail(MRP_FG_ERR_INVALID, "the destination buffer is smaller than mrp_fg_halo_query().bytes");
    } else if ((rc = c->b_halo.ensure((size_t)c->halo_info.bytes))) {
        return rc;
    }
    const int64_t K = c->halo_info.n_ticks, T = c->halo_info.n_trades;
    const HaloLayout lo = halo_layout(H, K, T, L);
    unsigned char* dst = static_cast<unsigned char*>(dst_device ? dst_device : c->b_halo.p);
    const DevInputs& in = c->in;
    auto d2d = [&](size_t off, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst + off, src, bytes, cudaMemcpyDeviceToDevice, st) : cudaSuccess; };
    auto d2d_levels = [&](size_t off, const void* src, size_t elem) {   // [L][N] rows a.. -> dense [L][H]
        return cudaMemcpy2DAsync(dst + off, (size_t)H * elem, static_cast<const unsigned char*>(src) + (size_t)a * elem, (size_t)N * elem,
                                 (size_t)H * elem, (size_t)L, cudaMemcpyDeviceToDevice, st);
    };
    CU(d2d(lo.ts, in.ts + a, (size_t)H * 8));
    CU(d2d_levels(lo.bpx, in.bid_px, 8));
    CU(d2d_levels(lo.apx, in.ask_px, 8));
    CU(d2d_levels(lo.bsz, in.bid_sz, 4));
    CU(d2d_levels(lo.asz, in.ask_sz, 4));
    CU(d2d(lo.buy, in.buy + a, (size_t)H * 8));
    CU(d2d(lo.sell, in.sell + a, (size_t)H * 8));
    CU(d2d(lo.adds, in.adds + a, (size_t)H * 4));
    CU(launch_rebase(in.kcur + a, reinterpret_cast<uint32_t*>(dst + lo.kcur), H, (uint32_t)c->halo_t0, st));
    CU(launch_rebase(in.jcur + a, reinterpret_cast<uint32_t*>(dst + lo.jcur), H, (uint32_t)c->halo_j0, st));
    CU(d2d(lo.mid, in.mid + c->halo_t0, (size_t)K * 8));
    CU(d2d(lo.spread, in.spread + c->halo_t0, (size_t)K * 8));
    CU(d2d(lo.tdir, in.tdir + c->halo_j0, (size_t)T));
    c->launches += 2;
    if (packed_device) *packed_device = dst;
    *info = c->halo_info;
    return MRP_FG_OK;
}

int mrp_fg_halo_import(mrp_fg_ctx* c, const void* packed_device, const mrp_fg_halo_info* info, void* stream) {
    if (!c || !packed_device || !info) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->bound) return fail(MRP_FG_ERR_STATE, "no inputs bound");
    if (c->n_streams != 1) return fail(MRP_FG_ERR_INVALID, "mrp_fg_halo_import needs ONE bound stream (a time shard)");
    const int64_t N = c->n_rows, H = info->n_rows, K = info->n_ticks, T = info->n_trades;
    const int L = c->p.depth_levels;
    if (H < 0 || K < 0 || T < 0 || H > N || K > c->n_ticks || T > c->n_trades) return fail(MRP_FG_ERR_INVALID, "the halo does not fit the headroom of the bound stream");
    const HaloLayout lo = halo_layout(H, K, T, L);
    if ((int64_t)lo.total != info->bytes) return fail(MRP_FG_ERR_INVALID, "halo info does not match this context's depth_levels / layout");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned char* src = static_cast<const unsigned char*>(packed_device);
    const DevInputs& in = c->in;   // the bound arrays are the caller's; the headroom in front of them is overwritten
    auto d2d = [&](const void* dstp, size_t off, size_t bytes) { return bytes ? cudaMemcpyAsync(const_cast<void*>(dstp), src + off, bytes, cudaMemcpyDeviceToDevice, st) : cudaSuccess; };
    auto d2d_levels = [&](const void* dstp, size_t off, size_t elem) {   // dense [L][H] -> the first H rows of [L][N]
        return cudaMemcpy2DAsync(const_cast<void*>(dstp), (size_t)N * elem, src + off, (size_t)H * elem, (size_t)H * elem, (size_t)L, cudaMemcpyDeviceToDevice, st);
    };
    if (H > 0) {
        CU(d2d(in.ts, lo.ts, (size_t)H * 8));
        CU(d2d_levels(in.bid_px, lo.bpx, 8));
        CU(d2d_levels(in.ask_px, lo.apx, 8));
        CU(d2d_levels(in.bid_sz, lo.bsz, 4));
        CU(d2d_levels(in.ask_sz, lo.asz, 4));
        CU(d2d(in.buy, lo.buy, (size_t)H * 8));
        CU(d2d(in.sell, lo.sell, (size_t)H * 8));
        CU(d2d(in.adds, lo.adds, (size_t)H * 4));
        CU(d2d(in.kcur, lo.kcur, (size_t)H * 4));
        CU(d2d(in.jcur, lo.jcur, (size_t)H * 4));
    }
    CU(d2d(in.mid, lo.mid, (size_t)K * 8));
    CU(d2d(in.spread, lo.spread, (size_t)K * 8));
    CU(d2d(in.tdir, lo.tdir, (size_t)T));
    return MRP_FG_OK;
}

int mrp_fg_sync(mrp_fg_ctx* c, void* stream) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    CU(cudaGetLastError());
    return MRP_FG_OK;
}

int mrp_fg_num_kernels(void) { return kNumKernels; }
const char* mrp_fg_kernel_name(int i) { return (i >= 0 && i < kNumKernels) ? kKernelNames[i] : ""; }

int mrp_fg_enable_timing(mrp_fg_ctx* c, int on) {
    if (!c) return fail(MRP_FG_ERR_INVALID, "ctx is NULL");
    c->timing = on != 0;
    c->have_times = false;
    return MRP_FG_OK;
}

int mrp_fg_kernel_times_ms(mrp_fg_ctx* c, float* ms) {
    if (!c || !ms) return fail(MRP_FG_ERR_INVALID, "NULL argument");
    if (!c->have_times) return fail(MRP_FG_ERR_STATE, "no timed run yet (mrp_fg_enable_timing, then mrp_fg_run)");
    CU(cudaEventSynchronize(c->ev[kNumKernels]));
    for (int i = 0; i < kNumKernels; ++i) CU(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
    if (c->two_streams && c->ev_aux[0]) CU(cudaEventElapsedTime(&ms[1], c->ev_aux[0], c->ev_aux[1]));   // the book pass ran beside the window pass
    return MRP_FG_OK;
}

int64_t mrp_fg_launch_count(const mrp_fg_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
