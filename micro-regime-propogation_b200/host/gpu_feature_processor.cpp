#include "gpu_feature_processor.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace microregime {

[[noreturn]] void GpuFeatureProcessor::fail(const char* what) const {
    throw std::runtime_error(std::string("GpuFeatureProcessor(") + symbol_ + "): " + what + ": " + mrp_fg_last_error());
}

GpuFeatureProcessor::GpuFeatureProcessor(std::string symbol, int device, const mrp_fg_params* params) : symbol_(std::move(symbol)) {
    mrp_fg_params p;
    if (params) p = *params;
    else {
        mrp_fg_default_params(&p);
        p.depth_levels = (int)mrp_host::kDepthLevels;
    }
    if (p.depth_levels != (int)mrp_host::kDepthLevels)
        throw std::runtime_error("GpuFeatureProcessor: params.depth_levels must equal the snapshot type's DEPTH_LEVELS");
    levels_ = p.depth_levels;
    params_ = p;
    if (mrp_fg_create(&p, device, &ctx_) != MRP_FG_OK) fail("mrp_fg_create");
}

GpuFeatureProcessor::~GpuFeatureProcessor() {
    detach(this);
    if (ctx_) mrp_fg_destroy(ctx_);
}

void GpuFeatureProcessor::on_tick(double midprice, double spread) {
    mid_.push_back(midprice);
    spread_.push_back(spread);
}

void GpuFeatureProcessor::on_trade(std::int8_t direction) { tdir_.push_back(direction); }

void GpuFeatureProcessor::stage(const FeatureInputSnapshot& s, std::uint64_t grid_timestamp_ns) {
    if (mid_.size() > 0xffff0000ull || tdir_.size() > 0xffff0000ull) throw std::runtime_error("GpuFeatureProcessor: more than 2^32-65536 ticks / trades in one stream");
    ts_.push_back(s.timestamp_ns);
    grid_.push_back(grid_timestamp_ns);
    instrument_.push_back(s.instrument);
    for (int i = 0; i < levels_; ++i) {
        bid_px_.push_back(s.bid_prices[i]);
        ask_px_.push_back(s.ask_prices[i]);
        bid_sz_.push_back(s.bid_sizes[i]);
        ask_sz_.push_back(s.ask_sizes[i]);
    }
    buy_.push_back(s.rolling_buy_volume);
    sell_.push_back(s.rolling_sell_volume);
    adds_.push_back(s.adds_since_last_snapshot);
    kcur_.push_back((std::uint32_t)mid_.size());
    jcur_.push_back((std::uint32_t)tdir_.size());
}

// ---- one row per call, through the reference's own signatures
FeatureSet GpuFeatureProcessor::GetRawFeatureSet(const FeatureInputSnapshot& snapshot) {
    stage(snapshot, snapshot.timestamp_ns);
    const std::int64_t t = (std::int64_t)ts_.size() - 1;
    // rows that can still influence row t: its z-score window, the EMA warm-up of the window's first row, the lag-21 midprice
    // and the previous row of the book deltas — the engine's own rule (mrp_fg_halo_rows_for)
    const std::int64_t H = mrp_fg_halo_rows_for(&params_);
    const std::int64_t a = std::max<std::int64_t>(0, t - H);
    const std::int64_t n = t - a + 1;
    // ticks / trades those rows can see
    const std::int64_t wmax = std::max<std::int64_t>(params_.mid_window, (std::int64_t)params_.dir_window + 1);
    const std::int64_t k_first = a > 0 ? (std::int64_t)kcur_[a - 1] : 0, j_first = a > 0 ? (std::int64_t)jcur_[a - 1] : 0;
    const std::int64_t tick0 = a > 0 ? std::max<std::int64_t>(0, k_first - wmax - 2) : 0;
    const std::int64_t trade0 = a > 0 ? std::max<std::int64_t>(0, j_first - params_.dir_window - 2) : 0;
    const int L = levels_;
    std::vector<double> bp((std::size_t)n * L), ap((std::size_t)n * L);
    std::vector<std::int32_t> bs((std::size_t)n * L), as((std::size_t)n * L);
    std::vector<std::uint32_t> kc((std::size_t)n), jc((std::size_t)n);
    for (std::int64_t r = 0; r < n; ++r) {
        const std::size_t g = (std::size_t)(a + r);
        for (int i = 0; i < L; ++i) {
            bp[(std::size_t)i * n + r] = bid_px_[g * L + i];
            ap[(std::size_t)i * n + r] = ask_px_[g * L + i];
            bs[(std::size_t)i * n + r] = bid_sz_[g * L + i];
            as[(std::size_t)i * n + r] = ask_sz_[g * L + i];
        }
        kc[(std::size_t)r] = (std::uint32_t)((std::int64_t)kcur_[g] - tick0);
        jc[(std::size_t)r] = (std::uint32_t)((std::int64_t)jcur_[g] - trade0);
    }
    const std::int64_t row_off[2] = {0, n}, tick_off[2] = {0, (std::int64_t)mid_.size() - tick0},
                       trade_off[2] = {0, (std::int64_t)tdir_.size() - trade0};
    mrp_fg_soa_view v{};
    v.n_streams = 1;
    v.memory = MRP_MEM_HOST;
    v.row_offsets = row_off;
    v.tick_offsets = tick_off;
    v.trade_offsets = trade_off;
    v.ts_ns = ts_.data() + a;
    v.bid_px = bp.data();
    v.ask_px = ap.data();
    v.bid_sz = bs.data();
    v.ask_sz = as.data();
    v.buy_vol = buy_.data() + a;
    v.sell_vol = sell_.data() + a;
    v.adds = adds_.data() + a;
    v.tick_cursor = kc.data();
    v.trade_cursor = jc.data();
    v.mid = mid_.data() + tick0;
    v.spread = spread_.data() + tick0;
    v.trade_dir = tdir_.data() + trade0;
    double raw[MRP_FG_NUM_FEATURES];
    if (mrp_fg_upload(ctx_, &v, nullptr) != MRP_FG_OK) fail("mrp_fg_upload");
    if (mrp_fg_run(ctx_, nullptr) != MRP_FG_OK) fail("mrp_fg_run");
    if (mrp_fg_download_rows(ctx_, n - 1, n, raw, last_norm_, nullptr) != MRP_FG_OK) fail("mrp_fg_download_rows");
    if (mrp_fg_sync(ctx_, nullptr) != MRP_FG_OK) fail("mrp_fg_sync");
    last_valid_ = true;
    FeatureSet f;
    f.timestamp_ns = snapshot.timestamp_ns;   // feature_processor.cpp:13
    f.instrument = snapshot.instrument;       // :14
    for (int c = 0; c < MRP_FG_NUM_FEATURES; ++c) f.*(mrp_host::kFeatureMembers[c]) = raw[c];
    return f;
}

FeatureSet GpuFeatureProcessor::GetProcessedFeatureSet(const FeatureSet& raw_feature_set) {
    if (!last_valid_) throw std::runtime_error("GpuFeatureProcessor::GetProcessedFeatureSet: no row has been computed (call GetRawFeatureSet first)");
    FeatureSet f;
    f.timestamp_ns = raw_feature_set.timestamp_ns;   // feature_normalizer.cpp:57-58
    f.instrument = raw_feature_set.instrument;
    for (int c = 0; c < MRP_FG_NUM_FEATURES; ++c) f.*(mrp_host::kFeatureMembers[c]) = last_norm_[c];
    return f;
}

void GpuFeatureProcessor::clear() {
    ts_.clear(); grid_.clear(); instrument_.clear();
    bid_px_.clear(); ask_px_.clear(); bid_sz_.clear(); ask_sz_.clear();
    buy_.clear(); sell_.clear(); adds_.clear(); kcur_.clear(); jcur_.clear();
    mid_.clear(); spread_.clear(); tdir_.clear();
}

void GpuFeatureProcessor::run(std::vector<double>& raw, std::vector<double>& norm) {
    const std::int64_t n = (std::int64_t)ts_.size();
    raw.assign((std::size_t)n * MRP_FG_NUM_FEATURES, 0.0);
    norm.assign((std::size_t)n * MRP_FG_NUM_FEATURES, 0.0);
    if (n == 0) return;
    const int L = levels_;
    // level-major [L][n] as the C ABI wants it
    std::vector<double> bp((std::size_t)n * L), ap((std::size_t)n * L);
    std::vector<std::int32_t> bs((std::size_t)n * L), as((std::size_t)n * L);
    for (std::int64_t r = 0; r < n; ++r)
        for (int i = 0; i < L; ++i) {
            bp[(std::size_t)i * n + r] = bid_px_[(std::size_t)r * L + i];
            ap[(std::size_t)i * n + r] = ask_px_[(std::size_t)r * L + i];
            bs[(std::size_t)i * n + r] = bid_sz_[(std::size_t)r * L + i];
            as[(std::size_t)i * n + r] = ask_sz_[(std::size_t)r * L + i];
        }
    const std::int64_t row_off[2] = {0, n}, tick_off[2] = {0, (std::int64_t)mid_.size()}, trade_off[2] = {0, (std::int64_t)tdir_.size()};
    mrp_fg_soa_view v{};
    v.n_streams = 1;
    v.memory = MRP_MEM_HOST;
    v.row_offsets = row_off;
    v.tick_offsets = tick_off;
    v.trade_offsets = trade_off;
    v.ts_ns = ts_.data();
    v.bid_px = bp.data();
    v.ask_px = ap.data();
    v.bid_sz = bs.data();
    v.ask_sz = as.data();
    v.buy_vol = buy_.data();
    v.sell_vol = sell_.data();
    v.adds = adds_.data();
    v.tick_cursor = kcur_.data();
    v.trade_cursor = jcur_.data();
    v.mid = mid_.data();
    v.spread = spread_.data();
    v.trade_dir = tdir_.data();
    // host arrays in, host features out: chunks of the day are uploaded / computed / downloaded on two alternating lanes
    if (mrp_fg_process_host(ctx_, &v, raw.data(), norm.data(), 0) != MRP_FG_OK) fail("mrp_fg_process_host");
}

void GpuFeatureProcessor::flush_columns(std::vector<double>& raw, std::vector<double>& norm, std::vector<std::uint64_t>& grid_ts) {
    run(raw, norm);
    grid_ts = grid_;
    clear();
}

namespace {
void fill(FeatureSet& f, const std::vector<double>& cols, std::int64_t n, std::int64_t r) {
    for (int c = 0; c < MRP_FG_NUM_FEATURES; ++c) f.*(mrp_host::kFeatureMembers[c]) = cols[(std::size_t)c * n + r];
}
}  // namespace

void GpuFeatureProcessor::flush(DataReciever& sink) {
    std::vector<double> raw, norm;
    run(raw, norm);
    const std::int64_t n = (std::int64_t)ts_.size();
    FeatureSet a, b;
    for (std::int64_t r = 0; r < n; ++r) {
        a.timestamp_ns = b.timestamp_ns = ts_[r];       // feature_processor.cpp:13, feature_normalizer.cpp:57-58
        a.instrument = b.instrument = instrument_[r];   // :14
        fill(a, raw, n, r);
        fill(b, norm, n, r);
        sink.ingest_feature_set(symbol_, grid_[r], a, b);
    }
    clear();
}

void GpuFeatureProcessor::flush_pair(GpuFeatureProcessor& base, GpuFeatureProcessor& future, DataReciever& base_sink, DataReciever& future_sink) {
    if (base.ts_.size() != future.ts_.size()) throw std::runtime_error("GpuFeatureProcessor::flush_pair: the two instruments staged different row counts");
    std::vector<double> braw, bnorm, fraw, fnorm;
    base.run(braw, bnorm);
    future.run(fraw, fnorm);
    const std::int64_t n = (std::int64_t)base.ts_.size();
    FeatureSet a, b;
    for (std::int64_t r = 0; r < n; ++r) {
        a.timestamp_ns = b.timestamp_ns = base.ts_[r];
        a.instrument = b.instrument = base.instrument_[r];
        fill(a, braw, n, r);
        fill(b, bnorm, n, r);
        base_sink.ingest_feature_set(base.symbol_, base.grid_[r], a, b);
        a.timestamp_ns = b.timestamp_ns = future.ts_[r];
        a.instrument = b.instrument = future.instrument_[r];
        fill(a, fraw, n, r);
        fill(b, fnorm, n, r);
        future_sink.ingest_feature_set(future.symbol_, future.grid_[r], a, b);
    }
    base.clear();
    future.clear();
}

// ---- trade hook registry (a pipeline holds two engines; a flat list is enough)
namespace {
struct Attached {
    const void* engine;
    GpuFeatureProcessor* processor;
};
std::vector<Attached>& attached() {
    static std::vector<Attached> v;
    return v;
}
}  // namespace

void GpuFeatureProcessor::attach(const void* feature_engine, GpuFeatureProcessor* processor) {
    for (Attached& a : attached())
        if (a.engine == feature_engine) {
            a.processor = processor;
            return;
        }
    attached().push_back({feature_engine, processor});
}

void GpuFeatureProcessor::detach(const GpuFeatureProcessor* processor) {
    for (Attached& a : attached())
        if (a.processor == processor) a.processor = nullptr;
}

void mrp_trade_hook(const void* feature_engine, std::int8_t direction) {
    for (const Attached& a : attached())
        if (a.engine == feature_engine && a.processor) a.processor->on_trade(direction);
}

}  // namespace microregime
