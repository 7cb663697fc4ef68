// mbo_synth_reader.cpp — link-time substitute for the reference's databento-backed DbnMboReader
// (data_ingestion/include/dbn_reader.hpp:9-36; the real one, data_ingestion/src/data/dbn_reader.cpp, needs databento-cpp, which is
// not vendored).  TEST INFRASTRUCTURE ONLY (see fg_oracle.h): it lets the reference's OWN pipeline — EventParser, OrderEngine,
// OrderBookManager, FeatureEngine, DualFeaturePipeline::run, CsvWriter — run end to end on a seeded synthetic MBO day
// (BASELINE.json configs[0]).  No feature arithmetic here: only an event source.
//
// The stream: a limit-order book round a slowly walking mid (ticks of 0.25), orders added 1..12 ticks from the mid, cancelled at
// random, levels that the mid walks into are cancelled first (the book never crosses), trades at the touch; event times advance
// by a random 0.2-4.4 ms from 60 s before the NYSE open of the pipeline's date (dual_feature_pipeline.cpp:147-158), so a full
// session is ~10 M events per instrument.  Environment: MRP_SYN_EVENTS (events per instrument, default 400000), MRP_SYN_SEED.
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <deque>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "dbn_reader.hpp"

namespace {
struct Rng {
    uint64_t s;
    uint64_t next() {   // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
};
uint64_t nyse_open_ns(const std::string& path) {
    // the reader is built from ".../<venue>-YYYYMMDD.mbo.dbn.zst" (dual_feature_pipeline.cpp:128-137)
    std::string date = "20250505";
    const size_t p = path.find(".mbo");
    if (p != std::string::npos && p >= 8) date = path.substr(p - 8, 8);
    std::tm t = {};
    t.tm_year = std::stoi(date.substr(0, 4)) - 1900;
    t.tm_mon = std::stoi(date.substr(4, 2)) - 1;
    t.tm_mday = std::stoi(date.substr(6, 2));
    t.tm_hour = 13;
    t.tm_min = 30;
    return (uint64_t)timegm(&t) * 1000000000ull;
}
}  // namespace

struct DbnMboReader::Impl {
    Rng rng;
    std::string instrument;
    uint64_t ts = 0, n_left = 0, next_id = 1;
    int mid = 20000;                                  // in ticks of 0.25
    std::map<int, std::vector<uint64_t>> bids, asks;  // price tick -> live order ids
    std::deque<MarketEvent> pending;                  // cancels that keep the book uncrossed, emitted before the next random event
    uint32_t seq = 0;

    MarketEvent make(char action, char side, int price_tick, int size, uint64_t id) {
        MarketEvent e{};
        ts += 200000 + rng.below(4200000);
        e.timestamp_ns = ts;
        e.instrument = instrument;
        e.action = action;
        e.side = side;
        e.price = price_tick * 0.25;
        e.size = size;
        e.order_id = id;
        e.flags = 0;
        e.instrument_id = 4916;                      // order_engine.cpp:40: ES events must carry this id
        e.channel_id = 0;
        e.sequence = ++seq;
        return e;
    }
    void drop_level(std::map<int, std::vector<uint64_t>>& book, int tick, char side) {
        auto it = book.find(tick);
        if (it == book.end()) return;
        for (uint64_t id : it->second) pending.push_back(make('C', side, tick, 0, id));
        book.erase(it);
    }
    MarketEvent step() {
        if (!pending.empty()) {
            MarketEvent e = pending.front();
            pending.pop_front();
            e.timestamp_ns = (ts += 1000);
            return e;
        }
        const uint32_t u = rng.below(1000);
        if (u < 6) {   // the mid walks one tick: clear the level it walks into first
            const int d = (rng.next() & 1) ? 1 : -1;
            if (d > 0) drop_level(asks, mid + 1, 'A');
            else drop_level(bids, mid - 1, 'B');
            mid += d;
            if (mid < 4000) mid = 4000;
            if (!pending.empty()) return step();
        }
        if (u < 620) {   // add
            const bool bid = rng.next() & 1;
            const int k = 1 + (int)rng.below(12);
            const int tick = bid ? mid - k : mid + k;
            const uint64_t id = next_id++;
            (bid ? bids : asks)[tick].push_back(id);
            return make('A', bid ? 'B' : 'A', tick, 1 + (int)rng.below(200), id);
        }
        if (u < 920) {   // cancel a random live order
            auto& book = (rng.next() & 1) ? bids : asks;
            if (!book.empty()) {
                auto it = book.begin();
                std::advance(it, rng.below((uint32_t)book.size()));
                if (!it->second.empty()) {
                    const size_t i = rng.below((uint32_t)it->second.size());
                    const uint64_t id = it->second[i];
                    it->second.erase(it->second.begin() + i);
                    const int tick = it->first;
                    if (it->second.empty()) book.erase(it);
                    return make('C', &book == &bids ? 'B' : 'A', tick, 0, id);
                }
            }
            return step();
        }
        // trade at the touch
        const bool buy = rng.below(1000) < 540;
        return make('T', buy ? 'B' : 'A', buy ? mid + 1 : mid - 1, 1 + (int)rng.below(50), 0);
    }
};

DbnMboReader::DbnMboReader(const std::string& filepath, const std::string& instrument) : pimpl_(new Impl), instrument_(instrument) {
    const char* ne = std::getenv("MRP_SYN_EVENTS");
    const char* se = std::getenv("MRP_SYN_SEED");
    uint64_t seed = se ? std::strtoull(se, nullptr, 10) : 0x5EEDull;
    for (char c : instrument) seed = seed * 1315423911ull + (unsigned char)c;
    pimpl_->rng.s = seed;
    pimpl_->instrument = instrument;
    pimpl_->n_left = ne ? std::strtoull(ne, nullptr, 10) : 400000ull;
    pimpl_->ts = nyse_open_ns(filepath) - 60ull * 1000000000ull;
    // order ids are unique across the two instruments of a pipeline: OrderEngine tracks them in ONE map (order_engine.cpp:12-18)
    pimpl_->next_id = ((seed >> 8) & 0xfffffull) << 40 | 1ull;
}
DbnMboReader::~DbnMboReader() = default;
bool DbnMboReader::has_next() const { return pimpl_->n_left > 0; }
MarketEvent DbnMboReader::next_event() {
    if (pimpl_->n_left == 0) throw std::runtime_error("DbnMboReader: end of the synthetic stream");
    --pimpl_->n_left;
    ++event_count_;
    return pimpl_->step();
}
