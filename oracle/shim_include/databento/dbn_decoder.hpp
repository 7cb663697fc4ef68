// Intentionally empty: stands in for the un-vendored databento-cpp header that data_ingestion/include/dbn_reader.hpp:6
// includes.  DbnMboReader is PIMPL (dbn_reader.hpp:26-29), so nothing of databento is needed to compile its users.
#pragma once
