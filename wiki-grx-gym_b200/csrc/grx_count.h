// Launch accounting: every kernel launch the library issues bumps one process-wide counter (bench.py reports the difference over its
// timed region as `gpu_launches`).  Launches recorded into a CUDA graph are counted when the graph is REPLAYED, not when it is captured.
#pragma once
#include <atomic>
extern std::atomic<unsigned long long> g_grx_launches;
inline void grx_count_launch(unsigned long long n = 1) { g_grx_launches.fetch_add(n, std::memory_order_relaxed); }
