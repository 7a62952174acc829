// GEMV scan, variant 1: cp.async.bulk (TMA 1-D) slabs into an mbarrier-guarded smem ring.
// (under construction)
#include "cab_internal.h"

namespace cab {
void launch_gemv_bulk_scan(const ScanArgs &, const GemvConfig &, int, cudaStream_t) {}
}  // namespace cab
