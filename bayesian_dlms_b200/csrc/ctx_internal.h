// ctx_internal.h -- what comm.cu (the multi-GPU communicator) needs from a context beyond the
// public ABI.  Not exported from libbdlm.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bdlm.h"

namespace bdlm {

// The CUDA stream device-mode calls of this context are enqueued on.
cudaStream_t ctx_stream(bdlm_ctx *c);
// Restrict the NEXT batched call on this context to series [lo, hi) of the problem it is given
// (the arrays keep the full batch pitch B; status, per-series parameters and the RNG subsequences
// are indexed globally, so a batch cut over several contexts gives the results of one call).
// lo < 0 clears the restriction; the restriction is consumed by one call.
void ctx_set_range(bdlm_ctx *c, int64_t lo, int64_t hi);
void ctx_count_launches(bdlm_ctx *c, int64_t n);
// Peer mailboxes the bdlm_scan_dist_* calls of this context use until cleared (peers == nullptr).
struct ScanPeers;
void scan_set_peers(bdlm_ctx *c, const ScanPeers *peers, const unsigned long long *epoch_dev);

}  // namespace bdlm
