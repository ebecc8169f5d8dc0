// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#ifndef PB_HOST_HPP_
#define PB_HOST_HPP_

#include <cuda_runtime.h>

#include "../../include/phoebus_b200.h"

namespace pb {

int set_error(int code, const char *msg);
int set_cuda_error(cudaError_t e, const char *where);
// cudaPeekAtLastError after a launch (does not synchronise)
int check_launch(const char *kernel);
// common argument validation for tasks
int check_args(const pb_pack_t *pack, const pb_cfg_t *cfg);
// tunables (environment overrides, read once): PB_FLUX_THREADS, PB_ZONE_THREADS
int flux_threads_target();
int zone_threads();
bool flux_split_enabled(); // PB_FLUX_MODE=fused|split (default split when the pack carries ql/qr scratch)

// per-task device timing (pb_profile_enable): events on the launching stream, no-ops when off
void prof_begin(int task, cudaStream_t s);
void prof_end(int task, cudaStream_t s);

} // namespace pb
#endif
