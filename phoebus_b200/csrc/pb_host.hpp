// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#ifndef PB_HOST_HPP_
#define PB_HOST_HPP_

#include <cuda_runtime.h>

#include <initializer_list>
#include <utility>

#include "../../include/phoebus_b200.h"

namespace pb {

int set_error(int code, const char *msg);
int set_cuda_error(cudaError_t e, const char *where);
// cudaPeekAtLastError after a launch (does not synchronise)
int check_launch(const char *kernel);
// common argument validation for tasks
int check_args(const pb_pack_t *pack, const pb_cfg_t *cfg);
// the pack fields one entry point dereferences must be there: PB_EINVAL naming the missing one
int need_fields(const char *entry, std::initializer_list<std::pair<const void *, const char *>> fields);
#define PB_NEED(entry, ...)                                   \
  do {                                                        \
    const int need_rc_ = need_fields(entry, {__VA_ARGS__});   \
    if (need_rc_) return need_rc_;                            \
  } while (0)
#define PB_F(pack, name) std::pair<const void *, const char *>((pack)->name, #name)
// tunables (environment overrides, read once): PB_FLUX_THREADS, PB_ZONE_THREADS
int flux_threads_target();
int zone_threads();
bool riemann_block_split_enabled(); // option riemann_blocks / PB_RIEMANN_BLOCKS=1|0: unmagnetised MeshBlocks get their own Riemann launch
int recon_cells_per_thread();   // option recon_cpt / PB_RECON_CPT=1|2: cells per thread in the reconstruction launch
bool flag_compaction_enabled(); // option flag_compact / PB_FLAG_PASS=compact|packed: flag-driven fixup / floor passes
bool flux_tma_enabled();   // option flux_tma / PB_FLUX_STAGE=tma|cpasync: how the fused flux kernel stages its stencil tile
bool flux_split_enabled(); // option flux_split / PB_FLUX_MODE=fused|split (default split when the pack carries ql/qr scratch)
int c2p_arith_mode();      // option c2p_arith / PB_C2P_ARITH: 0 exact, 1 inline, 2 one launch + parking + redo queue (default), 3 split
int sm_count();            // SMs of the current device (grid sizing of the persistent / grid-stride launches)

// Device queue of work items (linear thread indices) that the deferred-branch arithmetic (ArithFast, pb_device.cuh)
// rejected; a *_redo_kernel recomputes exactly those items with the compiler's divide / sqrt right after, on the same
// stream. Library-owned workspace, grown on demand (first call / larger pack), shared by all users because every use
// is stream-ordered: memset count -> fast kernel -> redo kernel.
struct RedoQueue {
  unsigned *count;           // entries appended by the current launch
  unsigned *list;            // linear thread index of each rejected item
  unsigned long long *total; // accumulated over the life of the library (diagnostics: pb_arith_redo_count)
};
int redo_queue(size_t entries, RedoQueue &out, cudaStream_t s);
// one device and one stream at a time own the library's work spaces; PB_EINVAL otherwise
int workspace_claim(cudaStream_t s);
unsigned long long redo_total_host();

// per-task device timing (pb_profile_enable): events on the launching stream, no-ops when off
void prof_begin(int task, cudaStream_t s);
void prof_end(int task, cudaStream_t s);

} // namespace pb
#endif
