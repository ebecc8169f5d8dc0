// Error reporting, argument checks, tunables, and the stage sequencer of the C ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <utility>
#include <vector>

#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {

static thread_local char g_err[512] = "";

int set_error(int code, const char *msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char *where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return PB_ECUDA;
}
int check_launch(const char *kernel) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_cuda_error(e, kernel);
  }
  return PB_OK;
}
int check_args(const pb_pack_t *pack, const pb_cfg_t *cfg) {
  if (!pack || !cfg) return set_error(PB_EINVAL, "NULL pack or cfg");
  if (pack->nblocks <= 0) return set_error(PB_EINVAL, "pack has no blocks");
  if (cfg->ndim < 1 || cfg->ndim > 3) return set_error(PB_EINVAL, "ndim must be 1, 2 or 3");
  if (cfg->ng < 1) return set_error(PB_EINVAL, "nghost must be >= 1");
  if (!pack->blocks || !pack->geom || !pack->floor_tab || !pack->prim) return set_error(PB_EINVAL, "NULL pack field");
  const Lay L(*cfg);
  if ((long long)L.ncell * pack->nblocks * 24 >= (1LL << 40)) return set_error(PB_EINVAL, "pack too large");
  if (cfg->floor_type < 0 || cfg->floor_type > 4) return set_error(PB_EINVAL, "unknown floor type");
  return PB_OK;
}
// entry points validate the pack fields they dereference: a partly filled pb_pack_t is an argument error, not an
// asynchronous illegal-address fault
int need_fields(const char *entry, std::initializer_list<std::pair<const void *, const char *>> fields) {
  for (const auto &f : fields)
    if (!f.first) {
      char msg[160];
      snprintf(msg, sizeof msg, "%s: pack->%s is NULL", entry, f.second);
      return set_error(PB_EINVAL, msg);
    }
  return PB_OK;
}
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  const int x = atoi(v);
  return x > 0 ? x : dflt;
}
int flux_threads_target() {
  static int v = env_int("PB_FLUX_THREADS", 272);
  return v;
}
// run-time options (pb_set_option); initial values from the environment: PB_FLUX_MODE=fused|split,
// PB_C2P_ARITH=queue|exact|inline
static int g_flux_split = [] {
  const char *e = getenv("PB_FLUX_MODE");
  return (e && e[0] == 'f') ? 0 : 1;
}();
static int g_c2p_arith = [] {
  const char *e = getenv("PB_C2P_ARITH");
  return (e && e[0] == 'e') ? 0 : ((e && e[0] == 'i') ? 1 : ((e && e[0] == 's') ? 3 : 2));
}();
int sm_count() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
  }();
  return n;
}
static int g_flux_tma = [] {
  const char *e = getenv("PB_FLUX_STAGE");
  return (e && e[0] == 'c') ? 0 : 1; // cpasync | tma (default)
}();
static int g_flag_compact = [] {
  const char *e = getenv("PB_FLAG_PASS");
  return (e && e[0] == 'p') ? 0 : 1; // packed (16 zones per thread, serial) | compact (queue, one zone per thread; default)
}();
bool flag_compaction_enabled() { return g_flag_compact != 0; }
static int g_recon_cpt = [] {
  const char *e = getenv("PB_RECON_CPT");
  return (e && e[0] == '2') ? 2 : 1;
}();
int recon_cells_per_thread() { return g_recon_cpt; }
static int g_riemann_blocks = [] {
  const char *e = getenv("PB_RIEMANN_BLOCKS");
  return (e && e[0] == '1') ? 1 : 0; // measured slower than the per-warp decision (profiles/r02e_variants.txt): off by default
}();
bool riemann_block_split_enabled() { return g_riemann_blocks != 0; }
static int g_ghost_plan = [] {
  const char *e = getenv("PB_GHOST_PLAN");
  return (e && e[0] == '0') ? 0 : 1; // use pack->ghost_mask / c2p_zones when the pack carries them
}();
bool flux_split_enabled() { return g_flux_split != 0; }
bool flux_tma_enabled() { return g_flux_tma != 0; }
int c2p_arith_mode() { return g_c2p_arith; }
int zone_threads() {
  static int v = env_int("PB_ZONE_THREADS", 128);
  return v > 128 ? 128 : v; // zone kernels are compiled with __launch_bounds__(128, n)
}

} // namespace pb

using namespace pb;

extern "C" const char *pb_version(void) { return "phoebus_b200 0.1 (sm_100a, fp64, fmad=off)"; }
extern "C" const char *pb_last_error(void) { return g_err; }
extern "C" int pb_nprim(const pb_cfg_t *cfg) { return cfg ? 7 + 3 * cfg->mhd : set_error(PB_EINVAL, "NULL cfg"); }
extern "C" int pb_ncons(const pb_cfg_t *cfg) { return cfg ? 5 + 3 * cfg->mhd : set_error(PB_EINVAL, "NULL cfg"); }
extern "C" long long pb_cells_per_block(const pb_cfg_t *cfg) { return cfg ? Lay(*cfg).ncell : set_error(PB_EINVAL, "NULL cfg"); }
extern "C" int pb_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "cudaGetDevice");
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return set_cuda_error(cudaGetLastError(), "cudaDeviceGetAttribute");
  return n;
}

// ---- optional per-task device timing of the sequencer (bench.py): CUDA events recorded on the
// launching stream around each task group; read back (and reset) with pb_profile_read.
namespace pb {
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev[PB_NTASK][2];
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    cudaEvent_t e;
    if (!pool.empty()) {
      e = pool.back();
      pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
};
static Profiler g_prof;
struct ProfScope {
  int id;
  cudaStream_t s;
  ProfScope(int id_, cudaStream_t s_) : id(id_), s(s_) {
    if (g_prof.on) {
      cudaEvent_t e = g_prof.get();
      cudaEventRecord(e, s);
      g_prof.ev[id][0].push_back(e);
    }
  }
  ~ProfScope() {
    if (g_prof.on) {
      cudaEvent_t e = g_prof.get();
      cudaEventRecord(e, s);
      g_prof.ev[id][1].push_back(e);
    }
  }
};
} // namespace pb

namespace pb {
static RedoQueue g_rq = {nullptr, nullptr, nullptr};
static size_t g_rq_cap = 0;
// The library's work spaces (redo queue, con2prim scratch, error counter) are shared by all calls and used in stream
// order: they belong to ONE device and to one stream at a time. workspace_claim() enforces that: a call from another
// device, or on another stream while the previous stream still has work in flight, is refused (PB_EINVAL) instead of
// racing on the queue counters. Streams that take part in one CUDA-graph capture count as one (the capture orders them).
static int g_ws_device = -1;
static cudaStream_t g_ws_stream = nullptr;
static bool g_ws_have_stream = false;
int workspace_claim(cudaStream_t s) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "cudaGetDevice");
  if (g_ws_device >= 0 && dev != g_ws_device)
    return set_error(PB_EINVAL, "phoebus_b200 work spaces belong to another device of this process (one process per GPU)");
  g_ws_device = dev;
  if (g_ws_have_stream && s != g_ws_stream) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      const cudaError_t q = cudaStreamQuery(g_ws_stream);
      if (q == cudaErrorNotReady)
        return set_error(PB_EINVAL, "phoebus_b200 work spaces are in use on another stream (one stage in flight at a time)");
      if (q != cudaSuccess) cudaGetLastError(); // e.g. the old stream is being captured or was destroyed: not our business
    }
  }
  g_ws_stream = s;
  g_ws_have_stream = true;
  return PB_OK;
}
int redo_queue(size_t entries, RedoQueue &out, cudaStream_t s) {
  int rc = workspace_claim(s);
  if (rc) return rc;
  if (!g_rq.count) {
    if (cudaMalloc(&g_rq.count, 2 * sizeof(unsigned) + sizeof(unsigned long long) + 8) != cudaSuccess)
      return set_cuda_error(cudaGetLastError(), "redo queue counters");
    cudaMemset(g_rq.count, 0, 2 * sizeof(unsigned) + sizeof(unsigned long long) + 8);
    g_rq.total = reinterpret_cast<unsigned long long *>(g_rq.count + 2);
  }
  if (entries > g_rq_cap) {
    cudaDeviceSynchronize(); // an earlier launch may still be reading the old list
    if (g_rq.list) cudaFree(g_rq.list);
    g_rq.list = nullptr;
    if (cudaMalloc(&g_rq.list, entries * sizeof(unsigned)) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "redo queue");
    g_rq_cap = entries;
  }
  out = g_rq;
  return PB_OK;
}
unsigned long long redo_total_host() {
  unsigned long long v = 0;
  if (g_rq.total) cudaMemcpy(&v, g_rq.total, sizeof v, cudaMemcpyDeviceToHost);
  return v;
}

void prof_begin(int id, cudaStream_t s) {
  if (g_prof.on) {
    cudaEvent_t e = g_prof.get();
    cudaEventRecord(e, s);
    g_prof.ev[id][0].push_back(e);
  }
}
void prof_end(int id, cudaStream_t s) {
  if (g_prof.on) {
    cudaEvent_t e = g_prof.get();
    cudaEventRecord(e, s);
    g_prof.ev[id][1].push_back(e);
  }
}
} // namespace pb

extern "C" int pb_set_option(const char *key, int value) {
  if (key && !strcmp(key, "flux_split")) g_flux_split = value != 0;
  else if (key && !strcmp(key, "flux_tma")) g_flux_tma = value != 0;
  else if (key && !strcmp(key, "flag_compact")) g_flag_compact = value != 0;
  else if (key && !strcmp(key, "recon_cpt") && (value == 1 || value == 2)) g_recon_cpt = value;
  else if (key && !strcmp(key, "riemann_blocks")) g_riemann_blocks = value != 0;
  else if (key && !strcmp(key, "ghost_plan")) g_ghost_plan = value != 0;
  else if (key && !strcmp(key, "c2p_arith") && value >= 0 && value <= 3) g_c2p_arith = value;
  else return set_error(PB_EINVAL, "pb_set_option: unknown key or value");
  return PB_OK;
}
extern "C" int pb_get_option(const char *key) {
  if (key && !strcmp(key, "flux_split")) return g_flux_split;
  if (key && !strcmp(key, "flux_tma")) return g_flux_tma;
  if (key && !strcmp(key, "flag_compact")) return g_flag_compact;
  if (key && !strcmp(key, "recon_cpt")) return g_recon_cpt;
  if (key && !strcmp(key, "riemann_blocks")) return g_riemann_blocks;
  if (key && !strcmp(key, "ghost_plan")) return g_ghost_plan;
  if (key && !strcmp(key, "c2p_arith")) return g_c2p_arith;
  return set_error(PB_EINVAL, "pb_get_option: unknown key");
}

extern "C" int pb_profile_enable(int on) {
  pb::g_prof.on = on != 0;
  return PB_OK;
}
extern "C" int pb_profile_read(double *ms, int *calls) {
  using namespace pb;
  for (int t = 0; t < PB_NTASK; ++t) {
    double tot = 0;
    const size_t n = g_prof.ev[t][1].size();
    for (size_t i = 0; i < n; ++i) {
      float x = 0;
      cudaEventSynchronize(g_prof.ev[t][1][i]);
      if (cudaEventElapsedTime(&x, g_prof.ev[t][0][i], g_prof.ev[t][1][i]) == cudaSuccess) tot += x;
      g_prof.pool.push_back(g_prof.ev[t][0][i]);
      g_prof.pool.push_back(g_prof.ev[t][1][i]);
    }
    g_prof.ev[t][0].clear();
    g_prof.ev[t][1].clear();
    if (ms) ms[t] = tot;
    if (calls) calls[t] = (int)n;
  }
  return PB_OK;
}

namespace pb {
int launch_c2p(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int ib[2], const int jb[2], const int kb[2],
               unsigned char *zflags, cudaStream_t s);
int launch_c2p_list(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int *blocks, int nlist, int which,
                    unsigned char *zflags, cudaStream_t s);
int launch_c2p_plan(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int *blocks, int nlist, bool with_ghosts,
                    unsigned char *zflags, cudaStream_t s);
int launch_fixup(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, int mode, unsigned char *zflags, cudaStream_t s);
int launch_floors(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, unsigned long long *counters, int mode,
                  unsigned char *zflags, const double *lazy_u0, const double *lazy_ubase, double beta, cudaStream_t s);
} // namespace pb

// ---- stage sequencing, ref: src/phoebus_driver.cpp:204-738 (fluid-only sub-graph).
// With pack->zflags the fixup / floor tasks only visit the zones they can change (flags maintained by
// the c2p, fixup and floor kernels themselves) and the ghost zones of the stage-output container are
// not materialised before the halo exchange overwrites them; results are identical either way
// (tests/test_gpu_parity.py runs both).
extern "C" int pb_rk_stage_pre(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1,
                               double beta, double dt, unsigned long long *counters, void *stream) {
  int rc;
  cudaStream_t s = (cudaStream_t)stream;
  const bool flagged = pack->zflags != nullptr;
  // region 2: CalculateFluxes -> FluxCT -> FixFluxes
  {
    ProfScope p(PB_TASK_FLUX, s);
    if ((rc = pb_calculate_fluxes(pack, cfg, stream))) return rc;
  }
  {
    ProfScope p(PB_TASK_CT_FIX, s);
    if ((rc = pb_flux_ct(pack, cfg, stream))) return rc;
    if ((rc = pb_fix_fluxes(pack, cfg, stream))) return rc;
  }
  // regions 1 + 6: sources, FluxDivergence, SumData, Average, Update
  if (pack->gsrc) {
    {
      ProfScope p(PB_TASK_SOURCE, s);
      if ((rc = pb_source_terms(pack, cfg, pack->gsrc, stream))) return rc;
    }
    ProfScope p(PB_TASK_UPDATE, s);
    if ((rc = pb_update_from_sources(pack, cfg, sc0, base, beta, beta * dt, sc1, pack->gsrc, nullptr, !flagged, stream))) return rc;
  } else {
    ProfScope p(PB_TASK_UPDATE, s);
    if ((rc = pb_update_fused(pack, cfg, sc0, base, beta, beta * dt, sc1, nullptr, nullptr, !flagged, stream))) return rc;
  }
  // region 7: fixup + floors acting on the previous stage's prims / fail flags
  ProfScope p7(PB_TASK_STALE_FIXUP_FLOORS, s);
  if (flagged) {
    if ((rc = launch_fixup(pack, cfg, sc1, 2, pack->zflags, s))) return rc;
    if ((rc = launch_floors(pack, cfg, sc1, counters, 2, pack->zflags, sc0, base, beta, s))) return rc;
  } else {
    if ((rc = pb_c2p_fixup(pack, cfg, sc1, stream))) return rc;
    if ((rc = pb_apply_floors(pack, cfg, sc1, counters, stream))) return rc;
  }
  return PB_OK;
}

// Region 9 in two parts, for callers that run an inter-GPU halo exchange: pb_rk_stage_post_early may be issued
// while the transfer is in flight (after the same-GPU halo copy): it solves con2prim on the INTERIOR zones of the
// MeshBlocks that touch no mesh boundary (pack->inner_blocks) — zones no ghost fill reads or writes. Blocks on a mesh
// boundary must wait: with bc_vars = primitive the boundary fill copies the previous stage's primitives of their
// interior. pb_rk_stage_post2(early_done = 1) then does the rest: BCs, con2prim on the ghost shells of the inner
// blocks and on the entire boundary blocks, fixup, floors. Zones are independent, so the bits do not change.
// Whether starting con2prim under an inter-GPU transfer pays for this pack: the split turns one con2prim task into three
// (interiors of the inner MeshBlocks, their ghost shells, the boundary blocks); with few inner blocks (strong scaling: 16
// blocks per GPU) the extra launches cost more than the hidden transfer (measured at 8 GPUs).
extern "C" int pb_early_c2p_pays(const pb_pack_t *pack) {
  if (!pack || !pack->inner_blocks || !pack->bc_blocks || pack->n_inner_blocks <= 0) return 0;
  if (pack->n_inner_blocks + pack->n_bc_blocks != pack->nblocks) return 0;
  return 3 * pack->n_inner_blocks >= pack->nblocks ? 1 : 0;
}

// Whether the stage sequencer uses the pack's ghost plan (ghost_plan_host.cpp): callers that issue the same-GPU ghost copy
// themselves (inter-GPU exchange) pass phase = pb_ghost_plan_active(...) to pb_halo_exchange_local_ex.
extern "C" int pb_ghost_plan_active(const pb_pack_t *pack, const pb_cfg_t *cfg) {
  if (!pack || !cfg || !g_ghost_plan || !cfg->bc_prim) return 0;
  return (pack->ghost_mask && pack->c2p_zones) ? 1 : 0;
}

extern "C" int pb_rk_stage_post_early(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc1, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!pack->inner_blocks || pack->n_inner_blocks <= 0) return PB_OK;
  if (!pack->bc_blocks || pack->n_inner_blocks + pack->n_bc_blocks != pack->nblocks)
    return set_error(PB_EINVAL, "pb_rk_stage_post_early: inner_blocks and bc_blocks must partition the pack");
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope p(PB_TASK_C2P, s);
  if (pb_ghost_plan_active(pack, cfg)) return launch_c2p_plan(pack, cfg, sc1, pack->inner_blocks, pack->n_inner_blocks, false, pack->zflags, s);
  return launch_c2p_list(pack, cfg, sc1, pack->inner_blocks, pack->n_inner_blocks, 0, pack->zflags, s);
}

extern "C" int pb_rk_stage_post2(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc1, unsigned long long *counters,
                                 int early_done, void *stream) {
  int rc;
  cudaStream_t s = (cudaStream_t)stream;
  // the two lists must partition the pack, otherwise some block would get no con2prim
  if (early_done && pack->inner_blocks && pack->n_inner_blocks > 0 &&
      (!pack->bc_blocks || pack->n_inner_blocks + pack->n_bc_blocks != pack->nblocks))
    return set_error(PB_EINVAL, "pb_rk_stage_post2: inner_blocks and bc_blocks must partition the pack");
  const bool split = early_done && pack->inner_blocks && pack->n_inner_blocks > 0;
  // region 9: BCs -> c2p over the entire block -> fixup -> floors
  {
    ProfScope p(PB_TASK_BCS, s);
    if ((rc = pb_apply_bcs(pack, cfg, sc1, stream))) return rc;
  }
  const Lay L(*cfg);
  const int ib[2] = {0, L.ni - 1}, jb[2] = {0, L.nj - 1}, kb[2] = {0, L.nk - 1};
  const bool plan = pb_ghost_plan_active(pack, cfg) != 0;
  if (plan) {
    // con2prim where it has to be solved, then the deferred ghost zones take the neighbour's result (before the fixup,
    // which reads ghost primitives and fail flags as con2prim left them)
    {
      ProfScope p(PB_TASK_C2P, s);
      if (split) rc = launch_c2p_plan(pack, cfg, sc1, pack->bc_blocks, pack->n_bc_blocks, true, pack->zflags, s);
      else rc = launch_c2p_plan(pack, cfg, sc1, nullptr, pack->nblocks, true, pack->zflags, s);
      if (rc) return rc;
    }
    ProfScope p(PB_TASK_HALO, s);
    if ((rc = pb_halo_exchange_local_ex(pack, cfg, sc1, 2, stream))) return rc;
  } else {
    ProfScope p(PB_TASK_C2P, s);
    if (split) {
      if ((rc = launch_c2p_list(pack, cfg, sc1, pack->inner_blocks, pack->n_inner_blocks, 1, pack->zflags, s))) return rc;
      if ((rc = launch_c2p_list(pack, cfg, sc1, pack->bc_blocks, pack->n_bc_blocks, 2, pack->zflags, s))) return rc;
    } else if (pack->zflags) {
      if ((rc = launch_c2p(pack, cfg, sc1, ib, jb, kb, pack->zflags, s))) return rc;
    } else {
      if ((rc = pb_cons_to_prim(pack, cfg, sc1, ib, jb, kb, stream))) return rc;
    }
  }
  ProfScope p(PB_TASK_FIXUP_FLOORS, s);
  if (pack->zflags) {
    if ((rc = launch_fixup(pack, cfg, sc1, 1, pack->zflags, s))) return rc;
    if ((rc = launch_floors(pack, cfg, sc1, counters, 1, pack->zflags, nullptr, nullptr, 1.0, s))) return rc;
  } else {
    if ((rc = pb_c2p_fixup(pack, cfg, sc1, stream))) return rc;
    if ((rc = pb_apply_floors(pack, cfg, sc1, counters, stream))) return rc;
  }
  return PB_OK;
}

extern "C" int pb_rk_stage_post(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc1, unsigned long long *counters,
                                void *stream) {
  return pb_rk_stage_post2(pack, cfg, sc1, counters, 0, stream);
}

extern "C" int pb_rk_stage(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1,
                           double beta, double dt, unsigned long long *counters, void *stream) {
  int rc;
  if ((rc = pb_rk_stage_pre(pack, cfg, sc0, base, sc1, beta, dt, counters, stream))) return rc;
  // region 8: ghost exchange
  {
    ProfScope p(PB_TASK_HALO, (cudaStream_t)stream);
    if ((rc = pb_halo_exchange_local_ex(pack, cfg, sc1, pb_ghost_plan_active(pack, cfg), stream))) return rc;
  }
  return pb_rk_stage_post(pack, cfg, sc1, counters, stream);
}
