// Error reporting, argument checks, tunables, and the stage sequencer of the C ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
int zone_threads() {
  static int v = env_int("PB_ZONE_THREADS", 128);
  return v > 128 ? 128 : v; // zone kernels are compiled with __launch_bounds__(128, n)
}

} // namespace pb

using namespace pb;

extern "C" const char *pb_version(void) { return "phoebus_b200 0.1 (sm_100a, fp64, fmad=off)"; }
extern "C" const char *pb_last_error(void) { return g_err; }
extern "C" int pb_nprim(const pb_cfg_t *cfg) { return 7 + 3 * cfg->mhd; }
extern "C" int pb_ncons(const pb_cfg_t *cfg) { return 5 + 3 * cfg->mhd; }
extern "C" long long pb_cells_per_block(const pb_cfg_t *cfg) { return Lay(*cfg).ncell; }
extern "C" int pb_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "cudaGetDevice");
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return set_cuda_error(cudaGetLastError(), "cudaDeviceGetAttribute");
  return n;
}

// ---- stage sequencing, ref: src/phoebus_driver.cpp:204-738 (fluid-only sub-graph)
extern "C" int pb_rk_stage_pre(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1,
                               double beta, double dt, unsigned long long *counters, void *stream) {
  int rc;
  // region 2: CalculateFluxes -> FluxCT -> FixFluxes
  if ((rc = pb_calculate_fluxes(pack, cfg, stream))) return rc;
  if ((rc = pb_flux_ct(pack, cfg, stream))) return rc;
  if ((rc = pb_fix_fluxes(pack, cfg, stream))) return rc;
  // regions 1 + 6: sources, FluxDivergence, SumData, Average, Update (fused)
  if ((rc = pb_update_fused(pack, cfg, sc0, base, beta, beta * dt, sc1, nullptr, nullptr, stream))) return rc;
  // region 7: fixup + floors acting on the previous stage's prims / fail flags
  if ((rc = pb_c2p_fixup(pack, cfg, sc1, stream))) return rc;
  if ((rc = pb_apply_floors(pack, cfg, sc1, counters, stream))) return rc;
  return PB_OK;
}

extern "C" int pb_rk_stage_post(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc1, unsigned long long *counters,
                                void *stream) {
  int rc;
  // region 9: BCs -> c2p over the entire block -> fixup -> floors
  if ((rc = pb_apply_bcs(pack, cfg, sc1, stream))) return rc;
  const Lay L(*cfg);
  const int ib[2] = {0, L.ni - 1}, jb[2] = {0, L.nj - 1}, kb[2] = {0, L.nk - 1};
  if ((rc = pb_cons_to_prim(pack, cfg, sc1, ib, jb, kb, stream))) return rc;
  if ((rc = pb_c2p_fixup(pack, cfg, sc1, stream))) return rc;
  if ((rc = pb_apply_floors(pack, cfg, sc1, counters, stream))) return rc;
  return PB_OK;
}

extern "C" int pb_rk_stage(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1,
                           double beta, double dt, unsigned long long *counters, void *stream) {
  int rc;
  if ((rc = pb_rk_stage_pre(pack, cfg, sc0, base, sc1, beta, dt, counters, stream))) return rc;
  // region 8: ghost exchange
  if ((rc = pb_halo_exchange_local(pack, cfg, sc1, stream))) return rc;
  return pb_rk_stage_post(pack, cfg, sc1, counters, stream);
}
