// FluxCT, FixFluxes, FluxDivergence, CalculateFluidSourceTerms, WeightedSumData and their fusion
// (divergence + geometric sources + RK update in one pass over the interior zones of a pack).
#include <vector>

#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {

PB_DEV bool interior_zone(const Lay &L, int nblocks, int ei, int ej, int ek, int &b, int &k, int &j, int &i) {
  // zones [is, ie+ei] x [js, je+ej] x [ks, ke+ek]
  const int ni = L.ie - L.is + 1 + ei, nj = L.je - L.js + 1 + ej, nk = L.ke - L.ks + 1 + ek;
  const long long nreg = (long long)ni * nj * nk;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nreg * nblocks) return false;
  b = (int)(t / nreg);
  int rem = (int)(t % nreg);
  i = L.is + rem % ni;
  rem /= ni;
  j = L.js + rem % nj;
  k = L.ks + rem / nj;
  return true;
}

// ------------------------------------------------------------------ FluxCT
// ref: src/fluid/fluid.cpp:775-841 (Toth flux-CT). Two launches: the B-flux update reads EMFs of
// neighbouring corners, and it overwrites the fluxes the EMFs are built from.
template <int NDIM> __global__ void emf_kernel(const pb_pack_t pack, const pb_cfg_t cfg) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!interior_zone(L, pack.nblocks, 1, 1, NDIM == 3 ? 1 : 0, b, k, j, i)) return;
  constexpr int NCn = 8;
  const int N = L.ncell, o = L.idx(k, j, i);
  const int sj = L.ni, sk = L.ni * L.nj;
  const Real *fb = pack.flux + (size_t)b * L.ndim * NCn * N;
  auto F = [&](int dir, int bc) { return fb + ((size_t)(dir - 1) * NCn + CB + bc) * N + o; };
  Real *emf = pack.emf + (size_t)b * 3 * N + o;
  if (NDIM == 2) {
    emf[0] = 0.25 * (F(1, 1)[0] + F(1, 1)[-sj] - F(2, 0)[0] - F(2, 0)[-1]);
  } else {
    emf[0] = 0.25 * (F(2, 2)[0] + F(2, 2)[-sk] - F(3, 1)[0] - F(3, 1)[-sj]);
    emf[(size_t)N] = -0.25 * (F(1, 2)[0] + F(1, 2)[-sk] - F(3, 0)[0] - F(3, 0)[-1]);
    emf[(size_t)2 * N] = 0.25 * (F(1, 1)[0] + F(1, 1)[-sj] - F(2, 0)[0] - F(2, 0)[-1]);
  }
}
template <int NDIM> __global__ void ct_flux_kernel(const pb_pack_t pack, const pb_cfg_t cfg) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!interior_zone(L, pack.nblocks, 1, 1, NDIM == 3 ? 1 : 0, b, k, j, i)) return;
  constexpr int NCn = 8;
  const int N = L.ncell, o = L.idx(k, j, i);
  const int sj = L.ni, sk = L.ni * L.nj;
  Real *fb = pack.flux + (size_t)b * L.ndim * NCn * N;
  auto F = [&](int dir, int bc) -> Real & { return fb[((size_t)(dir - 1) * NCn + CB + bc) * N + o]; };
  const Real *emf = pack.emf + (size_t)b * 3 * N + o;
  if (NDIM == 2) {
    F(1, 0) = 0.0;
    F(1, 1) = 0.5 * (emf[0] + emf[sj]);
    F(2, 0) = -0.5 * (emf[0] + emf[1]);
    F(2, 1) = 0.0;
  } else {
    const Real *e0 = emf, *e1 = emf + (size_t)N, *e2 = emf + (size_t)2 * N;
    F(1, 0) = 0.0;
    F(1, 1) = 0.5 * (e2[0] + e2[sj]);
    F(1, 2) = -0.5 * (e1[0] + e1[sk]);
    F(2, 0) = -0.5 * (e2[0] + e2[1]);
    F(2, 1) = 0.0;
    F(2, 2) = 0.5 * (e0[0] + e0[sk]);
    F(3, 0) = 0.5 * (e1[0] + e1[1]);
    F(3, 1) = -0.5 * (e0[0] + e0[sj]);
    F(3, 2) = 0.0;
  }
}

// ------------------------------------------------------------------ FixFluxes
// ref: src/fixup/fixup.cpp:728-1078 (fluid variables). One thread per boundary-face zone; the six
// faces touch disjoint flux entries except polar's B2 mirror, which only reads the face plane its own
// thread column owns.
template <bool MHD> __global__ void fix_fluxes_kernel(const pb_pack_t pack, const pb_cfg_t cfg, int nface_max) {
  constexpr int NCn = NC<MHD>::ncons;
  const Lay L(cfg);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_block = (long long)2 * L.ndim * nface_max;
  if (t >= per_block * pack.nblocks) return;
  const int b = (int)(t / per_block);
  int rem = (int)(t % per_block);
  const int f = rem / nface_max; // face id 0..2*ndim-1
  const int c = rem % nface_max;
  const pb_block_t &blk = pack.blocks[b];
  const int bc = blk.bc[f];
  if (bc <= 0) return;
  const int dir = f / 2, side = f % 2;
  const int n1 = L.ie - L.is + 1, n2 = L.je - L.js + 1, n3 = L.ke - L.ks + 1;
  int i, j, k;
  if (dir == 0) {
    if (c >= n2 * n3) return;
    i = side == 0 ? L.is : L.ie + 1;
    j = L.js + c % n2;
    k = L.ks + c / n2;
  } else if (dir == 1) {
    if (c >= n1 * n3) return;
    j = side == 0 ? L.js : L.je + 1;
    i = L.is + c % n1;
    k = L.ks + c / n1;
  } else {
    if (c >= n1 * n2) return;
    k = side == 0 ? L.ks : L.ke + 1;
    i = L.is + c % n1;
    j = L.js + c / n1;
  }
  const int N = L.ncell, o = L.idx(k, j, i);
  Real *fb = pack.flux + (size_t)b * L.ndim * NCn * N;
  auto F = [&](int d, int m, int off) -> Real & { return fb[((size_t)(d - 1) * NCn + m) * N + o + off]; };
  const int d1 = dir + 1;
  if (dir != 1) { // X1 and X3 faces: outflow / reflect only
    if (bc == PB_BC_OUTFLOW) {
      F(d1, CRHO, 0) = side == 0 ? dmin(F(d1, CRHO, 0), 0.0) : dmax(F(d1, CRHO, 0), 0.0);
    } else if (bc == PB_BC_REFLECT) {
      F(d1, CRHO, 0) = 0.0;
      F(d1, CENG, 0) = 0.0;
    }
    return;
  }
  if (bc == PB_BC_OUTFLOW) {
    F(2, CRHO, 0) = side == 0 ? dmin(F(2, CRHO, 0), 0.0) : dmax(F(2, CRHO, 0), 0.0);
  } else if (bc == PB_BC_POLAR) {
    F(2, CRHO, 0) = 0.0;
    F(2, CENG, 0) = 0.0;
    for (int ii = 0; ii < 3; ++ii) {
      F(2, CMOM + ii, 0) = 0.0;
      if (MHD) F(2, CB + ii, 0) = 0.0;
    }
    if (MHD) {
      const int sj = L.ni;
      if (side == 0) {
        F(1, CB + 1, -sj) = -F(1, CB + 1, 0);
        if (L.ndim == 3) F(3, CB + 1, -sj) = -F(3, CB + 1, 0);
      } else {
        F(1, CB + 1, 0) = -F(1, CB + 1, -sj);
        if (L.ndim == 3) F(3, CB + 1, 0) = -F(3, CB + 1, -sj);
      }
    }
  } else if (bc == PB_BC_REFLECT) {
    F(2, CRHO, 0) = 0.0;
    F(2, CENG, 0) = 0.0;
    F(2, CMOM + 0, 0) = 0.0;
    F(2, CMOM + 2, 0) = 0.0;
  }
}

// ------------------------------------------------------------------ divergence / sources
// ref: external/parthenon/src/interface/update.hpp:43-58 (FluxDivHelper)
template <int NCn, int NDIM> PB_DEV void flux_div_nd(const Lay &L, const pb_block_t &blk, const Real *fb, int o, Real *du) {
  // uniform_cartesian.hpp:30-40: face areas and the cell volume are products of dx
  const Real area[3] = {blk.dx[1] * blk.dx[2], blk.dx[0] * blk.dx[2], blk.dx[0] * blk.dx[1]};
  const Real vol = blk.dx[0] * blk.dx[1] * blk.dx[2];
  const int st[3] = {1, L.ni, L.ni * L.nj};
  const int N = L.ncell;
  // NDIM is a template parameter so that the 2 * NDIM loads of a component are issued together
#pragma unroll
  for (int l = 0; l < NCn; ++l) {
    Real lo[NDIM], hi[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
      const Real *f = fb + ((size_t)d * NCn + l) * N + o;
      lo[d] = __ldg(f);
      hi[d] = __ldg(f + st[d]);
    }
    Real acc = 0;
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
      const Real t = (area[d] * hi[d] - area[d] * lo[d]);
      acc = (d == 0) ? t : acc + t;
    }
    du[l] = -acc / vol;
  }
}
template <int NCn> PB_DEV void flux_div(const Lay &L, const pb_block_t &blk, const Real *fb, int o, Real *du) {
  if (L.ndim == 3) flux_div_nd<NCn, 3>(L, blk, fb, o, du);
  else if (L.ndim == 2) flux_div_nd<NCn, 2>(L, blk, fb, o, du);
  else flux_div_nd<NCn, 1>(L, blk, fb, o, du);
}

// ref: src/fluid/fluid.cpp:593-654 with src/fluid/tmunu.hpp:63-131 and
// src/geometry/geometry_utils.hpp:73-80 (connection from metric derivatives); src[4] = S1,S2,S3,tau
template <bool MHD>
PB_DEV void source_zone(const GeomTab &gt, int gcell, const Real *prim, int N, Real src[4]) {
  constexpr int prs = NC<MHD>::prs;
  PointGeom g;
  g.load(gt, 0, gcell);
  // --- Tmunu (contravariant)
  Real vp[3] = {prim[(size_t)PVEL * N], prim[(size_t)(PVEL + 1) * N], prim[(size_t)(PVEL + 2) * N]};
  Real B[3] = {0, 0, 0};
  if (MHD)
    for (int a = 0; a < 3; ++a) B[a] = prim[(size_t)(PB_ + a) * N];
  const Real W = lorentz_factor(vp, g); // gamma_ij is the spatial block of g_mu_nu
  Real Bdotv = 0.0, Bsq = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) {
      Bdotv += B[a] * (vp[bb] / W) * g.g4(a + 1, bb + 1);
      Bsq += B[a] * B[bb] * g.g4(a + 1, bb + 1);
    }
  const Real iW = ratio(1., W);
  const Real alpha = g.alpha;
  Real u[4], bv[4];
  u[0] = ratio(W, fabs(alpha));
  bv[0] = u[0] * Bdotv;
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    u[l] = vp[l - 1] - u[0] * g.beta[l - 1];
    bv[l] = iW * (B[l - 1] + alpha * bv[0] * u[l]);
  }
  const Real bsq = Bsq * iW * iW + Bdotv * Bdotv;
  // SpacetimeMetricInverse, ref: cached_system.hpp:330-349
  Real gg[4][4];
  {
    Real alpha2 = alpha;
    alpha2 *= alpha2;
    gg[0][0] = -ratio(1, alpha2);
#pragma unroll
    for (int m = 0; m < 3; ++m) gg[m + 1][0] = gg[0][m + 1] = ratio(g.beta[m], alpha2);
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        gg[m + 1][n + 1] = g.gc3(m, n);
        gg[m + 1][n + 1] -= gg[0][m + 1] * gg[0][n + 1] * alpha2;
      }
  }
  const Real rho = prim[(size_t)PRHO * N], uu = prim[(size_t)PENG * N], P = prim[(size_t)prs * N];
  const Real A = rho + uu + P + bsq;
  const Real Bc = P + 0.5 * bsq;
  Real T[4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) T[m][n] = A * u[m] * u[n] + Bc * gg[m][n] - bv[m] * bv[n];
  // --- metric derivatives and connection
  const Real gdet = alpha * g.detgam;
  Real dg[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      dg[m][n][0] = 0.0;
#pragma unroll
      for (int s = 1; s < 4; ++s) dg[m][n][s] = gt.der(G_DG + 3 * flat4(m, n) + (s - 1), gcell);
    }
  Real gam[4][4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) gam[a][b][c] = 0.5 * (dg[b][a][c] + dg[c][a][b] - dg[b][c][a]);
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    Real src_mom = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) src_mom += T[m][n] * (dg[n][l + 1][m] - gam[l + 1][n][m]);
    src[l] = gdet * src_mom;
  }
  Real TGam = 0.0;
  const Real inv_alpha2 = ratio(1, alpha * alpha);
  const Real gcon0[4] = {-inv_alpha2, inv_alpha2 * g.beta[0], inv_alpha2 * g.beta[1], inv_alpha2 * g.beta[2]};
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int n = 0; n < 4; n++) {
      Real gam0 = 0;
#pragma unroll
      for (int r = 0; r < 4; r++) gam0 += gcon0[r] * gam[r][m][n];
      TGam += T[m][n] * gam0;
    }
  Real Ta = 0.0;
  Real da[4] = {0.0, gt.der(G_DALPHA + 0, gcell), gt.der(G_DALPHA + 1, gcell), gt.der(G_DALPHA + 2, gcell)};
#pragma unroll
  for (int m = 0; m < 4; m++) Ta += T[m][0] * da[m];
  src[3] = gdet * alpha * (Ta - TGam);
}

template <bool MHD> __global__ void divergence_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *dudt) {
  constexpr int NCn = NC<MHD>::ncons;
  const Lay L(cfg);
  int b, k, j, i;
  if (!interior_zone(L, pack.nblocks, 0, 0, 0, b, k, j, i)) return;
  const int N = L.ncell, o = L.idx(k, j, i);
  Real du[NCn];
  flux_div<NCn>(L, pack.blocks[b], pack.flux + (size_t)b * L.ndim * NCn * N, o, du);
#pragma unroll
  for (int l = 0; l < NCn; ++l) dudt[((size_t)b * NCn + l) * N + o] = du[l];
}

template <bool MHD> __global__ void source_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *gsrc) {
  constexpr int NP = NC<MHD>::nprim;
  const Lay L(cfg);
  int b, k, j, i;
  if (!interior_zone(L, pack.nblocks, 0, 0, 0, b, k, j, i)) return;
  const int N = L.ncell, o = L.idx(k, j, i);
  const GeomTab gt(cfg, pack.geom, pack.blocks[b].geom_slot);
  Real src[4];
  source_zone<MHD>(gt, gt.cell(k, j, i), pack.prim + (size_t)b * NP * N + o, N, src);
#pragma unroll
  for (int l = 0; l < 4; ++l) gsrc[((size_t)b * 4 + l) * N + o] = src[l];
}

// Fused: dudt = -div F; dudt[S,tau] = 1*dudt + 1*gsrc (SumData); u_avg = beta*u0 + (1-beta)*ubase
// (AverageIndependentData); uout = 1*u_avg + bdt*dudt (UpdateIndependentData).
// ref: src/phoebus_driver.cpp:560-576, external/parthenon/src/interface/update.hpp:71-136
template <bool MHD>
__global__ void __launch_bounds__(128, 5) update_fused_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *u0, const Real *ubase,
                                    Real beta, Real bdt, Real *uout, Real *dudt_out, Real *gsrc_out, const Real *dt_dev) {
  if (dt_dev) bdt = beta * __ldg(dt_dev); // graph replay: the step size lives in device memory
  constexpr int NCn = NC<MHD>::ncons, NP = NC<MHD>::nprim;
  const Lay L(cfg);
  int b, k, j, i;
  if (!interior_zone(L, pack.nblocks, 0, 0, 0, b, k, j, i)) return;
  const int N = L.ncell, o = L.idx(k, j, i);
  Real du[NCn];
  flux_div<NCn>(L, pack.blocks[b], pack.flux + (size_t)b * L.ndim * NCn * N, o, du);
  const GeomTab gt(cfg, pack.geom, pack.blocks[b].geom_slot);
  Real src[4];
  source_zone<MHD>(gt, gt.cell(k, j, i), pack.prim + (size_t)b * NP * N + o, N, src);
  if (gsrc_out) {
#pragma unroll
    for (int l = 0; l < 4; ++l) gsrc_out[((size_t)b * 4 + l) * N + o] = src[l];
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) du[CMOM + l] = 1.0 * du[CMOM + l] + 1.0 * src[l];
  const Real w2 = 1.0 - beta;
#pragma unroll
  for (int l = 0; l < NCn; ++l) {
    const size_t a = ((size_t)b * NCn + l) * N + o;
    const Real uavg = beta * u0[a] + w2 * ubase[a];
    uout[a] = 1.0 * uavg + bdt * du[l];
    if (dudt_out) dudt_out[a] = du[l];
  }
}

// Same update with the geometric sources already in gsrc (CalculateFluidSourceTerms ran as its own
// launch): a pure streaming kernel (24 flux + 4 source + 16 cons words in, 8 out per zone), few registers.
template <bool MHD>
__global__ void __launch_bounds__(256, 4) update_stream_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *u0, const Real *ubase,
                                                               Real beta, Real bdt, Real *uout, const Real *gsrc, Real *dudt_out,
                                                               const Real *dt_dev) {
  if (dt_dev) bdt = beta * __ldg(dt_dev); // graph replay: the step size lives in device memory
  constexpr int NCn = NC<MHD>::ncons;
  const Lay L(cfg);
  int b, k, j, i;
  if (!interior_zone(L, pack.nblocks, 0, 0, 0, b, k, j, i)) return;
  const int N = L.ncell, o = L.idx(k, j, i);
  Real du[NCn];
  flux_div<NCn>(L, pack.blocks[b], pack.flux + (size_t)b * L.ndim * NCn * N, o, du);
#pragma unroll
  for (int l = 0; l < 4; ++l) du[CMOM + l] = 1.0 * du[CMOM + l] + 1.0 * __ldg(gsrc + ((size_t)b * 4 + l) * N + o);
  const Real w2 = 1.0 - beta;
#pragma unroll
  for (int l = 0; l < NCn; ++l) {
    const size_t a = ((size_t)b * NCn + l) * N + o;
    const Real uavg = beta * u0[a] + w2 * ubase[a];
    uout[a] = 1.0 * uavg + bdt * du[l];
    if (dudt_out) dudt_out[a] = du[l];
  }
}

// Ghost zones of the same update: WeightedSumData runs over ENTIRE arrays (update.hpp:77-80) and
// dUdt is zero outside the interior, so uout = beta*u0 + (1-beta)*ubase there. The values matter
// only to the stale-prim ApplyFloors of region 7 (which may re-invert a ghost zone and thereby
// advance its c2p_mu history) and are overwritten by the halo exchange / BCs afterwards.
__global__ void update_ghost_kernel(const pb_cfg_t cfg, int nblocks, int ncons, const Real *u0, const Real *ubase,
                                    Real beta, Real *uout) {
  const Lay L(cfg);
  const int N = L.ncell;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * nblocks) return;
  const int b = (int)(t / N), o = (int)(t % N);
  const int i = o % L.ni, j = (o / L.ni) % L.nj, k = o / (L.ni * L.nj);
  if (i >= L.is && i <= L.ie && j >= L.js && j <= L.je && k >= L.ks && k <= L.ke) return;
  const Real w2 = 1.0 - beta;
  for (int l = 0; l < ncons; ++l) {
    const size_t a = ((size_t)b * ncons + l) * N + o;
    const Real uavg = beta * u0[a] + w2 * ubase[a];
    uout[a] = 1.0 * uavg + 0.0;
  }
}

__global__ void weighted_sum_kernel(const Real *x, const Real *y, Real w1, Real w2, Real *z, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) z[t] = w1 * x[t] + w2 * y[t];
}

static unsigned grid_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// set while a stage is being captured into a CUDA graph (pb_stage_graph_create): the update kernels then read dt from
// device memory instead of their by-value argument
static const double *g_stage_dt_dev = nullptr;
void set_stage_dt_pointer(const double *dt_dev) { g_stage_dt_dev = dt_dev; }

__global__ void dt_control_kernel(const Real *est, Real cfl, Real *dt) {
  const Real a = 2.0 * *dt, b = cfl * *est;
  *dt = b < a ? b : a; // std::min(2 dt, cfl est), ref: parthenon driver.cpp:216-240
}

} // namespace pb

using namespace pb;

extern "C" int pb_flux_ct(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cfg->mhd || cfg->ndim == 1) return PB_OK; // ref: fluid.cpp:776-781
  PB_NEED("pb_flux_ct", PB_F(pack, flux), PB_F(pack, emf));
  const Lay L(*cfg);
  const long long n = (long long)(cfg->nx[0] + 1) * (cfg->nx[1] + 1) * (cfg->ndim == 3 ? cfg->nx[2] + 1 : 1) * pack->nblocks;
  const int T = 256;
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->ndim == 2) {
    emf_kernel<2><<<grid_for(n, T), T, 0, s>>>(*pack, *cfg);
    ct_flux_kernel<2><<<grid_for(n, T), T, 0, s>>>(*pack, *cfg);
  } else {
    emf_kernel<3><<<grid_for(n, T), T, 0, s>>>(*pack, *cfg);
    ct_flux_kernel<3><<<grid_for(n, T), T, 0, s>>>(*pack, *cfg);
  }
  return check_launch("flux_ct");
}

extern "C" int pb_fix_fluxes(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cfg->enable_flux_fixup) return PB_OK; // ref: fixup.cpp:734-736
  PB_NEED("pb_fix_fluxes", PB_F(pack, flux));
  const int n1 = cfg->nx[0], n2 = cfg->ndim >= 2 ? cfg->nx[1] : 1, n3 = cfg->ndim >= 3 ? cfg->nx[2] : 1;
  int nface_max = n2 * n3;
  if (n1 * n3 > nface_max) nface_max = n1 * n3;
  if (n1 * n2 > nface_max) nface_max = n1 * n2;
  const long long n = (long long)2 * cfg->ndim * nface_max * pack->nblocks;
  const int T = 128;
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd) fix_fluxes_kernel<true><<<grid_for(n, T), T, 0, s>>>(*pack, *cfg, nface_max);
  else fix_fluxes_kernel<false><<<grid_for(n, T), T, 0, s>>>(*pack, *cfg, nface_max);
  return check_launch("fix_fluxes_kernel");
}

static long long n_interior(const pb_pack_t *pack, const pb_cfg_t *cfg) {
  return (long long)cfg->nx[0] * (cfg->ndim >= 2 ? cfg->nx[1] : 1) * (cfg->ndim >= 3 ? cfg->nx[2] : 1) * pack->nblocks;
}

extern "C" int pb_flux_divergence(const pb_pack_t *pack, const pb_cfg_t *cfg, double *dudt, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!dudt) return set_error(PB_EINVAL, "pb_flux_divergence: NULL output");
  PB_NEED("pb_flux_divergence", PB_F(pack, flux));
  const int T = 256;
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd) divergence_kernel<true><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, dudt);
  else divergence_kernel<false><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, dudt);
  return check_launch("divergence_kernel");
}

extern "C" int pb_source_terms(const pb_pack_t *pack, const pb_cfg_t *cfg, double *gsrc, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!gsrc) return set_error(PB_EINVAL, "pb_source_terms: NULL output");
  const int T = zone_threads();
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd) source_kernel<true><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, gsrc);
  else source_kernel<false><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, gsrc);
  return check_launch("source_kernel");
}

// ref: fluid::CalculateDivB, src/fluid/fluid.cpp:843-897 — corner-centred div B of the conserved field on
// interior zones; sums left to right as the reference writes them
template <int NDIM> __global__ void divb_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *cons, Real *divb) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!interior_zone(L, pack.nblocks, 0, 0, 0, b, k, j, i)) return;
  const int N = L.ncell, o = L.idx(k, j, i);
  const int sj = L.ni, sk = L.ni * L.nj;
  const pb_block_t &blk = pack.blocks[b];
  const Real *b0 = cons + ((size_t)b * 8 + CB) * N, *b1 = b0 + N, *b2 = b1 + N;
  Real v;
  if (NDIM == 2) {
    v = 0.5 * (b0[o] + b0[o - sj] - b0[o - 1] - b0[o - sj - 1]) / blk.dx[0] +
        0.5 * (b1[o] + b1[o - 1] - b1[o - sj] - b1[o - sj - 1]) / blk.dx[1];
  } else {
    v = 0.25 *
            (b0[o] + b0[o - sj] + b0[o - sk] + b0[o - sk - sj] - b0[o - 1] - b0[o - sj - 1] - b0[o - sk - 1] -
             b0[o - sk - sj - 1]) /
            blk.dx[0] +
        0.25 *
            (b1[o] + b1[o - 1] + b1[o - sk] + b1[o - sk - 1] - b1[o - sj] - b1[o - sj - 1] - b1[o - sk - sj] -
             b1[o - sk - sj - 1]) /
            blk.dx[1] +
        0.25 *
            (b2[o] + b2[o - 1] + b2[o - sj] + b2[o - sj - 1] - b2[o - sk] - b2[o - sk - 1] - b2[o - sk - sj] -
             b2[o - sk - sj - 1]) /
            blk.dx[2];
  }
  divb[(size_t)b * N + o] = v;
}

// ---- history reductions of the fluid package (ref: src/fluid/fluid.cpp:325-347, History::ReduceOneVar,
// src/analysis/history.hpp:50-86): max of p.density, sums of c.density, c.energy, c.momentum[3] over interior zones.
// The reference's volume weighting compares Reducer_t with Kokkos::Sum<Real, ExecSpace> and therefore never applies
// (Reducer_t is Kokkos::Sum<Real>): the sums are plain sums, as its .hst files show. One CTA per MeshBlock reduces in a
// fixed tree order; the caller adds the per-block partials in block order, so the result is run-to-run reproducible.
__global__ void __launch_bounds__(256) history_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *cons, Real *partial) {
  const Lay L(cfg);
  const int b = blockIdx.x, N = L.ncell;
  const int nxi = L.ie - L.is + 1, nxj = L.je - L.js + 1, nxk = L.ke - L.ks + 1;
  const int nz = nxi * nxj * nxk;
  const int ncons = 5 + 3 * cfg.mhd, nprim = 7 + 3 * cfg.mhd;
  Real acc[6] = {-DBL_MAX, 0.0, 0.0, 0.0, 0.0, 0.0}; // max rho, D, tau, S1, S2, S3
  for (int z = threadIdx.x; z < nz; z += blockDim.x) {
    const int i = L.is + z % nxi, j = L.js + (z / nxi) % nxj, k = L.ks + z / (nxi * nxj);
    const int o = L.idx(k, j, i);
    const Real *u = cons + (size_t)b * ncons * N + o;
    acc[0] = dmax(acc[0], pack.prim[(size_t)b * nprim * N + o]);
    acc[1] += u[(size_t)CRHO * N];
    acc[2] += u[(size_t)CENG * N];
    acc[3] += u[(size_t)CMOM * N];
    acc[4] += u[(size_t)(CMOM + 1) * N];
    acc[5] += u[(size_t)(CMOM + 2) * N];
  }
  __shared__ Real sh[6][256];
  for (int q = 0; q < 6; ++q) sh[q][threadIdx.x] = acc[q];
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      sh[0][threadIdx.x] = dmax(sh[0][threadIdx.x], sh[0][threadIdx.x + off]);
      for (int q = 1; q < 6; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x < 6) partial[(size_t)b * 6 + threadIdx.x] = sh[threadIdx.x][0];
}

extern "C" int pb_history_fluid(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, double *partial_dev,
                                double out_host[6], void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons || !partial_dev || !out_host) return set_error(PB_EINVAL, "pb_history_fluid: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  history_kernel<<<pack->nblocks, 256, 0, s>>>(*pack, *cfg, cons, partial_dev);
  rc = check_launch("history_kernel");
  if (rc) return rc;
  std::vector<double> h((size_t)pack->nblocks * 6);
  cudaError_t e = cudaMemcpyAsync(h.data(), partial_dev, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return set_cuda_error(e, "pb_history_fluid");
  out_host[0] = -DBL_MAX;
  for (int q = 1; q < 6; ++q) out_host[q] = 0.0;
  for (int b = 0; b < pack->nblocks; ++b) {
    out_host[0] = h[(size_t)b * 6] > out_host[0] ? h[(size_t)b * 6] : out_host[0];
    for (int q = 1; q < 6; ++q) out_host[q] += h[(size_t)b * 6 + q];
  }
  return PB_OK;
}

// ---- horizon-flux history of the torus (ref: src/analysis/history.cpp:25-222, history_utils.hpp:35-180; registered in
// src/geometry/geometry.cpp:67-99 when the coordinate system has a horizon parameter "xh"): mdot, jet energy and
// momentum flux (zones with magnetisation > fluid/sigma_cutoff), magnetic flux Phi. Each is the cell-centred flux of the
// zone column that contains x1 = xh, linearly interpolated to xh with a centred slope, times dx2 dx3.
struct HistCell { // what the four flux functions read at (k, j, i)
  PointGeom g;
  Real rho, v[3], B[3];
};
template <bool MHD> PB_DEV void hist_load(const pb_pack_t &pack, const pb_cfg_t &cfg, const Lay &L, int b, int k, int j, int i, HistCell &c) {
  constexpr int NP = NC<MHD>::nprim;
  const GeomTab gt(cfg, pack.geom, pack.blocks[b].geom_slot);
  c.g.load(gt, 0, gt.cell(k, j, i));
  const Real *p = pack.prim + (size_t)b * NP * L.ncell + L.idx(k, j, i);
  c.rho = p[0];
  for (int a = 0; a < 3; ++a) {
    c.v[a] = p[(size_t)(PVEL + a) * L.ncell];
    c.B[a] = MHD ? p[(size_t)(PB_ + a) * L.ncell] : 0.0;
  }
}
PB_DEV Real hist_bsq(const HistCell &c, Real W) { // phoebus::GetMagneticFieldSquared, relativity_utils.hpp:131-141
  Real Bsq = 0, Bdotv = 0;
  for (int m = 0; m < 3; ++m)
    for (int n = 0; n < 3; ++n) {
      Bdotv += c.g.g4(m + 1, n + 1) * ratio(c.v[m], W) * c.B[n];
      Bsq += c.g.g4(m + 1, n + 1) * c.B[m] * c.B[n];
    }
  return ratio(Bsq, W * W) + Bdotv * Bdotv;
}
PB_DEV Real hist_mass_flux(const HistCell &c) { // CalcMassFlux
  const Real W = lorentz_factor(c.v, c.g);
  const Real ucon = c.v[0] - c.g.beta[0] * W / c.g.alpha;
  return -c.g.alpha * c.g.detgam * c.rho * ucon;
}
PB_DEV Real hist_sigma(const HistCell &c) { // CalcMagnetization
  const Real W = lorentz_factor(c.v, c.g);
  return hist_bsq(c, W) / c.rho;
}
PB_DEV void hist_em_fluxes(const HistCell &c, Real &en, Real &mom) { // CalcEMEnergyFlux, CalcEMMomentumFlux
  const Real W = lorentz_factor(c.v, c.g);
  Real Bdotv = 0.0, vcov[3] = {0, 0, 0}, Bcov[3] = {0, 0, 0};
  for (int m = 0; m < 3; ++m)
    for (int n = 0; n < 3; ++n) {
      vcov[m] += c.g.g4(m + 1, n + 1) * c.v[n];
      Bcov[m] += c.g.g4(m + 1, n + 1) * c.B[n];
      Bdotv += c.g.g4(m + 1, n + 1) * c.B[m] * ratio(c.v[n], W);
    }
  const Real lapse = c.g.alpha;
  const Real b0 = Bdotv * W / lapse;
  const Real ucov0 = -W * lapse;
  const Real ucon1 = c.v[0] - c.g.beta[0] * W / lapse;
  const Real ucov1 = vcov[0];
  const Real bcon1 = ratio((c.B[0] + lapse * b0 * ucon1), W);
  const Real bcov1 = ratio((Bcov[0] + lapse * b0 * ucov1), W);
  const Real bsq = hist_bsq(c, W);
  const Real T_EM_10 = bsq * ucon1 * ucov0 - bcon1 * b0;
  const Real T_EM_11 = bsq * ucon1 * ucov1 - bcon1 * bcov1 + 0.5 * bsq;
  en = +T_EM_10 * lapse * c.g.detgam;
  mom = -T_EM_11 * lapse * c.g.detgam;
}
template <bool MHD>
__global__ void __launch_bounds__(128) history_torus_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *cons, Real xh,
                                                            Real sigma_cutoff, Real *partial) {
  const Lay L(cfg);
  const int b = blockIdx.x, N = L.ncell;
  const pb_block_t &blk = pack.blocks[b];
  const int nxj = L.je - L.js + 1, nxk = L.ke - L.ks + 1;
  Real acc[4] = {0.0, 0.0, 0.0, 0.0}; // mdot, jet energy flux, jet momentum flux, Phi (before the 0.5)
  // the interior zone column whose faces bracket xh (at most one per block)
  int ih = -1;
  for (int i = L.is; i <= L.ie; ++i)
    if (Xf(blk, 0, i) <= xh && xh < Xf(blk, 0, i + 1)) ih = i;
  if (ih >= 0) {
    const Real dx1 = blk.dx[0], dx2 = blk.dx[1], dx3 = blk.dx[2];
    for (int z = threadIdx.x; z < nxj * nxk; z += blockDim.x) {
      const int j = L.js + z % nxj, k = L.ks + z / nxj;
      HistCell c0, cm, cp;
      hist_load<MHD>(pack, cfg, L, b, k, j, ih, c0);
      hist_load<MHD>(pack, cfg, L, b, k, j, ih - 1, cm);
      hist_load<MHD>(pack, cfg, L, b, k, j, ih + 1, cp);
      {
        const Real m = (hist_mass_flux(cp) - hist_mass_flux(cm)) / (2.0 * dx1);
        acc[0] += (hist_mass_flux(c0) + (xh - Xc(blk, 0, ih)) * m) * dx2 * dx3;
      }
      if (MHD) {
        if (hist_sigma(c0) > sigma_cutoff) {
          Real e0, p0, em, pm, ep, pp;
          hist_em_fluxes(c0, e0, p0);
          hist_em_fluxes(cm, em, pm);
          hist_em_fluxes(cp, ep, pp);
          acc[1] += (e0 + (xh - Xc(blk, 0, ih)) * ((ep - em) / (2.0 * dx1))) * dx2 * dx3;
          acc[2] += (p0 + (xh - Xc(blk, 0, ih)) * ((pp - pm) / (2.0 * dx1))) * dx2 * dx3;
        }
        // Phi is evaluated one zone further out (i + 1), ref: history.cpp:208-218
        const Real *cb = cons + ((size_t)b * 8 + CB) * N;
        auto phi = [&](int i) {
          const GeomTab gt(cfg, pack.geom, blk.geom_slot);
          return fabs(cb[L.idx(k, j, i)]) * gt.at(0, G_ALPHA, gt.cell(k, j, i));
        };
        const Real m = (phi(ih + 2) - phi(ih)) / (2.0 * dx1);
        acc[3] += (phi(ih + 1) + (xh - Xc(blk, 0, ih + 1)) * m) * dx2 * dx3;
      }
    }
  }
  __shared__ Real sh[4][128];
  for (int q = 0; q < 4; ++q) sh[q][threadIdx.x] = acc[q];
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off)
      for (int q = 0; q < 4; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x < 4) partial[(size_t)b * 4 + threadIdx.x] = sh[threadIdx.x][0];
}

extern "C" int pb_history_torus(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, double xh, double sigma_cutoff,
                                double *partial_dev, double out_host[4], void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons || !partial_dev || !out_host) return set_error(PB_EINVAL, "pb_history_torus: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd) history_torus_kernel<true><<<pack->nblocks, 128, 0, s>>>(*pack, *cfg, cons, xh, sigma_cutoff, partial_dev);
  else history_torus_kernel<false><<<pack->nblocks, 128, 0, s>>>(*pack, *cfg, cons, xh, sigma_cutoff, partial_dev);
  rc = check_launch("history_torus_kernel");
  if (rc) return rc;
  std::vector<double> h((size_t)pack->nblocks * 4);
  cudaError_t e = cudaMemcpyAsync(h.data(), partial_dev, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return set_cuda_error(e, "pb_history_torus");
  for (int q = 0; q < 4; ++q) out_host[q] = 0.0;
  for (int b = 0; b < pack->nblocks; ++b)
    for (int q = 0; q < 4; ++q) out_host[q] += h[(size_t)b * 4 + q];
  out_host[3] *= 0.5; // 0.5 int detg B^r dx2 dx3
  return PB_OK;
}

extern "C" int pb_calculate_divb(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, double *divb, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons || !divb) return set_error(PB_EINVAL, "pb_calculate_divb: NULL argument");
  if (!cfg->mhd || cfg->ndim < 2) return PB_OK; // ref: fluid.cpp:846, and no 1-D branch there
  const int T = 256;
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->ndim == 2) divb_kernel<2><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, cons, divb);
  else divb_kernel<3><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, cons, divb);
  return check_launch("divb_kernel");
}

// ------------------------------------------------------------------ per-MeshBlock allocations <-> slabs
// Parthenon allocates every variable per MeshBlock (ref: external/parthenon/src/interface/variable.cpp, one ParArray per
// block and variable); pb_pack_t takes one slab per field. A host that cannot place its variables in such slabs gathers them
// before a cycle and scatters them after it. gridDim.y = block, grid-stride over the block's words, 16-byte accesses when
// both sides are aligned.
namespace pb {
template <bool GATHER>
__global__ void block_copy_kernel(double *slab, double *const *block_ptrs, const long long words) {
  const int b = blockIdx.y;
  double *blk = block_ptrs[b];
  double *sl = slab + (size_t)b * words;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<size_t>(blk) | reinterpret_cast<size_t>(sl)) & 15) == 0) {
    const long long n2 = words / 2;
    double2 *a = reinterpret_cast<double2 *>(GATHER ? sl : blk);
    const double2 *c = reinterpret_cast<const double2 *>(GATHER ? blk : sl);
    for (long long q = t; q < n2; q += stride) a[q] = c[q];
    if ((words & 1) && t == 0) (GATHER ? sl : blk)[words - 1] = (GATHER ? blk : sl)[words - 1];
  } else {
    for (long long q = t; q < words; q += stride) (GATHER ? sl : blk)[q] = (GATHER ? blk : sl)[q];
  }
}
static int launch_block_copy(bool gather, double *slab, double *const *block_ptrs, int nblocks, long long words, cudaStream_t s) {
  if (!slab || !block_ptrs) return set_error(PB_EINVAL, "pb_blocks_gather / scatter: NULL argument");
  if (nblocks <= 0 || words <= 0) return PB_OK;
  if (nblocks > 65535) return set_error(PB_EINVAL, "pb_blocks_gather / scatter: more than 65535 blocks per call");
  long long gx = (words / 2 + 255) / 256;
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  const dim3 grid((unsigned)gx, (unsigned)nblocks);
  if (gather) block_copy_kernel<true><<<grid, 256, 0, s>>>(slab, block_ptrs, words);
  else block_copy_kernel<false><<<grid, 256, 0, s>>>(slab, block_ptrs, words);
  return check_launch("block_copy_kernel");
}
} // namespace pb

extern "C" int pb_blocks_gather(double *slab, const double *const *block_ptrs_dev, int nblocks, long long words_per_block,
                                void *stream) {
  return launch_block_copy(true, slab, const_cast<double *const *>(reinterpret_cast<const double *const *>(block_ptrs_dev)), nblocks,
                           words_per_block, (cudaStream_t)stream);
}
extern "C" int pb_blocks_scatter(const double *slab, double *const *block_ptrs_dev, int nblocks, long long words_per_block,
                                 void *stream) {
  return launch_block_copy(false, const_cast<double *>(slab), block_ptrs_dev, nblocks, words_per_block, (cudaStream_t)stream);
}

extern "C" int pb_weighted_sum(const double *x, const double *y, double w1, double w2, double *z, long long n,
                               void *stream) {
  if (n <= 0) return PB_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  weighted_sum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, w1, w2, z, n);
  return check_launch("weighted_sum_kernel");
}

extern "C" int pb_update_from_sources(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *u0, const double *ubase,
                                      double beta, double bdt, double *uout, const double *gsrc, double *dudt_out,
                                      int update_ghosts, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!u0 || !ubase || !uout || !gsrc) return set_error(PB_EINVAL, "NULL container");
  PB_NEED("pb_update_from_sources", PB_F(pack, flux));
  const int T = 256;
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd)
    update_stream_kernel<true><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, u0, ubase, beta, bdt, uout, gsrc, dudt_out, g_stage_dt_dev);
  else
    update_stream_kernel<false><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, u0, ubase, beta, bdt, uout, gsrc, dudt_out, g_stage_dt_dev);
  if (update_ghosts) {
    const Lay L(*cfg);
    const long long ntot = (long long)L.ncell * pack->nblocks;
    update_ghost_kernel<<<grid_for(ntot, 256), 256, 0, s>>>(*cfg, pack->nblocks, 5 + 3 * cfg->mhd, u0, ubase, beta, uout);
  }
  return check_launch("update_stream_kernel");
}

extern "C" int pb_update_fused(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *u0, const double *ubase,
                               double beta, double bdt, double *uout, double *dudt_out, double *gsrc_out,
                               int update_ghosts, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!u0 || !ubase || !uout) return set_error(PB_EINVAL, "NULL conserved container");
  PB_NEED("pb_update_fused", PB_F(pack, flux));
  const int T = zone_threads();
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd)
    update_fused_kernel<true><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, u0, ubase, beta, bdt, uout, dudt_out, gsrc_out, g_stage_dt_dev);
  else
    update_fused_kernel<false><<<grid_for(n_interior(pack, cfg), T), T, 0, s>>>(*pack, *cfg, u0, ubase, beta, bdt, uout, dudt_out, gsrc_out, g_stage_dt_dev);
  if (update_ghosts) {
    const Lay L(*cfg);
    const long long ntot = (long long)L.ncell * pack->nblocks;
    update_ghost_kernel<<<grid_for(ntot, 256), 256, 0, s>>>(*cfg, pack->nblocks, 5 + 3 * cfg->mhd, u0, ubase, beta, uout);
  }
  return check_launch("update_fused_kernel");
}

extern "C" int pb_dt_control(const double *est_dev, double cfl, double *dt_dev, void *stream) {
  if (!est_dev || !dt_dev) return set_error(PB_EINVAL, "pb_dt_control: NULL argument");
  dt_control_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(est_dev, cfl, dt_dev);
  return check_launch("dt_control_kernel");
}
