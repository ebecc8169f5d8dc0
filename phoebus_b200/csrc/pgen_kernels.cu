// Device side of the Fishbone-Moncrief torus problem generator (ref: src/pgen/torus.cpp:285-474,525-545,849-930).
// The transcendental part of the solution (rho, u, W v^i as functions of x1, x2) is tabulated on the host
// (pb_pgen_torus_table_host, geometry_host.cpp); everything here is +,-,*,/ in the reference's order:
//   pgen_torus_fill    energy jitter, atmosphere at the floors, T / P / gamma1 from the EOS (torus.cpp:334-362)
//   pgen_torus_bfield  vector potential A_phi = max(<rho> - 0.2, 0) at cell corners and its curl (torus.cpp:435-470)
//   pgen_betas         the reductions of ComputeBetas (torus.cpp:849-930): beta_min, max b^2, max P over interior zones
//   scale_bfield       B *= sqrt(beta_norm / beta_target) on interior zones (torus.cpp:721-728)
// The caller sequences them with PrimitiveToConserved / ApplyFloors / the ghost exchange / FillDerived exactly as
// Mesh::Initialize and PostInitializationModifier do (phoebus_b200/pgen.py: init_torus).
#include <vector>

#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {

// Counter-based uniform deviate in [0, 1) from the zone's coordinates: decomposition-independent and reproducible.
// (The reference draws from a Kokkos XorShift64 pool, whose stream depends on the thread count; with
// torus/u_jitter = 0, as in the parity fixtures, the deviate is multiplied by zero.)
PB_DEV Real zone_drand(Real x1, Real x2, Real x3, unsigned long long seed) {
  unsigned long long h = seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
  const unsigned long long w[3] = {(unsigned long long)__double_as_longlong(x1), (unsigned long long)__double_as_longlong(x2),
                                   (unsigned long long)__double_as_longlong(x3)};
#pragma unroll
  for (int q = 0; q < 3; ++q) { // splitmix64 finaliser per word
    h ^= w[q];
    h ^= h >> 30;
    h *= 0xBF58476D1CE4E5B9ULL;
    h ^= h >> 27;
    h *= 0x94D049BB133111EBULL;
    h ^= h >> 31;
  }
  return (Real)(h >> 11) * (1.0 / 9007199254740992.0);
}

template <bool MHD>
__global__ void pgen_torus_fill_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *__restrict__ table, Real u_jitter,
                                       unsigned long long seed) {
  constexpr int NP = NC<MHD>::nprim, prs = NC<MHD>::prs, gm1 = NC<MHD>::gm1;
  const Lay L(cfg);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)L.ncell * pack.nblocks) return;
  const int b = (int)(t / L.ncell), o = (int)(t - (long long)b * L.ncell);
  const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
  const pb_block_t &blk = pack.blocks[b];
  const int tk = cfg.gk > 1 ? cfg.gk : 1; // 2-D table for an axisymmetric geometry cache
  const size_t N2 = (size_t)tk * L.nj * L.ni;
  const Real *tab = table + (size_t)blk.geom_slot * 5 * N2 + ((size_t)(tk > 1 ? k : 0) * L.nj + j) * L.ni + i;
  Real rho = __ldg(tab), u = __ldg(tab + N2);
  const Real v[3] = {__ldg(tab + 2 * N2), __ldg(tab + 3 * N2), __ldg(tab + 4 * N2)};
  const Real x1v = Xc(blk, 0, i), x2v = Xc(blk, 1, j), x3 = Xc(blk, 2, k);
  const Real drand = u_jitter != 0.0 ? zone_drand(x1v, x2v, x3, seed) : 0.5;
  u *= (1. + u_jitter * (drand - 0.5));
  Real rhoflr = 0, epsflr, A, B;
  load_floor(pack.floor_tab, blk.floor_slot, L.ni, 0, i, A, B);
  get_floors(cfg, A, B, x1v, x2v, x3, rhoflr, epsflr);
  rho = rho < rhoflr ? rhoflr : rho;
  u = u / rho < epsflr ? rho * epsflr : u;
  const Eos eos(cfg);
  const Real T = eos.T_rho_sie(rho, u / rho);
  const Real P = eos.P_rho_T(rho, T);
  Real *p = pack.prim + (size_t)b * NP * L.ncell + o;
  const size_t N = L.ncell;
  p[(size_t)PRHO * N] = rho;
#pragma unroll
  for (int d = 0; d < 3; ++d) p[(size_t)(PVEL + d) * N] = v[d];
  p[(size_t)PENG * N] = u;
  if (MHD) {
#pragma unroll
    for (int d = 0; d < 3; ++d) p[(size_t)(PB_ + d) * N] = 0.0;
  }
  p[(size_t)prs * N] = P;
  p[(size_t)gm1 * N] = eos.K_rho_T(rho, T) / P;
  pack.temp[(size_t)b * N + o] = T;
}

// A_phi at the lower-left corner of cell (j, i), from the density of plane k = 0 (kb.s of the entire domain);
// zero in the first row / column (the reference's array is zero-initialised there), torus.cpp:435-451
PB_DEV Real torus_aphi(const Real *rho0, const Lay &L, int j, int i) {
  if (j < 1 || i < 1) return 0.0;
  const Real *q = rho0 + (size_t)j * L.ni + i;
  const Real rho_av = 0.25 * (q[0] + q[-1] + q[-L.ni] + q[-L.ni - 1]);
  const Real a = rho_av - 0.2;
  return a > 0 ? a : 0.0;
}
__global__ void pgen_torus_bfield_kernel(const pb_pack_t pack, const pb_cfg_t cfg) {
  constexpr int NP = NC<true>::nprim;
  const Lay L(cfg);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)L.ncell * pack.nblocks) return;
  const int b = (int)(t / L.ncell), o = (int)(t - (long long)b * L.ncell);
  const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
  if (j > L.nj - 2 || i > L.ni - 2) return; // loop bounds jb.e - 1, ib.e - 1 (torus.cpp:455-456)
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  const Real gamdet = gt.at(0, G_DETGAM, gt.cell(k, j, i));
  const size_t N = L.ncell;
  Real *p = pack.prim + (size_t)b * NP * N;
  const Real *rho0 = p + (size_t)PRHO * N; // plane k = 0
  const Real a00 = torus_aphi(rho0, L, j, i), a10 = torus_aphi(rho0, L, j + 1, i), a01 = torus_aphi(rho0, L, j, i + 1),
             a11 = torus_aphi(rho0, L, j + 1, i + 1);
  p[(size_t)PB_ * N + o] = -(a00 - a10 + a01 - a11) / (2.0 * blk.dx[1] * gamdet);
  p[(size_t)(PB_ + 1) * N + o] = (a00 + a10 - a01 - a11) / (2.0 * blk.dx[0] * gamdet);
  p[(size_t)(PB_ + 2) * N + o] = 0.0;
}

// ComputeBetas, torus.cpp:849-930: one CTA per MeshBlock; min / max are order independent, so a tree is exact
__global__ void __launch_bounds__(256) pgen_betas_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real rho_min_bnorm, Real *partial) {
  constexpr int NP = NC<true>::nprim, prs = NC<true>::prs;
  const Lay L(cfg);
  const int b = blockIdx.x;
  const size_t N = L.ncell;
  const int nxi = L.ie - L.is + 1, nxj = L.je - L.js + 1, nxk = L.ke - L.ks + 1;
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  Real beta_min = INFINITY, bsq_max = -INFINITY, press_max = -INFINITY;
  for (int z = threadIdx.x; z < nxi * nxj * nxk; z += blockDim.x) {
    const int i = L.is + z % nxi, j = L.js + (z / nxi) % nxj, k = L.ks + z / (nxi * nxj);
    const int o = L.idx(k, j, i);
    const Real *p = pack.prim + (size_t)b * NP * N + o;
    PointGeom g;
    g.load(gt, 0, gt.cell(k, j, i));
    const Real vp[3] = {p[(size_t)PVEL * N], p[(size_t)(PVEL + 1) * N], p[(size_t)(PVEL + 2) * N]};
    const Real Bp[3] = {p[(size_t)PB_ * N], p[(size_t)(PB_ + 1) * N], p[(size_t)(PB_ + 2) * N]};
    const Real W = lorentz_factor(vp, g);
    Real Bsq = 0, Bdotv = 0; // GetMagneticFieldSquared, relativity_utils.hpp:131-141
    for (int m = 0; m < 3; ++m)
      for (int n = 0; n < 3; ++n) {
        Bdotv += g.g4(m + 1, n + 1) * ratio(vp[m], W) * Bp[n];
        Bsq += g.g4(m + 1, n + 1) * Bp[m] * Bp[n];
      }
    const Real bsq = ratio(Bsq, W * W) + Bdotv * Bdotv;
    const Real P = p[(size_t)prs * N];
    const Real beta = ratio(P, 0.5 * bsq);
    if (p[(size_t)PRHO * N] > rho_min_bnorm && beta < beta_min) beta_min = beta;
    bsq_max = fmax(bsq, bsq_max); // std::max(bsq, bsq_max) differs from fmax only for NaN
    press_max = fmax(P, press_max);
  }
  __shared__ Real sh[3][256];
  sh[0][threadIdx.x] = beta_min;
  sh[1][threadIdx.x] = bsq_max;
  sh[2][threadIdx.x] = press_max;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      sh[0][threadIdx.x] = fmin(sh[0][threadIdx.x], sh[0][threadIdx.x + off]);
      sh[1][threadIdx.x] = fmax(sh[1][threadIdx.x], sh[1][threadIdx.x + off]);
      sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + off]);
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[(size_t)b * 3 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void scale_bfield_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real fac) {
  constexpr int NP = NC<true>::nprim;
  const Lay L(cfg);
  const int nxi = L.ie - L.is + 1, nxj = L.je - L.js + 1, nxk = L.ke - L.ks + 1;
  const long long nz = (long long)nxi * nxj * nxk;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nz * pack.nblocks) return;
  const int b = (int)(t / nz);
  const int z = (int)(t - (long long)b * nz);
  const int i = L.is + z % nxi, j = L.js + (z / nxi) % nxj, k = L.ks + z / (nxi * nxj);
  Real *p = pack.prim + (size_t)b * NP * L.ncell + L.idx(k, j, i);
#pragma unroll
  for (int d = 0; d < 3; ++d) p[(size_t)(PB_ + d) * L.ncell] *= fac;
}

} // namespace pb

using namespace pb;

extern "C" int pb_pgen_torus_fill(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *table_dev, double u_jitter,
                                  unsigned long long seed, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!table_dev || !pack->temp) return set_error(PB_EINVAL, "pb_pgen_torus_fill: NULL table or temperature field");
  if (cfg->ndim < 2) return set_error(PB_EINVAL, "the torus problem needs at least two dimensions");
  const Lay L(*cfg);
  const long long n = (long long)L.ncell * pack->nblocks;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (cfg->mhd) pgen_torus_fill_kernel<true><<<grid, 256, 0, s>>>(*pack, *cfg, table_dev, u_jitter, seed);
  else pgen_torus_fill_kernel<false><<<grid, 256, 0, s>>>(*pack, *cfg, table_dev, u_jitter, seed);
  return check_launch("pgen_torus_fill_kernel");
}

extern "C" int pb_pgen_torus_bfield(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cfg->mhd) return PB_OK; // ref: torus.cpp:453 (ibhi > 0)
  if (cfg->ndim < 2) return set_error(PB_EINVAL, "the torus problem needs at least two dimensions");
  const Lay L(*cfg);
  const long long n = (long long)L.ncell * pack->nblocks;
  pgen_torus_bfield_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*pack, *cfg);
  return check_launch("pgen_torus_bfield_kernel");
}

extern "C" int pb_pgen_betas(const pb_pack_t *pack, const pb_cfg_t *cfg, double rho_min_bnorm, double *partial_dev,
                             double out_host[3], void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!partial_dev || !out_host) return set_error(PB_EINVAL, "pb_pgen_betas: NULL argument");
  if (!cfg->mhd) return set_error(PB_EINVAL, "pb_pgen_betas needs fluid/mhd = true");
  cudaStream_t s = (cudaStream_t)stream;
  pgen_betas_kernel<<<pack->nblocks, 256, 0, s>>>(*pack, *cfg, rho_min_bnorm, partial_dev);
  rc = check_launch("pgen_betas_kernel");
  if (rc) return rc;
  std::vector<double> h((size_t)pack->nblocks * 3);
  cudaError_t e = cudaMemcpyAsync(h.data(), partial_dev, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return set_cuda_error(e, "pb_pgen_betas");
  out_host[0] = INFINITY;
  out_host[1] = out_host[2] = -INFINITY;
  for (int b = 0; b < pack->nblocks; ++b) {
    out_host[0] = h[(size_t)b * 3] < out_host[0] ? h[(size_t)b * 3] : out_host[0];
    out_host[1] = h[(size_t)b * 3 + 1] > out_host[1] ? h[(size_t)b * 3 + 1] : out_host[1];
    out_host[2] = h[(size_t)b * 3 + 2] > out_host[2] ? h[(size_t)b * 3 + 2] : out_host[2];
  }
  return PB_OK;
}

extern "C" int pb_scale_bfield(const pb_pack_t *pack, const pb_cfg_t *cfg, double factor, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cfg->mhd) return PB_OK;
  const long long n = (long long)cfg->nx[0] * (cfg->ndim >= 2 ? cfg->nx[1] : 1) * (cfg->ndim >= 3 ? cfg->nx[2] : 1) * pack->nblocks;
  scale_bfield_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*pack, *cfg, factor);
  return check_launch("scale_bfield_kernel");
}
