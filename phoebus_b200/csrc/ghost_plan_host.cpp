// Which ghost zones need their own con2prim.
//
// The reference converts conserved to primitive over the ENTIRE MeshBlock after every ghost exchange
// (FillDerived, ref: src/fluid/fluid.cpp:477-517 called with the entire index range from the stage task list,
// src/phoebus_driver.cpp:640-663): 40^3 zones for a 32^3 block, 49 % of them ghosts. A ghost zone that was filled by
// the exchange holds the neighbour's conserved state and — with bc_vars = primitive — the neighbour's previous
// primitives as the root finder's guess, so the reference computes there exactly what the
// neighbour computes for its own interior zone, PROVIDED the two blocks' tables agree bit for bit at that zone (metric,
// floor factors: block coordinates are xmin_block + (i + 1/2) dx_block and can differ in the last place between
// blocks). For those zones the B200 path copies the neighbour's result (all fields con2prim writes, plus the conserved
// state it was computed from) after con2prim instead of solving again. This function decides, zone by zone, where that
// is exact:
//   * bc_vars = primitive (the guess travels), floors not coordinate-dependent beyond the tabulated x1 factors;
//   * the zone lies only in faces towards another MeshBlock of the same pack (no physical, periodic or remote face:
//     ConvertBoundaryConditions rewrites the conserved state of every ghost zone on a mesh boundary, periodic ones
//     included, ref: src/phoebus_boundaries/phoebus_boundaries.cpp:303-448);
//   * the zone is not read by a physical ghost fill of its own block (the first ng interior layers next to a physical
//     face, across the whole transverse extent, ref: phoebus_boundaries.cpp:44-270);
//   * the cell-centred geometry words and the cell floor factors of the ghost zone equal the neighbour's, bitwise.
// Everything else keeps its own con2prim. Tests: tests/test_ghost_plan.py (CPU), tests/test_gpu_parity.py (every
// fixture runs with the plan on and off).
#include <cstdint>
#include <cstring>

#include "../../include/phoebus_b200.h"

namespace {
struct HostLay {
  int ndim, ng, ni, nj, nk, lo[3], hi[3], ncell;
  explicit HostLay(const pb_cfg_t &c) {
    ndim = c.ndim;
    ng = c.ng;
    ni = c.nx[0] + 2 * ng;
    nj = ndim >= 2 ? c.nx[1] + 2 * ng : 1;
    nk = ndim >= 3 ? c.nx[2] + 2 * ng : 1;
    lo[0] = ng;
    hi[0] = ng + c.nx[0] - 1;
    lo[1] = ndim >= 2 ? ng : 0;
    hi[1] = ndim >= 2 ? ng + c.nx[1] - 1 : 0;
    lo[2] = ndim >= 3 ? ng : 0;
    hi[2] = ndim >= 3 ? ng + c.nx[2] - 1 : 0;
    ncell = ni * nj * nk;
  }
  int idx(int k, int j, int i) const { return (k * nj + j) * ni + i; }
};
inline bool same_bits(double a, double b) { return std::memcmp(&a, &b, sizeof(double)) == 0; }
} // namespace

extern "C" long long pb_twin_entries(const pb_cfg_t *cfg, int nblocks) {
  if (!cfg || nblocks <= 0) return 0;
  const HostLay L(*cfg);
  return (long long)nblocks * 2 * L.ng * L.nk * L.nj;
}

extern "C" int pb_ghost_plan_host(const pb_cfg_t *cfg, const pb_block_t *blocks, int nblocks, const double *geom_host,
                                  const double *floor_host, unsigned char *mask_out, unsigned *zones_out,
                                  long long *nzones_out, unsigned *twins_out, long long *ntwins_out) {
  if (!cfg || !blocks || nblocks <= 0 || !geom_host || !floor_host || !mask_out || !zones_out || !nzones_out) return PB_EINVAL;
  const HostLay L(*cfg);
  if ((long long)nblocks * L.ncell >= (1ll << 32)) return PB_EINVAL;
  const size_t gcell = (size_t)cfg->gk * cfg->gj * cfg->gi;
  auto gcell_of = [&](int k, int j, int i) {
    const int kk = cfg->gk == 1 ? 0 : k, jj = cfg->gj == 1 ? 0 : j, ii = cfg->gi == 1 ? 0 : i;
    return ((size_t)kk * cfg->gj + jj) * cfg->gi + ii;
  };
  std::memset(mask_out, 0, (size_t)nblocks * L.ncell);
  const bool possible = cfg->bc_prim && cfg->floor_type != 4;
  // twin entries live in a side buffer of nblocks * 2 ng * nk * nj entries addressed with 32-bit strides
  const bool twins = twins_out && ntwins_out && (long long)nblocks * 2 * L.ng * L.nk * L.nj < (1ll << 31);
  if (ntwins_out) *ntwins_out = 0;
  const int nx[3] = {cfg->nx[0], cfg->nx[1], cfg->nx[2]};
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblocks; ++b) {
    if (!possible) continue;
    const pb_block_t &blk = blocks[b];
    for (int k = 0; k < L.nk; ++k)
      for (int j = 0; j < L.nj; ++j)
        for (int i = 0; i < L.ni; ++i) {
          const int c[3] = {i, j, k};
          int off[3] = {0, 0, 0};
          bool ok = true, ghost = false;
          for (int d = 0; d < L.ndim && ok; ++d) {
            if (c[d] < L.lo[d]) off[d] = -1;
            else if (c[d] > L.hi[d]) off[d] = 1;
            if (off[d]) {
              ghost = true;
              if (blk.bc[2 * d + (off[d] > 0)] != PB_BC_BLOCK) ok = false;
            }
            // source layers of a physical ghost fill
            if (blk.bc[2 * d] > 0 && c[d] >= L.lo[d] && c[d] < L.lo[d] + L.ng) ok = false;
            if (blk.bc[2 * d + 1] > 0 && c[d] <= L.hi[d] && c[d] > L.hi[d] - L.ng) ok = false;
          }
          if (!ghost || !ok) continue;
          const int nb = blk.nbr[(off[0] + 1) + 3 * (off[1] + 1) + 9 * (off[2] + 1)];
          if (nb < 0 || nb >= nblocks) continue;
          const pb_block_t &src = blocks[nb];
          // the neighbour must name this block at the opposite offset (the push form of the copy starts from the source)
          if (src.nbr[(-off[0] + 1) + 3 * (-off[1] + 1) + 9 * (-off[2] + 1)] != b) continue;
          const int si = i - off[0] * nx[0], sj = j - off[1] * nx[1], sk = k - off[2] * nx[2];
          // tables
          const size_t g0 = gcell_of(k, j, i), g1 = gcell_of(sk, sj, si);
          if (!(blk.geom_slot == src.geom_slot && g0 == g1)) {
            const double *t0 = geom_host + (size_t)blk.geom_slot * PB_GEOM_WORDS * gcell;
            const double *t1 = geom_host + (size_t)src.geom_slot * PB_GEOM_WORDS * gcell;
            for (int w = 0; w < PB_NGLOC && ok; ++w) ok = same_bits(t0[(size_t)w * gcell + g0], t1[(size_t)w * gcell + g1]);
          }
          if (ok && !(blk.floor_slot == src.floor_slot && i == si)) {
            const double *f0 = floor_host + (size_t)blk.floor_slot * 4 * L.ni;
            const double *f1 = floor_host + (size_t)src.floor_slot * 4 * L.ni;
            ok = same_bits(f0[i], f1[si]) && same_bits(f0[L.ni + i], f1[L.ni + si]);
          }
          // 1: the neighbour's result is the ghost zone's result. 2 ("twin"): same inputs, but this block's tables differ
          // from the neighbour's in the last place (x1 ghost columns: block coordinates are xmin_block + (i + 1/2) dx_block)
          // — solved from the neighbour's inputs with THIS block's tables in a side buffer (zone_kernels.cu: twin kernels),
          // so that neither a sparse pre-con2prim copy nor 8-byte holes in the deferred copy remain
          if (ok) mask_out[(size_t)b * L.ncell + L.idx(k, j, i)] = 1;
          else if (twins && off[0] != 0) mask_out[(size_t)b * L.ncell + L.idx(k, j, i)] = 2;
        }
  }
  if (twins) { // ordered (block, x1 ghost column, k, j): consecutive entries are consecutive in the side buffer
    long long m = 0;
    for (int b = 0; b < nblocks; ++b)
      for (int c = 0; c < 2 * L.ng; ++c) {
        const int i = c < L.ng ? c : c + nx[0];
        for (int k = 0; k < L.nk; ++k)
          for (int j = 0; j < L.nj; ++j) {
            const size_t z = (size_t)b * L.ncell + L.idx(k, j, i);
            if (mask_out[z] == 2) twins_out[m++] = (unsigned)z;
          }
      }
    *ntwins_out = m;
  }
  // the ghost zones that keep their own con2prim (interior zones are enumerated arithmetically by the kernels)
  long long n = 0;
  for (int b = 0; b < nblocks; ++b) {
    const unsigned base = (unsigned)((size_t)b * L.ncell);
    for (int k = 0; k < L.nk; ++k)
      for (int j = 0; j < L.nj; ++j)
        for (int i = 0; i < L.ni; ++i) {
          const bool in = i >= L.lo[0] && i <= L.hi[0] && j >= L.lo[1] && j <= L.hi[1] && k >= L.lo[2] && k <= L.hi[2];
          const int o = L.idx(k, j, i);
          if (!in && !mask_out[(size_t)base + o]) zones_out[n++] = base + (unsigned)o;
        }
  }
  *nzones_out = n;
  return PB_OK;
}
