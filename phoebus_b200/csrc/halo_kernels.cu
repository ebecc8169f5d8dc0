// Ghost-zone exchange. Same-GPU neighbours are read directly (no staging buffers, unlike the
// reference's pack -> MPI buffer -> unpack path for same-rank neighbours,
// ref: external/parthenon/src/bvals/comms/boundary_communication.cpp:49-346); neighbours on other
// GPUs go through per-peer aggregated buffers filled/drained by pb_halo_pack / pb_halo_unpack and
// moved by NCCL send/recv in the caller.
//
// Variable set: every Metadata::FillGhost field — the conserved
// variables, plus, with bc_vars = primitive, all primitives, temperature, c2p_mu and fail
// (ref: src/fluid/fluid.cpp:156-175,198-236).
#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {
int launch_twins(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, cudaStream_t s); // zone_kernels.cu

struct HaloFields {
  Real *ptr[5];  // cons, prim, temp, c2p_mu, fail
  int ncomp[5];
  int nfield;
};

static HaloFields halo_fields(const pb_pack_t &pack, const pb_cfg_t &cfg, Real *cons) {
  HaloFields f;
  const int NP = 7 + 3 * cfg.mhd, NCn = 5 + 3 * cfg.mhd;
  f.ptr[0] = cons;
  f.ncomp[0] = NCn;
  f.nfield = 1;
  if (cfg.bc_prim) {
    f.ptr[1] = pack.prim;
    f.ncomp[1] = NP;
    f.ptr[2] = pack.temp;
    f.ncomp[2] = 1;
    f.ptr[3] = pack.c2p_mu;
    f.ncomp[3] = 1;
    f.ptr[4] = pack.fail;
    f.ncomp[4] = 1;
    f.nfield = 5;
  }
  return f;
}

// Copy of one ghost zone: all components of a field group are loaded before any is stored — source and destination live
// in the same arrays, so the compiler cannot hoist loads over stores by itself, and a zone's 8 (conserved) or 8 + 13 (with
// bc_vars = primitive) words would otherwise be that many serial DRAM round trips (measured: 1.63 -> 1.31 ms per exchange
// at 256^3). Measured and dropped (profiles/r02i_variants_halo_mapping.txt, r02j, r02n): one load group for everything, 16-byte
// accesses for pairs of zones, a precomputed (destination, source) list instead of the index decoding, higher occupancy —
// the copy is bound by the memory system (128-byte line fills for 32-byte row ends), not by instructions or latency.
template <bool MHD, bool PRIM, int PHASE>
PB_DEV void halo_copy_zone(const pb_pack_t &pack, const HaloFields &hf, const int ndim, const int N, const size_t b, const size_t nb,
                           const int o, const int so) {
  constexpr int NCn = NC<MHD>::ncons, NP = NC<MHD>::nprim;
  {
    const Real *src = hf.ptr[0] + nb * NCn * N + so;
    Real *dst = hf.ptr[0] + b * NCn * N + o;
    Real v[NCn];
#pragma unroll
    for (int n = 0; n < NCn; ++n) v[n] = src[(size_t)n * N];
#pragma unroll
    for (int n = 0; n < NCn; ++n) dst[(size_t)n * N] = v[n];
  }
  if (PRIM) {
    const Real *src = hf.ptr[1] + nb * NP * N + so;
    Real *dst = hf.ptr[1] + b * NP * N + o;
    Real q[NP], w[3];
#pragma unroll
    for (int n = 0; n < NP; ++n) q[n] = src[(size_t)n * N];
#pragma unroll
    for (int f = 0; f < 3; ++f) w[f] = hf.ptr[2 + f][nb * N + so];
#pragma unroll
    for (int n = 0; n < NP; ++n) dst[(size_t)n * N] = q[n];
#pragma unroll
    for (int f = 0; f < 3; ++f) hf.ptr[2 + f][b * N + o] = w[f];
  }
  if (PHASE == 2) {
    Real c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d < ndim) c[d] = pack.csig[(nb * ndim + d) * N + so];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d < ndim) pack.csig[(b * ndim + d) * N + o] = c[d];
    if (pack.zflags) pack.zflags[b * N + o] = pack.zflags[nb * N + so];
  }
}

// Phase 2 for a twin zone: every field from the side-buffer entry (twin_kernel, zone_kernels.cu) instead of the neighbour
template <bool MHD>
PB_DEV void halo_copy_twin(const pb_pack_t &pack, const HaloFields &hf, const pb_cfg_t &cfg, const Lay &L, const int b, const int k,
                           const int j, const int i) {
  constexpr int NCn = NC<MHD>::ncons, NP = NC<MHD>::nprim;
  const TwinBuf tb(pack, L, cfg);
  const int e = tb.entry(b, k, j, i);
  const size_t S = (size_t)tb.S, N = (size_t)L.ncell, o = (size_t)L.idx(k, j, i);
  const Real *src = tb.base + e;
  Real v[NCn], q[NP], w[3], c[3];
#pragma unroll
  for (int n = 0; n < NCn; ++n) v[n] = src[n * S];
#pragma unroll
  for (int n = 0; n < NP; ++n) q[n] = src[(NCn + n) * S];
#pragma unroll
  for (int f = 0; f < 3; ++f) w[f] = src[(NCn + NP + f) * S];
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (d < L.ndim) c[d] = src[(NCn + NP + 3 + d) * S];
  const unsigned char zf = tb.flags[e];
  Real *cd = hf.ptr[0] + (size_t)b * NCn * N + o;
#pragma unroll
  for (int n = 0; n < NCn; ++n) cd[n * N] = v[n];
  Real *pd = hf.ptr[1] + (size_t)b * NP * N + o;
#pragma unroll
  for (int n = 0; n < NP; ++n) pd[n * N] = q[n];
#pragma unroll
  for (int f = 0; f < 3; ++f) hf.ptr[2 + f][(size_t)b * N + o] = w[f];
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (d < L.ndim) pack.csig[((size_t)b * L.ndim + d) * N + o] = c[d];
  if (pack.zflags) pack.zflags[(size_t)b * N + o] = zf;
}

// One thread per ghost zone (Shell enumeration: rows along x1 stay contiguous).
// PHASE (ghost plan, ghost_plan_host.cpp): 0 every ghost zone; 1 skip the zones flagged in pack.ghost_mask; 2 only
// those, after con2prim, with the cell signal speeds and the zone flags on top of the exchanged fields.
template <bool MHD, bool PRIM, int PHASE>
__global__ void __launch_bounds__(256) halo_local_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const HaloFields hf) {
  const Lay L(cfg);
  const Shell sh(L);
  const int N = L.ncell;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)sh.n * pack.nblocks) return;
  const int b = (int)(t / sh.n);
  int i, j, k;
  sh.cell(L, (int)(t % sh.n), k, j, i);
  const int o = L.idx(k, j, i);
  if (PHASE != 0) {
    const unsigned char m = pack.ghost_mask[(size_t)b * N + o];
    if ((m != 0) != (PHASE == 2)) return;
    if (PHASE == 2 && m == 2) { // twin: the zone's own con2prim result waits in the side buffer (zone_kernels.cu)
      halo_copy_twin<MHD>(pack, hf, cfg, L, b, k, j, i);
      return;
    }
  }
  const int o1 = i < L.is ? -1 : (i > L.ie ? 1 : 0);
  const int o2 = j < L.js ? -1 : (j > L.je ? 1 : 0);
  const int o3 = k < L.ks ? -1 : (k > L.ke ? 1 : 0);
  if (o1 == 0 && o2 == 0 && o3 == 0) return;
  const int nb = pack.blocks[b].nbr[(o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)];
  if (nb < 0) return;
  const int si = i - o1 * cfg.nx[0], sj = j - o2 * cfg.nx[1], sk = k - o3 * cfg.nx[2];
  halo_copy_zone<MHD, PRIM, PHASE>(pack, hf, L.ndim, N, b, nb, o, L.idx(sk, sj, si));
}

// plan entry: {block, i0, i1, j0, j1, k0, k1, off_lo, off_hi, reserved}; offset in doubles
template <bool PACK>
__global__ void halo_buffer_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const HaloFields hf, const int *plan,
                                   Real *buf) {
  const Lay L(cfg);
  const int N = L.ncell;
  const int *p = plan + (size_t)blockIdx.x * 10;
  const int b = p[0], i0 = p[1], i1 = p[2], j0 = p[3], j1 = p[4], k0 = p[5], k1 = p[6];
  const unsigned long long off = (unsigned long long)(unsigned)p[7] | ((unsigned long long)(unsigned)p[8] << 32);
  const int ni = i1 - i0 + 1, nj = j1 - j0 + 1, nk = k1 - k0 + 1;
  const int nz = ni * nj * nk;
  Real *seg = buf + off;
  int var0 = 0;
  for (int f = 0; f < hf.nfield; ++f) {
    const int nc = hf.ncomp[f];
    Real *fld = hf.ptr[f] + (size_t)b * nc * N;
    // gridDim.y CTAs share one segment; each thread moves all components of its zones, four loads in flight
    for (int z = blockIdx.y * blockDim.x + threadIdx.x; z < nz; z += gridDim.y * blockDim.x) {
      const int i = i0 + z % ni, j = j0 + (z / ni) % nj, k = k0 + z / (ni * nj);
      Real *g = fld + L.idx(k, j, i);
      Real *s = seg + (size_t)var0 * nz + z;
      int n = 0;
      for (; n + 4 <= nc; n += 4) {
        Real v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = PACK ? g[(size_t)(n + u) * N] : s[(size_t)(n + u) * nz];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (PACK) s[(size_t)(n + u) * nz] = v[u];
          else g[(size_t)(n + u) * N] = v[u];
        }
      }
      for (; n < nc; ++n) {
        if (PACK) s[(size_t)n * nz] = g[(size_t)n * N];
        else g[(size_t)n * N] = s[(size_t)n * nz];
      }
    }
    var0 += nc;
  }
}

} // namespace pb

using namespace pb;

extern "C" int pb_halo_nvar(const pb_cfg_t *cfg) {
  const int NP = 7 + 3 * cfg->mhd, NCn = 5 + 3 * cfg->mhd;
  return NCn + (cfg->bc_prim ? NP + 3 : 0);
}

extern "C" int pb_halo_exchange_local_ex(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, int phase, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (phase < 0 || phase > 2) return set_error(PB_EINVAL, "pb_halo_exchange_local_ex: phase must be 0, 1 or 2");
  if (phase != 0 && !pack->ghost_mask) return set_error(PB_EINVAL, "pb_halo_exchange_local_ex: phase 1 / 2 need pack->ghost_mask");
  if (phase == 2 && !cfg->bc_prim) return PB_OK; // no zone is ever deferred without bc_vars = primitive
  if (phase == 2) PB_NEED("pb_halo_exchange_local_ex", PB_F(pack, csig));
  const Lay L(*cfg);
  const Shell sh(L);
  const long long n = (long long)sh.n * pack->nblocks;
  if (n <= 0) return PB_OK;
  const int T = 256;
  const unsigned G = (unsigned)((n + T - 1) / T);
  const HaloFields hf = halo_fields(*pack, *cfg, cons);
  cudaStream_t st = (cudaStream_t)stream;
#define PB_HALO(M, P)                                                                   \
  do {                                                                                  \
    if (phase == 0) halo_local_kernel<M, P, 0><<<G, T, 0, st>>>(*pack, *cfg, hf);       \
    else if (phase == 1) halo_local_kernel<M, P, 1><<<G, T, 0, st>>>(*pack, *cfg, hf);  \
    else halo_local_kernel<M, P, 2><<<G, T, 0, st>>>(*pack, *cfg, hf);                  \
  } while (0)
  if (cfg->mhd) {
    if (cfg->bc_prim) PB_HALO(true, true);
    else PB_HALO(true, false);
  } else {
    if (cfg->bc_prim) PB_HALO(false, true);
    else PB_HALO(false, false);
  }
#undef PB_HALO
  rc = check_launch("halo_local_kernel");
  if (rc) return rc;
  // the twins are solved with every pre-con2prim copy, from their neighbours' inputs before any con2prim overwrites them
  // (phase 0 too: a full exchange may precede a planned post phase)
  if (phase != 2 && pack->ghost_mask && cfg->bc_prim) return launch_twins(pack, cfg, cons, st);
  return PB_OK;
}

extern "C" int pb_halo_exchange_local(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, void *stream) {
  return pb_halo_exchange_local_ex(pack, cfg, cons, 0, stream);
}

extern "C" int pb_halo_pack(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, const int *plan, int nseg,
                            double *buf, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (nseg <= 0) return PB_OK;
  halo_buffer_kernel<true><<<dim3(nseg, 8), 256, 0, (cudaStream_t)stream>>>(*pack, *cfg, halo_fields(*pack, *cfg, const_cast<double *>(cons)), plan, buf);
  return check_launch("halo_pack");
}

extern "C" int pb_halo_unpack(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int *plan, int nseg,
                              const double *buf, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (nseg <= 0) return PB_OK;
  halo_buffer_kernel<false><<<dim3(nseg, 8), 256, 0, (cudaStream_t)stream>>>(*pack, *cfg, halo_fields(*pack, *cfg, cons), plan, const_cast<double *>(buf));
  return check_launch("halo_unpack");
}
