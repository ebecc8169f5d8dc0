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

__global__ void halo_local_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const HaloFields hf) {
  const Lay L(cfg);
  const Shell sh(L);
  const int N = L.ncell;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)sh.n * pack.nblocks) return;
  const int b = (int)(t / sh.n);
  int i, j, k;
  sh.cell(L, (int)(t % sh.n), k, j, i);
  const int o = L.idx(k, j, i);
  const int o1 = i < L.is ? -1 : (i > L.ie ? 1 : 0);
  const int o2 = j < L.js ? -1 : (j > L.je ? 1 : 0);
  const int o3 = k < L.ks ? -1 : (k > L.ke ? 1 : 0);
  if (o1 == 0 && o2 == 0 && o3 == 0) return;
  const int nb = pack.blocks[b].nbr[(o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)];
  if (nb < 0) return;
  const int si = i - o1 * cfg.nx[0], sj = j - o2 * cfg.nx[1], sk = k - o3 * cfg.nx[2];
  const int so = L.idx(sk, sj, si);
  for (int f = 0; f < hf.nfield; ++f) {
    const int nc = hf.ncomp[f];
    const Real *src = hf.ptr[f] + (size_t)nb * nc * N + so;
    Real *dst = hf.ptr[f] + (size_t)b * nc * N + o;
    int n = 0;
#ifndef PB_HALO_ILP
#define PB_HALO_ILP 4
#endif
    for (; n + PB_HALO_ILP <= nc; n += PB_HALO_ILP) { // independent loads in flight per thread (the copy is latency bound otherwise)
      Real v[PB_HALO_ILP];
#pragma unroll
      for (int u = 0; u < PB_HALO_ILP; ++u) v[u] = src[(size_t)(n + u) * N];
#pragma unroll
      for (int u = 0; u < PB_HALO_ILP; ++u) dst[(size_t)(n + u) * N] = v[u];
    }
    for (; n < nc; ++n) dst[(size_t)n * N] = src[(size_t)n * N];
  }
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

extern "C" int pb_halo_exchange_local(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  const Lay L(*cfg);
  const Shell sh(L);
  const long long n = (long long)sh.n * pack->nblocks;
  if (n <= 0) return PB_OK;
  const int T = 256;
  halo_local_kernel<<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(*pack, *cfg, halo_fields(*pack, *cfg, cons));
  return check_launch("halo_local_kernel");
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
