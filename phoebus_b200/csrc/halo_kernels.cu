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

// Copy of one ghost zone (T = double) or of two zones adjacent along x1 (T = double2; offsets even): all components of a
// field group are loaded before any is stored — source and destination live in the same arrays, so the compiler cannot
// hoist loads over stores by itself, and a zone's 8 (conserved) or 8 + 13 (with bc_vars = primitive) words would
// otherwise be that many serial DRAM round trips. HOIST: one group for everything (more registers, one round trip).
template <class T, bool MHD, bool PRIM, int PHASE, bool HOIST>
PB_DEV void halo_copy_zone(const pb_pack_t &pack, const HaloFields &hf, const int ndim, const int N, const size_t b, const size_t nb,
                           const int o, const int so) {
  constexpr int NCn = NC<MHD>::ncons, NP = NC<MHD>::nprim;
  auto at = [](Real *p) { return reinterpret_cast<T *>(p); };
  T *csrc = at(hf.ptr[0] + nb * NCn * N + so), *cdst = at(hf.ptr[0] + b * NCn * N + o);
  constexpr int S = sizeof(T) / sizeof(Real);
  const size_t NS = (size_t)N / S; // component stride in units of T
  T v[NCn], q[PRIM ? NP : 1], w[3], c[3];
#pragma unroll
  for (int n = 0; n < NCn; ++n) v[n] = csrc[n * NS];
  if (!HOIST) {
#pragma unroll
    for (int n = 0; n < NCn; ++n) cdst[n * NS] = v[n];
  }
  if (PRIM) {
    T *src = at(hf.ptr[1] + nb * NP * N + so), *dst = at(hf.ptr[1] + b * NP * N + o);
#pragma unroll
    for (int n = 0; n < NP; ++n) q[n] = src[n * NS];
#pragma unroll
    for (int f = 0; f < 3; ++f) w[f] = *at(hf.ptr[2 + f] + nb * N + so);
    if (PHASE == 2 && HOIST) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (d < ndim) c[d] = *at(pack.csig + (nb * ndim + d) * N + so);
    }
    if (HOIST) {
#pragma unroll
      for (int n = 0; n < NCn; ++n) cdst[n * NS] = v[n];
    }
#pragma unroll
    for (int n = 0; n < NP; ++n) dst[n * NS] = q[n];
#pragma unroll
    for (int f = 0; f < 3; ++f) *at(hf.ptr[2 + f] + b * N + o) = w[f];
  } else if (HOIST) {
#pragma unroll
    for (int n = 0; n < NCn; ++n) cdst[n * NS] = v[n];
  }
  if (PHASE == 2) {
    if (!(PRIM && HOIST)) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (d < ndim) c[d] = *at(pack.csig + (nb * ndim + d) * N + so);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (d < ndim) *at(pack.csig + (b * ndim + d) * N + o) = c[d];
    if (pack.zflags) {
      for (int u = 0; u < S; ++u) pack.zflags[b * N + o + u] = pack.zflags[nb * N + so + u];
    }
  }
}

// PHASE (ghost plan, ghost_plan_host.cpp): 0 every ghost zone; 1 skip the zones flagged in pack.ghost_mask; 2 only
// those, after con2prim, with the cell signal speeds and the zone flags on top of the exchanged fields.
// VAR 0: one thread per ghost zone; 1: the same with a single load group; 2: one thread per pair of zones adjacent along
// x1 (16-byte accesses; needs even nx1 and ng — ghost / interior boundaries then fall between pairs).
template <bool MHD, bool PRIM, int PHASE, int VAR>
__global__ void __launch_bounds__(256) halo_local_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const HaloFields hf) {
  const Lay L(cfg);
  Lay Lh = L;
  if (VAR == 2) {
    Lh.ni = L.ni / 2;
    Lh.is = L.is / 2;
    Lh.ie = (L.ie + 1) / 2 - 1;
  }
  const Shell sh(Lh);
  const int N = L.ncell;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)sh.n * pack.nblocks) return;
  const int b = (int)(t / sh.n);
  int i, j, k;
  sh.cell(Lh, (int)(t % sh.n), k, j, i);
  if (VAR == 2) i *= 2;
  const int o = L.idx(k, j, i);
  bool first = true, second = VAR == 2;
  if (PHASE != 0) {
    const unsigned char *m = pack.ghost_mask + (size_t)b * N + o;
    first = (m[0] != 0) == (PHASE == 2);
    if (VAR == 2) second = (m[1] != 0) == (PHASE == 2);
    if (!first && !second) return;
  }
  const int o1 = i < L.is ? -1 : (i > L.ie ? 1 : 0);
  const int o2 = j < L.js ? -1 : (j > L.je ? 1 : 0);
  const int o3 = k < L.ks ? -1 : (k > L.ke ? 1 : 0);
  if (o1 == 0 && o2 == 0 && o3 == 0) return;
  const int nb = pack.blocks[b].nbr[(o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)];
  if (nb < 0) return;
  const int si = i - o1 * cfg.nx[0], sj = j - o2 * cfg.nx[1], sk = k - o3 * cfg.nx[2];
  const int so = L.idx(sk, sj, si);
  if (VAR == 2) {
    if (first && second) halo_copy_zone<double2, MHD, PRIM, PHASE, false>(pack, hf, L.ndim, N, b, nb, o, so);
    else halo_copy_zone<double, MHD, PRIM, PHASE, false>(pack, hf, L.ndim, N, b, nb, o + (second ? 1 : 0), so + (second ? 1 : 0));
  } else {
    halo_copy_zone<double, MHD, PRIM, PHASE, VAR == 1>(pack, hf, L.ndim, N, b, nb, o, so);
  }
}

// The same copy driven by the ghost plan's (destination, source) list: no index decoding, no idle lanes — the two phases
// of a planned stage each visit only their own zones.
template <bool MHD, bool PRIM, int PHASE>
__global__ void __launch_bounds__(256) halo_list_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const HaloFields hf,
                                                        const uint2 *__restrict__ pairs, const long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const Lay L(cfg);
  const unsigned N = (unsigned)L.ncell;
  const uint2 p = __ldg(pairs + t);
  const unsigned b = p.x / N, nb = p.y / N;
  halo_copy_zone<double, MHD, PRIM, PHASE, false>(pack, hf, L.ndim, (int)N, b, nb, (int)(p.x - b * N), (int)(p.y - nb * N));
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
  static const bool use_list = [] {
    const char *e = getenv("PB_HALO_LIST");
    return !(e && e[0] == '0');
  }();
  if (phase != 0 && pack->halo_pairs && use_list) {
    const long long n = pack->halo_seg[phase] - pack->halo_seg[phase - 1];
    if (n <= 0) return PB_OK;
    const uint2 *pairs = reinterpret_cast<const uint2 *>(pack->halo_pairs) + pack->halo_seg[phase - 1];
    const unsigned G = (unsigned)((n + 255) / 256);
    const HaloFields hf = halo_fields(*pack, *cfg, cons);
    cudaStream_t st = (cudaStream_t)stream;
#define PB_HALO_L(M, P)                                                                          \
  do {                                                                                           \
    if (phase == 1) halo_list_kernel<M, P, 1><<<G, 256, 0, st>>>(*pack, *cfg, hf, pairs, n);     \
    else halo_list_kernel<M, P, 2><<<G, 256, 0, st>>>(*pack, *cfg, hf, pairs, n);                \
  } while (0)
    if (cfg->mhd) {
      if (cfg->bc_prim) PB_HALO_L(true, true);
      else PB_HALO_L(true, false);
    } else {
      if (cfg->bc_prim) PB_HALO_L(false, true);
      else PB_HALO_L(false, false);
    }
#undef PB_HALO_L
    return check_launch("halo_list_kernel");
  }
  // thread mapping: PB_HALO_VARIANT=zone|hoist|pair (default zone: the three measured the same, profiles/r02i_variants.txt)
  static const int want = [] {
    const char *e = getenv("PB_HALO_VARIANT");
    return (e && e[0] == 'p') ? 2 : ((e && e[0] == 'h') ? 1 : 0);
  }();
  const bool pairs_ok = cfg->nx[0] % 2 == 0 && cfg->ng % 2 == 0 && L.ncell % 2 == 0;
  const int var = (want == 2 && !pairs_ok) ? 0 : want;
  Lay Lh = L;
  if (var == 2) {
    Lh.ni = L.ni / 2;
    Lh.is = L.is / 2;
    Lh.ie = (L.ie + 1) / 2 - 1;
  }
  const Shell sh(Lh);
  const long long n = (long long)sh.n * pack->nblocks;
  if (n <= 0) return PB_OK;
  const int T = 256;
  const unsigned G = (unsigned)((n + T - 1) / T);
  const HaloFields hf = halo_fields(*pack, *cfg, cons);
  cudaStream_t st = (cudaStream_t)stream;
#define PB_HALO_V(M, P, PH)                                                              \
  do {                                                                                   \
    if (var == 0) halo_local_kernel<M, P, PH, 0><<<G, T, 0, st>>>(*pack, *cfg, hf);      \
    else if (var == 1) halo_local_kernel<M, P, PH, 1><<<G, T, 0, st>>>(*pack, *cfg, hf); \
    else halo_local_kernel<M, P, PH, 2><<<G, T, 0, st>>>(*pack, *cfg, hf);               \
  } while (0)
#define PB_HALO(M, P)                        \
  do {                                       \
    if (phase == 0) PB_HALO_V(M, P, 0);      \
    else if (phase == 1) PB_HALO_V(M, P, 1); \
    else PB_HALO_V(M, P, 2);                 \
  } while (0)
  if (cfg->mhd) {
    if (cfg->bc_prim) PB_HALO(true, true);
    else PB_HALO(true, false);
  } else {
    if (cfg->bc_prim) PB_HALO(false, true);
    else PB_HALO(false, false);
  }
#undef PB_HALO
#undef PB_HALO_V
  return check_launch("halo_local_kernel");
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
