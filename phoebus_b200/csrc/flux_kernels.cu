// fluid::CalculateFluxes for a whole pack: fused reconstruction + Riemann solve, one launch per
// direction (ref: src/fluid/fluid.cpp:659-773, src/fluid/riemann.hpp:256-301).
//
// The reference runs two kernels per MeshBlock ("Reconstruct" writes ql/qr for all directions to
// global memory, "CalculateFluxes" reads them back: 120 words/zone of round trip). Here a CTA owns a
// tile of stencil pencils: every thread reconstructs the 7/10 primitives of ONE cell (exactly one
// reconstruction per cell and direction, as in the reference), hands its left-face state to the
// neighbouring thread through shared memory, keeps its right-face state in registers and then solves
// the Riemann problem at its own lower face. ql/qr never touch HBM.
#include <cuda.h> // CUtensorMap and the enums of cuTensorMapEncodeTiled (types only: the entry point is fetched at run time)

#include <cstdint>
#include <cstring>

#define PB_FLOORS_OUT_OF_LINE 1 // see get_floors (pb_device.cuh)
#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {

// ---- TMA staging of the stencil tile (option flux_tma): the tile of one CTA is a dense 5-D box of the primitive slab
// [block][component][k][j][i] — NP components x (NS + 4) cells along the stencil axis x TT pencils — so ONE
// cp.async.bulk.tensor load issued by one thread replaces ~50 per-thread copies; completion is tracked by an mbarrier.
PB_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
PB_DEV void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PB_DEV void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
PB_DEV void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
PB_DEV void tma_load_5d(void *dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// debugging switches: -DPB_FLUX_RECON_EXACT / -DPB_FLUX_RIEMANN_EXACT build the phase with the
// compiler's own divide / sqrt expansion (same bits, slower)
#ifdef PB_FLUX_RECON_EXACT
typedef ArithExact ArithRecon;
#else
typedef ArithFast ArithRecon;
#endif
#ifdef PB_FLUX_RIEMANN_EXACT
typedef ArithExact ArithRiemann;
#else
typedef ArithFast ArithRiemann;
#endif

struct FluxTile {
  int NS;      // cells along the stencil axis handled per pencil (n + 2)
  int TT;      // pencils per CTA
  int nT1, nT; // transverse extents: fast one, total pencils per block
  int tiles;   // CTAs per block
  int tiles1;  // TMA staging: tiles along the fast transverse axis (a tile is a box: it never wraps to the next row)
  int row;     // length of a row of the staged tile in shared memory (NS + 4 along x1, TT across; TMA boxes are padded so that
  int pad;     // they start on a 16-byte boundary of the slab: `pad` leading cells, measured requirement, profiles/r02_tma_probe.txt)
};

// ---- asynchronous global -> shared staging (LDGSTS): the copies bypass registers, so a CTA has its
// whole stencil tile and its face geometry in flight at once instead of 50 dependent loads per thread
PB_DEV void cp_async8(Real *smem_dst, const Real *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
PB_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PB_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int FLUX_NGEO = 19; // alpha, detgam, beta[3], gcov[10], gamma^{dd}, floor A, floor B, (pad)

// Shared memory of one CTA (doubles):
//   s_in  [NP][(NS+4)*TT]  primitives of the tile's stencil pencils, 2 extra cells at either end
//   s_ql  [NP][(NS+1)*TT]  left-face states handed to the neighbouring thread
//   s_geo [FLUX_NGEO][NS*TT] this thread's face geometry + floor-table entries
// 2 CTAs of <= 288 threads per SM: 96 registers (measured optimum on B200, profiles/r01_sweep2*)
template <int DIR, int RECON, bool MHD, bool TMA>
__global__ void __launch_bounds__(288, 2) flux_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const FluxTile ft,
                                                      const __grid_constant__ CUtensorMap tmap) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  extern __shared__ __align__(128) Real smem[];
  __shared__ uint64_t s_bar;
  const Lay L(cfg);
  const int b = blockIdx.x / ft.tiles, tile = blockIdx.x % ft.tiles;
  const int NS = ft.NS, TT = ft.TT;
  const int tid = threadIdx.x, nthr = NS * TT;
  const int RS = ft.row;
  const int plane_in = (DIR == 1) ? RS * TT : RS * (NS + 4), plane = (NS + 1) * TT;
  Real *s_in = smem, *s_ql = smem + NP * plane_in, *s_geo = s_ql + NP * plane;
  int sI, tI;
  if (DIR == 1) {
    sI = tid % NS;
    tI = tid / NS;
  } else {
    tI = tid % TT;
    sI = tid / TT;
  }
  int t1, t2;
  bool valid;
  if (TMA) { // rectangular tiles: TT pencils of one row
    t1 = (tile % ft.tiles1) * TT + tI;
    t2 = tile / ft.tiles1;
    valid = (t1 < ft.nT1) && (sI < NS) && (tI < TT);
  } else {
    const int col = tile * TT + tI;
    valid = (col < ft.nT) && (sI < NS) && (tI < TT);
    t1 = col % ft.nT1, t2 = col / ft.nT1;
  }
  int i, j, k, st;
  if (DIR == 1) {
    i = L.is - 1 + sI;
    j = L.js - L.dj + t1;
    k = L.ks - L.dk + t2;
    st = 1;
  } else if (DIR == 2) {
    i = L.is - 1 + t1;
    j = L.js - 1 + sI;
    k = L.ks - L.dk + t2;
    st = L.ni;
  } else {
    i = L.is - 1 + t1;
    j = L.js - 1 + t2;
    k = L.ks - 1 + sI;
    st = L.ni * L.nj;
  }
  // entry e of pencil tI in s_in <-> cell (sI = e - 2); my own cell is entry sI + 2
  const int in0 = ((DIR == 1) ? tI * RS + sI : sI * RS + tI) + ft.pad;
  const int in_st = (DIR == 1) ? 1 : RS;
  const int slot_w = (DIR == 1) ? tI * (NS + 1) + sI + 1 : (sI + 1) * TT + tI; // my ql goes to face s+1
  const int slot_r = (DIR == 1) ? tI * (NS + 1) + sI : sI * TT + tI;           // face s reads cell s-1's ql
  const int o = L.idx(k, j, i);
  const pb_block_t &blk = pack.blocks[b];
  if (TMA) {
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
      const int r0 = (tile % ft.tiles1) * TT, r1 = tile / ft.tiles1;
      int c0, c1, c2;
      if (DIR == 1) c0 = L.is - 3 - ft.pad, c1 = L.js - L.dj + r0, c2 = L.ks - L.dk + r1;
      else if (DIR == 2) c0 = L.is - 1 + r0 - ft.pad, c1 = L.js - 3, c2 = L.ks - L.dk + r1;
      else c0 = L.is - 1 + r0 - ft.pad, c1 = L.js - 1 + r1, c2 = L.ks - 3;
      mbar_expect_tx(&s_bar, (unsigned)(NP * plane_in * sizeof(Real)));
      tma_load_5d(s_in, &tmap, &s_bar, c0, c1, c2, 0, b);
    }
  } else if (valid) {
    const Real *v = pack.prim + (size_t)b * NP * L.ncell + o;
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      const Real *p = v + (size_t)n * L.ncell;
      Real *d = s_in + n * plane_in + in0;
      cp_async8(d + 2 * in_st, p);
      // only the cells the scheme reads: none beyond the tile for constant, one for linear, two for weno5 / mp5
      // (with nghost = 2 and linear reconstruction the second one would lie outside the block)
      if (sI < 2 && (RECON >= 2 || (RECON == 1 && sI == 1))) cp_async8(d, p - 2 * st);
      if (sI >= NS - 2 && (RECON >= 2 || (RECON == 1 && sI == NS - 2))) cp_async8(d + 4 * in_st, p + 2 * st);
    }
  }
  cp_async_commit();
  if (valid && sI > 0) {
    // FaceGeom (ref: riemann.hpp:32-54): only gamma^{dd} of the inverse metric is needed
    const GeomTab gt(cfg, pack.geom, blk.geom_slot);
    const int gcell = gt.cell(k, j, i);
    const Real *gp = gt.t + (size_t)(DIR * PB_NGLOC) * gt.gcell + gcell;
    Real *d = s_geo + tid;
    cp_async8(d, gp + (size_t)G_ALPHA * gt.gcell);
    cp_async8(d + nthr, gp + (size_t)G_DETGAM * gt.gcell);
#pragma unroll
    for (int m = 0; m < 3; ++m) cp_async8(d + (2 + m) * nthr, gp + (size_t)(G_BCON + m) * gt.gcell);
#pragma unroll
    for (int m = 0; m < 10; ++m) cp_async8(d + (5 + m) * nthr, gp + (size_t)(G_GCOV + m) * gt.gcell);
    cp_async8(d + 15 * nthr, gp + (size_t)(G_GAMCON + flat4(DIR - 1, DIR - 1)) * gt.gcell);
    const Real *fp = pack.floor_tab + ((size_t)blk.floor_slot * 4 + (DIR == 1 ? 2 : 0)) * L.ni + i;
    cp_async8(d + 16 * nthr, fp);
    cp_async8(d + 17 * nthr, fp + L.ni);
  }
  cp_async_commit();
  if (TMA) {
    mbar_wait(&s_bar, 0); // the box has landed (complete_tx)
  } else {
    cp_async_wait<1>(); // the stencil tile has landed (the geometry may still be in flight)
    __syncthreads();
  }

  Real qR[NP];
  if (valid) {
    const Real *c = s_in + in0;
    ArithRecon af;
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      const Real *q = c + n * plane_in;
      Real ql;
      recon5<RECON>(af, q[0], q[in_st], q[2 * in_st], q[3 * in_st], q[4 * in_st], ql, qR[n]);
      s_ql[n * plane + slot_w] = ql;
    }
    if (af.failed()) { // an operand outside the fast divide's exponent range: redo with the compiler's a/b
      ArithExact ax;
#pragma unroll 1
      for (int n = 0; n < NP; ++n) {
        const Real *q = c + n * plane_in;
        Real ql, qr;
        recon5<RECON>(ax, q[0], q[in_st], q[2 * in_st], q[3 * in_st], q[4 * in_st], ql, qr);
        s_ql[n * plane + slot_w] = ql;
        qR[n] = qr;
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  if (!valid || sI == 0) return;
  Real qL[NP];
#pragma unroll
  for (int n = 0; n < NP; ++n) qL[n] = s_ql[n * plane + slot_r];

  PointGeom g;
  const Real *sg = s_geo + tid;
  g.alpha = sg[0];
  g.detgam = sg[nthr];
#pragma unroll
  for (int d = 0; d < 3; ++d) g.beta[d] = sg[(2 + d) * nthr];
#pragma unroll
  for (int m = 0; m < 10; ++m) g.gcov[m] = sg[(5 + m) * nthr];
  const Real gdd = sg[15 * nthr];
  const Real gdet = g.alpha * g.detgam; // DetG = alpha * sqrt(gamma), cached_system.hpp:371-374
  FloorCtx fc;
  fc.A = sg[16 * nthr];
  fc.B = sg[17 * nthr];
  fc.x1 = (DIR == 1) ? Xf(blk, 0, i) : Xc(blk, 0, i);
  fc.x2 = (DIR == 2) ? Xf(blk, 1, j) : Xc(blk, 1, j);
  fc.x3 = (DIR == 3) ? Xf(blk, 2, k) : Xc(blk, 2, k);

  Real Ul[NCn], Fl[NCn], Ur[NCn], Fr[NCn], vml, vpl, vmr, vpr;
  {
    ArithRiemann af;
    prim_to_flux<DIR, MHD>(af, cfg, fc, g, gdd, qL, vml, vpl, Ul, Fl);
    prim_to_flux<DIR, MHD>(af, cfg, fc, g, gdd, qR, vmr, vpr, Ur, Fr);
    if (af.failed()) {
      ArithExact ax;
      prim_to_flux<DIR, MHD>(ax, cfg, fc, g, gdd, qL, vml, vpl, Ul, Fl);
      prim_to_flux<DIR, MHD>(ax, cfg, fc, g, gdd, qR, vmr, vpr, Ur, Fr);
    }
  }
  Real *fl = pack.flux + ((size_t)b * L.ndim + (DIR - 1)) * NCn * L.ncell + o;
  Real speed;
  if (cfg.riemann == 0) { // llf, ref: riemann.hpp:256-276
    const Real cmax = dmax(dmax(-vml, vpl), dmax(-vmr, vpr));
#pragma unroll
    for (int m = 0; m < NCn; m++)
      fl[(size_t)m * L.ncell] = 0.5 * ((Fl[m] + Fr[m]) * gdet - cmax * ((Ur[m] - Ul[m]) * g.detgam));
    speed = cmax;
  } else { // hll, ref: riemann.hpp:278-301
    const Real cl = dmin(dmin(vml, vmr), 0.0);
    const Real cr = dmax(dmax(vpl, vpr), 0.0);
    const Real crcl = cr * cl;
    const Real inv_cr_cl = ratio(1.0, cr - cl);
#pragma unroll
    for (int m = 0; m < NCn; m++)
      fl[(size_t)m * L.ncell] = ((cr * Fl[m] - cl * Fr[m]) * gdet + crcl * (Ur[m] - Ul[m]) * g.detgam) * inv_cr_cl;
    speed = dmax(-cl, cr);
  }
  pack.fsig[((size_t)b * L.ndim + (DIR - 1)) * L.ncell + o] = speed;
}

// cuTensorMapEncodeTiled over the primitive slab as a 5-D tensor (i, j, k, component, block); the driver entry point is
// resolved through the runtime (no link-time dependency on libcuda). Returns false if TMA cannot be used.
static bool encode_prim_map(const pb_pack_t &pack, const pb_cfg_t &cfg, int nprim, const cuuint32_t box[5], CUtensorMap *out) {
  typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode enc = [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<Encode>(fn);
  }();
  if (!enc) return false;
  const Lay L(cfg);
  const cuuint64_t dims[5] = {(cuuint64_t)L.ni, (cuuint64_t)L.nj, (cuuint64_t)L.nk, (cuuint64_t)nprim, (cuuint64_t)pack.nblocks};
  const cuuint64_t strides[4] = {(cuuint64_t)L.ni * 8, (cuuint64_t)L.ni * L.nj * 8, (cuuint64_t)L.ncell * 8, (cuuint64_t)nprim * L.ncell * 8};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(pack.prim) & 15) != 0) return false;
  for (int d = 0; d < 5; ++d)
    if (box[d] > 256 || box[d] > dims[d]) return false;
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, pack.prim, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int DIR, int RECON, bool MHD>
static int launch_dir(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s) {
  const Lay L(cfg);
  FluxTile ft;
  int n1, n2;
  if (DIR == 1) {
    ft.NS = cfg.nx[0] + 2;
    n1 = L.je - L.js + 1 + 2 * L.dj;
    n2 = L.ke - L.ks + 1 + 2 * L.dk;
  } else if (DIR == 2) {
    ft.NS = cfg.nx[1] + 2;
    n1 = cfg.nx[0] + 2;
    n2 = L.ke - L.ks + 1 + 2 * L.dk;
  } else {
    ft.NS = cfg.nx[2] + 2;
    n1 = cfg.nx[0] + 2;
    n2 = cfg.nx[1] + 2;
  }
  ft.nT1 = n1;
  ft.nT = n1 * n2;
  int target = flux_threads_target();
  ft.TT = target / ft.NS;
  if (ft.TT < 1) ft.TT = 1;
  if (ft.TT > ft.nT) ft.TT = ft.nT;
  while (ft.NS * ft.TT > 288 && ft.TT > 1) ft.TT--;
  if (ft.NS * ft.TT > 288) return set_error(PB_EINVAL, "block too large for the flux tile");
  // TMA staging needs a rectangular tile: TT pencils of ONE row, and a box that starts on a 16-byte boundary of the slab
  // (even cell index) with an even number of cells per row
  bool tma = flux_tma_enabled() && L.ni % 2 == 0;
  int TTt = ft.TT > n1 ? n1 : ft.TT;
  int row, pad;
  if (DIR == 1) {
    const int start = L.is - 3;
    pad = ((start % 2) + 2) % 2;
    row = (ft.NS + 4 + pad + 1) & ~1;
  } else {
    TTt &= ~1;
    tma = tma && TTt >= 2;
    pad = (L.is - 1) & 1;
    row = TTt + 2 * pad;
  }
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  if (tma) {
    const cuuint32_t S = (cuuint32_t)(ft.NS + 4), R = (cuuint32_t)row;
    const cuuint32_t box[5] = {R, DIR == 1 ? (cuuint32_t)TTt : (DIR == 2 ? S : 1u), DIR == 3 ? S : 1u, (cuuint32_t)NC<MHD>::nprim, 1u};
    tma = encode_prim_map(pack, cfg, NC<MHD>::nprim, box, &tmap);
  }
  if (tma) {
    ft.TT = TTt;
    ft.row = row;
    ft.pad = pad;
  } else {
    ft.row = DIR == 1 ? ft.NS + 4 : ft.TT;
    ft.pad = 0;
  }
  ft.tiles1 = tma ? (n1 + ft.TT - 1) / ft.TT : 0;
  ft.tiles = tma ? ft.tiles1 * n2 : (ft.nT + ft.TT - 1) / ft.TT;
  const int threads = ft.NS * ft.TT;
  const size_t plane_in = DIR == 1 ? (size_t)ft.row * ft.TT : (size_t)ft.row * (ft.NS + 4);
  const size_t smem = ((size_t)NC<MHD>::nprim * ((size_t)(ft.NS + 1) * ft.TT + plane_in) + (size_t)FLUX_NGEO * threads) * sizeof(Real);
  if (tma) {
    auto kern = flux_kernel<DIR, RECON, MHD, true>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_cuda_error(e, "flux smem attribute");
    }
    kern<<<(unsigned)(pack.nblocks * ft.tiles), threads, smem, s>>>(pack, cfg, ft, tmap);
  } else {
    auto kern = flux_kernel<DIR, RECON, MHD, false>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_cuda_error(e, "flux smem attribute");
    }
    kern<<<(unsigned)(pack.nblocks * ft.tiles), threads, smem, s>>>(pack, cfg, ft, tmap);
  }
  return check_launch("flux_kernel");
}

template <int RECON, bool MHD> static int launch_all(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s) {
  int rc = launch_dir<1, RECON, MHD>(pack, cfg, s);
  if (rc == 0 && cfg.ndim >= 2) rc = launch_dir<2, RECON, MHD>(pack, cfg, s);
  if (rc == 0 && cfg.ndim >= 3) rc = launch_dir<3, RECON, MHD>(pack, cfg, s);
  return rc;
}

template <bool MHD> static int launch_recon(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s) {
  switch (cfg.recon) {
  case 0: return launch_all<0, MHD>(pack, cfg, s);
  case 1: return launch_all<1, MHD>(pack, cfg, s);
  case 2: return launch_all<2, MHD>(pack, cfg, s);
  case 3: return launch_all<3, MHD>(pack, cfg, s);
  }
  return set_error(PB_EINVAL, "unknown reconstruction");
}

} // namespace pb

namespace pb {
int calculate_fluxes_split(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s); // flux_split_kernels.cu
unsigned long long redo_total_host();
}

extern "C" unsigned long long pb_arith_redo_count(void) { return pb::redo_total_host(); }

extern "C" int pb_calculate_fluxes(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream) {
  using namespace pb;
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (cfg->ng < 3 && cfg->recon >= 2) return set_error(PB_EINVAL, "weno5/mp5 need nghost >= 3");
  if (cfg->riemann != 0 && cfg->riemann != 1) return set_error(PB_EINVAL, "unknown riemann solver");
  cudaStream_t s = (cudaStream_t)stream;
  if (pack->recon_lr && flux_split_enabled()) return calculate_fluxes_split(*pack, *cfg, s);
  return cfg->mhd ? launch_recon<true>(*pack, *cfg, s) : launch_recon<false>(*pack, *cfg, s);
}
