// fluid::CalculateFluxes as two launches per direction (ref: src/fluid/fluid.cpp:659-773):
//   recon_kernel   one cell per thread: reconstructs the 7/10 primitives along DIR and writes the
//                  left/right face states into the pack's ql/qr scratch (the reference's ql / qr
//                  fields, fluid.cpp:288-299, one direction at a time);
//   riemann_kernel one face per thread: prim_to_flux on both states + LLF/HLL (riemann.hpp:87-301).
// Why two launches when flux_kernels.cu fuses them: the path is FP64-latency bound, not HBM bound.
// Reconstruction is 75 % of the FP64 work and needs < 64 registers, the Riemann solve needs > 110;
// fused, the whole kernel runs at the Riemann solve's occupancy (17 warps/SM, FP64 pipe 50 % busy).
// Split, reconstruction runs at 2-3x the occupancy; the ql/qr round trip (40 words per zone and
// direction) hides under the arithmetic (measured: profiles/r01_flux_split*.md).
#define PB_FLOORS_OUT_OF_LINE 1 // see get_floors (pb_device.cuh)
#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {

#ifdef PB_AF_DIAG
__device__ unsigned g_diag = 0;
#endif

// one int per MeshBlock, library-owned: "this direction's face states of the block carry a non-zero B"
static int *g_bflag = nullptr;
static int g_bflag_cap = 0;
static int *block_flags(int nblocks) {
  if (nblocks > g_bflag_cap) {
    cudaDeviceSynchronize();
    if (g_bflag) cudaFree(g_bflag);
    g_bflag = nullptr;
    if (cudaMalloc(&g_bflag, sizeof(int) * (size_t)nblocks) != cudaSuccess) return nullptr;
    g_bflag_cap = nblocks;
  }
  return g_bflag;
}

struct Box { // inclusive cell ranges of one block
  int i0, i1, j0, j1, k0, k1;
  __host__ __device__ int ni() const { return i1 - i0 + 1; }
  __host__ __device__ int nj() const { return j1 - j0 + 1; }
  __host__ __device__ long long n() const { return (long long)ni() * nj() * (k1 - k0 + 1); }
};
PB_DEV bool cell_of_thread(const Box &r, int nblocks, int &b, int &k, int &j, int &i) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nreg = r.n();
  if (t >= nreg * nblocks) return false;
  b = (int)(t / nreg);
  int rem = (int)(t - (long long)b * nreg);
  const int q = rem / r.ni();
  i = r.i0 + (rem - q * r.ni());
  const int q2 = q / r.nj();
  j = r.j0 + (q - q2 * r.nj());
  k = r.k0 + q2;
  return true;
}

#ifndef PB_RECON_MINB
#define PB_RECON_MINB 4
#endif
// riemann_kernel: 128 threads x 3 CTAs/SM = 168 registers, no spills: measured 10.3 ms per cycle vs 11.2 (256 x 2, 128 regs,
// 100 B spills), 10.7 (128 x 4), 15.4 (256 x 3, 80 regs)
#ifndef PB_RIEMANN_MINB
#define PB_RIEMANN_MINB 3
#endif
#ifndef PB_RIEMANN_T
#define PB_RIEMANN_T 128
#endif
#ifndef PB_RECON_T
#define PB_RECON_T 256
#endif

// the cells the scheme reads and no others: with nghost = 2 and linear reconstruction (the reference's default deck
// values) the cells at distance 2 of the first / last reconstructed cell lie outside the block
template <int RECON> PB_DEV void load_stencil(const Real *p, int st, Real q[5]) {
  q[2] = __ldg(p);
  q[1] = RECON >= 1 ? __ldg(p - st) : q[2];
  q[3] = RECON >= 1 ? __ldg(p + st) : q[2];
  q[0] = RECON >= 2 ? __ldg(p - 2 * st) : q[2];
  q[4] = RECON >= 2 ? __ldg(p + 2 * st) : q[2];
}

// bflag (may be NULL; MHD packs only): bflag[b] is set when the block gets a face state with a non-zero B component, i.e. when
// the Riemann solve of this direction cannot take the unmagnetised variant for the whole block
template <int DIR, int RECON, int NP>
__global__ void __launch_bounds__(PB_RECON_T, PB_RECON_MINB) recon_kernel(const Real *__restrict__ prim, Real *__restrict__ ql,
                                                                   Real *__restrict__ qr, const pb_cfg_t cfg, const Box reg,
                                                                   const int nblocks, int *__restrict__ bflag) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!cell_of_thread(reg, nblocks, b, k, j, i)) return;
  const int st = DIR == 1 ? 1 : (DIR == 2 ? L.ni : L.ni * L.nj);
  const size_t o = (size_t)b * NP * L.ncell + L.idx(k, j, i);
  const Real *v = prim + o;
  Real *l = ql + o + st; // my left state belongs to the face of cell s+1
  Real *r = qr + o;
  ArithFast af;
  long long anyB = 0;
#pragma unroll
  for (int n = 0; n < NP; ++n) {
    const Real *p = v + (size_t)n * L.ncell;
    Real a, c;
    Real q[5];
    load_stencil<RECON>(p, st, q);
    recon5z<RECON>(af, q[0], q[1], q[2], q[3], q[4], a, c);
    l[(size_t)n * L.ncell] = a;
    r[(size_t)n * L.ncell] = c;
    if (NP == 10 && n >= PB_ && n < PB_ + 3) anyB |= __double_as_longlong(a) | __double_as_longlong(c);
  }
  if (NP == 10 && bflag && (anyB << 1) != 0) bflag[b] = 1; // (racing stores of the same value)
  if (af.failed()) { // operands outside the fast divide's range: redo with the compiler's a/b
    ArithExact ax;
#pragma unroll 1
    for (int n = 0; n < NP; ++n) {
      const Real *p = v + (size_t)n * L.ncell;
      Real a, c;
      Real q[5];
      load_stencil<RECON>(p, st, q);
      recon5z<RECON>(ax, q[0], q[1], q[2], q[3], q[4], a, c);
      l[(size_t)n * L.ncell] = a;
      r[(size_t)n * L.ncell] = c;
      if (NP == 10 && bflag && n >= PB_ && n < PB_ + 3 && ((__double_as_longlong(a) | __double_as_longlong(c)) << 1) != 0) bflag[b] = 1;
    }
  }
}

// Two cells per thread (option recon_cpt = 2): the pair is taken ALONG the stencil for DIR >= 2 — six loads per primitive
// instead of ten, rows along i stay coalesced — and along x2 for DIR == 1 (no shared loads there, but two independent
// reconstructions per thread to hide the FP64 and load latencies). Same arithmetic per cell, same bits.
#ifndef PB_RECON2_T
#define PB_RECON2_T 128
#endif
#ifndef PB_RECON2_MINB
#define PB_RECON2_MINB 4
#endif
template <int DIR, int RECON, int NP>
__global__ void __launch_bounds__(PB_RECON2_T, PB_RECON2_MINB) recon2_kernel(const Real *__restrict__ prim, Real *__restrict__ ql,
                                                                      Real *__restrict__ qr, const pb_cfg_t cfg, const Box reg,
                                                                      const int nblocks, int *__restrict__ bflag) {
  constexpr int PD = DIR == 1 ? 2 : DIR; // the paired dimension
  const Lay L(cfg);
  Box half = reg; // one thread per PAIR along PD
  const int npd = (PD == 2 ? reg.j1 - reg.j0 : reg.k1 - reg.k0) + 1;
  if (PD == 2) half.j1 = half.j0 + (npd + 1) / 2 - 1;
  else half.k1 = half.k0 + (npd + 1) / 2 - 1;
  int b, k, j, i;
  if (!cell_of_thread(half, nblocks, b, k, j, i)) return;
  if (PD == 2) j = reg.j0 + 2 * (j - reg.j0);
  else k = reg.k0 + 2 * (k - reg.k0);
  const bool second = PD == 2 ? (j + 1 <= reg.j1) : (k + 1 <= reg.k1);
  const int st = DIR == 1 ? 1 : (DIR == 2 ? L.ni : L.ni * L.nj);
  const int sp = PD == 2 ? L.ni : L.ni * L.nj; // stride between the two cells of the pair
  const size_t o = (size_t)b * NP * L.ncell + L.idx(k, j, i);
  const Real *v = prim + o;
  Real *l = ql + o + st;
  Real *r = qr + o;
  ArithFast af;
  long long anyB = 0;
#pragma unroll
  for (int n = 0; n < NP; ++n) {
    const Real *p = v + (size_t)n * L.ncell;
    Real qa[5], qb[5];
    if (DIR == 1) {
      load_stencil<RECON>(p, st, qa);
      if (second) load_stencil<RECON>(p + sp, st, qb);
      else
        for (int e = 0; e < 5; ++e) qb[e] = qa[e];
    } else { // sp == st: cell B's stencil is cell A's shifted by one
      load_stencil<RECON>(p, st, qa);
      qb[0] = qa[1], qb[1] = qa[2], qb[2] = qa[3], qb[3] = qa[4];
      qb[4] = (RECON >= 2 && second) ? __ldg(p + 3 * st) : qa[4];
      if (RECON < 2) { // load_stencil filled the unused slots with the centre value: rebuild B's own
        qb[2] = second ? __ldg(p + st) : qa[2];
        qb[1] = RECON >= 1 ? qa[2] : qb[2];
        qb[3] = (RECON >= 1 && second) ? __ldg(p + 2 * st) : qb[2];
        qb[0] = qb[4] = qb[2];
      }
    }
    Real a0, c0, a1, c1;
    recon5z<RECON>(af, qa[0], qa[1], qa[2], qa[3], qa[4], a0, c0);
    recon5z<RECON>(af, qb[0], qb[1], qb[2], qb[3], qb[4], a1, c1);
    l[(size_t)n * L.ncell] = a0;
    r[(size_t)n * L.ncell] = c0;
    if (second) {
      l[(size_t)n * L.ncell + sp] = a1;
      r[(size_t)n * L.ncell + sp] = c1;
    }
    if (NP == 10 && n >= PB_ && n < PB_ + 3)
      anyB |= __double_as_longlong(a0) | __double_as_longlong(c0) | __double_as_longlong(a1) | __double_as_longlong(c1);
  }
  if (NP == 10 && bflag && (anyB << 1) != 0) bflag[b] = 1;
  if (af.failed()) { // operands outside the fast divide's range: redo both cells with the compiler's a/b
    ArithExact ax;
    for (int c = 0; c < (second ? 2 : 1); ++c) {
#pragma unroll 1
      for (int n = 0; n < NP; ++n) {
        const Real *p = v + (size_t)n * L.ncell + (size_t)c * sp;
        Real a, cc;
        Real q[5];
        load_stencil<RECON>(p, st, q);
        recon5z<RECON>(ax, q[0], q[1], q[2], q[3], q[4], a, cc);
        l[(size_t)n * L.ncell + (size_t)c * sp] = a;
        r[(size_t)n * L.ncell + (size_t)c * sp] = cc;
        if (NP == 10 && bflag && n >= PB_ && n < PB_ + 3 && ((__double_as_longlong(a) | __double_as_longlong(cc)) << 1) != 0) bflag[b] = 1;
      }
    }
  }
}

// One face: load both states and the face geometry, prim_to_flux on each, LLF / HLL, store.
// Returns the arithmetic policy's failure flag (always false for ArithExact).
#ifndef PB_RIEMANN_BZ
#define PB_RIEMANN_BZ 1
#endif
// BMODE: 0 = decide per warp from the loaded states, 1 = the caller knows that B == 0 (whole block), 2 = general path only
template <int DIR, bool MHD, class AR, int BMODE = (PB_RIEMANN_BZ != 0 ? 0 : 2)>
PB_DEV bool riemann_face(const pb_pack_t &pack, const pb_cfg_t &cfg, const Lay &L, const Real *__restrict__ ql,
                         const Real *__restrict__ qr, int b, int k, int j, int i) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  const int o = L.idx(k, j, i);
  const pb_block_t &blk = pack.blocks[b];
  Real qL[NP], qR[NP];
  {
    const size_t q0 = (size_t)b * NP * L.ncell + o;
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      qL[n] = __ldg(ql + q0 + (size_t)n * L.ncell);
      qR[n] = __ldg(qr + q0 + (size_t)n * L.ncell);
    }
  }
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  PointGeom g;
  const int gcell = gt.cell(k, j, i);
  // FaceGeom (ref: riemann.hpp:32-54): only gamma^{dd} of the inverse metric is needed
  g.alpha = gt.at(DIR, G_ALPHA, gcell);
  g.detgam = gt.at(DIR, G_DETGAM, gcell);
#pragma unroll
  for (int d = 0; d < 3; ++d) g.beta[d] = gt.at(DIR, G_BCON + d, gcell);
#pragma unroll
  for (int m = 0; m < 10; ++m) g.gcov[m] = gt.at(DIR, G_GCOV + m, gcell);
  const Real gdd = gt.at(DIR, G_GAMCON + flat4(DIR - 1, DIR - 1), gcell);
  const Real gdet = g.alpha * g.detgam; // DetG = alpha * sqrt(gamma), cached_system.hpp:371-374
  FloorCtx fc;
  load_floor(pack.floor_tab, blk.floor_slot, L.ni, DIR == 1 ? 1 : 0, i, fc.A, fc.B);
  fc.x1 = (DIR == 1) ? Xf(blk, 0, i) : Xc(blk, 0, i);
  fc.x2 = (DIR == 2) ? Xf(blk, 1, j) : Xc(blk, 1, j);
  fc.x3 = (DIR == 3) ? Xf(blk, 2, k) : Xc(blk, 2, k);

  Real Ul[NCn], Fl[NCn], Ur[NCn], Fr[NCn], vml, vpl, vmr, vpr;
  AR ar;
  // unmagnetised faces (both states have B == +-0 in all components) take the exact hydrodynamic variant of
  // prim_to_flux; the decision is made per warp so that the two code paths never run side by side
  bool bzero = MHD && BMODE == 1;
  if (MHD && BMODE == 0) {
    long long any = 0;
#pragma unroll
    for (int m = 0; m < 3; ++m) any |= __double_as_longlong(qL[PB_ + m]) | __double_as_longlong(qR[PB_ + m]);
    bzero = __all_sync(__activemask(), (any << 1) == 0);
  }
  if (MHD && BMODE != 2 && bzero) {
    prim_to_flux<DIR, MHD, true>(ar, cfg, fc, g, gdd, qL, vml, vpl, Ul, Fl);
    prim_to_flux<DIR, MHD, true>(ar, cfg, fc, g, gdd, qR, vmr, vpr, Ur, Fr);
  } else if (BMODE != 1 || !MHD) {
    prim_to_flux<DIR, MHD, false>(ar, cfg, fc, g, gdd, qL, vml, vpl, Ul, Fl);
    prim_to_flux<DIR, MHD, false>(ar, cfg, fc, g, gdd, qR, vmr, vpr, Ur, Fr);
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
  return ar.failed();
}

// Faces whose operands fall outside the fast divide's exponent range are not redone in place (keeping
// both code paths in one kernel costs 300 B of spills and doubles its size): their linear index goes
// to a device queue and riemann_redo_kernel recomputes exactly those faces with the compiler's a/b,
// 1/x, sqrt right after, on the same stream, overwriting what the fast pass stored (RedoQueue: pb_host.hpp).
PB_DEV void face_of_index(const Box &r, long long t, int &b, int &k, int &j, int &i) {
  const long long nreg = r.n();
  b = (int)(t / nreg);
  int rem = (int)(t - (long long)b * nreg);
  const int q = rem / r.ni();
  i = r.i0 + (rem - q * r.ni());
  const int q2 = q / r.nj();
  j = r.j0 + (q - q2 * r.nj());
  k = r.k0 + q2;
}

// bflag == NULL: every face, unmagnetised variant chosen per warp. bflag != NULL: the kernel is launched twice over the same
// faces — WANT = 0 takes the MeshBlocks without any non-zero B face state (flag written by recon_kernel) through the
// unmagnetised variant only, compiled for a higher occupancy; WANT = 1 takes the others through the per-warp decision.
#ifndef PB_RIEMANN_BZ_MINB
#define PB_RIEMANN_BZ_MINB 5
#endif
template <int DIR, bool MHD, int WANT>
__global__ void __launch_bounds__(PB_RIEMANN_T, (WANT == 0 ? PB_RIEMANN_BZ_MINB : PB_RIEMANN_MINB))
    riemann_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Box reg, const Real *__restrict__ ql, const Real *__restrict__ qr,
                   const RedoQueue rq, const int *__restrict__ bflag) {
  const Lay L(cfg);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= reg.n() * pack.nblocks) return;
  int b, k, j, i;
  face_of_index(reg, t, b, k, j, i);
  bool bad;
  if (WANT == 0) {
    if (__ldg(bflag + b) != 0) return;
    bad = riemann_face<DIR, MHD, ArithFast, 1>(pack, cfg, L, ql, qr, b, k, j, i);
  } else {
    if (bflag && __ldg(bflag + b) == 0) return;
    bad = riemann_face<DIR, MHD, ArithFast>(pack, cfg, L, ql, qr, b, k, j, i);
  }
  if (bad) rq.list[atomicAdd(rq.count, 1u)] = (unsigned)t;
}

template <int DIR, bool MHD>
__global__ void __launch_bounds__(128) riemann_redo_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Box reg,
                                                           const Real *__restrict__ ql, const Real *__restrict__ qr,
                                                           const RedoQueue rq) {
  const Lay L(cfg);
  const unsigned n = *rq.count;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(rq.total, (unsigned long long)n);
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int b, k, j, i;
    face_of_index(reg, (long long)rq.list[e], b, k, j, i);
    riemann_face<DIR, MHD, ArithExact, 2>(pack, cfg, L, ql, qr, b, k, j, i);
  }
}

template <int DIR, int RECON, bool MHD>
static int launch_dir_split(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s) {
  constexpr int NP = NC<MHD>::nprim;
  const Lay L(cfg);
  // cells reconstructed: one layer beyond the interior along DIR; transverse extent as the
  // reference's flux loops (faces in the first ghost layer feed FluxCT), ref: fluid.cpp:700-720
  Box rc;
  if (DIR == 1) rc = Box{L.is - 1, L.ie + 1, L.js - L.dj, L.je + L.dj, L.ks - L.dk, L.ke + L.dk};
  else if (DIR == 2) rc = Box{L.is - 1, L.ie + 1, L.js - 1, L.je + 1, L.ks - L.dk, L.ke + L.dk};
  else rc = Box{L.is - 1, L.ie + 1, L.js - 1, L.je + 1, L.ks - 1, L.ke + 1};
  Box rf = rc; // faces: drop the first cell along DIR
  if (DIR == 1) rf.i0 += 1;
  else if (DIR == 2) rf.j0 += 1;
  else rf.k0 += 1;
  Real *ql = pack.recon_lr, *qr = pack.recon_lr + (size_t)pack.nblocks * NP * L.ncell;
  const long long n1 = rc.n() * pack.nblocks, n2 = rf.n() * pack.nblocks;
  int *bflag = nullptr;
  if (MHD && riemann_block_split_enabled()) {
    if (!(bflag = block_flags(pack.nblocks))) return set_cuda_error(cudaGetLastError(), "block flags");
    cudaMemsetAsync(bflag, 0, sizeof(int) * pack.nblocks, s);
  }
  prof_begin(PB_TASK_RECON, s);
  if (recon_cells_per_thread() == 2 && L.ndim >= 2) {
    const int npd = (DIR == 3 ? rc.k1 - rc.k0 : rc.j1 - rc.j0) + 1;
    const long long nthr = n1 / npd * ((npd + 1) / 2);
    recon2_kernel<DIR, RECON, NP><<<(unsigned)((nthr + PB_RECON2_T - 1) / PB_RECON2_T), PB_RECON2_T, 0, s>>>(pack.prim, ql, qr, cfg, rc, pack.nblocks, bflag);
  } else {
    recon_kernel<DIR, RECON, NP><<<(unsigned)((n1 + PB_RECON_T - 1) / PB_RECON_T), PB_RECON_T, 0, s>>>(pack.prim, ql, qr, cfg, rc, pack.nblocks, bflag);
  }
  prof_end(PB_TASK_RECON, s);
  int rcode = check_launch("recon_kernel");
  if (rcode) return rcode;
  if (n2 >= (1ll << 32)) return set_error(PB_EINVAL, "pack too large for the 32-bit redo queue");
  RedoQueue rq;
  if ((rcode = redo_queue((size_t)n2, rq, s))) return rcode;
  prof_begin(PB_TASK_RIEMANN, s);
  cudaMemsetAsync(rq.count, 0, sizeof(unsigned), s);
  const unsigned gr = (unsigned)((n2 + PB_RIEMANN_T - 1) / PB_RIEMANN_T);
  if (bflag) riemann_kernel<DIR, MHD, 0><<<gr, PB_RIEMANN_T, 0, s>>>(pack, cfg, rf, ql, qr, rq, bflag);
  riemann_kernel<DIR, MHD, 1><<<gr, PB_RIEMANN_T, 0, s>>>(pack, cfg, rf, ql, qr, rq, bflag);
  riemann_redo_kernel<DIR, MHD><<<sm_count() * 4, 128, 0, s>>>(pack, cfg, rf, ql, qr, rq);
  prof_end(PB_TASK_RIEMANN, s);
  return check_launch("riemann_kernel");
}

template <int RECON, bool MHD> static int launch_all_split(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s) {
  int rc = launch_dir_split<1, RECON, MHD>(pack, cfg, s);
  if (rc == 0 && cfg.ndim >= 2) rc = launch_dir_split<2, RECON, MHD>(pack, cfg, s);
  if (rc == 0 && cfg.ndim >= 3) rc = launch_dir_split<3, RECON, MHD>(pack, cfg, s);
  return rc;
}
template <bool MHD> static int launch_recon_split(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s) {
  switch (cfg.recon) {
  case 0: return launch_all_split<0, MHD>(pack, cfg, s);
  case 1: return launch_all_split<1, MHD>(pack, cfg, s);
  case 2: return launch_all_split<2, MHD>(pack, cfg, s);
  case 3: return launch_all_split<3, MHD>(pack, cfg, s);
  }
  return set_error(PB_EINVAL, "unknown reconstruction");
}

int presize_block_flags(int nblocks) { return block_flags(nblocks) ? PB_OK : set_cuda_error(cudaGetLastError(), "block flags"); }

int calculate_fluxes_split(const pb_pack_t &pack, const pb_cfg_t &cfg, cudaStream_t s) {
  return cfg.mhd ? launch_recon_split<true>(pack, cfg, s) : launch_recon_split<false>(pack, cfg, s);
}

} // namespace pb
