// One-zone-per-thread tasks over a whole pack: ConservedToPrimitive (robust), its failure fixup,
// ApplyFloors, PrimitiveToConserved, physical boundary conditions, the CFL estimate.
#include <cstdlib>

#include "pb_device.cuh"
#include "pb_host.hpp"

namespace pb {

// ------------------------------------------------------------------ helpers
struct Region {
  int i0, i1, j0, j1, k0, k1;
  __host__ __device__ int ni() const { return i1 - i0 + 1; }
  __host__ __device__ int nj() const { return j1 - j0 + 1; }
  __host__ __device__ int nk() const { return k1 - k0 + 1; }
  __host__ __device__ long long n() const { return (long long)ni() * nj() * nk(); }
};

PB_DEV bool zone_of_thread(const Region &r, int nblocks, int &b, int &k, int &j, int &i) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nreg = r.n();
  if (t >= nreg * nblocks) return false;
  b = (int)(t / nreg);
  int rem = (int)(t % nreg);
  i = r.i0 + rem % r.ni();
  rem /= r.ni();
  j = r.j0 + rem % r.nj();
  k = r.k0 + rem / r.nj();
  return true;
}

PB_DEV FloorCtx cell_floor_ctx(const pb_pack_t &pack, const pb_block_t &blk, const Lay &L, int k, int j, int i) {
  FloorCtx fc;
  load_floor(pack.floor_tab, blk.floor_slot, L.ni, 0, i, fc.A, fc.B);
  fc.x1 = Xc(blk, 0, i);
  fc.x2 = Xc(blk, 1, j);
  fc.x3 = Xc(blk, 2, k);
  return fc;
}

template <int N> PB_DEV void load_comps(const Real *base, int ncell, Real *out) {
#pragma unroll
  for (int n = 0; n < N; ++n) out[n] = base[(size_t)n * ncell];
}
template <int N> PB_DEV void store_comps(Real *base, int ncell, const Real *in) {
#pragma unroll
  for (int n = 0; n < N; ++n) base[(size_t)n * ncell] = in[n];
}

// Field pointers of one zone
template <bool MHD> struct ZonePtr {
  Real *cons, *prim, *temp, *mu, *fail, *csig;
  int ncell;
  PB_DEV ZonePtr(const pb_pack_t &pack, Real *cons_slab, const Lay &L, int b, int o) : ncell(L.ncell) {
    cons = cons_slab + (size_t)b * NC<MHD>::ncons * L.ncell + o;
    prim = pack.prim + (size_t)b * NC<MHD>::nprim * L.ncell + o;
    temp = pack.temp + (size_t)b * L.ncell + o;
    mu = pack.c2p_mu + (size_t)b * L.ncell + o;
    fail = pack.fail + (size_t)b * L.ncell + o;
    csig = pack.csig + (size_t)b * L.ndim * L.ncell + o;
  }
  // a zone whose fields live in the twin side buffer (entry e of S)
  PB_DEV ZonePtr(const TwinBuf &tb, int e) : ncell(tb.S) {
    const size_t S = (size_t)tb.S;
    cons = tb.base + e;
    prim = cons + NC<MHD>::ncons * S;
    temp = prim + NC<MHD>::nprim * S;
    mu = temp + S;
    fail = mu + S;
    csig = fail + S;
  }
};

// c2p on one zone with the reference's write pattern: nothing is written when the NaN guard trips,
// only p.bfield when the root is not bracketed (the reference aborts there; we count it).
// ARITH 0: the compiler's divide / sqrt. 1: ArithFast with an inline exact re-solve. 2: ArithFast only; a zone
// whose operands leave the fast path's exponent range writes NOTHING and reports `rejected` (the caller queues
// it for c2p_redo_kernel).
template <bool MHD, int ARITH, class G>
PB_DEV bool c2p_zone_global(const pb_cfg_t &cfg, const FloorCtx &fc, const G &g, const ZonePtr<MHD> &z, int ndim,
                            Real *U, Real *P, unsigned long long *err_count, bool *rejected = nullptr) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  Real temp, sig[3];
  Real mu = *z.mu;
  const Real D0 = U[CRHO], S0 = U[CMOM], S1 = U[CMOM + 1], S2 = U[CMOM + 2], T0 = U[CENG];
  bool unbracketed;
  const bool nan_in = isnan(U[CRHO]) + isnan(U[CMOM]) + isnan(U[CENG]) > 0;
  bool ok;
  if (ARITH == 2) {
    ArithFast af;
    ok = c2p_zone<MHD>(af, cfg, fc, g, U, P, temp, mu, sig, unbracketed);
    if (af.failed()) {
      *rejected = true;
      return false;
    }
  } else if (ARITH == 1) {
    // deferred-branch divide / sqrt (ArithFast); a zone with an operand outside its exponent range is
    // solved again from the saved inputs with the compiler's own sequences
    const Real mu0 = mu;
    ArithFast af;
    ok = c2p_zone<MHD>(af, cfg, fc, g, U, P, temp, mu, sig, unbracketed);
    if (af.failed()) {
      U[CRHO] = D0, U[CMOM] = S0, U[CMOM + 1] = S1, U[CMOM + 2] = S2, U[CENG] = T0;
      mu = mu0;
      ArithExact ax;
      ok = c2p_zone<MHD>(ax, cfg, fc, g, U, P, temp, mu, sig, unbracketed);
    }
  } else {
    ArithExact ax;
    ok = c2p_zone<MHD>(ax, cfg, fc, g, U, P, temp, mu, sig, unbracketed);
  }
  if (nan_in) return false;
  if (unbracketed) {
    if (MHD) {
#pragma unroll
      for (int a = 0; a < 3; ++a) z.prim[(size_t)(PB_ + a) * z.ncell] = P[PB_ + a];
    }
    if (err_count) atomicAdd(err_count, 1ULL);
    return false;
  }
  store_comps<NP>(z.prim, z.ncell, P);
  *z.temp = temp;
  *z.mu = mu;
  for (int a = 0; a < ndim; ++a) z.csig[(size_t)a * z.ncell] = sig[a];
  if (U[CRHO] != D0 || U[CMOM] != S0 || U[CMOM + 1] != S1 || U[CMOM + 2] != S2 || U[CENG] != T0 ||
      isnan(U[CRHO]) || isnan(U[CMOM]) || isnan(U[CMOM + 1]) || isnan(U[CMOM + 2]) || isnan(U[CENG])) {
#pragma unroll
    for (int n = 0; n < 5; ++n) z.cons[(size_t)n * z.ncell] = U[n];
  }
  (void)NCn;
  return ok;
}

// ------------------------------------------------------------------ zone flags
// One byte per zone (ghosts included) lets the sequencer skip the zones on which the fixup and
// floor tasks are no-ops (almost all of them) without re-reading their primitives:
//   ZF_FAIL  the last entire-domain c2p failed here (fail == 0)
//   ZF_FLOOR floors/ceilings would change the zone's current primitives (post-c2p pass pending)
//   ZF_STALE floors/ceilings would change the zone's FINAL primitives of the last stage, i.e. the
//            stale-prim ApplyFloors of the next stage's region 7 fires here
enum { ZF_FAIL = 1, ZF_FLOOR = 2, ZF_STALE = 4 };

// first part of ApplyFloors (ref: src/fixup/fixup.cpp:370-440): decides whether the floors apply.
template <bool MHD, class G>
PB_DEV bool floors_decide(const pb_cfg_t &cfg, const FloorCtx &fc, const G &g, const Real *P, Real &drho,
                          Real &du, Real &sie_floor) {
  Real rho_floor = 0.0; // ref passes an uninitialised rho_floor; only ExpX1RhoU reads it
  get_floors(cfg, fc.A, fc.B, fc.x1, fc.x2, fc.x3, rho_floor, sie_floor);
  const Real e_max = cfg.sie_max;
  Real rho_floor_max = rho_floor;
  Real u_floor_max = rho_floor * sie_floor;
  if (cfg.enable_mhd_floors) {
    Real Bsq = 0.0, Bdotv = 0.0;
    const Real vp[3] = {P[PVEL], P[PVEL + 1], P[PVEL + 2]};
    Real bp[3] = {0, 0, 0};
    if (MHD)
      for (int a = 0; a < 3; ++a) bp[a] = P[PB_ + a];
    const Real W = lorentz_factor(vp, g);
    const Real iW = 1.0 / W;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Bsq += g.g4(a + 1, c + 1) * bp[a] * bp[c];
        Bdotv += g.g4(a + 1, c + 1) * bp[a] * vp[c]; // vp, not vp/W: ref fixup.cpp:420-425
      }
    const Real bcon0 = W * Bdotv / g.alpha;
    const Real bsq = (Bsq + g.alpha * g.alpha * bcon0 * bcon0) * iW * iW;
    rho_floor_max = dmax(rho_floor_max, bsq / cfg.bsqorho_max);
    u_floor_max = dmax(u_floor_max, bsq / cfg.bsqou_max);
    rho_floor_max = dmax(rho_floor_max, dmax(P[PENG], u_floor_max) / e_max);
  }
  drho = rho_floor_max - P[PRHO];
  du = u_floor_max - P[PENG];
  return (drho > 0. || du > 0.);
}
// ceiling test of ApplyFloors (ref: fixup.cpp:505-507) on the given primitives
template <class G> PB_DEV bool ceiling_decide(const pb_cfg_t &cfg, const G &g, const Real *P, Real &W) {
  const Real vpcon[3] = {P[PVEL], P[PVEL + 1], P[PVEL + 2]};
  W = lorentz_factor(vpcon, g);
  return (W > cfg.gam_max || P[PENG] / P[PRHO] > cfg.sie_max);
}
// would ApplyFloors change this zone?
template <bool MHD, class G> PB_DEV bool floors_needed(const pb_cfg_t &cfg, const FloorCtx &fc, const G &g, const Real *P) {
  Real drho, du, sf, W;
  if (floors_decide<MHD>(cfg, fc, g, P, drho, du, sf)) return true;
  return ceiling_decide(cfg, g, P, W);
}

// Conserved variables of a zone as ApplyFloors sees them. In the stale-prim pass of region 7 the
// ghost zones of the stage-output container hold beta*u0 + (1-beta)*ubase (WeightedSumData runs over
// entire arrays and dUdt is zero there); the sequencer does not materialise that, it is formed here.
struct LazyGhost {
  const Real *u0, *ubase; // NULL: read the container
  Real beta;
};

// ApplyFloors on one zone, ref: src/fixup/fixup.cpp:370-684 (fluid part). Returns true if the zone
// was changed.
template <bool MHD>
PB_DEV bool floors_zone(const pb_cfg_t &cfg, const FloorCtx &fc, const PointGeom &g, const ZonePtr<MHD> &z, int ndim,
                        const Real *lazy_u0, const Real *lazy_ubase, Real beta, unsigned long long *counters) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons, prs = NC<MHD>::prs, gm1 = NC<MHD>::gm1;
  const int N = z.ncell;
  const Eos eos(cfg);
  Real P[NP];
  load_comps<NP>(z.prim, N, P);
  Real bp[3] = {0, 0, 0};
  if (MHD)
    for (int a = 0; a < 3; ++a) bp[a] = P[PB_ + a];
  Real drho, du, sie_floor;
  bool floor_applied = floors_decide<MHD>(cfg, fc, g, P, drho, du, sie_floor), ceiling_applied = false;
  if (floor_applied) {
    drho = dmax(drho, du / sie_floor);
    du = dmax(du, sie_floor * drho);
  }
  Real U[NCn];
  bool have_U = false, store_P = false, store_U = false;
  auto load_U = [&]() {
    if (lazy_u0) {
      const Real w2 = 1.0 - beta;
#pragma unroll
      for (int n = 0; n < NCn; ++n) U[n] = 1.0 * (beta * lazy_u0[(size_t)n * N] + w2 * lazy_ubase[(size_t)n * N]) + 0.0;
    } else {
      load_comps<NCn>(z.cons, N, U);
    }
    have_U = true;
  };
  Real temp_out = 0.0;
  if (floor_applied) {
    if (counters) atomicAdd(&counters[0], 1ULL);
    load_U();
    const Real vp_normalobs[3] = {0, 0, 0};
    const Real dprs = eos.P_rho_sie(drho, ratio(du, drho));
    const Real dgm1 = ratio(eos.K_rho_sie(drho, ratio(du, drho)), dprs);
    Real dcrho, dS[3], dBcons[3], dtau, nosig[3];
    p2c<false>(drho, vp_normalobs, bp, du, dprs, dgm1, g, dcrho, dS, dBcons, dtau, nosig);
    U[CRHO] += dcrho;
    for (int a = 0; a < 3; ++a) U[CMOM + a] += dS[a];
    U[CENG] += dtau;
    // the in-kernel c2p reads the updated cons and writes prims/temp/mu/csig/cons itself
#pragma unroll
    for (int n = 0; n < 5; ++n) z.cons[(size_t)n * N] = U[n];
    Real Pn[NP];
    const bool ok = c2p_zone_global<MHD, 0>(cfg, fc, g, z, ndim, U, Pn, nullptr);
    // refresh the local view of the zone from what c2p wrote (it never touches c.bfield)
    load_comps<NP>(z.prim, N, P);
    load_comps<5>(z.cons, N, U);
    if (!ok) {
      if (counters) atomicAdd(&counters[2], 1ULL);
      P[PRHO] = drho;
      for (int a = 0; a < 3; ++a) P[PVEL + a] = vp_normalobs[a];
      P[PENG] = du;
      temp_out = eos.T_rho_sie(P[PRHO], P[PENG] / P[PRHO]);
      *z.temp = temp_out;
      p2c<false>(drho, vp_normalobs, bp, du, dprs, dgm1, g, U[CRHO], dS, dBcons, U[CENG], nosig);
      for (int a = 0; a < 3; ++a) U[CMOM + a] = dS[a];
      store_P = store_U = true;
    }
    // bp keeps the pre-c2p p.bfield, as in the reference (fixup.cpp:441-447)
  }
  Real W;
  ceiling_applied = ceiling_decide(cfg, g, P, W);
  if (ceiling_applied) {
    if (counters) atomicAdd(&counters[1], 1ULL);
    if (!have_U) load_U();
    Real vpcon[3] = {P[PVEL], P[PVEL + 1], P[PVEL + 2]};
    const Real rescale = sqrt(cfg.gam_max * cfg.gam_max - 1.) / (W * W - 1.); // sic, ref: fixup.cpp:510
    for (int a = 0; a < 3; ++a) vpcon[a] *= rescale;
    for (int a = 0; a < 3; ++a) P[PVEL + a] = vpcon[a];
    Real dS[3], dBcons[3], nosig[3];
    p2c<false>(P[PRHO], vpcon, bp, P[PENG], P[prs], P[gm1], g, U[CRHO], dS, dBcons, U[CENG], nosig);
    for (int a = 0; a < 3; ++a) U[CMOM + a] = dS[a];
    store_P = store_U = true;
  }
  if (floor_applied || ceiling_applied) {
    temp_out = eos.T_rho_sie(P[PRHO], ratio(P[PENG], P[PRHO]));
    P[prs] = eos.P_rho_T(P[PRHO], temp_out);
    P[gm1] = ratio(eos.K_rho_T(P[PRHO], temp_out), P[prs]);
    *z.temp = temp_out;
    store_P = true;
  }
  if (store_P) store_comps<NP>(z.prim, N, P);
  if (store_U) {
#pragma unroll
    for (int n = 0; n < 5; ++n) z.cons[(size_t)n * N] = U[n];
  }
  return floor_applied || ceiling_applied;
}

// ConservedToPrimitiveFixup on one failed zone, ref: src/fixup/fixup_c2p.cpp:170-353
template <bool MHD>
PB_DEV void fixup_zone(const pb_cfg_t &cfg, const Lay &L, const PointGeom &g, const ZonePtr<MHD> &z) {
  constexpr int NP = NC<MHD>::nprim, prs = NC<MHD>::prs, gm1 = NC<MHD>::gm1;
  const int N = L.ncell;
  const int st1 = L.ni, st2 = L.ni * L.nj;
  const Real *fl = z.fail;
  const Eos eos(cfg);
  const Real gamma_max = cfg.gam_max;
  auto fixup = [&](int iv, Real inv_mask_sum) {
    Real *v = z.prim + (size_t)iv * N;
    Real acc = fl[-1] * v[-1] + fl[1] * v[1];
    if (L.ndim > 1) {
      acc += fl[-st1] * v[-st1] + fl[st1] * v[st1];
      if (L.ndim == 3) acc += fl[-st2] * v[-st2] + fl[st2] * v[st2];
    }
    return inv_mask_sum * acc;
  };
  Real num_valid = fl[-1] + fl[1];
  if (L.ndim > 1) num_valid += fl[-st1] + fl[st1];
  if (L.ndim == 3) num_valid += fl[-st2] + fl[st2];
  Real P[NP];
  load_comps<NP>(z.prim, N, P);
  if (num_valid > 0.5 && cfg.c2p_failure_strategy == 0) {
    const Real norm = 1.0 / num_valid;
    P[PRHO] = fixup(PRHO, norm);
    for (int pv = PVEL; pv <= PVEL + 2; pv++) P[pv] = fixup(pv, norm);
    P[PENG] = fixup(PENG, norm);
    P[PRHO] = dmax(P[PRHO], 100. * R_SMALL);
    P[PENG] = dmax(P[PENG], 100. * R_SMALL);
  } else {
    P[PRHO] = 100. * R_SMALL;
    P[PENG] = 100. * R_SMALL;
    for (int ii = 0; ii < 3; ++ii) P[PVEL + ii] = 0.;
  }
  Real vpcon[3] = {P[PVEL], P[PVEL + 1], P[PVEL + 2]};
  Real W = lorentz_factor(vpcon, g);
  if (W > gamma_max) {
    const Real rescale = sqrt(gamma_max * gamma_max - 1.) / (W * W - 1.); // sic, ref: fixup_c2p.cpp:271
    for (int ii = 0; ii < 3; ++ii) vpcon[ii] *= rescale;
    for (int ii = 0; ii < 3; ++ii) P[PVEL + ii] = vpcon[ii];
    W = gamma_max;
  }
  const Real temp = eos.T_rho_sie(P[PRHO], ratio(P[PENG], P[PRHO]));
  P[prs] = eos.P_rho_T(P[PRHO], temp);
  P[gm1] = ratio(eos.K_rho_T(P[PRHO], temp), P[prs]);
  Real S[3], bcons[3], bp[3] = {0.0, 0.0, 0.0}, sig[3], D, tau;
  if (MHD)
    for (int a = 0; a < 3; ++a) bp[a] = P[PB_ + a];
  p2c<true>(P[PRHO], vpcon, bp, P[PENG], P[prs], P[gm1], g, D, S, bcons, tau, sig);
  store_comps<NP>(z.prim, N, P);
  *z.temp = temp;
  z.cons[(size_t)CRHO * N] = D;
  for (int a = 0; a < 3; ++a) z.cons[(size_t)(CMOM + a) * N] = S[a];
  z.cons[(size_t)CENG * N] = tau;
  for (int a = 0; a < L.ndim; ++a) z.csig[(size_t)a * N] = sig[a];
}

struct ZoneCtx { // everything a zone task needs, built from the thread's (b,k,j,i)
  int b, k, j, i, o;
  bool interior;
};

// ------------------------------------------------------------------ ConToPrim::Solve
// ref: src/fluid/fluid.cpp:477-517. With zflags != NULL the kernel also records, per zone, whether
// the c2p failed and whether ApplyFloors will have to touch the zone (one extra byte written instead
// of a second full pass over the primitives).
#ifndef PB_C2P_MINB
#define PB_C2P_MINB 5 // 96 registers, 20 warps/SM: measured 11.65 ms vs 11.81 (4 CTAs) and 12.4 (6 CTAs) per cycle
#endif
template <bool MHD, int ARITH>
PB_DEV bool c2p_thread(const pb_pack_t &pack, const pb_cfg_t &cfg, Real *cons, const Lay &L, int b, int k, int j, int i,
                       unsigned long long *err_count, unsigned char *zflags) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  const int o = L.idx(k, j, i);
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  PointGeom g; // (a shared-memory metric, SmemGeom, was measured: same spills, no gain)
  g.load(gt, 0, gt.cell(k, j, i));
  const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
  const ZonePtr<MHD> z(pack, cons, L, b, o);
  Real U[NCn], P[NP];
  load_comps<NCn>(z.cons, L.ncell, U);
  bool rejected = false;
  const bool ok = c2p_zone_global<MHD, ARITH>(cfg, fc, g, z, L.ndim, U, P, err_count, &rejected);
  if (ARITH == 2 && rejected) return true;
  *z.fail = ok ? 1.0 : 0.0;
  if (zflags) {
    unsigned char f = ok ? 0 : ZF_FAIL;
    if (cfg.enable_floors) {
      // the primitives ApplyFloors will see: what c2p just wrote, or (NaN guard / unbracketed root)
      // whatever the zone held before
      const bool nan_in = isnan(U[CRHO]) + isnan(U[CMOM]) + isnan(U[CENG]) > 0;
      if (!ok) load_comps<NP>(z.prim, L.ncell, P);
      (void)nan_in;
      if (floors_needed<MHD>(cfg, fc, g, P)) f |= ZF_FLOOR;
    }
    zflags[(size_t)b * L.ncell + o] = f;
  }
  return false;
}

// Which zones a c2p launch covers: a box of every block (the task as the reference runs it), or — for the stage
// sequencer that hides the inter-GPU halo transfer under interior work — a list of blocks combined with either a
// box or the ghost shell. Queue entries are block * ncell + cell so that the redo kernel needs no enumeration.
struct C2PWork {
  Region reg;        // box (used unless shell)
  const int *blocks; // block list or NULL (= all blocks of the pack)
  int nlist;         // blocks covered
  int shell;         // 1: ghost shell instead of the box
  const unsigned *zones; // ghost plan: the box of the nlist blocks (enumerated arithmetically: no dependent load in front of
  long long nzones;      // the zone's own loads), followed by nzones explicit zone ids (block * ncell + cell)
};
PB_DEV bool c2p_zone_of_list(const C2PWork &w, const Lay &L, long long t, int &b, int &k, int &j, int &i) {
  const long long nreg = w.reg.n(), dense = nreg * w.nlist;
  if (t < dense) {
    const int q = (int)(t / nreg);
    b = w.blocks ? w.blocks[q] : q;
    int rem = (int)(t - (long long)q * nreg);
    i = w.reg.i0 + rem % w.reg.ni();
    rem /= w.reg.ni();
    j = w.reg.j0 + rem % w.reg.nj();
    k = w.reg.k0 + rem / w.reg.nj();
    return true;
  }
  t -= dense;
  if (t >= w.nzones) return false;
  const unsigned z = __ldg(w.zones + t);
  b = (int)(z / (unsigned)L.ncell);
  int rem = (int)(z - (unsigned)b * (unsigned)L.ncell);
  i = rem % L.ni;
  rem /= L.ni;
  j = rem % L.nj;
  k = rem / L.nj;
  return true;
}
PB_DEV bool c2p_zone_of_thread(const C2PWork &w, const Lay &L, int &b, int &k, int &j, int &i) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w.zones) return c2p_zone_of_list(w, L, t, b, k, j, i);
  if (w.shell) {
    const Shell sh(L);
    if (t >= (long long)sh.n * w.nlist) return false;
    const int q = (int)(t / sh.n);
    b = w.blocks ? w.blocks[q] : q;
    sh.cell(L, (int)(t - (long long)q * sh.n), k, j, i);
    return true;
  }
  const long long nreg = w.reg.n();
  if (t >= nreg * w.nlist) return false;
  const int q = (int)(t / nreg);
  b = w.blocks ? w.blocks[q] : q;
  int rem = (int)(t - (long long)q * nreg);
  i = w.reg.i0 + rem % w.reg.ni();
  rem /= w.reg.ni();
  j = w.reg.j0 + rem % w.reg.nj();
  k = w.reg.k0 + rem / w.reg.nj();
  return true;
}
template <bool MHD, int ARITH>
__global__ void __launch_bounds__(128, PB_C2P_MINB) c2p_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const C2PWork work,
                                                               unsigned long long *err_count, unsigned char *zflags, const RedoQueue rq) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!c2p_zone_of_thread(work, L, b, k, j, i)) return;
  if (c2p_thread<MHD, ARITH>(pack, cfg, cons, L, b, k, j, i, err_count, zflags))
    rq.list[atomicAdd(rq.count, 1u)] = (unsigned)((long long)b * L.ncell + L.idx(k, j, i));
}
// zones the fast arithmetic rejected, solved with the compiler's divide / sqrt (see RedoQueue, pb_host.hpp)
template <bool MHD>
__global__ void __launch_bounds__(128, 4) c2p_redo_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons,
                                                          unsigned long long *err_count, unsigned char *zflags, const RedoQueue rq) {
  const Lay L(cfg);
  const unsigned n = *rq.count;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(rq.total, (unsigned long long)n);
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned t = rq.list[e];
    const int b = (int)(t / (unsigned)L.ncell), o = (int)(t % (unsigned)L.ncell);
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    c2p_thread<MHD, 0>(pack, cfg, cons, L, b, k, j, i, err_count, zflags);
  }
}

// ------------------------------------------------------------------ ConToPrim::Solve as separate launches
// The pack-level con2prim in three phases (option c2p_arith = 3, the default):
//   c2p_solve_kernel    the root find only (setup + regula falsi): the long, divergent part. Small code, no primitive
//                       state in registers. Writes the root into c2p_mu and one status byte per zone.
//                       WARP-LEVEL HANDLING OF SLOWLY CONVERGING ZONES: after every iteration the warp votes; once at
//                       most PB_C2P_DEFER lanes are still iterating, those lanes park their bracket state in a device
//                       queue and the warp retires. c2p_slow_kernel resumes the parked zones, 32 slow zones per warp,
//                       from exactly the saved state (RFState), so the bits do not depend on where a zone converged.
//   c2p_unwrap_kernel   setup again + unwrap at the root + p2c (bounds used) or signal speeds, stores, fail flag, zone
//                       flags: uniform work for every zone.
//   c2p_redo_kernel     zones the fast arithmetic rejected: the whole solve (queued by the solve) or the unwrap only
//                       (queued by the unwrap, bit 31 set) with the compiler's divide / sqrt.
// Workspace (library-owned, grown on demand): status bytes, slow queue, iteration histogram.
enum { C2S_OK = 0, C2S_NAN = 1, C2S_UNBRACKETED = 2, C2S_REDO = 3, C2S_SLOW = 4 };
struct SlowEntry {
  unsigned zone;
  int it, b1, b2;
  Real a, b, ya, yb, sign;
};
struct C2PWorkspace {
  unsigned char *status;     // [nblocks * ncell]
  SlowEntry *slow;           // parked root finds
  unsigned *slow_count;      // entries appended by the current launch
  unsigned slow_cap;
  unsigned long long *stats; // optional (pb_c2p_stats): [0..31] iterations per solve, [32] solves that used a bound,
                             // [33] failed solves, [34] zones parked, [35] zones solved
};
#ifndef PB_C2P_DEFER
#define PB_C2P_DEFER 4 // park when at most this many lanes of the warp are still iterating ...
#endif
#ifndef PB_C2P_DEFER_IT
#define PB_C2P_DEFER_IT 4 // ... and they have done at least this many iterations (1 - 3 iterations are the norm, measured)
#endif
// statistics are summed per CTA in shared memory and flushed with one atomic per non-empty counter
constexpr int C2P_NSTAT = 36;
PB_DEV void stats_begin(unsigned *s_stat) {
  for (int q = threadIdx.x; q < C2P_NSTAT; q += blockDim.x) s_stat[q] = 0;
  __syncthreads();
}
PB_DEV void stats_end(const unsigned *s_stat, unsigned long long *stats) {
  __syncthreads();
  if (stats)
    for (int q = threadIdx.x; q < C2P_NSTAT; q += blockDim.x)
      if (s_stat[q]) atomicAdd(&stats[q], (unsigned long long)s_stat[q]);
}
#ifndef PB_C2P_SOLVE_MINB
#define PB_C2P_SOLVE_MINB 6
#endif

// geometry a root find needs: inverse 3-metric, spatial metric, sqrt(gamma) (no lapse / shift)
struct SolveGeom {
  Real detgam, gcov[10], gamcon[6];
  PB_DEV Real g4(int m, int n) const { return gcov[flat4(m, n)]; }
  PB_DEV Real gc3(int m, int n) const { return gamcon[flat4(m, n)]; }
  PB_DEV void load(const GeomTab &g, int cell) {
    detgam = g.at(0, G_DETGAM, cell);
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int n = m; n < 3; ++n) gcov[flat4(m + 1, n + 1)] = g.at(0, G_GCOV + flat4(m + 1, n + 1), cell);
#pragma unroll
    for (int m = 0; m < 6; ++m) gamcon[m] = g.at(0, G_GAMCON + m, cell);
  }
};

PB_DEV bool c2p_zone_of_index(const C2PWork &w, const Lay &L, long long t, int &b, int &k, int &j, int &i) {
  if (w.zones) return c2p_zone_of_list(w, L, t, b, k, j, i);
  if (w.shell) {
    const Shell sh(L);
    const int q = (int)(t / sh.n);
    b = w.blocks ? w.blocks[q] : q;
    sh.cell(L, (int)(t - (long long)q * sh.n), k, j, i);
    return true;
  }
  const long long nreg = w.reg.n();
  const int q = (int)(t / nreg);
  b = w.blocks ? w.blocks[q] : q;
  int rem = (int)(t - (long long)q * nreg);
  i = w.reg.i0 + rem % w.reg.ni();
  rem /= w.reg.ni();
  j = w.reg.j0 + rem % w.reg.nj();
  k = w.reg.k0 + rem / w.reg.nj();
  return true;
}

template <bool MHD>
__global__ void __launch_bounds__(128, PB_C2P_SOLVE_MINB) c2p_solve_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *__restrict__ cons,
                                                                            const C2PWork work, const long long ntot, const C2PWorkspace ws,
                                                                            const RedoQueue rq) {
  constexpr int NCn = NC<MHD>::ncons;
  __shared__ unsigned s_stat[C2P_NSTAT];
  stats_begin(s_stat);
  const Lay L(cfg);
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < ntot;
  if (!valid) t = ntot - 1; // lanes past the end shadow the last zone (they take part in the warp votes, never store)
  int b, k, j, i;
  c2p_zone_of_index(work, L, t, b, k, j, i);
  const int o = L.idx(k, j, i);
  const unsigned zone = (unsigned)((long long)b * L.ncell + o);
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  SolveGeom g;
  g.load(gt, gt.cell(k, j, i));
  const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
  Real U[NCn];
  load_comps<NCn>(cons + (size_t)b * NCn * L.ncell + o, L.ncell, U);
  const bool nan_in = isnan(U[CRHO]) + isnan(U[CMOM]) + isnan(U[CENG]) > 0;
  ArithFast af;
  C2PVars<MHD> s;
  c2p_setup<MHD>(af, cfg, fc, g, U, s);
  Residual<ArithFast> res = c2p_residual<MHD>(af, cfg, fc, s);
  RFState st;
  const Real guess = pack.c2p_mu[zone];
  const bool bracketed = rf_begin(af, res, 0.0, 1.0, guess, st);
  bool active = valid && !nan_in && bracketed && rf_running(st, cfg.c2p_tol, cfg.c2p_max_iter);
  bool parked = false;
  while (__any_sync(0xffffffffu, active)) {
    if (active) {
      rf_iterate(af, res, st);
      active = rf_running(st, cfg.c2p_tol, cfg.c2p_max_iter);
    }
    const int nact = __popc(__ballot_sync(0xffffffffu, active));
    if (nact > 0 && nact <= PB_C2P_DEFER && st.it >= PB_C2P_DEFER_IT) {
      // a few stragglers hold the warp: park them (unless the fast arithmetic already rejected the zone, or the queue
      // is full: then they simply finish here)
      if (active && !af.failed()) {
        const unsigned e = atomicAdd(ws.slow_count, 1u);
        if (e < ws.slow_cap) {
          SlowEntry &q = ws.slow[e];
          q.zone = zone;
          q.it = st.it, q.b1 = st.b1, q.b2 = st.b2;
          q.a = st.a, q.b = st.b, q.ya = st.ya, q.yb = st.yb, q.sign = st.sign;
          parked = true;
          active = false;
        }
      }
    }
  }
  if (valid) {
    unsigned char status;
    if (nan_in) status = C2S_NAN;
    else if (parked) status = C2S_SLOW;
    else if (af.failed()) {
      status = C2S_REDO;
      rq.list[atomicAdd(rq.count, 1u)] = zone;
    } else if (!bracketed) status = C2S_UNBRACKETED;
    else {
      status = C2S_OK;
      pack.c2p_mu[zone] = 0.5 * (st.a + st.b);
      if (ws.stats) atomicAdd(&s_stat[st.it < 31 ? st.it : 31], 1u);
    }
    ws.status[zone] = status;
    if (ws.stats && parked) atomicAdd(&s_stat[34], 1u);
  }
  stats_end(s_stat, ws.stats);
}

template <bool MHD>
__global__ void __launch_bounds__(128, PB_C2P_SOLVE_MINB) c2p_slow_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *__restrict__ cons,
                                                                           const C2PWorkspace ws, const RedoQueue rq) {
  constexpr int NCn = NC<MHD>::ncons;
  const Lay L(cfg);
  unsigned n = *ws.slow_count;
  if (n > ws.slow_cap) n = ws.slow_cap;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const SlowEntry q = ws.slow[e];
    const int b = (int)(q.zone / (unsigned)L.ncell), o = (int)(q.zone % (unsigned)L.ncell);
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    const pb_block_t &blk = pack.blocks[b];
    const GeomTab gt(cfg, pack.geom, blk.geom_slot);
    SolveGeom g;
    g.load(gt, gt.cell(k, j, i));
    const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
    Real U[NCn];
    load_comps<NCn>(cons + (size_t)b * NCn * L.ncell + o, L.ncell, U);
    ArithFast af;
    C2PVars<MHD> s;
    c2p_setup<MHD>(af, cfg, fc, g, U, s);
    Residual<ArithFast> res = c2p_residual<MHD>(af, cfg, fc, s);
    RFState st;
    st.a = q.a, st.b = q.b, st.ya = q.ya, st.yb = q.yb, st.sign = q.sign;
    st.it = q.it, st.b1 = q.b1, st.b2 = q.b2;
    while (rf_running(st, cfg.c2p_tol, cfg.c2p_max_iter)) rf_iterate(af, res, st);
    if (af.failed()) {
      ws.status[q.zone] = C2S_REDO;
      rq.list[atomicAdd(rq.count, 1u)] = q.zone;
    } else {
      ws.status[q.zone] = C2S_OK;
      pack.c2p_mu[q.zone] = 0.5 * (st.a + st.b);
      if (ws.stats) atomicAdd(&ws.stats[st.it < 31 ? st.it : 31], 1ULL);
    }
  }
}

// the unwrap of one solved zone with arithmetic policy AR; returns false if AR rejected an operand (nothing stored)
template <bool MHD, class AR>
PB_DEV bool c2p_unwrap_zone(const pb_pack_t &pack, const pb_cfg_t &cfg, Real *cons, const Lay &L, int b, int k, int j, int i,
                            unsigned char *zflags, unsigned *stats) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  const int o = L.idx(k, j, i);
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  PointGeom g;
  g.load(gt, 0, gt.cell(k, j, i));
  const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
  const ZonePtr<MHD> z(pack, cons, L, b, o);
  Real U[NCn], P[NP], temp, sig[3];
  load_comps<NCn>(z.cons, L.ncell, U);
  const Real D0 = U[CRHO], S0 = U[CMOM], S1 = U[CMOM + 1], S2 = U[CMOM + 2], T0 = U[CENG];
  AR ar;
  C2PVars<MHD> s;
  c2p_setup<MHD>(ar, cfg, fc, g, U, s);
  Residual<AR> res = c2p_residual<MHD>(ar, cfg, fc, s);
  const bool ok = c2p_unwrap<MHD>(ar, cfg, g, s, res, *z.mu, U, P, temp, sig);
  if (ar.failed()) return false;
  store_comps<NP>(z.prim, z.ncell, P);
  *z.temp = temp;
  for (int a = 0; a < L.ndim; ++a) z.csig[(size_t)a * z.ncell] = sig[a];
  if (U[CRHO] != D0 || U[CMOM] != S0 || U[CMOM + 1] != S1 || U[CMOM + 2] != S2 || U[CENG] != T0 ||
      isnan(U[CRHO]) || isnan(U[CMOM]) || isnan(U[CMOM + 1]) || isnan(U[CMOM + 2]) || isnan(U[CENG])) {
#pragma unroll
    for (int n = 0; n < 5; ++n) z.cons[(size_t)n * z.ncell] = U[n];
  }
  *z.fail = ok ? 1.0 : 0.0;
  if (stats) { // shared-memory counters of the CTA
    if (res.used_bounds()) atomicAdd(&stats[32], 1u);
    if (!ok) atomicAdd(&stats[33], 1u);
    atomicAdd(&stats[35], 1u);
  }
  if (zflags) {
    unsigned char f = ok ? 0 : ZF_FAIL;
    if (cfg.enable_floors && floors_needed<MHD>(cfg, fc, g, P)) f |= ZF_FLOOR;
    zflags[(size_t)b * L.ncell + o] = f;
  }
  return true;
}

// one thread of the unwrap launch: dispatch on the status byte the solve left
template <bool MHD>
PB_DEV void c2p_unwrap_thread(const pb_pack_t &pack, const pb_cfg_t &cfg, Real *cons, const C2PWork &work, const Lay &L, long long t,
                              const C2PWorkspace &ws, unsigned long long *err_count, unsigned char *zflags, const RedoQueue &rq,
                              unsigned *s_stat) {
  constexpr int NP = NC<MHD>::nprim;
  int b, k, j, i;
  c2p_zone_of_index(work, L, t, b, k, j, i);
  const int o = L.idx(k, j, i);
  const unsigned zone = (unsigned)((long long)b * L.ncell + o);
  const unsigned char status = ws.status[zone];
  if (status == C2S_REDO) return;
  if (status == C2S_OK) {
    if (!c2p_unwrap_zone<MHD, ArithFast>(pack, cfg, cons, L, b, k, j, i, zflags, s_stat))
      rq.list[atomicAdd(rq.count, 1u)] = zone | 0x80000000u;
    return;
  }
  // NaN guard: nothing written. Unbracketed root: p.bfield only (the reference aborts there; counted)
  const ZonePtr<MHD> z(pack, cons, L, b, o);
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  if (status == C2S_UNBRACKETED) {
    if (MHD) {
      const Real igdet = 1.0 / gt.at(0, G_DETGAM, gt.cell(k, j, i));
#pragma unroll
      for (int a = 0; a < 3; ++a) z.prim[(size_t)(PB_ + a) * z.ncell] = z.cons[(size_t)(CB + a) * z.ncell] * igdet;
    }
    if (err_count) atomicAdd(err_count, 1ULL);
  }
  *z.fail = 0.0;
  if (s_stat) atomicAdd(&s_stat[33], 1u);
  if (zflags) {
    unsigned char f = ZF_FAIL;
    if (cfg.enable_floors) {
      PointGeom g;
      g.load(gt, 0, gt.cell(k, j, i));
      const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
      Real P[NP];
      load_comps<NP>(z.prim, L.ncell, P);
      if (floors_needed<MHD>(cfg, fc, g, P)) f |= ZF_FLOOR;
    }
    zflags[(size_t)b * L.ncell + o] = f;
  }
}

#ifndef PB_C2P_UNWRAP_MINB
#define PB_C2P_UNWRAP_MINB 4
#endif
template <bool MHD>
__global__ void __launch_bounds__(128, PB_C2P_UNWRAP_MINB) c2p_unwrap_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const C2PWork work,
                                                                              const long long ntot, const C2PWorkspace ws, unsigned long long *err_count,
                                                                              unsigned char *zflags, const RedoQueue rq) {
  __shared__ unsigned s_stat[C2P_NSTAT];
  stats_begin(s_stat);
  const Lay L(cfg);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntot) c2p_unwrap_thread<MHD>(pack, cfg, cons, work, L, t, ws, err_count, zflags, rq, ws.stats ? s_stat : nullptr);
  stats_end(s_stat, ws.stats);
}

// rejected work of the split path: bit 31 clear = the whole zone (solve + unwrap) from the untouched guess, bit 31
// set = the unwrap only (the root in c2p_mu is already exact), both with the compiler's divide / sqrt
template <bool MHD>
__global__ void __launch_bounds__(128, 4) c2p_split_redo_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons,
                                                                unsigned long long *err_count, unsigned char *zflags, const RedoQueue rq,
                                                                unsigned long long *stats) {
  const Lay L(cfg);
  const unsigned n = *rq.count;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(rq.total, (unsigned long long)n);
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned t = rq.list[e] & 0x7fffffffu;
    const bool unwrap_only = (rq.list[e] & 0x80000000u) != 0;
    const int b = (int)(t / (unsigned)L.ncell), o = (int)(t % (unsigned)L.ncell);
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    if (unwrap_only) c2p_unwrap_zone<MHD, ArithExact>(pack, cfg, cons, L, b, k, j, i, zflags, nullptr);
    else c2p_thread<MHD, 0>(pack, cfg, cons, L, b, k, j, i, err_count, zflags);
    if (stats) atomicAdd(&stats[35], 1ULL); // redone zones: counted as solved, not in the histogram
  }
}

// ------------------------------------------------------------------ ConToPrim::Solve, one launch + warp-level parking
// Option c2p_arith = 2 (the default): solve and unwrap in ONE kernel (measured faster than the three-launch split on the
// torus, where the guess from the previous stage makes 1 - 3 regula-falsi iterations the norm and the unwrap + p2c tail is
// the larger part), with the same warp-level handling of slowly converging zones: after every iteration the warp votes,
// and once at most PB_C2P_DEFER lanes are still iterating (after >= PB_C2P_DEFER_IT iterations) those lanes park their
// bracket state; c2p_finish_kernel resumes them, 32 slow zones per warp, and unwraps them. Zones the fast arithmetic
// rejects go to the redo queue as before.
template <bool MHD>
PB_DEV void c2p_store_zone(const pb_cfg_t &cfg, const FloorCtx &fc, const PointGeom &g, const ZonePtr<MHD> &z, const Lay &L, int b, int o,
                           const Real *U, const Real *U0, const Real *P, Real temp, Real mu, const Real sig[3], bool ok, unsigned char *zflags) {
  constexpr int NP = NC<MHD>::nprim;
  store_comps<NP>(z.prim, z.ncell, P);
  *z.temp = temp;
  *z.mu = mu;
  for (int a = 0; a < L.ndim; ++a) z.csig[(size_t)a * z.ncell] = sig[a];
  bool changed = false;
#pragma unroll
  for (int n = 0; n < 5; ++n) changed = changed || U[n] != U0[n] || isnan(U[n]);
  if (changed) {
#pragma unroll
    for (int n = 0; n < 5; ++n) z.cons[(size_t)n * z.ncell] = U[n];
  }
  *z.fail = ok ? 1.0 : 0.0;
  if (zflags) {
    unsigned char f = ok ? 0 : ZF_FAIL;
    if (cfg.enable_floors && floors_needed<MHD>(cfg, fc, g, P)) f |= ZF_FLOOR;
    zflags[(size_t)b * L.ncell + o] = f;
  }
}

template <bool MHD>
__global__ void __launch_bounds__(128, PB_C2P_MINB) c2p_park_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const C2PWork work,
                                                                    unsigned long long *err_count, unsigned char *zflags, const RedoQueue rq,
                                                                    const C2PWorkspace ws) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  __shared__ unsigned s_stat[C2P_NSTAT];
  stats_begin(s_stat);
  const Lay L(cfg);
  int b, k, j, i;
  if (c2p_zone_of_thread(work, L, b, k, j, i)) {
    const unsigned lanes = __activemask(); // the lanes of this warp that own a zone
    const int o = L.idx(k, j, i);
    const unsigned zone = (unsigned)((long long)b * L.ncell + o);
    const pb_block_t &blk = pack.blocks[b];
    const GeomTab gt(cfg, pack.geom, blk.geom_slot);
    PointGeom g;
    g.load(gt, 0, gt.cell(k, j, i));
    const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
    const ZonePtr<MHD> z(pack, cons, L, b, o);
    Real U[NCn], U0[5], P[NP];
    load_comps<NCn>(z.cons, L.ncell, U);
#pragma unroll
    for (int n = 0; n < 5; ++n) U0[n] = U[n];
    const bool nan_in = isnan(U[CRHO]) + isnan(U[CMOM]) + isnan(U[CENG]) > 0;
    ArithFast af;
    C2PVars<MHD> s;
    c2p_setup<MHD>(af, cfg, fc, g, U, s);
    Residual<ArithFast> res = c2p_residual<MHD>(af, cfg, fc, s);
    RFState st;
    const bool bracketed = rf_begin(af, res, 0.0, 1.0, *z.mu, st);
    bool active = !nan_in && bracketed && rf_running(st, cfg.c2p_tol, cfg.c2p_max_iter);
    bool parked = false;
    while (__any_sync(lanes, active)) {
      if (active) {
        rf_iterate(af, res, st);
        active = rf_running(st, cfg.c2p_tol, cfg.c2p_max_iter);
      }
      const int nact = __popc(__ballot_sync(lanes, active));
      if (nact > 0 && nact <= PB_C2P_DEFER && st.it >= PB_C2P_DEFER_IT && active && !af.failed()) {
        const unsigned e = atomicAdd(ws.slow_count, 1u);
        if (e < ws.slow_cap) {
          SlowEntry &q = ws.slow[e];
          q.zone = zone;
          q.it = st.it, q.b1 = st.b1, q.b2 = st.b2;
          q.a = st.a, q.b = st.b, q.ya = st.ya, q.yb = st.yb, q.sign = st.sign;
          parked = true;
          active = false;
        }
      }
    }
    if (parked) {
      if (ws.stats) atomicAdd(&s_stat[34], 1u);
    } else if (nan_in) { // the NaN guard: nothing but the failure flag is written (ref: con2prim_robust.hpp:318-324)
      *z.fail = 0.0;
      if (zflags) {
        load_comps<NP>(z.prim, L.ncell, P);
        zflags[(size_t)b * L.ncell + o] = ZF_FAIL | ((cfg.enable_floors && floors_needed<MHD>(cfg, fc, g, P)) ? ZF_FLOOR : 0);
      }
      if (ws.stats) atomicAdd(&s_stat[33], 1u);
    } else if (!bracketed && !af.failed()) { // the reference aborts here; p.bfield is all it has written by then
      if (MHD) {
#pragma unroll
        for (int a = 0; a < 3; ++a) z.prim[(size_t)(PB_ + a) * z.ncell] = s.Bp[a];
      }
      if (err_count) atomicAdd(err_count, 1ULL);
      *z.fail = 0.0;
      if (zflags) {
        load_comps<NP>(z.prim, L.ncell, P);
        zflags[(size_t)b * L.ncell + o] = ZF_FAIL | ((cfg.enable_floors && floors_needed<MHD>(cfg, fc, g, P)) ? ZF_FLOOR : 0);
      }
      if (ws.stats) atomicAdd(&s_stat[33], 1u);
    } else {
      const Real mu = 0.5 * (st.a + st.b);
      Real temp, sig[3];
      const bool ok = c2p_unwrap<MHD>(af, cfg, g, s, res, mu, U, P, temp, sig);
      if (af.failed()) {
        rq.list[atomicAdd(rq.count, 1u)] = zone; // nothing stored: c2p_redo_kernel solves the zone with the compiler's divide / sqrt
      } else {
        c2p_store_zone<MHD>(cfg, fc, g, z, L, b, o, U, U0, P, temp, mu, sig, ok, zflags);
        if (ws.stats) {
          atomicAdd(&s_stat[st.it < 31 ? st.it : 31], 1u);
          atomicAdd(&s_stat[35], 1u);
          if (res.used_bounds()) atomicAdd(&s_stat[32], 1u);
          if (!ok) atomicAdd(&s_stat[33], 1u);
        }
      }
    }
  }
  stats_end(s_stat, ws.stats);
}

// the parked zones of c2p_park_kernel: resume the root find from the saved bracket, unwrap, store
template <bool MHD>
__global__ void __launch_bounds__(128, PB_C2P_MINB) c2p_finish_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, unsigned char *zflags,
                                                                      const RedoQueue rq, const C2PWorkspace ws) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  const Lay L(cfg);
  unsigned n = *ws.slow_count;
  if (n > ws.slow_cap) n = ws.slow_cap;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const SlowEntry q = ws.slow[e];
    const int b = (int)(q.zone / (unsigned)L.ncell), o = (int)(q.zone % (unsigned)L.ncell);
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    const pb_block_t &blk = pack.blocks[b];
    const GeomTab gt(cfg, pack.geom, blk.geom_slot);
    PointGeom g;
    g.load(gt, 0, gt.cell(k, j, i));
    const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
    const ZonePtr<MHD> z(pack, cons, L, b, o);
    Real U[NCn], U0[5], P[NP];
    load_comps<NCn>(z.cons, L.ncell, U);
#pragma unroll
    for (int m = 0; m < 5; ++m) U0[m] = U[m];
    ArithFast af;
    C2PVars<MHD> s;
    c2p_setup<MHD>(af, cfg, fc, g, U, s);
    Residual<ArithFast> res = c2p_residual<MHD>(af, cfg, fc, s);
    RFState st;
    st.a = q.a, st.b = q.b, st.ya = q.ya, st.yb = q.yb, st.sign = q.sign;
    st.it = q.it, st.b1 = q.b1, st.b2 = q.b2;
    while (rf_running(st, cfg.c2p_tol, cfg.c2p_max_iter)) rf_iterate(af, res, st);
    const Real mu = 0.5 * (st.a + st.b);
    Real temp, sig[3];
    const bool ok = c2p_unwrap<MHD>(af, cfg, g, s, res, mu, U, P, temp, sig);
    if (af.failed()) {
      rq.list[atomicAdd(rq.count, 1u)] = q.zone;
    } else {
      c2p_store_zone<MHD>(cfg, fc, g, z, L, b, o, U, U0, P, temp, mu, sig, ok, zflags);
      if (ws.stats) {
        atomicAdd(&ws.stats[st.it < 31 ? st.it : 31], 1ULL);
        atomicAdd(&ws.stats[35], 1ULL);
        if (res.used_bounds()) atomicAdd(&ws.stats[32], 1ULL);
        if (!ok) atomicAdd(&ws.stats[33], 1ULL);
      }
    }
  }
}

// ------------------------------------------------------------------ ConToPrim::Solve fixup
// MODE 0: the task as the reference runs it (every interior zone reads its fail flag).
// MODE 1 / 2: driven by the zone flags (post-c2p pass / stale-prim pass of region 7); the zone's
// floor flag is refreshed because the fixup rewrites its primitives.
template <bool MHD, int MODE>
PB_DEV void fixup_body(const pb_pack_t &pack, const pb_cfg_t &cfg, Real *cons, unsigned char *zflags, const Lay &L, int b, int k,
                       int j, int i, int o) {
  const ZonePtr<MHD> z(pack, cons, L, b, o);
  if (MODE == 0 && *z.fail != 0.0) return;
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  PointGeom g;
  g.load(gt, 0, gt.cell(k, j, i));
  fixup_zone<MHD>(cfg, L, g, z);
  if (MODE != 0 && cfg.enable_floors) {
    constexpr int NP = NC<MHD>::nprim;
    Real P[NP];
    load_comps<NP>(z.prim, L.ncell, P);
    const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
    const bool need = floors_needed<MHD>(cfg, fc, g, P);
    unsigned char f = zflags[(size_t)b * L.ncell + o];
    const unsigned char bit = MODE == 1 ? ZF_FLOOR : ZF_STALE;
    f = need ? (f | bit) : (f & ~bit);
    zflags[(size_t)b * L.ncell + o] = f;
  }
}
template <bool MHD, int MODE>
__global__ void __launch_bounds__(128, 4) c2p_fixup_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const Region reg,
                                                           unsigned char *zflags) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!zone_of_thread(reg, pack.nblocks, b, k, j, i)) return;
  const int o = L.idx(k, j, i);
  if (MODE != 0) {
    if (!(zflags[(size_t)b * L.ncell + o] & ZF_FAIL)) return;
  }
  fixup_body<MHD, MODE>(pack, cfg, cons, zflags, L, b, k, j, i, o);
}

// Flag-driven passes touch << 1 % of the zones: a thread reads 16 flag bytes at once (one uint4) and
// only then visits the flagged zones, so the pass costs ncell/16 threads instead of ncell.
constexpr int FLAG_VEC = 16;
PB_DEV bool flagged_cells(const unsigned char *zflags, long long nvec_total, unsigned char bit, long long &base, unsigned &mask) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nvec_total) return false;
  const uint4 w = __ldg(reinterpret_cast<const uint4 *>(zflags) + t);
  const unsigned rep = 0x01010101u * bit;
  if (((w.x | w.y | w.z | w.w) & rep) == 0) return false;
  const unsigned words[4] = {w.x, w.y, w.z, w.w};
  mask = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if ((words[q] >> (8 * c)) & bit) mask |= 1u << (4 * q + c);
  base = t * FLAG_VEC;
  return true;
}
template <bool MHD, int MODE>
__global__ void __launch_bounds__(128, 4) c2p_fixup_packed_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons,
                                                                  const Region reg, unsigned char *zflags) {
  const Lay L(cfg);
  long long base;
  unsigned mask;
  if (!flagged_cells(zflags, (long long)pack.nblocks * L.ncell / FLAG_VEC, ZF_FAIL, base, mask)) return;
  const int b = (int)(base / L.ncell);
  const int o0 = (int)(base - (long long)b * L.ncell);
  for (int c = 0; c < FLAG_VEC; ++c) {
    if (!(mask >> c & 1u)) continue;
    const int o = o0 + c;
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    if (i < reg.i0 || i > reg.i1 || j < reg.j0 || j > reg.j1 || k < reg.k0 || k > reg.k1) continue;
    fixup_body<MHD, MODE>(pack, cfg, cons, zflags, L, b, k, j, i, o);
  }
}

// ------------------------------------------------------------------ ApplyFloors
// MODE 0: full scan (the task as the reference runs it). MODE 1: zones flagged ZF_FLOOR after the
// entire-domain c2p; leaves ZF_STALE = "floors would fire again on the final primitives".
// MODE 2: zones flagged ZF_STALE (region 7, stale primitives), ghost cons formed lazily.
template <bool MHD, int MODE>
PB_DEV void floors_body(const pb_pack_t &pack, const pb_cfg_t &cfg, Real *cons, unsigned long long *counters, unsigned char *zflags,
                        const LazyGhost &lazy, const Lay &L, int b, int k, int j, int i, int o, unsigned char f) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  const pb_block_t &blk = pack.blocks[b];
  const ZonePtr<MHD> z(pack, cons, L, b, o);
  const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  PointGeom g;
  g.load(gt, 0, gt.cell(k, j, i));
  const bool ghost = !(i >= L.is && i <= L.ie && j >= L.js && j <= L.je && k >= L.ks && k <= L.ke);
  const Real *l0 = nullptr, *lb = nullptr;
  if (MODE == 2 && ghost && lazy.u0) {
    l0 = lazy.u0 + (size_t)b * NCn * L.ncell + o;
    lb = lazy.ubase + (size_t)b * NCn * L.ncell + o;
  }
  floors_zone<MHD>(cfg, fc, g, z, L.ndim, l0, lb, lazy.beta, counters);
  if (MODE == 1) {
    Real P[NP];
    load_comps<NP>(z.prim, L.ncell, P);
    f &= ~(ZF_FLOOR | ZF_STALE);
    if (floors_needed<MHD>(cfg, fc, g, P)) f |= ZF_STALE;
    zflags[(size_t)b * L.ncell + o] = f;
  }
}
template <bool MHD, int MODE>
__global__ void __launch_bounds__(128, 4) floors_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const Region reg,
                                                        unsigned long long *counters, unsigned char *zflags, const LazyGhost lazy) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!zone_of_thread(reg, pack.nblocks, b, k, j, i)) return;
  const int o = L.idx(k, j, i);
  unsigned char f = 0;
  if (MODE != 0) {
    f = zflags[(size_t)b * L.ncell + o];
    if (!(f & (MODE == 1 ? ZF_FLOOR : ZF_STALE))) return;
  }
  floors_body<MHD, MODE>(pack, cfg, cons, counters, zflags, lazy, L, b, k, j, i, o, f);
}
template <bool MHD, int MODE>
__global__ void __launch_bounds__(128, 4) floors_packed_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons,
                                                               unsigned long long *counters, unsigned char *zflags,
                                                               const LazyGhost lazy) {
  const Lay L(cfg);
  long long base;
  unsigned mask;
  if (!flagged_cells(zflags, (long long)pack.nblocks * L.ncell / FLAG_VEC, MODE == 1 ? ZF_FLOOR : ZF_STALE, base, mask)) return;
  const int b = (int)(base / L.ncell);
  const int o0 = (int)(base - (long long)b * L.ncell);
  for (int c = 0; c < FLAG_VEC; ++c) { // the whole block is the region (ApplyFloors runs over the entire domain)
    if (!(mask >> c & 1u)) continue;
    const int o = o0 + c;
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    floors_body<MHD, MODE>(pack, cfg, cons, counters, zflags, lazy, L, b, k, j, i, o, zflags[(size_t)b * L.ncell + o]);
  }
}

// Compaction of the flagged zones (the general case: a run where a large share of the domain floors every stage would
// make the 16-zones-per-thread walk above serial): one pass over the flag bytes appends the linear index of every flagged
// zone to a device queue (16 flags per thread, one atomic per warp), then floors_queue_kernel / fixup_queue_kernel handle
// ONE zone per thread from the queue — the cost of the pass is that of con2prim over the same number of zones.
__global__ void __launch_bounds__(256) flag_compact_kernel(const unsigned char *zflags, long long nvec_total, unsigned char bit,
                                                           const RedoQueue q) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned mask = 0;
  if (t < nvec_total) {
    const uint4 w = __ldg(reinterpret_cast<const uint4 *>(zflags) + t);
    const unsigned rep = 0x01010101u * bit;
    if (((w.x | w.y | w.z | w.w) & rep) != 0) {
      const unsigned words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int qq = 0; qq < 4; ++qq)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if ((words[qq] >> (8 * c)) & bit) mask |= 1u << (4 * qq + c);
    }
  }
  const int mine = __popc(mask);
  int incl = mine; // inclusive prefix sum over the warp
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(q.count, (unsigned)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  unsigned at = base + (unsigned)(incl - mine);
  for (int c = 0; c < FLAG_VEC; ++c)
    if (mask >> c & 1u) q.list[at++] = (unsigned)(t * FLAG_VEC + c);
}
template <bool MHD, int MODE>
__global__ void __launch_bounds__(128, 4) floors_queue_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, unsigned long long *counters,
                                                              unsigned char *zflags, const LazyGhost lazy, const RedoQueue q) {
  const Lay L(cfg);
  const unsigned n = *q.count;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned t = q.list[e];
    const int b = (int)(t / (unsigned)L.ncell), o = (int)(t % (unsigned)L.ncell);
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    floors_body<MHD, MODE>(pack, cfg, cons, counters, zflags, lazy, L, b, k, j, i, o, zflags[(size_t)b * L.ncell + o]);
  }
}
template <bool MHD, int MODE>
__global__ void __launch_bounds__(128, 4) fixup_queue_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const Region reg,
                                                             unsigned char *zflags, const RedoQueue q) {
  const Lay L(cfg);
  const unsigned n = *q.count;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned t = q.list[e];
    const int b = (int)(t / (unsigned)L.ncell), o = (int)(t % (unsigned)L.ncell);
    const int i = o % L.ni, r = o / L.ni, j = r % L.nj, k = r / L.nj;
    if (i < reg.i0 || i > reg.i1 || j < reg.j0 || j > reg.j1 || k < reg.k0 || k > reg.k1) continue;
    fixup_body<MHD, MODE>(pack, cfg, cons, zflags, L, b, k, j, i, o);
  }
}

// (Re)build the zone flags from the current primitives and fail field: needed once before the first
// flag-driven stage and whenever something outside the sequencer has changed p.* or fail.
template <bool MHD>
__global__ void __launch_bounds__(128, 4) flags_scan_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Region reg,
                                                            unsigned char *zflags) {
  constexpr int NP = NC<MHD>::nprim;
  const Lay L(cfg);
  int b, k, j, i;
  if (!zone_of_thread(reg, pack.nblocks, b, k, j, i)) return;
  const int o = L.idx(k, j, i);
  const pb_block_t &blk = pack.blocks[b];
  unsigned char f = pack.fail[(size_t)b * L.ncell + o] == 0.0 ? ZF_FAIL : 0;
  if (cfg.enable_floors) {
    const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
    const GeomTab gt(cfg, pack.geom, blk.geom_slot);
    PointGeom g;
    g.load(gt, 0, gt.cell(k, j, i));
    Real P[NP];
    load_comps<NP>(pack.prim + (size_t)b * NP * L.ncell + o, L.ncell, P);
    if (floors_needed<MHD>(cfg, fc, g, P)) f |= ZF_STALE;
  }
  zflags[(size_t)b * L.ncell + o] = f;
}

// ------------------------------------------------------------------ PrimToCons
// ref: src/fluid/fluid.cpp:368-449
template <bool MHD>
PB_DEV void p2c_zone_global(const pb_cfg_t &cfg, const PointGeom &g, const ZonePtr<MHD> &z, int ndim) {
  constexpr int NP = NC<MHD>::nprim, prs = NC<MHD>::prs, gm1 = NC<MHD>::gm1;
  const int N = z.ncell;
  Real P[NP];
  load_comps<NP>(z.prim, N, P);
  Real S[3], bcons[3], sig[3], bp[3] = {0, 0, 0}, D, tau;
  const Real vel[3] = {P[PVEL], P[PVEL + 1], P[PVEL + 2]};
  if (MHD)
    for (int a = 0; a < 3; ++a) bp[a] = P[PB_ + a];
  p2c<true>(P[PRHO], vel, bp, P[PENG], P[prs], P[gm1], g, D, S, bcons, tau, sig);
  z.cons[(size_t)CRHO * N] = D;
  for (int a = 0; a < 3; ++a) z.cons[(size_t)(CMOM + a) * N] = S[a];
  z.cons[(size_t)CENG * N] = tau;
  if (MHD)
    for (int a = 0; a < 3; ++a) z.cons[(size_t)(CB + a) * N] = bcons[a];
  for (int a = 0; a < ndim; ++a) z.csig[(size_t)a * N] = sig[a];
}

template <bool MHD>
__global__ void p2c_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const Region reg) {
  const Lay L(cfg);
  int b, k, j, i;
  if (!zone_of_thread(reg, pack.nblocks, b, k, j, i)) return;
  const int o = L.idx(k, j, i);
  const pb_block_t &blk = pack.blocks[b];
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  PointGeom g;
  g.load(gt, 0, gt.cell(k, j, i));
  const ZonePtr<MHD> z(pack, cons, L, b, o);
  p2c_zone_global<MHD>(cfg, g, z, L.ndim);
}

// ------------------------------------------------------------------ physical boundaries
// Ghost fill (ref: src/phoebus_boundaries/phoebus_boundaries.cpp:44-270) + ConvertBoundaryConditions
// (ref: :303-448) in ONE pass. The reference applies the faces ix1,ox1,ix2,ox2,ix3,ox3 one after the
// other over the entire transverse extent, so an edge/corner ghost zone ends up holding the value
// its LAST face copied from a zone that an EARLIER face had filled. Each thread here composes that
// chain for its own ghost zone (x3 map, then x2, then x1 to find the source; factors applied in the
// reference's order x1, x2, x3), which makes the fill race-free and single-launch. The conversion
// (inflow check, p2c of the prim-filled ghosts) is pointwise and follows in the same thread.
template <bool MHD>
// what: bit 0 = ghost fill, bit 1 = ConvertBoundaryConditions (pb_apply_bcs does both; Mesh::Initialize's
// CommunicateBoundaries fills without converting, external/parthenon/src/mesh/mesh.cpp:692-706).
__global__ void __launch_bounds__(128, 6) bc_kernel(const pb_pack_t pack, const pb_cfg_t cfg, Real *cons, const int nlist,
                                                    unsigned long long *err_count, const int what) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  const Lay L(cfg);
  const Shell sh(L);
  // threads enumerate the ghost shell of the blocks that touch a mesh boundary
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)sh.n * nlist) return;
  const int b = pack.bc_blocks ? pack.bc_blocks[t / sh.n] : (int)(t / sh.n);
  int k, j, i;
  sh.cell(L, (int)(t % sh.n), k, j, i);
  const pb_block_t &blk = pack.blocks[b];
  const int lo[3] = {L.is, L.js, L.ks}, hi[3] = {L.ie, L.je, L.ke};
  const int idx[3] = {i, j, k};
  // which face region (if any) contains this zone in each direction
  int face[3], nphys = 0, nconv = 0;
  for (int d = 0; d < 3; ++d) {
    face[d] = -1;
    if (d >= L.ndim) continue;
    if (idx[d] < lo[d]) face[d] = 2 * d;
    else if (idx[d] > hi[d]) face[d] = 2 * d + 1;
    if (face[d] >= 0) {
      if (blk.bc[face[d]] > 0) nphys++;
      if (blk.bc[face[d]] != 0) nconv++;
    }
  }
  if (nconv == 0) return;
  const int N = L.ncell;
  const int o = L.idx(k, j, i);
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  const ZonePtr<MHD> z(pack, cons, L, b, o);
  if (nphys > 0 && (what & 1)) {
    // source chain: cur[3] = zone read by the x3 step, cur[2] by the x2 step, ...
    int c3[3] = {i, j, k}, c2[3], c1[3], c0[3];
    auto map = [&](int d, const int in[3], int out[3]) {
      out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
      const int f = face[d];
      if (f < 0 || blk.bc[f] <= 0) return false;
      // membership is decided on the ORIGINAL zone: maps only change coordinate d
      const int outer = f & 1;
      const int ref = outer ? hi[d] : lo[d];
      if (blk.bc[f] == PB_BC_OUTFLOW) out[d] = ref;
      else out[d] = 2 * ref + (outer ? 1 : -1) - in[d];
      return true;
    };
    const bool m3 = map(2, c3, c2);
    const bool m2 = map(1, c2, c1);
    const bool m1 = map(0, c1, c0);
    const int src = L.idx(c0[2], c0[1], c0[0]);
    const bool act[3] = {m1, m2, m3};
    const int *dstc[3] = {c1, c2, c3}; // destination zone of the step in direction d
    const int *srcc[3] = {c0, c1, c2};
    if (cfg.bc_prim) {
      Real P[NP];
      load_comps<NP>(pack.prim + (size_t)b * NP * N + src, N, P);
      for (int d = 0; d < 3; ++d) {
        if (!act[d]) continue;
        const int bc = blk.bc[face[d]];
        for (int n = 0; n < NP; ++n) {
          Real sgn = 1.0;
          if (bc == PB_BC_REFLECT) {
            if ((n >= PVEL && n < PVEL + 3 && n - PVEL == d) || (MHD && n >= PB_ && n < PB_ + 3 && n - PB_ == d)) sgn = -1.0;
          } else if (bc == PB_BC_POLAR) {
            if (n == PVEL + 1 || (MHD && n == PB_ + 1)) sgn = -1.0;
          }
          P[n] = sgn * P[n];
        }
      }
      store_comps<NP>(z.prim, N, P);
      *z.temp = pack.temp[(size_t)b * N + src];
      *z.mu = pack.c2p_mu[(size_t)b * N + src];
      *z.fail = pack.fail[(size_t)b * N + src];
      // cons are copied too (no sqrt(gamma) rescale in primitive mode) but overwritten by the p2c of the conversion
      // (every physical face is "!= block"): only a fill WITHOUT conversion has to materialise them
      if (!(what & 2)) {
        Real U[NCn];
        load_comps<NCn>(cons + (size_t)b * NCn * N + src, N, U);
        for (int d = 0; d < 3; ++d) {
          if (!act[d] || blk.bc[face[d]] != PB_BC_REFLECT) continue;
          U[CMOM + d] = -1.0 * U[CMOM + d];
          if (MHD) U[CB + d] = -1.0 * U[CB + d];
        }
        store_comps<NCn>(z.cons, N, U);
      }
    } else {
      Real U[NCn];
      load_comps<NCn>(cons + (size_t)b * NCn * N + src, N, U);
      for (int d = 0; d < 3; ++d) {
        if (!act[d]) continue;
        const int bc = blk.bc[face[d]];
        const Real gd = gt.at(0, G_DETGAM, gt.cell(dstc[d][2], dstc[d][1], dstc[d][0]));
        const Real gs = gt.at(0, G_DETGAM, gt.cell(srcc[d][2], srcc[d][1], srcc[d][0]));
        const Real gratio = ratio(gd, gs);
        for (int n = 0; n < NCn; ++n) {
          Real sgn = 1.0;
          if (bc == PB_BC_REFLECT && ((n >= CMOM && n < CMOM + 3 && n - CMOM == d) || (MHD && n >= CB && n - CB == d))) sgn = -1.0;
          if (bc == PB_BC_OUTFLOW && d == 0) U[n] = gratio * U[n];
          else U[n] = sgn * gratio * U[n];
        }
      }
      store_comps<NCn>(z.cons, N, U);
    }
  }
  // ---- ConvertBoundaryConditions (pointwise)
  if (!(what & 2)) return;
  PointGeom g;
  g.load(gt, 0, gt.cell(k, j, i));
  bool did_p2c = false;
  if (cfg.ox1_inflow_check && blk.bc[1] != 0 && face[0] == 1) {
    if (!cfg.bc_prim) {
      const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
      Real U[NCn], P[NP];
      load_comps<NCn>(z.cons, N, U);
      const bool ok = c2p_zone_global<MHD, 0>(cfg, fc, g, z, L.ndim, U, P, err_count);
      *z.fail = ok ? 1.0 : 0.0;
    }
    Real vcon[3] = {z.prim[(size_t)PVEL * N], z.prim[(size_t)(PVEL + 1) * N], z.prim[(size_t)(PVEL + 2) * N]};
    Real W = lorentz_factor(vcon, g);
    const Real ucon1 = vcon[0] - W * g.beta[0] / g.alpha;
    if (ucon1 < 0) {
      for (int a = 0; a < 3; ++a) vcon[a] /= W;
      vcon[0] = g.beta[0] / g.alpha;
      Real vsq = 0.;
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) vsq += g.g4(a + 1, c + 1) * vcon[a] * vcon[c];
      W = 1. / sqrt(1. - vsq);
      for (int a = 0; a < 3; ++a) z.prim[(size_t)(PVEL + a) * N] = W * vcon[a];
    }
    p2c_zone_global<MHD>(cfg, g, z, L.ndim);
    did_p2c = true;
  }
  if (cfg.bc_prim && !did_p2c) p2c_zone_global<MHD>(cfg, g, z, L.ndim);
}

// ------------------------------------------------------------------ EstimateTimestep
// ref: src/fluid/fluid.cpp:899-944 (without cfl); positive doubles order like their bit patterns
__global__ void dt_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Region reg, int has_face_speeds,
                          Real *dt_out) {
  const Lay L(cfg);
  int b, k, j, i;
  Real my = DBL_MAX;
  if (zone_of_thread(reg, pack.nblocks, b, k, j, i)) {
    const int o = L.idx(k, j, i);
    const int N = L.ncell;
    const int st[3] = {1, L.ni, L.ni * L.nj};
    const pb_block_t &blk = pack.blocks[b];
    const Real *cs = pack.csig + (size_t)b * L.ndim * N + o;
    const Real *fs = pack.fsig + (size_t)b * L.ndim * N + o;
    Real ldt = 0.0;
    for (int d = 0; d < L.ndim; d++) {
      Real max_s = cs[(size_t)d * N];
      if (has_face_speeds) max_s = dmax(max_s, dmax(fs[(size_t)d * N], fs[(size_t)d * N + st[d]]));
      ldt += max_s / blk.dx[d];
    }
    my = 1.0 / (ldt + 1.e-50);
  }
  for (int off = 16; off > 0; off >>= 1) {
    const Real other = __shfl_xor_sync(0xffffffffu, my, off);
    my = dmin(my, other);
  }
  __shared__ Real s_min[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_min[warp] = my;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    my = lane < nw ? s_min[lane] : DBL_MAX;
    for (int off = 16; off > 0; off >>= 1) my = dmin(my, __shfl_xor_sync(0xffffffffu, my, off));
    if (lane == 0) atomicMin(reinterpret_cast<unsigned long long *>(dt_out), (unsigned long long)__double_as_longlong(my));
  }
}

// ------------------------------------------------------------------ twins of the ghost plan (ghost_plan_host.cpp)
// An x1 ghost zone whose tables differ from the neighbour's in the last place is solved from the neighbour's INPUTS (what
// the exchange would have put there: conserved state and c2p_mu; the previous primitives / temperature only on the paths
// where the reference leaves them untouched) with its own tables, exactly as c2p_thread solves a zone, BEFORE any con2prim
// of the stage overwrites those inputs: the kernel is launched with the pre-con2prim ghost copy. The result waits in the
// side buffer; the deferred copy (halo_kernels.cu, phase 2) moves it into the ghost zone together with its row.
PB_DEV void twin_decode(const Lay &L, unsigned z, int &b, int &k, int &j, int &i) {
  b = (int)(z / (unsigned)L.ncell);
  int rem = (int)(z - (unsigned)b * (unsigned)L.ncell);
  i = rem % L.ni;
  rem /= L.ni;
  j = rem % L.nj;
  k = rem / L.nj;
}
template <bool MHD>
__global__ void __launch_bounds__(128, 4) twin_kernel(const pb_pack_t pack, const pb_cfg_t cfg, const Real *__restrict__ cons,
                                                      unsigned long long *err_count) {
  constexpr int NP = NC<MHD>::nprim, NCn = NC<MHD>::ncons;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= pack.n_twin) return;
  const Lay L(cfg);
  const size_t N = (size_t)L.ncell;
  int b, k, j, i;
  twin_decode(L, __ldg(pack.twin_zones + e), b, k, j, i);
  const int o1 = i < L.is ? -1 : (i > L.ie ? 1 : 0);
  const int o2 = j < L.js ? -1 : (j > L.je ? 1 : 0);
  const int o3 = k < L.ks ? -1 : (k > L.ke ? 1 : 0);
  const pb_block_t &blk = pack.blocks[b];
  const size_t nb = (size_t)blk.nbr[(o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)];
  const size_t so = (size_t)L.idx(k - o3 * cfg.nx[2], j - o2 * cfg.nx[1], i - o1 * cfg.nx[0]);
  const GeomTab gt(cfg, pack.geom, blk.geom_slot);
  PointGeom g;
  g.load(gt, 0, gt.cell(k, j, i));
  const FloorCtx fc = cell_floor_ctx(pack, blk, L, k, j, i);
  const TwinBuf tb(pack, L, cfg);
  const int en = tb.entry(b, k, j, i);
  const ZonePtr<MHD> z(tb, en);
  const size_t S = (size_t)tb.S;
  Real U[NCn], P[NP];
#pragma unroll
  for (int n = 0; n < NCn; ++n) U[n] = cons[(nb * NCn + n) * N + so];
  *z.mu = pack.c2p_mu[nb * N + so];
#pragma unroll
  for (int n = 0; n < NCn; ++n) z.cons[n * S] = U[n];
  const bool nan_in = isnan(U[CRHO]) + isnan(U[CMOM]) + isnan(U[CENG]) > 0;
  // a NaN with a private payload marks "the solve wrote nothing here" (NaN guard, unbracketed root)
  const long long unwritten = 0x7ff8c2b0b2000001ll;
  *z.temp = __longlong_as_double(unwritten);
  const bool ok = c2p_zone_global<MHD, 1>(cfg, fc, g, z, L.ndim, U, P, err_count);
  if (__double_as_longlong(*z.temp) == unwritten) {
    // rare: the reference's ghost zone keeps the exchanged primitives and temperature and its own previous signal speeds
    // (they are not exchanged); an unbracketed root has written the magnetic field components only
#pragma unroll
    for (int n = 0; n < NP; ++n) {
      if (MHD && !nan_in && n >= PB_ && n < PB_ + 3) continue;
      z.prim[n * S] = pack.prim[(nb * NP + n) * N + so];
    }
    *z.temp = pack.temp[nb * N + so];
    const size_t o = (size_t)L.idx(k, j, i);
    for (int d = 0; d < L.ndim; ++d) z.csig[d * S] = pack.csig[((size_t)b * L.ndim + d) * N + o];
  }
  *z.fail = ok ? 1.0 : 0.0;
  unsigned char f = ok ? 0 : ZF_FAIL;
  if (cfg.enable_floors) {
    if (!ok) load_comps<NP>(z.prim, tb.S, P);
    if (floors_needed<MHD>(cfg, fc, g, P)) f |= ZF_FLOOR;
  }
  tb.flags[en] = f;
}

static unsigned grid_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

static Region entire(const Lay &L) { return Region{0, L.ni - 1, 0, L.nj - 1, 0, L.nk - 1}; }
static Region interior(const Lay &L) { return Region{L.is, L.ie, L.js, L.je, L.ks, L.ke}; }

static unsigned long long *g_err_count = nullptr; // device counter of unbracketed roots
unsigned long long *err_counter() {
  if (!g_err_count) {
    if (cudaMalloc(&g_err_count, sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    cudaMemset(g_err_count, 0, sizeof(unsigned long long));
  }
  return g_err_count;
}

} // namespace pb

using namespace pb;

namespace pb {
// library-owned scratch of the split con2prim (grown on demand; stream-ordered use like the redo queue)
static C2PWorkspace g_c2p_ws = {nullptr, nullptr, nullptr, 0, nullptr};
static size_t g_c2p_ws_cells = 0;
static unsigned long long *g_c2p_stats = nullptr; // allocated by pb_c2p_stats(enable)
static bool g_c2p_stats_on = false;
static int c2p_workspace(size_t cells, size_t nthr, C2PWorkspace &out) {
  if (!g_c2p_ws.slow_count) {
    if (cudaMalloc(&g_c2p_ws.slow_count, 16) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "c2p workspace");
    cudaMemset(g_c2p_ws.slow_count, 0, 16);
  }
  const size_t cap = cells / 4 + 1024; // parked zones beyond the capacity simply finish in the solve kernel
  if (cells > g_c2p_ws_cells) {
    cudaDeviceSynchronize();
    if (g_c2p_ws.status) cudaFree(g_c2p_ws.status);
    if (g_c2p_ws.slow) cudaFree(g_c2p_ws.slow);
    g_c2p_ws.status = nullptr;
    g_c2p_ws.slow = nullptr;
    if (cudaMalloc(&g_c2p_ws.status, cells) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "c2p status bytes");
    if (cudaMalloc(&g_c2p_ws.slow, cap * sizeof(SlowEntry)) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "c2p slow queue");
    g_c2p_ws.slow_cap = (unsigned)cap;
    g_c2p_ws_cells = cells;
  }
  out = g_c2p_ws;
  out.stats = g_c2p_stats_on ? g_c2p_stats : nullptr;
  (void)nthr;
  return PB_OK;
}
int presize_block_flags(int nblocks); // flux_split_kernels.cu
// everything the sequencer allocates lazily, sized for this pack up front (CUDA-graph capture may not allocate)
int presize_workspaces(const pb_pack_t *pack, const pb_cfg_t *cfg, cudaStream_t stream) {
  {
    const int rcb = presize_block_flags(pack->nblocks);
    if (rcb) return rcb;
  }
  const Lay L(*cfg);
  const size_t cells = (size_t)pack->nblocks * L.ncell;
  RedoQueue rq;
  int rc = redo_queue(cells, rq, stream);
  if (rc) return rc;
  C2PWorkspace ws;
  if ((rc = c2p_workspace(cells, cells, ws))) return rc;
  if (!err_counter()) return set_cuda_error(cudaGetLastError(), "error counter");
  return PB_OK;
}
// sequencer-internal launchers (abi_common.cu)
static int launch_c2p_work(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const C2PWork &work, long long nthr,
                           unsigned char *zflags, cudaStream_t s) {
  if (nthr <= 0) return PB_OK;
  const int T = zone_threads();
  const unsigned grid = grid_for(nthr, T);
  // option c2p_arith / PB_C2P_ARITH: queue (default) = ArithFast, rejected zones re-solved by c2p_redo_kernel; exact =
  // the compiler's divide / sqrt everywhere; inline = ArithFast with the exact re-solve inside the kernel
  const int mode = c2p_arith_mode();
  RedoQueue rq{nullptr, nullptr, nullptr};
  if (mode == 2) {
    if ((long long)pack->nblocks * Lay(*cfg).ncell >= (1ll << 32)) return set_error(PB_EINVAL, "pack too large for the 32-bit redo queue");
    int rc = redo_queue((size_t)nthr, rq, s);
    if (rc) return rc;
    cudaMemsetAsync(rq.count, 0, sizeof(unsigned), s);
  }
  if (mode == 3) {
    if ((long long)pack->nblocks * Lay(*cfg).ncell >= (1ll << 31)) return set_error(PB_EINVAL, "pack too large for the 31-bit redo queue");
    int rc = redo_queue((size_t)nthr, rq, s);
    if (rc) return rc;
    C2PWorkspace ws;
    if ((rc = c2p_workspace((size_t)pack->nblocks * Lay(*cfg).ncell, (size_t)nthr, ws))) return rc;
    cudaMemsetAsync(rq.count, 0, sizeof(unsigned), s);
    cudaMemsetAsync(ws.slow_count, 0, sizeof(unsigned), s);
    const int nsm = sm_count();
    if (cfg->mhd) {
      c2p_solve_kernel<true><<<grid, T, 0, s>>>(*pack, *cfg, cons, work, nthr, ws, rq);
      c2p_slow_kernel<true><<<nsm * PB_C2P_SOLVE_MINB, T, 0, s>>>(*pack, *cfg, cons, ws, rq);
      c2p_unwrap_kernel<true><<<grid, T, 0, s>>>(*pack, *cfg, cons, work, nthr, ws, err_counter(), zflags, rq);
      c2p_split_redo_kernel<true><<<nsm * 4, 128, 0, s>>>(*pack, *cfg, cons, err_counter(), zflags, rq, ws.stats);
    } else {
      c2p_solve_kernel<false><<<grid, T, 0, s>>>(*pack, *cfg, cons, work, nthr, ws, rq);
      c2p_slow_kernel<false><<<nsm * PB_C2P_SOLVE_MINB, T, 0, s>>>(*pack, *cfg, cons, ws, rq);
      c2p_unwrap_kernel<false><<<grid, T, 0, s>>>(*pack, *cfg, cons, work, nthr, ws, err_counter(), zflags, rq);
      c2p_split_redo_kernel<false><<<nsm * 4, 128, 0, s>>>(*pack, *cfg, cons, err_counter(), zflags, rq, ws.stats);
    }
    return check_launch("c2p_solve_kernel");
  }
  if (mode == 2) {
    if ((long long)pack->nblocks * Lay(*cfg).ncell >= (1ll << 32)) return set_error(PB_EINVAL, "pack too large for the 32-bit redo queue");
    int rc = redo_queue((size_t)nthr, rq, s);
    if (rc) return rc;
    C2PWorkspace ws;
    if ((rc = c2p_workspace((size_t)pack->nblocks * Lay(*cfg).ncell, (size_t)nthr, ws))) return rc;
    cudaMemsetAsync(rq.count, 0, sizeof(unsigned), s);
    cudaMemsetAsync(ws.slow_count, 0, sizeof(unsigned), s);
    const int nsm = sm_count();
    if (cfg->mhd) {
      c2p_park_kernel<true><<<grid, T, 0, s>>>(*pack, *cfg, cons, work, err_counter(), zflags, rq, ws);
      c2p_finish_kernel<true><<<nsm * 2, T, 0, s>>>(*pack, *cfg, cons, zflags, rq, ws);
      c2p_redo_kernel<true><<<nsm * 4, 128, 0, s>>>(*pack, *cfg, cons, err_counter(), zflags, rq);
    } else {
      c2p_park_kernel<false><<<grid, T, 0, s>>>(*pack, *cfg, cons, work, err_counter(), zflags, rq, ws);
      c2p_finish_kernel<false><<<nsm * 2, T, 0, s>>>(*pack, *cfg, cons, zflags, rq, ws);
      c2p_redo_kernel<false><<<nsm * 4, 128, 0, s>>>(*pack, *cfg, cons, err_counter(), zflags, rq);
    }
    return check_launch("c2p_park_kernel");
  }
#define PB_C2P(M, A) c2p_kernel<M, A><<<grid, T, 0, s>>>(*pack, *cfg, cons, work, err_counter(), zflags, rq)
  if (cfg->mhd) {
    if (mode == 0) PB_C2P(true, 0); else if (mode == 1) PB_C2P(true, 1); else PB_C2P(true, 2);
  } else {
    if (mode == 0) PB_C2P(false, 0); else if (mode == 1) PB_C2P(false, 1); else PB_C2P(false, 2);
  }
#undef PB_C2P
  if (mode == 2) {
    if (cfg->mhd) c2p_redo_kernel<true><<<sm_count() * 4, 128, 0, s>>>(*pack, *cfg, cons, err_counter(), zflags, rq);
    else c2p_redo_kernel<false><<<sm_count() * 4, 128, 0, s>>>(*pack, *cfg, cons, err_counter(), zflags, rq);
  }
  return check_launch("c2p_kernel");
}
int launch_c2p(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int ib[2], const int jb[2], const int kb[2],
               unsigned char *zflags, cudaStream_t s) {
  const Region reg{ib[0], ib[1], jb[0], jb[1], kb[0], kb[1]};
  if (reg.n() <= 0) return PB_OK;
  const C2PWork work{reg, nullptr, pack->nblocks, 0, nullptr, 0};
  return launch_c2p_work(pack, cfg, cons, work, reg.n() * pack->nblocks, zflags, s);
}
// ghost plan: c2p on the interior zones of a list of blocks (NULL: all blocks) and, with_ghosts, on the ghost zones the plan lists
int launch_c2p_plan(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int *blocks, int nlist, bool with_ghosts,
                    unsigned char *zflags, cudaStream_t s) {
  const Lay L(*cfg);
  const Region reg = interior(L);
  const long long nz = with_ghosts ? pack->n_c2p_zones : 0;
  const long long nthr = reg.n() * nlist + nz;
  if (nthr <= 0) return PB_OK;
  const C2PWork work{reg, blocks, nlist, 0, pack->c2p_zones, nz};
  return launch_c2p_work(pack, cfg, cons, work, nthr, zflags, s);
}
// launched with the pre-con2prim ghost copy (halo_kernels.cu): the twins read their neighbours' inputs before any con2prim runs
int launch_twins(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, cudaStream_t s) {
  if (pack->n_twin <= 0 || !pack->twin_zones) return PB_OK;
  if (!pack->twin_buf || !pack->twin_flags) return set_error(PB_EINVAL, "ghost plan: twin_zones without twin_buf / twin_flags");
  PB_NEED("ghost plan twins", PB_F(pack, temp), PB_F(pack, c2p_mu), PB_F(pack, csig));
  const unsigned grid = grid_for(pack->n_twin, 128);
  if (cfg->mhd) twin_kernel<true><<<grid, 128, 0, s>>>(*pack, *cfg, cons, err_counter());
  else twin_kernel<false><<<grid, 128, 0, s>>>(*pack, *cfg, cons, err_counter());
  return check_launch("twin_kernel");
}
// c2p on a list of blocks: which = 0 interior zones, 1 ghost shell, 2 entire block
int launch_c2p_list(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int *blocks, int nlist, int which,
                    unsigned char *zflags, cudaStream_t s) {
  if (nlist <= 0) return PB_OK;
  const Lay L(*cfg);
  const Shell sh(L);
  const Region reg = which == 0 ? interior(L) : entire(L);
  const C2PWork work{reg, blocks, nlist, which == 1 ? 1 : 0, nullptr, 0};
  const long long nthr = (which == 1 ? (long long)sh.n : reg.n()) * nlist;
  return launch_c2p_work(pack, cfg, cons, work, nthr, zflags, s);
}
int launch_fixup(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, int mode, unsigned char *zflags, cudaStream_t s) {
  if (!cfg->enable_c2p_fixup) return PB_OK; // ref: fixup_c2p.cpp:60-62
  const Lay L(*cfg);
  const Region reg = interior(L);
  const int T = zone_threads();
  const unsigned grid = grid_for(reg.n() * pack->nblocks, T);
  if (mode != 0 && L.ncell % FLAG_VEC == 0 && (reinterpret_cast<size_t>(zflags) & 15) == 0 &&
      (long long)pack->nblocks * L.ncell < (1ll << 32)) {
    const long long nvec = (long long)pack->nblocks * L.ncell / FLAG_VEC;
    if (flag_compaction_enabled()) {
      RedoQueue q;
      int rc = redo_queue((size_t)pack->nblocks * L.ncell, q, s);
      if (rc) return rc;
      cudaMemsetAsync(q.count, 0, sizeof(unsigned), s);
      flag_compact_kernel<<<grid_for(nvec, 256), 256, 0, s>>>(zflags, nvec, ZF_FAIL, q);
      const unsigned gq = (unsigned)(sm_count() * 8);
#define PB_FIXQ(M, MODE) fixup_queue_kernel<M, MODE><<<gq, T, 0, s>>>(*pack, *cfg, cons, reg, zflags, q)
      if (cfg->mhd) {
        if (mode == 1) PB_FIXQ(true, 1); else PB_FIXQ(true, 2);
      } else {
        if (mode == 1) PB_FIXQ(false, 1); else PB_FIXQ(false, 2);
      }
#undef PB_FIXQ
      return check_launch("fixup_queue_kernel");
    }
    const unsigned gp = grid_for(nvec, T);
#define PB_FIXP(M, MODE) c2p_fixup_packed_kernel<M, MODE><<<gp, T, 0, s>>>(*pack, *cfg, cons, reg, zflags)
    if (cfg->mhd) {
      if (mode == 1) PB_FIXP(true, 1); else PB_FIXP(true, 2);
    } else {
      if (mode == 1) PB_FIXP(false, 1); else PB_FIXP(false, 2);
    }
#undef PB_FIXP
    return check_launch("c2p_fixup_packed_kernel");
  }
#define PB_FIX(M, MODE) c2p_fixup_kernel<M, MODE><<<grid, T, 0, s>>>(*pack, *cfg, cons, reg, zflags)
  if (cfg->mhd) {
    if (mode == 0) PB_FIX(true, 0); else if (mode == 1) PB_FIX(true, 1); else PB_FIX(true, 2);
  } else {
    if (mode == 0) PB_FIX(false, 0); else if (mode == 1) PB_FIX(false, 1); else PB_FIX(false, 2);
  }
#undef PB_FIX
  return check_launch("c2p_fixup_kernel");
}
int launch_floors(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, unsigned long long *counters, int mode,
                  unsigned char *zflags, const double *lazy_u0, const double *lazy_ubase, double beta, cudaStream_t s) {
  if (!cfg->enable_floors) return PB_OK; // ref: fixup.cpp:293-295
  const Lay L(*cfg);
  const Region reg = entire(L);
  const int T = zone_threads();
  const unsigned grid = grid_for(reg.n() * pack->nblocks, T);
  const LazyGhost lazy{lazy_u0, lazy_ubase, beta};
  if (mode != 0 && L.ncell % FLAG_VEC == 0 && (reinterpret_cast<size_t>(zflags) & 15) == 0 &&
      (long long)pack->nblocks * L.ncell < (1ll << 32)) {
    const long long nvec = (long long)pack->nblocks * L.ncell / FLAG_VEC;
    if (flag_compaction_enabled()) {
      RedoQueue q;
      int rc = redo_queue((size_t)pack->nblocks * L.ncell, q, s);
      if (rc) return rc;
      cudaMemsetAsync(q.count, 0, sizeof(unsigned), s);
      flag_compact_kernel<<<grid_for(nvec, 256), 256, 0, s>>>(zflags, nvec, mode == 1 ? ZF_FLOOR : ZF_STALE, q);
      const unsigned gq = (unsigned)(sm_count() * 8);
#define PB_FLRQ(M, MODE) floors_queue_kernel<M, MODE><<<gq, T, 0, s>>>(*pack, *cfg, cons, counters, zflags, lazy, q)
      if (cfg->mhd) {
        if (mode == 1) PB_FLRQ(true, 1); else PB_FLRQ(true, 2);
      } else {
        if (mode == 1) PB_FLRQ(false, 1); else PB_FLRQ(false, 2);
      }
#undef PB_FLRQ
      return check_launch("floors_queue_kernel");
    }
    const unsigned gp = grid_for(nvec, T);
#define PB_FLRP(M, MODE) floors_packed_kernel<M, MODE><<<gp, T, 0, s>>>(*pack, *cfg, cons, counters, zflags, lazy)
    if (cfg->mhd) {
      if (mode == 1) PB_FLRP(true, 1); else PB_FLRP(true, 2);
    } else {
      if (mode == 1) PB_FLRP(false, 1); else PB_FLRP(false, 2);
    }
#undef PB_FLRP
    return check_launch("floors_packed_kernel");
  }
#define PB_FLR(M, MODE) floors_kernel<M, MODE><<<grid, T, 0, s>>>(*pack, *cfg, cons, reg, counters, zflags, lazy)
  if (cfg->mhd) {
    if (mode == 0) PB_FLR(true, 0); else if (mode == 1) PB_FLR(true, 1); else PB_FLR(true, 2);
  } else {
    if (mode == 0) PB_FLR(false, 0); else if (mode == 1) PB_FLR(false, 1); else PB_FLR(false, 2);
  }
#undef PB_FLR
  return check_launch("floors_kernel");
}
} // namespace pb

extern "C" int pb_cons_to_prim(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int ib[2],
                               const int jb[2], const int kb[2], void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons) return set_error(PB_EINVAL, "cons is NULL");
  PB_NEED("pb_cons_to_prim", PB_F(pack, temp), PB_F(pack, c2p_mu), PB_F(pack, fail), PB_F(pack, csig));
  return launch_c2p(pack, cfg, cons, ib, jb, kb, nullptr, (cudaStream_t)stream);
}

extern "C" int pb_c2p_fixup(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons) return set_error(PB_EINVAL, "cons is NULL");
  PB_NEED("pb_c2p_fixup", PB_F(pack, temp), PB_F(pack, fail), PB_F(pack, csig));
  return launch_fixup(pack, cfg, cons, 0, nullptr, (cudaStream_t)stream);
}

extern "C" int pb_apply_floors(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons,
                               unsigned long long *counters, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons) return set_error(PB_EINVAL, "cons is NULL");
  PB_NEED("pb_apply_floors", PB_F(pack, temp), PB_F(pack, c2p_mu), PB_F(pack, csig));
  return launch_floors(pack, cfg, cons, counters, 0, nullptr, nullptr, nullptr, 1.0, (cudaStream_t)stream);
}

extern "C" int pb_scan_zone_flags(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!pack->zflags) return set_error(PB_EINVAL, "pack has no zflags array");
  const Lay L(*cfg);
  const Region reg = entire(L);
  const int T = zone_threads();
  const unsigned grid = grid_for(reg.n() * pack->nblocks, T);
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd) flags_scan_kernel<true><<<grid, T, 0, s>>>(*pack, *cfg, reg, pack->zflags);
  else flags_scan_kernel<false><<<grid, T, 0, s>>>(*pack, *cfg, reg, pack->zflags);
  return check_launch("flags_scan_kernel");
}

extern "C" int pb_prim_to_cons(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int ib[2],
                               const int jb[2], const int kb[2], void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons) return set_error(PB_EINVAL, "cons is NULL");
  PB_NEED("pb_prim_to_cons", PB_F(pack, csig));
  const Region reg{ib[0], ib[1], jb[0], jb[1], kb[0], kb[1]};
  if (reg.n() <= 0) return PB_OK;
  const int T = zone_threads();
  const unsigned grid = grid_for(reg.n() * pack->nblocks, T);
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd) p2c_kernel<true><<<grid, T, 0, s>>>(*pack, *cfg, cons, reg);
  else p2c_kernel<false><<<grid, T, 0, s>>>(*pack, *cfg, cons, reg);
  return check_launch("p2c_kernel");
}

extern "C" int pb_apply_bcs(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, void *stream) {
  return pb_apply_bcs_ex(pack, cfg, cons, 3, stream);
}

extern "C" int pb_apply_bcs_ex(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, int what, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!cons || (what & 3) == 0) return set_error(PB_EINVAL, "pb_apply_bcs_ex: NULL container or empty task mask");
  PB_NEED("pb_apply_bcs", PB_F(pack, temp), PB_F(pack, c2p_mu), PB_F(pack, fail), PB_F(pack, csig));
  const Lay L(*cfg);
  const Shell sh(L);
  const int nlist = pack->bc_blocks ? pack->n_bc_blocks : pack->nblocks;
  if (nlist <= 0 || sh.n <= 0) return PB_OK;
  const int T = zone_threads();
  const unsigned grid = grid_for((long long)sh.n * nlist, T);
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->mhd) bc_kernel<true><<<grid, T, 0, s>>>(*pack, *cfg, cons, nlist, err_counter(), what);
  else bc_kernel<false><<<grid, T, 0, s>>>(*pack, *cfg, cons, nlist, err_counter(), what);
  return check_launch("bc_kernel");
}

extern "C" int pb_estimate_dt(const pb_pack_t *pack, const pb_cfg_t *cfg, int has_face_speeds, double *dt_out,
                              void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!dt_out) return set_error(PB_EINVAL, "pb_estimate_dt: NULL output");
  PB_NEED("pb_estimate_dt", PB_F(pack, csig), PB_F(pack, fsig));
  const Lay L(*cfg);
  const Region reg = interior(L);
  const int T = 256;
  const unsigned grid = grid_for(reg.n() * pack->nblocks, T);
  dt_kernel<<<grid, T, 0, (cudaStream_t)stream>>>(*pack, *cfg, reg, has_face_speeds, dt_out);
  return check_launch("dt_kernel");
}

// Con2prim statistics of the split path (the reference's CON2PRIM_STATISTICS histogram, ref: src/fluid/con2prim_statistics.hpp:22-41,
// con2prim_robust.hpp:377-379): enable != 0 starts counting (and zeroes the counters), enable == 0 stops. out_host (may be NULL):
// 36 counters accumulated since they were last zeroed: [0..31] regula-falsi iterations per solve (31 = 31 or more), [32] solves
// that used a floor / ceiling (conserved variables overwritten), [33] failed zones, [34] zones parked by the warp vote,
// [35] zones unwrapped. Synchronises the device.
extern "C" int pb_c2p_stats(int enable, unsigned long long out_host[36]) {
  if (!g_c2p_stats) {
    if (cudaMalloc(&g_c2p_stats, 36 * sizeof(unsigned long long)) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "pb_c2p_stats");
    cudaMemset(g_c2p_stats, 0, 36 * sizeof(unsigned long long));
  }
  if (out_host && cudaMemcpy(out_host, g_c2p_stats, 36 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return set_cuda_error(cudaGetLastError(), "pb_c2p_stats");
  if (enable && !g_c2p_stats_on) cudaMemset(g_c2p_stats, 0, 36 * sizeof(unsigned long long));
  g_c2p_stats_on = enable != 0;
  return PB_OK;
}

// Unbracketed con2prim roots since the library was loaded (or since the last reset): the reference aborts the run there
// (PARTHENON_FAIL, ref: src/phoebus_utils/root_find.hpp:115-118); here the zone is marked failed and counted, and the
// host layer decides (phoebus_b200.driver raises). Synchronises the device.
extern "C" long long pb_c2p_unbracketed_count(int reset) {
  unsigned long long *c = err_counter();
  if (!c) return set_cuda_error(cudaGetLastError(), "pb_c2p_unbracketed_count");
  unsigned long long v = 0;
  if (cudaMemcpy(&v, c, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "pb_c2p_unbracketed_count");
  if (reset) cudaMemset(c, 0, sizeof v);
  return (long long)v;
}
