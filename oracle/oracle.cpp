// TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's GRMHD hot path.
// See oracle.h for the contract. Plain scalar C++17, no Kokkos, no FMA contraction
// (compile with -ffp-contract=off to mirror the reference's generic x86-64 build,
// SURVEY.md Appendix A.14). Every function cites the reference file:line it restates.
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

namespace {

typedef double Real;

// ---------------------------------------------------------------- robust.hpp:9-41
constexpr Real R_SMALL = 10 * std::numeric_limits<Real>::min();
constexpr Real R_EPS = 10 * std::numeric_limits<Real>::epsilon();
inline Real make_positive(Real v) { return std::max(v, R_EPS); }
inline Real make_bounded(Real v, Real lo, Real hi) {
  return std::min(std::max(v, lo + R_EPS), hi * (1.0 - R_EPS));
}
inline Real ratio(Real a, Real b) {
  const Real sgn = b >= 0 ? 1 : -1;
  return a / (b + sgn * R_SMALL);
}

long long g_errors = 0; // counts conditions on which the reference would PARTHENON_FAIL

// ---------------------------------------------------------------- layout helpers
struct Lay {
  int ndim, ng, ni, nj, nk, is, ie, js, je, ks, ke, dj, dk;
  long ncell;
  int nprim, ncons, mhd;
  explicit Lay(const orc_cfg *c) {
    ndim = c->ndim;
    ng = c->ng;
    mhd = c->mhd;
    ni = c->nx[0] + 2 * ng;
    nj = ndim >= 2 ? c->nx[1] + 2 * ng : 1;
    nk = ndim >= 3 ? c->nx[2] + 2 * ng : 1;
    is = ng;
    ie = ng + c->nx[0] - 1;
    js = ndim >= 2 ? ng : 0;
    je = ndim >= 2 ? ng + c->nx[1] - 1 : 0;
    ks = ndim >= 3 ? ng : 0;
    ke = ndim >= 3 ? ng + c->nx[2] - 1 : 0;
    dj = ndim >= 2 ? 1 : 0;
    dk = ndim >= 3 ? 1 : 0;
    ncell = (long)ni * nj * nk;
    nprim = 7 + 3 * mhd;
    ncons = 5 + 3 * mhd;
  }
  inline long idx(int k, int j, int i) const { return ((long)k * nj + j) * ni + i; }
};
// primitive slots (fluid.cpp:272-286): rho, v1..3, u, [B1..3], P, gamma1
enum { PRHO = 0, PVEL = 1, PENG = 4, PB = 5 };
inline int PPRS(int mhd) { return 5 + 3 * mhd; }
inline int PGM1(int mhd) { return 6 + 3 * mhd; }
// conserved slots: D, S1..3, tau, [B1..3]
enum { CRHO = 0, CMOM = 1, CENG = 4, CB = 5 };

// ---------------------------------------------------------------- geometry tables
// component indices (cached_system.hpp:42-52); Flatten2 from geometry_utils.hpp:186-192
enum { G_ALPHA = 0, G_BCON = 1, G_GCOV = 4, G_GAMCON = 14, G_DETGAM = 20, G_DALPHA = 0, G_DG = 3 };
const int FLAT[4][4] = {{0, 1, 3, 6}, {1, 2, 4, 7}, {3, 4, 5, 8}, {6, 7, 8, 9}};

struct Geom {
  const Real *t;
  int gk, gj, gi;
  long gcell;
  Geom(const orc_cfg *c, const Real *tab) : t(tab), gk(c->gk), gj(c->gj), gi(c->gi) {
    gcell = (long)gk * gj * gi;
  }
  inline long cell(int k, int j, int i) const {
    const int kk = gk == 1 ? 0 : k, jj = gj == 1 ? 0 : j, ii = gi == 1 ? 0 : i;
    return ((long)kk * gj + jj) * gi + ii;
  }
  inline Real at(int loc, int comp, int k, int j, int i) const {
    return t[((long)loc * ORC_NGLOC + comp) * gcell + cell(k, j, i)];
  }
  inline Real der(int comp, int k, int j, int i) const {
    return t[((long)4 * ORC_NGLOC + comp) * gcell + cell(k, j, i)];
  }
  Real Lapse(int loc, int k, int j, int i) const { return at(loc, G_ALPHA, k, j, i); }
  Real DetGamma(int loc, int k, int j, int i) const { return at(loc, G_DETGAM, k, j, i); }
  // cached_system.hpp:371-374
  Real DetG(int loc, int k, int j, int i) const { return Lapse(loc, k, j, i) * DetGamma(loc, k, j, i); }
  void Shift(int loc, int k, int j, int i, Real b[3]) const {
    for (int d = 0; d < 3; ++d) b[d] = at(loc, G_BCON + d, k, j, i);
  }
  void SpacetimeMetric(int loc, int k, int j, int i, Real g[4][4]) const {
    for (int m = 0; m < 4; ++m)
      for (int n = 0; n < 4; ++n) g[m][n] = at(loc, G_GCOV + FLAT[m][n], k, j, i);
  }
  void Metric(int loc, int k, int j, int i, Real g[3][3]) const {
    for (int m = 0; m < 3; ++m)
      for (int n = 0; n < 3; ++n) g[m][n] = at(loc, G_GCOV + FLAT[m + 1][n + 1], k, j, i);
  }
  void MetricInverse(int loc, int k, int j, int i, Real g[3][3]) const {
    for (int m = 0; m < 3; ++m)
      for (int n = 0; n < 3; ++n) g[m][n] = at(loc, G_GAMCON + FLAT[m][n], k, j, i);
  }
  // cached_system.hpp:330-349
  void SpacetimeMetricInverse(int loc, int k, int j, int i, Real g[4][4]) const {
    Real alpha2 = Lapse(loc, k, j, i);
    alpha2 *= alpha2;
    g[0][0] = -ratio(1, alpha2);
    for (int m = 0; m < 3; ++m) g[m + 1][0] = g[0][m + 1] = ratio(at(loc, G_BCON + m, k, j, i), alpha2);
    for (int m = 0; m < 3; ++m)
      for (int n = 0; n < 3; ++n) {
        g[m + 1][n + 1] = at(loc, G_GAMCON + FLAT[m][n], k, j, i);
        g[m + 1][n + 1] -= g[0][m + 1] * g[0][n + 1] * alpha2;
      }
  }
  // cached_system.hpp:410-424
  void MetricDerivative(int k, int j, int i, Real dg[4][4][4]) const {
    for (int m = 0; m < 4; ++m)
      for (int n = 0; n < 4; ++n) {
        dg[m][n][0] = 0.0;
        for (int s = 1; s < 4; ++s) dg[m][n][s] = der(G_DG + 3 * FLAT[m][n] + (s - 1), k, j, i);
      }
  }
  void GradLnAlpha(int k, int j, int i, Real da[4]) const {
    da[0] = 0;
    for (int d = 1; d < 4; ++d) da[d] = der(G_DALPHA + d - 1, k, j, i);
  }
};

// uniform_cartesian.hpp:86-145
inline Real Xc(const orc_cfg *c, int d, int idx) { return c->xmin[d] + (idx + 0.5) * c->dx[d]; }
inline Real Xf(const orc_cfg *c, int d, int idx) { return c->xmin[d] + idx * c->dx[d]; }

// ---------------------------------------------------------------- fixup.hpp:72-147, 182-232
struct Bounds {
  const orc_cfg *c;
  explicit Bounds(const orc_cfg *cc) : c(cc) {}
  void GetFloors(Real x1, Real x2, Real x3, Real &rflr, Real &sflr) const {
    switch (c->floor_type) {
    case 0:
      rflr = c->rho0;
      sflr = c->sie0;
      break;
    case 1:
      rflr = c->rho0 * std::exp(c->rexp * x1);
      sflr = c->sie0 * std::exp(c->sexp * x1);
      break;
    case 2: {
      Real scratch = c->rho0 * std::exp(c->rexp * x1);
      sflr = c->sie0 * std::exp(c->sexp * x1) / std::max(rflr, scratch);
      rflr = scratch;
    } break;
    case 3:
      rflr = c->rho0 * std::min(1.0, std::pow(x1, c->rexp));
      sflr = c->sie0 * std::min(1.0, std::pow(x1, c->sexp));
      break;
    case 4: {
      Real r = std::sqrt(x1 * x1 + x2 * x2 + x3 * x3);
      rflr = c->rho0 * std::min(1.0, std::pow(r, c->rexp));
      sflr = c->sie0 * std::min(1.0, std::pow(r, c->sexp));
    } break;
    default:
      g_errors++;
    }
    rflr = std::max(rflr, c->rho_min_eos);
    sflr = std::max(sflr, c->sie_min_eos);
  }
  void GetCeilings(Real, Real, Real, Real &gmax, Real &smax) const {
    gmax = c->gam_max;
    smax = c->sie_max;
  }
  void GetMHDCeilings(Real, Real, Real, Real &bsqorho, Real &bsqou) const {
    bsqorho = c->bsqorho_max;
    bsqou = c->bsqou_max;
  }
};

// ---------------------------------------------------------------- eos_ideal.hpp:59-137
struct Eos {
  Real gm1, Cv;
  explicit Eos(const orc_cfg *c) : gm1(c->gm1), Cv(c->Cv) {}
  Real T_rho_sie(Real, Real sie) const { return std::max(0.0, sie / Cv); }
  Real P_rho_T(Real rho, Real T) const { return std::max(0.0, gm1 * rho * Cv * T); }
  Real P_rho_sie(Real rho, Real sie) const { return std::max(0.0, gm1 * rho * sie); }
  Real K_rho_T(Real rho, Real T) const { return std::max(0.0, (gm1 + 1) * gm1 * rho * Cv * T); }
  Real K_rho_sie(Real rho, Real sie) const { return std::max(0.0, (gm1 + 1) * gm1 * rho * sie); }
};

// ---------------------------------------------------------------- reconstruction.hpp:101-244
inline Real mc(Real dm, Real dp, Real alpha) {
  const Real dc = (dm * dp > 0.0) * 0.5 * (dm + dp);
  return std::copysign(std::min(std::fabs(dc), alpha * std::min(std::fabs(dm), std::fabs(dp))), dc);
}
inline void PiecewiseConstant(Real q0, Real &ql, Real &qr) {
  ql = q0;
  qr = q0;
}
inline void PiecewiseLinear(Real qm, Real q0, Real qp, Real &ql, Real &qr) {
  Real dq = qp - q0;
  dq = mc(q0 - qm, dq, 1.5);
  ql = q0 + 0.5 * dq;
  qr = q0 - 0.5 * dq;
}
inline void WENO5Z(Real q0, Real q1, Real q2, Real q3, Real q4, Real &ql, Real &qr) {
  constexpr Real w5alpha[3][3] = {{1.0 / 3.0, -7.0 / 6.0, 11.0 / 6.0},
                                  {-1.0 / 6.0, 5.0 / 6.0, 1.0 / 3.0},
                                  {1.0 / 3.0, 5.0 / 6.0, -1.0 / 6.0}};
  constexpr Real w5gamma[3] = {0.1, 0.6, 0.3};
  constexpr Real eps = 1e-100;
  constexpr Real thirteen_thirds = 13.0 / 3.0;
  Real a = q0 - 2 * q1 + q2;
  Real b = q0 - 4.0 * q1 + 3.0 * q2;
  Real beta0 = thirteen_thirds * a * a + b * b + eps;
  a = q1 - 2.0 * q2 + q3;
  b = q3 - q1;
  Real beta1 = thirteen_thirds * a * a + b * b + eps;
  a = q2 - 2.0 * q3 + q4;
  b = q4 - 4.0 * q3 + 3.0 * q2;
  Real beta2 = thirteen_thirds * a * a + b * b + eps;
  const Real tau5 = std::fabs(beta2 - beta0);
  beta0 = (beta0 + tau5) / beta0;
  beta1 = (beta1 + tau5) / beta1;
  beta2 = (beta2 + tau5) / beta2;
  Real w0 = w5gamma[0] * beta0 + eps;
  Real w1 = w5gamma[1] * beta1 + eps;
  Real w2 = w5gamma[2] * beta2 + eps;
  Real wsum = 1.0 / (w0 + w1 + w2);
  ql = w0 * (w5alpha[0][0] * q0 + w5alpha[0][1] * q1 + w5alpha[0][2] * q2);
  ql += w1 * (w5alpha[1][0] * q1 + w5alpha[1][1] * q2 + w5alpha[1][2] * q3);
  ql += w2 * (w5alpha[2][0] * q2 + w5alpha[2][1] * q3 + w5alpha[2][2] * q4);
  ql *= wsum;
  const Real alpha_l =
      3.0 * wsum * w0 * w1 * w2 / (w5gamma[2] * w0 * w1 + w5gamma[1] * w0 * w2 + w5gamma[0] * w1 * w2) + eps;
  w0 = w5gamma[0] * beta2 + eps;
  w1 = w5gamma[1] * beta1 + eps;
  w2 = w5gamma[2] * beta0 + eps;
  wsum = 1.0 / (w0 + w1 + w2);
  qr = w0 * (w5alpha[0][0] * q4 + w5alpha[0][1] * q3 + w5alpha[0][2] * q2);
  qr += w1 * (w5alpha[1][0] * q3 + w5alpha[1][1] * q2 + w5alpha[1][2] * q1);
  qr += w2 * (w5alpha[2][0] * q2 + w5alpha[2][1] * q1 + w5alpha[2][2] * q0);
  qr *= wsum;
  const Real alpha_r =
      3.0 * wsum * w0 * w1 * w2 / (w5gamma[2] * w0 * w1 + w5gamma[1] * w0 * w2 + w5gamma[0] * w1 * w2) + eps;
  Real dq = q3 - q2;
  dq = mc(q2 - q1, dq, 2.0);
  const Real alpha_lin = 2.0 * alpha_l * alpha_r / (alpha_l + alpha_r);
  ql = alpha_lin * ql + (1.0 - alpha_lin) * (q2 + 0.5 * dq);
  qr = alpha_lin * qr + (1.0 - alpha_lin) * (q2 - 0.5 * dq);
}
inline Real minmod2(Real a, Real b) { return a * b > 0.0 ? (std::fabs(a) < std::fabs(b) ? a : b) : 0.0; }
inline Real mp5_subcalc(Real Fjm2, Real Fjm1, Real Fj, Real Fjp1, Real Fjp2) {
  constexpr Real alpha = 4.0, epsm = 1.e-12;
  Real f = 2.0 * Fjm2 - 13.0 * Fjm1 + 47.0 * Fj + 27.0 * Fjp1 - 3.0 * Fjp2;
  f /= 60.0;
  const Real fMP = Fj + minmod2(Fjp1 - Fj, alpha * (Fj - Fjm1));
  if ((f - Fj) * (f - fMP) <= epsm) return f;
  const Real d2m = Fjm2 + Fj - 2.0 * Fjm1;
  const Real d2 = Fjm1 + Fjp1 - 2.0 * Fj;
  const Real d2p = Fj + Fjp2 - 2.0 * Fjp1;
  Real scrh1 = minmod2(4.0 * d2 - d2p, 4.0 * d2p - d2);
  Real scrh2 = minmod2(d2, d2p);
  const Real dMMp = minmod2(scrh1, scrh2);
  scrh1 = minmod2(4.0 * d2m - d2, 4.0 * d2 - d2m);
  scrh2 = minmod2(d2, d2m);
  const Real dMMm = minmod2(scrh1, scrh2);
  const Real fUL = Fj + alpha * (Fj - Fjm1);
  const Real fAV = 0.5 * (Fj + Fjp1);
  const Real fMD = fAV - 0.5 * dMMp;
  const Real fLC = 0.5 * (3.0 * Fj - Fjm1) + 4.0 / 3.0 * dMMm;
  scrh1 = std::fmin(Fj, Fjp1);
  scrh1 = std::fmin(scrh1, fMD);
  scrh2 = std::fmin(Fj, fUL);
  scrh2 = std::fmin(scrh2, fLC);
  const Real Fmin = std::fmax(scrh1, scrh2);
  scrh1 = std::fmax(Fj, Fjp1);
  scrh1 = std::fmax(scrh1, fMD);
  scrh2 = std::fmax(Fj, fUL);
  scrh2 = std::fmax(scrh2, fLC);
  const Real Fmax = std::fmin(scrh1, scrh2);
  return f + minmod2(Fmin - f, Fmax - f); // Median(f, Fmin, Fmax)
}
inline void MP5(Real q0, Real q1, Real q2, Real q3, Real q4, Real &ql, Real &qr) {
  ql = mp5_subcalc(q0, q1, q2, q3, q4);
  qr = mp5_subcalc(q4, q3, q2, q1, q0);
}
inline void recon5(int rt, Real q0, Real q1, Real q2, Real q3, Real q4, Real &ql, Real &qr) {
  switch (rt) {
  case 0: PiecewiseConstant(q2, ql, qr); break;
  case 1: PiecewiseLinear(q1, q2, q3, ql, qr); break;
  case 2: WENO5Z(q0, q1, q2, q3, q4, ql, qr); break;
  case 3: MP5(q0, q1, q2, q3, q4, ql, qr); break;
  default: g_errors++;
  }
}

// ---------------------------------------------------------------- relativity_utils.hpp:36-58
inline Real LorentzFactor4(const Real v[3], const Real g[4][4]) {
  Real vsq = 0.;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) vsq += g[a + 1][b + 1] * v[a] * v[b];
  return std::sqrt(1. + vsq);
}
inline Real LorentzFactor3(const Real v[3], const Real g[3][3]) {
  Real vsq = 0.;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) vsq += g[a][b] * v[a] * v[b];
  return std::sqrt(1. + vsq);
}

// ---------------------------------------------------------------- prim2con.hpp:25-102
void signal_speeds(Real rho, Real u, Real p, Real bsq, const Real v[3], Real vsq, Real gam1, Real alpha,
                   const Real beta[3], const Real gcon[3][3], Real *sig) {
  const Real rho_rel = rho + u + p;
  const Real vasq = bsq / (rho_rel + bsq);
  Real cssq = gam1 * p / rho_rel;
  cssq += vasq - cssq * vasq;
  const Real vcoff = alpha / (1.0 - vsq * cssq);
  for (int m = 0; m < 3; ++m) {
    const Real vpm = std::sqrt(cssq * (1.0 - vsq) * (gcon[m][m] * (1.0 - vsq * cssq) - v[m] * v[m] * (1.0 - cssq)));
    const Real vp = vcoff * (v[m] * (1.0 - cssq) + vpm) - beta[m];
    const Real vm = vcoff * (v[m] * (1.0 - cssq) - vpm) - beta[m];
    sig[m] = std::max(std::fabs(vp), std::fabs(vm));
  }
}
void p2c(Real rho, const Real vp[3], const Real b[3], Real u, Real p, Real gam1, const Real gcov[4][4],
         const Real gcon[3][3], const Real beta[3], Real alpha, Real gdet, Real &D, Real S[3], Real bcons[3],
         Real &tau, Real *sig) {
  Real vsq = 0.0, Bsq = 0.0, Bdotv = 0.0;
  const Real W = LorentzFactor4(vp, gcov);
  const Real iW = 1.0 / W;
  Real v[3] = {vp[0] * iW, vp[1] * iW, vp[2] * iW};
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) {
      vsq += gcov[a + 1][c + 1] * v[a] * v[c];
      Bsq += gcov[a + 1][c + 1] * b[a] * b[c];
      Bdotv += gcov[a + 1][c + 1] * b[a] * v[c];
    }
  const Real b0 = W * Bdotv;
  Real bcon[4] = {W * Bdotv / alpha, 0.0, 0.0, 0.0};
  for (int m = 0; m < 3; ++m) bcon[m + 1] = b[m] * iW + Bdotv * W * (v[m] - beta[m] / alpha);
  const Real bsq = Bsq * iW * iW + Bdotv * Bdotv;
  Real bcov[3] = {0.0, 0.0, 0.0};
  for (int m = 0; m < 3; ++m)
    for (int n = 0; n < 4; ++n) bcov[m] += gcov[m + 1][n] * bcon[n];
  D = gdet * rho * W;
  const Real rho_rel = (rho + u + p + bsq) * W * W;
  for (int m = 0; m < 3; ++m) {
    Real vcov = 0.0;
    for (int n = 0; n < 3; ++n) vcov += gcov[m + 1][n + 1] * v[n];
    S[m] = gdet * (rho_rel * vcov - b0 * bcov[m]);
    bcons[m] = gdet * b[m];
  }
  tau = gdet * (rho_rel - (p + 0.5 * bsq) - b0 * b0) - D; // USE_VALENCIA
  if (sig != nullptr) signal_speeds(rho, u, p, bsq, v, vsq, gam1, alpha, beta, gcon, sig);
}

// ---------------------------------------------------------------- riemann.hpp:32-54
struct FaceGeom {
  Real alpha, gdet, gammadet, gcov[4][4], beta[3], gdd, X[4];
  FaceGeom(const orc_cfg *c, const Geom &g, int d, int k, int j, int i) {
    const int loc = d; // Face_d
    alpha = g.Lapse(loc, k, j, i);
    gdet = g.DetG(loc, k, j, i);
    gammadet = g.DetGamma(loc, k, j, i);
    Real gcon[3][3];
    g.MetricInverse(loc, k, j, i, gcon);
    gdd = gcon[d - 1][d - 1];
    g.SpacetimeMetric(loc, k, j, i, gcov);
    g.Shift(loc, k, j, i, beta);
    X[0] = 0;
    X[1] = (d == 1 ? Xf(c, 0, i) : Xc(c, 0, i));
    X[2] = (d == 2 ? Xf(c, 1, j) : Xc(c, 1, j));
    X[3] = (d == 3 ? Xf(c, 2, k) : Xc(c, 2, k));
  }
};

// riemann.hpp:87-209; q points at the face state, component stride = cs
void prim_to_flux(const orc_cfg *c, const Bounds &bounds, int d, const FaceGeom &g, const Real *q, long cs,
                  Real &vm, Real &vp, Real *U, Real *F, Real sie_max, Real gam_max) {
  const int mhd = c->mhd;
  const int dir = d - 1;
  const int prs = PPRS(mhd), gm1 = PGM1(mhd);
  Real rho_floor = q[PRHO * cs];
  Real sie_floor;
  bounds.GetFloors(g.X[1], g.X[2], g.X[3], rho_floor, sie_floor);
  const Real rho = std::max(q[PRHO * cs], rho_floor);
  Real vpcon[] = {q[(PVEL)*cs], q[(PVEL + 1) * cs], q[(PVEL + 2) * cs]};
  Real W = LorentzFactor4(vpcon, g.gcov);
  if (W > gam_max) {
    const Real rescale = std::sqrt((gam_max * gam_max - 1.) / (W * W - 1.));
    for (int ii = 0; ii < 3; ++ii) vpcon[ii] *= rescale;
  }
  Real vcon[] = {vpcon[0] / W, vpcon[1] / W, vpcon[2] / W};
  const Real &vel = vcon[dir];
  Real Bcon[] = {0.0, 0.0, 0.0};
  const Real qe = q[PENG * cs];
  const Real u = (qe / rho > sie_floor ? (qe / rho > sie_max ? rho * sie_max : qe) : rho * sie_floor);
  const Real P = std::max(q[prs * cs], 0.0);
  const Real gamma1 = q[gm1 * cs];
  if (mhd)
    for (int m = 0; m < 3; ++m) Bcon[m] = q[(PB + m) * cs];
  Real BdotB = 0.0, Bdotv = 0.0, vsq = 0.0;
  for (int m = 0; m < 3; m++)
    for (int n = 0; n < 3; n++) {
      vsq += g.gcov[m + 1][n + 1] * vcon[m] * vcon[n];
      Bdotv += g.gcov[m + 1][n + 1] * Bcon[m] * vcon[n];
      BdotB += g.gcov[m + 1][n + 1] * Bcon[m] * Bcon[n];
    }
  const Real vsq_max = 1.0 - 1.0 / (gam_max * gam_max);
  if (vsq > vsq_max) {
    Real scale = vsq_max / vsq;
    vsq *= scale;
    scale = std::sqrt(scale);
    Bdotv *= scale;
    for (int m = 0; m < 3; m++) vcon[m] *= scale;
  }
  W = 1.0 / std::sqrt(1 - vsq);
  const Real vtil = vel - g.beta[dir] / g.alpha;
  U[CRHO] = rho * W;
  F[CRHO] = U[CRHO] * vtil;
  Real b0 = W * Bdotv;
  const Real bsq = (BdotB + b0 * b0) / (W * W);
  const Real rhohWsq = (rho + u + P) * W * W;
  const Real bsqWsq = bsq * W * W;
  for (int m = 0; m < 3; m++) {
    Real bcovm = g.gcov[m + 1][0] * b0 / g.alpha;
    Real vcovm = 0.0;
    for (int n = 1; n < 4; n++) {
      bcovm += g.gcov[m + 1][n] * (Bcon[n - 1] / W + b0 * (vcon[n - 1] - g.beta[n - 1] / g.alpha));
      vcovm += g.gcov[m + 1][n] * vcon[n - 1];
    }
    U[CMOM + m] = (rhohWsq + bsqWsq) * vcovm - b0 * bcovm;
    F[CMOM + m] = U[CMOM + m] * vtil + (P + 0.5 * bsq) * (dir == m) - bcovm * Bcon[dir] / W;
  }
  U[CENG] = rhohWsq + bsqWsq - (P + 0.5 * bsq) - b0 * b0 - U[CRHO];
  F[CENG] = U[CENG] * vtil + (P + 0.5 * bsq) * vel - Bdotv * Bcon[dir];
  if (mhd)
    for (int m = 0; m < 3; ++m) {
      U[CB + m] = Bcon[m];
      F[CB + m] = U[CB + m] * vtil - Bcon[dir] * (vcon[m] - g.beta[m] / g.alpha);
    }
  const Real vasq = bsqWsq / make_positive(rhohWsq + bsqWsq);
  const Real cssq = make_bounded(gamma1 * P * W * W / make_positive(rhohWsq), 0.0, 1.0);
  Real cmsq = make_bounded(cssq + vasq * (1.0 - cssq), 0.0, 1.0);
  const Real vcoff = g.alpha / (1.0 - vsq * cmsq);
  const Real v0 = vel * (1.0 - cmsq);
  const Real vpm = std::sqrt(make_positive(cmsq * (1.0 - vsq) * (g.gdd * (1.0 - vsq * cmsq) - vel * v0)));
  vp = vcoff * (v0 + vpm) - g.beta[dir];
  vm = vcoff * (v0 - vpm) - g.beta[dir];
}

// ---------------------------------------------------------------- con2prim_robust.hpp:46-180
class Residual {
 public:
  Residual(Real D, Real q, Real bsq, Real bsq_rpsq, Real rsq, Real rbsq, Real v0sq, const Eos &eos,
           const Bounds &bnds, Real x1, Real x2, Real x3, Real floor_scale_fac)
      : D_(D), q_(q), bsq_(bsq), bsq_rpsq_(bsq_rpsq), rsq_(rsq), rbsq_(rbsq), v0sq_(v0sq), eos_(eos),
        bounds_(bnds), x1_(x1), x2_(x2), x3_(x3), floor_scale_fac_(floor_scale_fac) {
    Real garbage = 0.0;
    bounds_.GetFloors(x1_, x2_, x3_, rho_floor_, garbage);
    bounds_.GetCeilings(x1_, x2_, x3_, gam_max_, e_max_);
    rho_floor_ *= floor_scale_fac_;
  }
  Real x_mu(Real mu) { return ratio(1.0, 1.0 + mu * bsq_); }
  Real rbarsq_mu(Real mu, Real x) { return x * (x * rsq_ + mu * (1.0 + x) * rbsq_); }
  Real qbar_mu(Real mu, Real x) {
    const Real mux = mu * x;
    return q_ - 0.5 * (bsq_ + mux * mux * bsq_rpsq_);
  }
  Real vhatsq_mu(Real mu, Real rbarsq) {
    const Real vsq_trial = mu * mu * rbarsq;
    if (vsq_trial > v0sq_) {
      used_gamma_max_ = true;
      return v0sq_;
    } else {
      used_gamma_max_ = false;
      return vsq_trial;
    }
  }
  Real iWhat_mu(Real vhatsq) { return std::sqrt(1.0 - vhatsq); }
  Real rhohat_mu(Real iWhat) {
    const Real rho_trial = D_ * iWhat;
    if (rho_trial <= rho_floor_) {
      used_density_floor_ = true;
      return rho_floor_;
    } else {
      used_density_floor_ = false;
      return rho_trial;
    }
  }
  Real ehat_mu(Real mu, Real qbar, Real rbarsq, Real vhatsq, Real What) {
    Real rho = rhohat_mu(ratio(1.0, What));
    bounds_.GetFloors(x1_, x2_, x3_, rho, e_floor_);
    e_floor_ *= floor_scale_fac_;
    const Real ehat_trial = What * (qbar - mu * rbarsq) + vhatsq * What * ratio(What, 1.0 + What);
    used_energy_floor_ = false;
    used_energy_max_ = false;
    if (ehat_trial <= e_floor_) {
      used_energy_floor_ = true;
      return e_floor_;
    } else if (ehat_trial > e_max_) {
      used_energy_max_ = true;
      return e_max_;
    } else {
      return ehat_trial;
    }
  }
  Real operator()(Real mu) {
    const Real x = x_mu(mu);
    const Real rbarsq = rbarsq_mu(mu, x);
    const Real qbar = qbar_mu(mu, x);
    const Real vhatsq = vhatsq_mu(mu, rbarsq);
    const Real iWhat = iWhat_mu(vhatsq);
    const Real What = ratio(1.0, iWhat);
    Real rhohat = rhohat_mu(iWhat);
    Real ehat = ehat_mu(mu, qbar, rbarsq, vhatsq, What);
    const Real Phat = eos_.P_rho_sie(rhohat, ehat);
    Real hhat = rhohat * (1.0 + ehat) + Phat;
    const Real ahat = Phat / make_positive(hhat - Phat);
    hhat /= make_positive(rhohat);
    const Real nua = (1.0 + ahat) * (1.0 + ehat) * iWhat;
    const Real nub = (1.0 + ahat) * (1.0 + qbar - mu * rbarsq);
    const Real nuhat = std::max(nua, nub);
    const Real muhat = 1.0 / (nuhat + mu * rbarsq);
    return mu - muhat;
  }
  bool used_density_floor() const { return used_density_floor_; }
  bool used_energy_floor() const { return used_energy_floor_; }
  bool used_gamma_max() const { return used_gamma_max_; }
  bool used_bounds() const {
    return (used_density_floor_ || used_energy_floor_ || used_energy_max_ || used_gamma_max_);
  }

 private:
  const Real D_, q_, bsq_, bsq_rpsq_, rsq_, rbsq_, v0sq_;
  const Eos &eos_;
  const Bounds &bounds_;
  const Real x1_, x2_, x3_;
  const Real floor_scale_fac_;
  Real rho_floor_, e_floor_, gam_max_, e_max_;
  bool used_density_floor_ = false, used_energy_floor_ = false, used_energy_max_ = false,
       used_gamma_max_ = false;
};

// ---------------------------------------------------------------- root_find.hpp:42-58,110-167
struct RootFind {
  int iteration_count = 0, max_iter;
  explicit RootFind(int mi) : max_iter(mi) {}
  template <typename F>
  void refine_bracket(F &func, Real guess, Real &a, Real &b, Real &ya, Real &yb) {
    if (a <= guess && b >= guess) {
      ya = func(guess);
      yb = func(b);
      if (ya * yb > 0.0) {
        yb = ya;
        b = guess;
        ya = func(a);
      } else {
        a = guess;
      }
    } else {
      ya = func(a);
      yb = func(b);
    }
  }
  template <typename F>
  Real regula_falsi(F &func, Real a, Real b, Real tol, Real guess, bool *unbracketed) {
    Real ya, yb;
    refine_bracket(func, guess, a, b, ya, yb);
    if (ya * yb > 0.0) { // check_bracket; the reference aborts here (status == nullptr)
      *unbracketed = true;
      return guess;
    }
    Real sign = (ya < 0 ? 1.0 : -1.0);
    ya *= sign;
    yb *= sign;
    int b1 = 0, b2 = 0;
    while (b - a > 2.0 * tol && iteration_count < max_iter) {
      Real c = (a * yb - b * ya) / (yb - ya);
      if (c == a) {
        b = a;
        iteration_count++;
        continue;
      } else if (c == b) {
        a = b;
        iteration_count++;
        continue;
      }
      Real yc = sign * func(c);
      if (yc > 0.0) {
        b = c;
        yb = yc;
        b1++;
        ya *= (b1 > 1 ? 0.5 : 1.0);
        b2 = 0;
      } else if (yc < 0.0) {
        a = c;
        ya = yc;
        b2++;
        yb *= (b2 > 1 ? 0.5 : 1.0);
        b1 = 0;
      } else {
        a = c;
        b = c;
      }
      iteration_count++;
    }
    return 0.5 * (a + b);
  }
};

// cell-centre geometry bundle, con2prim_robust.hpp:198-221
struct CellGeom {
  Real gcov4[4][4], gcon[3][3], beta[3], gdet, lapse;
  CellGeom(const Geom &g, int k, int j, int i) {
    gdet = g.DetGamma(0, k, j, i);
    lapse = g.Lapse(0, k, j, i);
    g.SpacetimeMetric(0, k, j, i, gcov4);
    g.MetricInverse(0, k, j, i, gcon);
    g.Shift(0, k, j, i, beta);
  }
};

// one zone of ConToPrim::solve, con2prim_robust.hpp:300-453. Returns true on success.
// Field pointers already offset to the zone; cs = component stride.
struct ZoneRefs {
  Real *cons, *prim, *temp, *mu, *csig;
  long cs;
};
bool c2p_zone(const orc_cfg *c, const Bounds &bounds, const Eos &eos, const CellGeom &g, Real x1, Real x2,
              Real x3, const ZoneRefs &z, int ndim) {
  const int mhd = c->mhd;
  const long cs = z.cs;
  Real *U = z.cons, *P = z.prim;
  const int prs = PPRS(mhd), gm1 = PGM1(mhd);
  int num_nans = std::isnan(U[CRHO * cs]) + std::isnan(U[CMOM * cs]) + std::isnan(U[CENG * cs]);
  if (num_nans > 0) return false;
  const Real igdet = 1.0 / g.gdet;
  Real rhoflr = 0.0, epsflr;
  bounds.GetFloors(x1, x2, x3, rhoflr, epsflr);
  rhoflr *= c->c2p_floor_scale_fac;
  epsflr *= c->c2p_floor_scale_fac;
  Real gam_max, eps_max;
  bounds.GetCeilings(x1, x2, x3, gam_max, eps_max);
  const Real D = U[CRHO * cs] * igdet;
  const Real tau = U[CENG * cs] * igdet;
  const Real q = ratio(tau, D);
  Real bsq = 0.0, bsq_rpsq = 0.0, rbsq = 0.0, rsq = 0.0, bdotr = 0.0;
  Real rcon[3];
  Real rcov[] = {ratio(U[(CMOM)*cs], U[CRHO * cs]), ratio(U[(CMOM + 1) * cs], U[CRHO * cs]),
                 ratio(U[(CMOM + 2) * cs], U[CRHO * cs])};
  for (int a = 0; a < 3; ++a) {
    rcon[a] = 0.0;
    for (int b = 0; b < 3; ++b) rcon[a] += g.gcon[a][b] * rcov[b];
    rsq += rcon[a] * rcov[a];
  }
  Real bu[] = {0.0, 0.0, 0.0};
  if (mhd) {
    for (int a = 0; a < 3; ++a) P[(PB + a) * cs] = U[(CB + a) * cs] * igdet;
    const Real sD = 1.0 / std::sqrt(std::max(D, rhoflr));
    for (int a = 0; a < 3; ++a) {
      bu[a] = P[(PB + a) * cs] * sD;
      bdotr += bu[a] * rcov[a];
    }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) bsq += g.gcov4[a + 1][b + 1] * bu[a] * bu[b];
    bsq = std::max(0.0, bsq);
    rbsq = bdotr * bdotr;
    bsq_rpsq = bsq * rsq - rbsq;
  }
  const Real zsq = rsq / 1.0; // h0sq_ = 1
  Real v0sq = std::min(zsq / (1.0 + zsq), 1.0 - 1.0 / (gam_max * gam_max));
  Residual res(D, q, bsq, bsq_rpsq, rsq, rbsq, v0sq, eos, bounds, x1, x2, x3, c->c2p_floor_scale_fac);
  RootFind root(c->c2p_max_iter);
  bool unbracketed = false;
  const Real mu = root.regula_falsi(res, 0.0, 1.0, c->c2p_tol, *z.mu, &unbracketed);
  if (unbracketed) {
    g_errors++; // reference: PARTHENON_FAIL("Aborting with unbracketed root.")
    return false;
  }
  *z.mu = mu;
  const Real x = res.x_mu(mu);
  const Real rbarsq = res.rbarsq_mu(mu, x);
  Real vsq = res.vhatsq_mu(mu, rbarsq);
  Real iW = res.iWhat_mu(vsq);
  P[PRHO * cs] = res.rhohat_mu(iW);
  const Real qbar = res.qbar_mu(mu, x);
  Real W = 1.0 / iW;
  P[PENG * cs] = res.ehat_mu(mu, qbar, rbarsq, vsq, W);
  *z.temp = eos.T_rho_sie(P[PRHO * cs], P[PENG * cs]);
  P[PENG * cs] *= P[PRHO * cs];
  P[prs * cs] = eos.P_rho_T(P[PRHO * cs], *z.temp);
  P[gm1 * cs] = eos.K_rho_T(P[PRHO * cs], *z.temp) / P[prs * cs];
  Real vel[3];
  for (int a = 0; a < 3; ++a) {
    vel[a] = W * mu * x * (rcon[a] + mu * bdotr * bu[a]);
    P[(PVEL + a) * cs] = vel[a];
  }
  if (mhd)
    for (int a = 0; a < 3; ++a) bu[a] = P[(PB + a) * cs];
  Real sig[3];
  if (res.used_bounds()) {
    Real S[3], bcons[3];
    p2c(P[PRHO * cs], vel, bu, P[PENG * cs], P[prs * cs], P[gm1 * cs], g.gcov4, g.gcon, g.beta, g.lapse, g.gdet,
        U[CRHO * cs], S, bcons, U[CENG * cs], sig);
    for (int a = 0; a < 3; ++a) U[(CMOM + a) * cs] = S[a];
  } else {
    for (int m = 0; m < 3; ++m) vel[m] *= iW;
    bsq = 0.0;
    if (mhd) {
      Real bdotv = 0.0;
      for (int m = 0; m < 3; ++m)
        for (int n = 0; n < 3; ++n) {
          bsq += bu[m] * bu[n] * g.gcov4[m + 1][n + 1];
          bdotv += bu[m] * vel[n] * g.gcov4[m + 1][n + 1];
        }
      bsq = iW * iW * bsq + bdotv * bdotv;
    }
    signal_speeds(P[PRHO * cs], P[PENG * cs], P[prs * cs], bsq, vel, vsq, P[gm1 * cs], g.lapse, g.beta, g.gcon,
                  sig);
  }
  for (int a = 0; a < ndim; ++a) z.csig[a * cs] = sig[a];
  num_nans = std::isnan(U[CRHO * cs]) + std::isnan(U[CMOM * cs]) + std::isnan(U[CENG * cs]);
  if (num_nans > 0) return false;
  if (c->c2p_fail_on_floors && (res.used_density_floor() || res.used_energy_floor())) return false;
  if (c->c2p_fail_on_ceilings && res.used_gamma_max()) return false;
  return true;
}

// ---------------------------------------------------------------- tmunu.hpp:63-131
void Tmunu(const orc_cfg *c, const Geom &g, const Real *prim, long cs, int k, int j, int i, Real T[4][4]) {
  const int mhd = c->mhd;
  Real u[4], b[4], gg[4][4], bsq;
  Real beta[3], Bdotv = 0.0, Bsq = 0.0;
  Real gammacov[3][3];
  g.Metric(0, k, j, i, gammacov);
  Real vp[3] = {prim[PVEL * cs], prim[(PVEL + 1) * cs], prim[(PVEL + 2) * cs]};
  Real B[3] = {0, 0, 0};
  if (mhd)
    for (int a = 0; a < 3; ++a) B[a] = prim[(PB + a) * cs];
  const Real W = LorentzFactor3(vp, gammacov);
  for (int a = 0; a < 3; ++a)
    for (int bb = 0; bb < 3; ++bb) {
      Bdotv += B[a] * (vp[bb] / W) * gammacov[a][bb];
      Bsq += B[a] * B[bb] * gammacov[a][bb];
    }
  const Real iW = ratio(1., W);
  Real alpha = g.Lapse(0, k, j, i);
  g.Shift(0, k, j, i, beta);
  u[0] = ratio(W, std::abs(alpha));
  b[0] = u[0] * Bdotv;
  for (int l = 1; l < 4; ++l) {
    u[l] = vp[l - 1] - u[0] * beta[l - 1];
    b[l] = iW * (B[l - 1] + alpha * b[0] * u[l]);
  }
  bsq = Bsq * iW * iW + Bdotv * Bdotv;
  g.SpacetimeMetricInverse(0, k, j, i, gg);
  Real rho = prim[PRHO * cs], uu = prim[PENG * cs], P = prim[PPRS(mhd) * cs];
  Real A = rho + uu + P + bsq;
  Real Bc = P + 0.5 * bsq;
  for (int m = 0; m < 4; ++m)
    for (int n = 0; n < 4; ++n) T[m][n] = A * u[m] * u[n] + Bc * gg[m][n] - b[m] * b[n];
}

} // namespace

// =====================================================================================
extern "C" {

long long orc_error_count() { return g_errors; }
void orc_error_reset() { g_errors = 0; }
int orc_nprim(const orc_cfg *c) { return 7 + 3 * c->mhd; }
int orc_ncons(const orc_cfg *c) { return 5 + 3 * c->mhd; }

void orc_recon5(int recon, const double q[5], double *ql, double *qr) {
  recon5(recon, q[0], q[1], q[2], q[3], q[4], *ql, *qr);
}

// fluid.cpp:659-773 + riemann.hpp:256-301
void orc_calculate_fluxes(const orc_cfg *c, const double *geom, const double *prim, double *flux,
                          double *fsig) {
  Lay L(c);
  Geom g(c, geom);
  Bounds bounds(c);
  const long N = L.ncell;
  const int nrecon = L.nprim;
  std::vector<Real> ql((size_t)L.ndim * nrecon * N, 0.0), qr((size_t)L.ndim * nrecon * N, 0.0);
  auto QL = [&](int d, int n, int k, int j, int i) -> Real & { return ql[((size_t)d * nrecon + n) * N + L.idx(k, j, i)]; };
  auto QR = [&](int d, int n, int k, int j, int i) -> Real & { return qr[((size_t)d * nrecon + n) * N + L.idx(k, j, i)]; };
  // kernel "Reconstruct", fluid.cpp:681-751
  for (int n = 0; n < nrecon; ++n) {
    const Real *v = prim + (size_t)n * N;
    for (int k = L.ks - L.dk; k <= L.ke + L.dk; ++k)
      for (int j = L.js - L.dj; j <= L.je + L.dj; ++j)
        for (int i = L.is - 1; i <= L.ie + 1; ++i) {
          const long o = L.idx(k, j, i);
          recon5(c->recon, v[o - 2], v[o - 1], v[o], v[o + 1], v[o + 2], QL(0, n, k, j, i + 1), QR(0, n, k, j, i));
          if (L.ndim > 1) {
            const long s = L.ni;
            recon5(c->recon, v[o - 2 * s], v[o - s], v[o], v[o + s], v[o + 2 * s], QL(1, n, k, j + 1, i),
                   QR(1, n, k, j, i));
          }
          if (L.ndim > 2) {
            const long s = (long)L.ni * L.nj;
            recon5(c->recon, v[o - 2 * s], v[o - s], v[o], v[o + s], v[o + 2 * s], QL(2, n, k + 1, j, i),
                   QR(2, n, k, j, i));
          }
        }
  }
  // kernel "CalculateFluxes", fluid.cpp:753-770; only faces whose ql AND qr were both
  // reconstructed are evaluated (the reference also evaluates the never-consumed
  // faces at index s-1 from stale ql)
  const long cs = N;
  for (int d = 1; d <= L.ndim; ++d) {
    const int klo = (d == 3) ? L.ks : L.ks - L.dk, khi = L.ke + L.dk;
    const int jlo = (d == 2) ? L.js : L.js - L.dj, jhi = L.je + L.dj;
    const int ilo = (d == 1) ? L.is : L.is - 1, ihi = L.ie + 1;
    for (int k = klo; k <= khi; ++k)
      for (int j = jlo; j <= jhi; ++j)
        for (int i = ilo; i <= ihi; ++i) {
          Real Ul[8], Ur[8], Fl[8], Fr[8], vml, vpl, vmr, vpr;
          FaceGeom fg(c, g, d, k, j, i);
          Real gam_max, sie_max;
          bounds.GetCeilings(fg.X[1], fg.X[2], fg.X[3], gam_max, sie_max);
          const long o = L.idx(k, j, i);
          prim_to_flux(c, bounds, d, fg, &ql[(size_t)(d - 1) * nrecon * N + o], cs, vml, vpl, Ul, Fl, sie_max, gam_max);
          prim_to_flux(c, bounds, d, fg, &qr[(size_t)(d - 1) * nrecon * N + o], cs, vmr, vpr, Ur, Fr, sie_max, gam_max);
          Real *fl = flux + (size_t)(d - 1) * L.ncons * N + o;
          Real speed;
          if (c->riemann == 0) { // llf, riemann.hpp:256-276
            const Real cmax = std::max(std::max(-vml, vpl), std::max(-vmr, vpr));
            for (int m = 0; m < L.ncons; m++)
              fl[m * cs] = 0.5 * ((Fl[m] + Fr[m]) * fg.gdet - cmax * ((Ur[m] - Ul[m]) * fg.gammadet));
            speed = cmax;
          } else { // hll, riemann.hpp:278-301
            const Real cl = std::min(std::min(vml, vmr), 0.0);
            const Real cr = std::max(std::max(vpl, vpr), 0.0);
            const Real crcl = cr * cl;
            const Real inv_cr_cl = ratio(1.0, cr - cl);
            for (int m = 0; m < L.ncons; m++)
              fl[m * cs] = ((cr * Fl[m] - cl * Fr[m]) * fg.gdet + crcl * (Ur[m] - Ul[m]) * fg.gammadet) * inv_cr_cl;
            speed = std::max(-cl, cr);
          }
          fsig[(size_t)(d - 1) * N + o] = speed;
        }
  }
}

// fluid.cpp:775-841
void orc_flux_ct(const orc_cfg *c, double *flux, double *emf) {
  Lay L(c);
  if (!c->mhd || L.ndim == 1) return;
  const long N = L.ncell;
  auto F = [&](int dir, int bcomp, int k, int j, int i) -> Real & {
    return flux[((size_t)(dir - 1) * L.ncons + CB + bcomp) * N + L.idx(k, j, i)];
  };
  if (L.ndim == 2) {
    const int k = 0;
    for (int j = L.js; j <= L.je + 1; ++j)
      for (int i = L.is; i <= L.ie + 1; ++i)
        emf[L.idx(k, j, i)] = 0.25 * (F(1, 1, k, j, i) + F(1, 1, k, j - 1, i) - F(2, 0, k, j, i) - F(2, 0, k, j, i - 1));
    // the reference reads emf(k,j+1,i), emf(k,j,i+1) while looping to je+1 / ie+1 (fluid.cpp:797-803)
    for (int j = L.js; j <= L.je + 1; ++j)
      for (int i = L.is; i <= L.ie + 1; ++i) {
        F(1, 0, k, j, i) = 0.0;
        F(1, 1, k, j, i) = 0.5 * (emf[L.idx(k, j, i)] + emf[L.idx(k, j + 1, i)]);
        F(2, 0, k, j, i) = -0.5 * (emf[L.idx(k, j, i)] + emf[L.idx(k, j, i + 1)]);
        F(2, 1, k, j, i) = 0.0;
      }
  } else {
    auto E = [&](int n, int k, int j, int i) -> Real & { return emf[(size_t)n * N + L.idx(k, j, i)]; };
    for (int k = L.ks; k <= L.ke + 1; ++k)
      for (int j = L.js; j <= L.je + 1; ++j)
        for (int i = L.is; i <= L.ie + 1; ++i) {
          E(0, k, j, i) = 0.25 * (F(2, 2, k, j, i) + F(2, 2, k - 1, j, i) - F(3, 1, k, j, i) - F(3, 1, k, j - 1, i));
          E(1, k, j, i) = -0.25 * (F(1, 2, k, j, i) + F(1, 2, k - 1, j, i) - F(3, 0, k, j, i) - F(3, 0, k, j, i - 1));
          E(2, k, j, i) = 0.25 * (F(1, 1, k, j, i) + F(1, 1, k, j - 1, i) - F(2, 0, k, j, i) - F(2, 0, k, j, i - 1));
        }
    for (int k = L.ks; k <= L.ke + 1; ++k)
      for (int j = L.js; j <= L.je + 1; ++j)
        for (int i = L.is; i <= L.ie + 1; ++i) {
          F(1, 0, k, j, i) = 0.0;
          F(1, 1, k, j, i) = 0.5 * (E(2, k, j, i) + E(2, k, j + 1, i));
          F(1, 2, k, j, i) = -0.5 * (E(1, k, j, i) + E(1, k + 1, j, i));
          F(2, 0, k, j, i) = -0.5 * (E(2, k, j, i) + E(2, k, j, i + 1));
          F(2, 1, k, j, i) = 0.0;
          F(2, 2, k, j, i) = 0.5 * (E(0, k, j, i) + E(0, k + 1, j, i));
          F(3, 0, k, j, i) = 0.5 * (E(1, k, j, i) + E(1, k, j, i + 1));
          F(3, 1, k, j, i) = -0.5 * (E(0, k, j, i) + E(0, k, j + 1, i));
          F(3, 2, k, j, i) = 0.0;
        }
  }
}

// fixup.cpp:728-1078 (fluid variables only)
void orc_fix_fluxes(const orc_cfg *c, double *flux) {
  Lay L(c);
  if (!c->enable_flux_fixup) return;
  const long N = L.ncell;
  auto F = [&](int dir, int m, int k, int j, int i) -> Real & {
    return flux[((size_t)(dir - 1) * L.ncons + m) * N + L.idx(k, j, i)];
  };
  // X1
  for (int side = 0; side < 2; ++side) {
    const int bc = c->bc[side];
    if (bc <= 0) continue;
    const int i = side == 0 ? L.is : L.ie + 1;
    for (int k = L.ks; k <= L.ke; ++k)
      for (int j = L.js; j <= L.je; ++j) {
        if (bc == 1) {
          F(1, CRHO, k, j, i) = side == 0 ? std::min(F(1, CRHO, k, j, i), 0.0) : std::max(F(1, CRHO, k, j, i), 0.0);
        } else if (bc == 2) {
          F(1, CRHO, k, j, i) = 0.0;
          F(1, CENG, k, j, i) = 0.0;
        }
      }
  }
  if (L.ndim == 1) return;
  for (int side = 0; side < 2; ++side) {
    const int bc = c->bc[2 + side];
    if (bc <= 0) continue;
    const int j = side == 0 ? L.js : L.je + 1;
    for (int k = L.ks; k <= L.ke; ++k)
      for (int i = L.is; i <= L.ie; ++i) {
        if (bc == 1) {
          F(2, CRHO, k, j, i) = side == 0 ? std::min(F(2, CRHO, k, j, i), 0.0) : std::max(F(2, CRHO, k, j, i), 0.0);
        } else if (bc == 3) { // polar
          F(2, CRHO, k, j, i) = 0.0;
          F(2, CENG, k, j, i) = 0.0;
          for (int ii = 0; ii < 3; ++ii) {
            F(2, CMOM + ii, k, j, i) = 0.0;
            if (c->mhd) F(2, CB + ii, k, j, i) = 0.0;
          }
          if (c->mhd) {
            if (side == 0) {
              F(1, CB + 1, k, j - 1, i) = -F(1, CB + 1, k, j, i);
              if (L.ndim == 3) F(3, CB + 1, k, j - 1, i) = -F(3, CB + 1, k, j, i);
            } else {
              F(1, CB + 1, k, j, i) = -F(1, CB + 1, k, j - 1, i);
              if (L.ndim == 3) F(3, CB + 1, k, j, i) = -F(3, CB + 1, k, j - 1, i);
            }
          }
        } else if (bc == 2) { // reflect
          F(2, CRHO, k, j, i) = 0.0;
          F(2, CENG, k, j, i) = 0.0;
          F(2, CMOM + 0, k, j, i) = 0.0;
          F(2, CMOM + 2, k, j, i) = 0.0;
        }
      }
  }
  if (L.ndim == 2) return;
  for (int side = 0; side < 2; ++side) {
    const int bc = c->bc[4 + side];
    if (bc <= 0) continue;
    const int k = side == 0 ? L.ks : L.ke + 1;
    for (int j = L.js; j <= L.je; ++j)
      for (int i = L.is; i <= L.ie; ++i) {
        if (bc == 1) {
          F(3, CRHO, k, j, i) = side == 0 ? std::min(F(3, CRHO, k, j, i), 0.0) : std::max(F(3, CRHO, k, j, i), 0.0);
        } else if (bc == 2) {
          F(3, CRHO, k, j, i) = 0.0;
          F(3, CENG, k, j, i) = 0.0;
        }
      }
  }
}

// update.hpp:43-58, update.cpp:63-85
void orc_flux_divergence(const orc_cfg *c, const double *flux, double *dudt) {
  Lay L(c);
  const long N = L.ncell;
  // uniform_cartesian.hpp:30-40
  const Real area[3] = {c->dx[1] * c->dx[2], c->dx[0] * c->dx[2], c->dx[0] * c->dx[1]};
  const Real vol = c->dx[0] * c->dx[1] * c->dx[2];
  const long st[3] = {1, L.ni, (long)L.ni * L.nj};
  for (int l = 0; l < L.ncons; ++l)
    for (int k = L.ks; k <= L.ke; ++k)
      for (int j = L.js; j <= L.je; ++j)
        for (int i = L.is; i <= L.ie; ++i) {
          const long o = L.idx(k, j, i);
          Real du = 0;
          for (int d = 0; d < L.ndim; ++d) {
            const Real *f = flux + ((size_t)d * L.ncons + l) * N;
            const Real t = (area[d] * f[o + st[d]] - area[d] * f[o]);
            du = (d == 0) ? t : du + t;
          }
          dudt[(size_t)l * N + o] = -du / vol;
        }
}

// ---------------------------------------------------------------- history (.hst columns)
// fluid.cpp:325-347, analysis/history.hpp:50-86. Reducer_t is Kokkos::Sum<Real> / Kokkos::Max<Real>; the
// volume_weighting test compares it with Kokkos::Sum<Real, ExecSpace> and is false: plain sums.
void orc_history_fluid(const orc_cfg *c, const double *prim, const double *cons, double out[6]) {
  Lay L(c);
  const long N = L.ncell;
  out[0] = -DBL_MAX;
  for (int q = 1; q < 6; ++q) out[q] = 0.0;
  for (int k = L.ks; k <= L.ke; ++k)
    for (int j = L.js; j <= L.je; ++j)
      for (int i = L.is; i <= L.ie; ++i) {
        const long o = L.idx(k, j, i);
        out[0] = std::max(out[0], prim[o]);
        out[1] += cons[o];
        out[2] += cons[4 * N + o];
        for (int d = 0; d < 3; ++d) out[3 + d] += cons[(1 + d) * N + o];
      }
}

namespace {
struct HistCell {
  Real lapse, gdet, shift[3], g4[4][4], gam[3][3], rho, v[3], B[3];
};
void hist_load(const orc_cfg *c, const Geom &g, const Lay &L, const double *prim, int k, int j, int i, HistCell &h) {
  const long N = L.ncell, o = L.idx(k, j, i);
  h.lapse = g.Lapse(0, k, j, i);
  h.gdet = g.DetGamma(0, k, j, i);
  g.Shift(0, k, j, i, h.shift);
  g.SpacetimeMetric(0, k, j, i, h.g4);
  g.Metric(0, k, j, i, h.gam);
  h.rho = prim[o];
  for (int d = 0; d < 3; ++d) {
    h.v[d] = prim[(1 + d) * N + o];
    h.B[d] = c->mhd ? prim[(5 + d) * N + o] : 0.0;
  }
}
// relativity_utils.hpp:131-141
Real hist_bsq(const HistCell &h, Real W) {
  Real Bsq = 0, Bdotv = 0;
  for (int m = 0; m < 3; ++m)
    for (int n = 0; n < 3; ++n) {
      Bdotv += h.gam[m][n] * ratio(h.v[m], W) * h.B[n];
      Bsq += h.gam[m][n] * h.B[m] * h.B[n];
    }
  return ratio(Bsq, W * W) + Bdotv * Bdotv;
}
// history_utils.hpp:35-56
Real hist_mass_flux(const HistCell &h) {
  const Real W = LorentzFactor4(h.v, h.g4);
  const Real ucon = h.v[0] - h.shift[0] * W / h.lapse;
  return -h.lapse * h.gdet * h.rho * ucon;
}
// history_utils.hpp:58-76
Real hist_sigma(const HistCell &h) {
  const Real W = LorentzFactor3(h.v, h.gam);
  return hist_bsq(h, W) / h.rho;
}
// history_utils.hpp:78-160
void hist_em(const HistCell &h, Real &en, Real &mom) {
  const Real W = LorentzFactor3(h.v, h.gam);
  Real Bdotv = 0.0, vcov[3] = {0, 0, 0}, Bcov[3] = {0, 0, 0};
  for (int m = 0; m < 3; ++m)
    for (int n = 0; n < 3; ++n) {
      vcov[m] += h.gam[m][n] * h.v[n];
      Bcov[m] += h.gam[m][n] * h.B[n];
      Bdotv += h.gam[m][n] * h.B[m] * ratio(h.v[n], W);
    }
  const Real b0 = Bdotv * W / h.lapse;
  const Real ucov0 = -W * h.lapse;
  const Real ucon1 = h.v[0] - h.shift[0] * W / h.lapse;
  const Real ucov1 = vcov[0];
  const Real bcon1 = ratio((h.B[0] + h.lapse * b0 * ucon1), W);
  const Real bcov1 = ratio((Bcov[0] + h.lapse * b0 * ucov1), W);
  const Real bsq = hist_bsq(h, W);
  en = +(bsq * ucon1 * ucov0 - bcon1 * b0) * h.lapse * h.gdet;
  mom = -(bsq * ucon1 * ucov1 - bcon1 * bcov1 + 0.5 * bsq) * h.lapse * h.gdet;
}
} // namespace

// analysis/history.cpp:25-222: fluxes through x1 = xh, interpolated from the zone column that contains it
void orc_history_torus(const orc_cfg *c, const double *geom, const double *prim, const double *cons, double xh,
                       double sigma_cutoff, double out[4]) {
  Lay L(c);
  Geom g(c, geom);
  const long N = L.ncell;
  for (int q = 0; q < 4; ++q) out[q] = 0.0;
  const Real dx1 = c->dx[0], dx2 = c->dx[1], dx3 = c->dx[2];
  for (int k = L.ks; k <= L.ke; ++k)
    for (int j = L.js; j <= L.je; ++j)
      for (int i = L.is; i <= L.ie; ++i) {
        if (!(Xf(c, 0, i) <= xh && xh < Xf(c, 0, i + 1))) continue;
        HistCell c0, cm, cp;
        hist_load(c, g, L, prim, k, j, i, c0);
        hist_load(c, g, L, prim, k, j, i - 1, cm);
        hist_load(c, g, L, prim, k, j, i + 1, cp);
        {
          const Real m = (hist_mass_flux(cp) - hist_mass_flux(cm)) / (2.0 * dx1);
          out[0] += (hist_mass_flux(c0) + (xh - Xc(c, 0, i)) * m) * dx2 * dx3;
        }
        if (!c->mhd) continue;
        if (hist_sigma(c0) > sigma_cutoff) {
          Real e0, p0, em, pm, ep, pp;
          hist_em(c0, e0, p0);
          hist_em(cm, em, pm);
          hist_em(cp, ep, pp);
          out[1] += (e0 + (xh - Xc(c, 0, i)) * ((ep - em) / (2.0 * dx1))) * dx2 * dx3;
          out[2] += (p0 + (xh - Xc(c, 0, i)) * ((pp - pm) / (2.0 * dx1))) * dx2 * dx3;
        }
        auto phi = [&](int ii) { return std::fabs(cons[5 * N + L.idx(k, j, ii)]) * g.Lapse(0, k, j, ii); };
        const Real m = (phi(i + 2) - phi(i)) / (2.0 * dx1);
        out[3] += (phi(i + 1) + (xh - Xc(c, 0, i + 1)) * m) * dx2 * dx3;
      }
}

// fluid.cpp:843-897 (CalculateDivB): the sums are evaluated left to right as written there
void orc_divb(const orc_cfg *c, const double *cons, double *divb) {
  Lay L(c);
  if (!c->mhd || L.ndim < 2) return;
  const long N = L.ncell;
  const Real *b0 = cons + (size_t)5 * N, *b1 = cons + (size_t)6 * N, *b2 = cons + (size_t)7 * N;
  const long sj = L.ni, sk = (long)L.ni * L.nj;
  for (int k = L.ks; k <= L.ke; ++k)
    for (int j = L.js; j <= L.je; ++j)
      for (int i = L.is; i <= L.ie; ++i) {
        const long o = L.idx(k, j, i);
        if (L.ndim == 2) {
          divb[o] = 0.5 * (b0[o] + b0[o - sj] - b0[o - 1] - b0[o - sj - 1]) / c->dx[0] +
                    0.5 * (b1[o] + b1[o - 1] - b1[o - sj] - b1[o - sj - 1]) / c->dx[1];
        } else {
          divb[o] = 0.25 *
                        (b0[o] + b0[o - sj] + b0[o - sk] + b0[o - sk - sj] - b0[o - 1] - b0[o - sj - 1] - b0[o - sk - 1] -
                         b0[o - sk - sj - 1]) /
                        c->dx[0] +
                    0.25 *
                        (b1[o] + b1[o - 1] + b1[o - sk] + b1[o - sk - 1] - b1[o - sj] - b1[o - sj - 1] - b1[o - sk - sj] -
                         b1[o - sk - sj - 1]) /
                        c->dx[1] +
                    0.25 *
                        (b2[o] + b2[o - 1] + b2[o - sj] + b2[o - sj - 1] - b2[o - sk] - b2[o - sk - 1] - b2[o - sk - sj] -
                         b2[o - sk - sj - 1]) /
                        c->dx[2];
        }
      }
}

// fluid.cpp:566-657
void orc_source_terms(const orc_cfg *c, const double *geom, const double *prim, double *gsrc) {
  Lay L(c);
  Geom g(c, geom);
  const long N = L.ncell;
  for (int k = L.ks; k <= L.ke; ++k)
    for (int j = L.js; j <= L.je; ++j)
      for (int i = L.is; i <= L.ie; ++i) {
        const long o = L.idx(k, j, i);
        Real T[4][4], dg[4][4][4], gam[4][4][4];
        Tmunu(c, g, prim + o, N, k, j, i, T);
        Real gdet = g.DetG(0, k, j, i);
        g.MetricDerivative(k, j, i, dg);
        // geometry_utils.hpp:73-80
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b)
            for (int cc = 0; cc < 4; ++cc) gam[a][b][cc] = 0.5 * (dg[b][a][cc] + dg[cc][a][b] - dg[b][cc][a]);
        for (int l = 0; l < 3; ++l) {
          Real src_mom = 0.0;
          for (int m = 0; m < 4; ++m)
            for (int n = 0; n < 4; ++n) src_mom += T[m][n] * (dg[n][l + 1][m] - gam[l + 1][n][m]);
          gsrc[(size_t)l * N + o] = gdet * src_mom;
        }
        Real TGam = 0.0;
        const Real alpha = g.Lapse(0, k, j, i);
        const Real inv_alpha2 = ratio(1, alpha * alpha);
        Real shift[3];
        g.Shift(0, k, j, i, shift);
        Real gcon0[4] = {-inv_alpha2, inv_alpha2 * shift[0], inv_alpha2 * shift[1], inv_alpha2 * shift[2]};
        for (int m = 0; m < 4; m++)
          for (int n = 0; n < 4; n++) {
            Real gam0 = 0;
            for (int r = 0; r < 4; r++) gam0 += gcon0[r] * gam[r][m][n];
            TGam += T[m][n] * gam0;
          }
        Real Ta = 0.0, da[4];
        g.GradLnAlpha(k, j, i, da);
        for (int m = 0; m < 4; m++) Ta += T[m][0] * da[m];
        gsrc[(size_t)3 * N + o] = gdet * alpha * (Ta - TGam);
      }
}

// update.hpp:71-136 as sequenced by phoebus_driver.cpp:560-576, over ENTIRE arrays
void orc_rk_update(const orc_cfg *c, double *dudt, const double *gsrc, double *u0, const double *ubase,
                   double beta, double bdt, double *uout) {
  Lay L(c);
  const long N = L.ncell;
  if (gsrc) // SumData(src_names = {c.momentum, c.energy}): z = 1*x + 1*y
    for (int l = 0; l < 4; ++l)
      for (long o = 0; o < N; ++o) dudt[(size_t)(CMOM + l) * N + o] = 1.0 * dudt[(size_t)(CMOM + l) * N + o] + 1.0 * gsrc[(size_t)l * N + o];
  const Real w2 = 1.0 - beta;
  for (long o = 0; o < (long)L.ncons * N; ++o) u0[o] = beta * u0[o] + w2 * ubase[o]; // AverageIndependentData (in place)
  for (long o = 0; o < (long)L.ncons * N; ++o) uout[o] = 1.0 * u0[o] + bdt * dudt[o]; // UpdateIndependentData
}

// fluid.cpp:477-517
void orc_cons_to_prim(const orc_cfg *c, const double *geom, double *cons, double *prim, double *temp,
                      double *c2p_mu, double *fail, double *csig, const int ib[2], const int jb[2],
                      const int kb[2]) {
  Lay L(c);
  Geom g(c, geom);
  Bounds bounds(c);
  Eos eos(c);
  const long N = L.ncell;
  for (int k = kb[0]; k <= kb[1]; ++k)
    for (int j = jb[0]; j <= jb[1]; ++j)
      for (int i = ib[0]; i <= ib[1]; ++i) {
        const long o = L.idx(k, j, i);
        CellGeom cg(g, k, j, i);
        ZoneRefs z{cons + o, prim + o, temp + o, c2p_mu + o, csig + o, N};
        const bool ok = c2p_zone(c, bounds, eos, cg, Xc(c, 0, i), Xc(c, 1, j), Xc(c, 2, k), z, L.ndim);
        fail[o] = ok ? 1.0 : 0.0;
      }
}

// fixup_c2p.cpp:170-353
void orc_c2p_fixup(const orc_cfg *c, const double *geom, double *cons, double *prim, double *temp,
                   const double *fail, double *csig) {
  Lay L(c);
  if (!c->enable_c2p_fixup) return;
  Geom g(c, geom);
  Bounds bounds(c);
  Eos eos(c);
  const long N = L.ncell;
  const int mhd = c->mhd, prs = PPRS(mhd), gm1 = PGM1(mhd);
  const long st[3] = {1, L.ni, (long)L.ni * L.nj};
  for (int k = L.ks; k <= L.ke; ++k)
    for (int j = L.js; j <= L.je; ++j)
      for (int i = L.is; i <= L.ie; ++i) {
        const long o = L.idx(k, j, i);
        if (fail[o] != 0.0) continue;
        Real gamma_max, e_max;
        bounds.GetCeilings(Xc(c, 0, i), Xc(c, 1, j), Xc(c, 2, k), gamma_max, e_max);
        auto fixup = [&](int iv, Real inv_mask_sum) {
          Real *v = prim + (size_t)iv * N;
          v[o] = fail[o - 1] * v[o - 1] + fail[o + 1] * v[o + 1];
          if (L.ndim > 1) {
            v[o] += fail[o - st[1]] * v[o - st[1]] + fail[o + st[1]] * v[o + st[1]];
            if (L.ndim == 3) v[o] += fail[o - st[2]] * v[o - st[2]] + fail[o + st[2]] * v[o + st[2]];
          }
          return inv_mask_sum * v[o];
        };
        Real num_valid = fail[o - 1] + fail[o + 1];
        if (L.ndim > 1) num_valid += fail[o - st[1]] + fail[o + st[1]];
        if (L.ndim == 3) num_valid += fail[o - st[2]] + fail[o + st[2]];
        Real &rho = prim[(size_t)PRHO * N + o], &u = prim[(size_t)PENG * N + o];
        if (num_valid > 0.5 && c->c2p_failure_strategy == 0) {
          const Real norm = 1.0 / num_valid;
          rho = fixup(PRHO, norm);
          for (int pv = PVEL; pv <= PVEL + 2; pv++) prim[(size_t)pv * N + o] = fixup(pv, norm);
          u = fixup(PENG, norm);
          rho = std::max(rho, 100. * R_SMALL);
          u = std::max(u, 100. * R_SMALL);
        } else {
          rho = 100. * R_SMALL;
          u = 100. * R_SMALL;
          for (int ii = 0; ii < 3; ++ii) prim[(size_t)(PVEL + ii) * N + o] = 0.;
        }
        CellGeom cg(g, k, j, i);
        Real vpcon[3] = {prim[(size_t)PVEL * N + o], prim[(size_t)(PVEL + 1) * N + o], prim[(size_t)(PVEL + 2) * N + o]};
        Real W = LorentzFactor4(vpcon, cg.gcov4);
        if (W > gamma_max) {
          const Real rescale = std::sqrt(gamma_max * gamma_max - 1.) / (W * W - 1.); // sic, fixup_c2p.cpp:271
          for (int ii = 0; ii < 3; ++ii) vpcon[ii] *= rescale;
          for (int ii = 0; ii < 3; ++ii) prim[(size_t)(PVEL + ii) * N + o] = vpcon[ii];
          W = gamma_max;
        }
        temp[o] = eos.T_rho_sie(rho, ratio(u, rho));
        prim[(size_t)prs * N + o] = eos.P_rho_T(rho, temp[o]);
        prim[(size_t)gm1 * N + o] = ratio(eos.K_rho_T(rho, temp[o]), prim[(size_t)prs * N + o]);
        Real S[3], bcons[3], bp[3] = {0.0, 0.0, 0.0}, sig[3];
        if (mhd)
          for (int a = 0; a < 3; ++a) bp[a] = prim[(size_t)(PB + a) * N + o];
        p2c(rho, vpcon, bp, u, prim[(size_t)prs * N + o], prim[(size_t)gm1 * N + o], cg.gcov4, cg.gcon, cg.beta,
            cg.lapse, cg.gdet, cons[(size_t)CRHO * N + o], S, bcons, cons[(size_t)CENG * N + o], sig);
        for (int a = 0; a < 3; ++a) cons[(size_t)(CMOM + a) * N + o] = S[a];
        for (int a = 0; a < L.ndim; ++a) csig[(size_t)a * N + o] = sig[a];
      }
}

// fixup.cpp:370-684 (fluid part)
void orc_apply_floors(const orc_cfg *c, const double *geom, double *cons, double *prim, double *temp,
                      double *c2p_mu, double *csig, long long *counters) {
  Lay L(c);
  if (!c->enable_floors) return;
  Geom g(c, geom);
  Bounds bounds(c);
  Eos eos(c);
  const long N = L.ncell;
  const int mhd = c->mhd, prs = PPRS(mhd), gm1 = PGM1(mhd);
  for (int k = 0; k < L.nk; ++k)
    for (int j = 0; j < L.nj; ++j)
      for (int i = 0; i < L.ni; ++i) {
        const long o = L.idx(k, j, i);
        auto PR = [&](int n) -> Real & { return prim[(size_t)n * N + o]; };
        auto CO = [&](int n) -> Real & { return cons[(size_t)n * N + o]; };
        const Real x1 = Xc(c, 0, i), x2 = Xc(c, 1, j), x3 = Xc(c, 2, k);
        Real rho_floor, sie_floor;
        bounds.GetFloors(x1, x2, x3, rho_floor, sie_floor);
        Real gamma_max, e_max;
        bounds.GetCeilings(x1, x2, x3, gamma_max, e_max);
        Real bsqorho_max, bsqou_max;
        bounds.GetMHDCeilings(x1, x2, x3, bsqorho_max, bsqou_max);
        Real rho_floor_max = rho_floor;
        Real u_floor_max = rho_floor * sie_floor;
        bool floor_applied = false, ceiling_applied = false;
        CellGeom cg(g, k, j, i);
        const Real alpha = cg.lapse;
        const Real sdetgam = cg.gdet;
        if (c->enable_mhd_floors) {
          Real Bsq = 0.0, Bdotv = 0.0;
          const Real vp[3] = {PR(PVEL), PR(PVEL + 1), PR(PVEL + 2)};
          // fixup.cpp:416-417 indexes components 1,2 of block 0: identical for per-block data
          const Real bp[3] = {PR(PB), PR(PB + 1), PR(PB + 2)};
          const Real W = LorentzFactor4(vp, cg.gcov4);
          const Real iW = 1.0 / W;
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
              Bsq += cg.gcov4[a + 1][b + 1] * bp[a] * bp[b];
              Bdotv += cg.gcov4[a + 1][b + 1] * bp[a] * vp[b];
            }
          Real bcon0 = W * Bdotv / alpha;
          const Real bsq = (Bsq + alpha * alpha * bcon0 * bcon0) * iW * iW;
          rho_floor_max = std::max(rho_floor_max, bsq / bsqorho_max);
          u_floor_max = std::max(u_floor_max, bsq / bsqou_max);
          rho_floor_max = std::max(rho_floor_max, std::max(PR(PENG), u_floor_max) / e_max);
        }
        Real drho = rho_floor_max - PR(PRHO);
        Real du = u_floor_max - PR(PENG);
        if (drho > 0. || du > 0.) {
          floor_applied = true;
          drho = std::max(drho, du / sie_floor);
          du = std::max(du, sie_floor * drho);
        }
        Real dcrho, dS[3], dBcons[3], dtau;
        Real bp[3] = {0, 0, 0};
        if (mhd)
          for (int a = 0; a < 3; ++a) bp[a] = PR(PB + a);
        if (floor_applied) {
          if (counters) counters[0]++;
          Real vp_normalobs[3] = {0, 0, 0};
          Real dprs = eos.P_rho_sie(drho, ratio(du, drho));
          Real dgm1 = ratio(eos.K_rho_sie(drho, ratio(du, drho)), dprs);
          p2c(drho, vp_normalobs, bp, du, dprs, dgm1, cg.gcov4, cg.gcon, cg.beta, alpha, sdetgam, dcrho, dS, dBcons,
              dtau, nullptr);
          CO(CRHO) += dcrho;
          for (int a = 0; a < 3; ++a) CO(CMOM + a) += dS[a];
          CO(CENG) += dtau;
          ZoneRefs z{cons + o, prim + o, temp + o, c2p_mu + o, csig + o, N};
          const bool ok = c2p_zone(c, bounds, eos, cg, x1, x2, x3, z, L.ndim);
          if (!ok) {
            if (counters) counters[2]++;
            PR(PRHO) = drho;
            for (int a = 0; a < 3; ++a) PR(PVEL + a) = vp_normalobs[a];
            PR(PENG) = du;
            temp[o] = eos.T_rho_sie(PR(PRHO), PR(PENG) / PR(PRHO));
            p2c(drho, vp_normalobs, bp, du, dprs, dgm1, cg.gcov4, cg.gcon, cg.beta, alpha, sdetgam, CO(CRHO), dS,
                dBcons, CO(CENG), nullptr);
            for (int a = 0; a < 3; ++a) CO(CMOM + a) = dS[a];
          }
        }
        Real vpcon[3] = {PR(PVEL), PR(PVEL + 1), PR(PVEL + 2)};
        const Real W = LorentzFactor4(vpcon, cg.gcov4);
        if (W > gamma_max || PR(PENG) / PR(PRHO) > e_max) ceiling_applied = true;
        if (ceiling_applied) {
          if (counters) counters[1]++;
          const Real rescale = std::sqrt(gamma_max * gamma_max - 1.) / (W * W - 1.); // sic, fixup.cpp:510
          for (int a = 0; a < 3; ++a) vpcon[a] *= rescale;
          for (int a = 0; a < 3; ++a) PR(PVEL + a) = vpcon[a];
          p2c(PR(PRHO), vpcon, bp, PR(PENG), PR(prs), PR(gm1), cg.gcov4, cg.gcon, cg.beta, alpha, sdetgam, CO(CRHO),
              dS, dBcons, CO(CENG), nullptr);
          for (int a = 0; a < 3; ++a) CO(CMOM + a) = dS[a];
        }
        if (floor_applied || ceiling_applied) {
          temp[o] = eos.T_rho_sie(PR(PRHO), ratio(PR(PENG), PR(PRHO)));
          PR(prs) = eos.P_rho_T(PR(PRHO), temp[o]);
          PR(gm1) = ratio(eos.K_rho_T(PR(PRHO), temp[o]), PR(prs));
        }
      }
}

// fluid.cpp:368-449
void orc_prim_to_cons(const orc_cfg *c, const double *geom, const double *prim, double *cons, double *csig,
                      const int ib[2], const int jb[2], const int kb[2]) {
  Lay L(c);
  Geom g(c, geom);
  const long N = L.ncell;
  const int mhd = c->mhd, prs = PPRS(mhd), gm1 = PGM1(mhd);
  for (int k = kb[0]; k <= kb[1]; ++k)
    for (int j = jb[0]; j <= jb[1]; ++j)
      for (int i = ib[0]; i <= ib[1]; ++i) {
        const long o = L.idx(k, j, i);
        CellGeom cg(g, k, j, i);
        Real S[3], bcons[3], sig[3], bp[3] = {0, 0, 0};
        const Real vel[3] = {prim[(size_t)PVEL * N + o], prim[(size_t)(PVEL + 1) * N + o], prim[(size_t)(PVEL + 2) * N + o]};
        if (mhd)
          for (int a = 0; a < 3; ++a) bp[a] = prim[(size_t)(PB + a) * N + o];
        p2c(prim[(size_t)PRHO * N + o], vel, bp, prim[(size_t)PENG * N + o], prim[(size_t)prs * N + o],
            prim[(size_t)gm1 * N + o], cg.gcov4, cg.gcon, cg.beta, cg.lapse, cg.gdet, cons[(size_t)CRHO * N + o], S,
            bcons, cons[(size_t)CENG * N + o], sig);
        for (int a = 0; a < 3; ++a) cons[(size_t)(CMOM + a) * N + o] = S[a];
        if (mhd)
          for (int a = 0; a < 3; ++a) cons[(size_t)(CB + a) * N + o] = bcons[a];
        for (int a = 0; a < L.ndim; ++a) csig[(size_t)a * N + o] = sig[a];
      }
}

// phoebus_boundaries.cpp:44-448: ghost fills at physical boundaries, applied face by face in
// the order ix1, ox1, ix2, ox2, ix3, ox3 over the ENTIRE transverse extent
// (external/parthenon/src/mesh/domain.hpp:180-240), followed by ConvertBoundaryConditions.
void orc_apply_bcs(const orc_cfg *c, const double *geom, double *cons, double *prim, double *temp,
                   double *c2p_mu, double *fail, double *csig) {
  orc_apply_bcs_ex(c, geom, cons, prim, temp, c2p_mu, fail, csig, 3);
}
// what: bit 0 = the ghost fills (parthenon::ApplyBoundaryConditions), bit 1 = ConvertBoundaryConditions
void orc_apply_bcs_ex(const orc_cfg *c, const double *geom, double *cons, double *prim, double *temp,
                      double *c2p_mu, double *fail, double *csig, int what) {
  Lay L(c);
  Geom g(c, geom);
  const long N = L.ncell;
  const int mhd = c->mhd;
  const int lo[3] = {L.is, L.js, L.ks}, hi[3] = {L.ie, L.je, L.ke}, nn[3] = {L.ni, L.nj, L.nk};
  for (int f = 0; f < 2 * L.ndim && (what & 1); ++f) {
    const int bc = c->bc[f];
    if (bc <= 0) continue; // 0: inter-block face, -1: periodic mesh face (filled by the halo exchange)
    const int dir = f / 2, outer = f % 2;
    int b0[3] = {0, 0, 0}, b1[3] = {L.ni - 1, L.nj - 1, L.nk - 1};
    if (outer) b0[dir] = hi[dir] + 1; else b1[dir] = lo[dir] - 1;
    const int ref = outer ? hi[dir] : lo[dir];
    const int offset = 2 * ref + (outer ? 1 : -1);
    for (int k = b0[2]; k <= b1[2]; ++k)
      for (int j = b0[1]; j <= b1[1]; ++j)
        for (int i = b0[0]; i <= b1[0]; ++i) {
          int idx[3] = {i, j, k}, rdx[3] = {i, j, k};
          if (bc == 1) rdx[dir] = ref;                 // outflow
          else rdx[dir] = offset - idx[dir];           // reflect / polar (jref = -j + 2 j0 - 1)
          const long o = L.idx(k, j, i), r = L.idx(rdx[2], rdx[1], rdx[0]);
          if (c->bc_prim) {
            // every FillGhost field is copied (fluid.cpp:156-175); cons are overwritten by p2c below
            for (int n = 0; n < L.nprim; ++n) {
              Real sgn = 1.0;
              if (bc == 2) { // GenericBC reflect: vector component along DIR flips
                if ((n >= PVEL && n < PVEL + 3 && n - PVEL == dir) || (mhd && n >= PB && n < PB + 3 && n - PB == dir)) sgn = -1.0;
              } else if (bc == 3) { // polar: only p.velocity(1), p.bfield(1)
                if (n == PVEL + 1 || (mhd && n == PB + 1)) sgn = -1.0;
              }
              prim[(size_t)n * N + o] = sgn * prim[(size_t)n * N + r];
            }
            temp[o] = temp[r];
            c2p_mu[o] = c2p_mu[r];
            fail[o] = fail[r];
            for (int n = 0; n < L.ncons; ++n) {
              Real sgn = 1.0;
              if (bc == 2 && ((n >= CMOM && n < CMOM + 3 && n - CMOM == dir) || (mhd && n >= CB && n - CB == dir))) sgn = -1.0;
              cons[(size_t)n * N + o] = sgn * cons[(size_t)n * N + r];
            }
          } else {
            const Real gratio = ratio(g.DetGamma(0, k, j, i), g.DetGamma(0, rdx[2], rdx[1], rdx[0]));
            for (int n = 0; n < L.ncons; ++n) {
              Real sgn = 1.0;
              if (bc == 2 && ((n >= CMOM && n < CMOM + 3 && n - CMOM == dir) || (mhd && n >= CB && n - CB == dir))) sgn = -1.0;
              if (bc == 1 && dir == 0) cons[(size_t)n * N + o] = gratio * cons[(size_t)n * N + r]; // OutflowInner/OuterX1
              else cons[(size_t)n * N + o] = sgn * gratio * cons[(size_t)n * N + r];                // GenericBC
            }
          }
        }
    (void)nn;
  }
  if (!(what & 2)) return;
  // ConvertBoundaryConditions, phoebus_boundaries.cpp:303-448
  // c->bc[f] != 0 marks faces with boundary_flag != block is NOT the same thing: the
  // reference tests "!= block", which also holds for periodic faces. Callers encode periodic
  // mesh faces as bc = -1 so they take part here but not in the ghost fills above.
  auto region = [&](int f, int ib[2], int jb[2], int kb[2]) {
    const int dir = f / 2, outer = f % 2;
    int b0[3] = {0, 0, 0}, b1[3] = {L.ni - 1, L.nj - 1, L.nk - 1};
    if (outer) b0[dir] = hi[dir] + 1; else b1[dir] = lo[dir] - 1;
    ib[0] = b0[0]; ib[1] = b1[0]; jb[0] = b0[1]; jb[1] = b1[1]; kb[0] = b0[2]; kb[1] = b1[2];
  };
  if (c->ox1_inflow_check && c->bc[1] != 0) {
    int ib[2], jb[2], kb[2];
    region(1, ib, jb, kb);
    if (!c->bc_prim) orc_cons_to_prim(c, geom, cons, prim, temp, c2p_mu, fail, csig, ib, jb, kb);
    for (int k = kb[0]; k <= kb[1]; ++k)
      for (int j = jb[0]; j <= jb[1]; ++j)
        for (int i = ib[0]; i <= ib[1]; ++i) {
          const long o = L.idx(k, j, i);
          Real vcon[3] = {prim[(size_t)PVEL * N + o], prim[(size_t)(PVEL + 1) * N + o], prim[(size_t)(PVEL + 2) * N + o]};
          Real gammacov[3][3];
          g.Metric(0, k, j, i, gammacov);
          Real W = LorentzFactor3(vcon, gammacov);
          const Real alpha = g.Lapse(0, k, j, i);
          Real beta[3];
          g.Shift(0, k, j, i, beta);
          Real ucon1 = vcon[0] - W * beta[0] / alpha;
          if (ucon1 < 0) {
            for (int a = 0; a < 3; ++a) vcon[a] /= W;
            vcon[0] = beta[0] / alpha;
            Real vsq = 0.;
            for (int a = 0; a < 3; ++a)
              for (int b = 0; b < 3; ++b) vsq += gammacov[a][b] * vcon[a] * vcon[b];
            W = 1. / std::sqrt(1. - vsq);
            for (int a = 0; a < 3; ++a) prim[(size_t)(PVEL + a) * N + o] = W * vcon[a];
          }
        }
    orc_prim_to_cons(c, geom, prim, cons, csig, ib, jb, kb);
  }
  if (c->bc_prim) {
    for (int f = 0; f < 2 * L.ndim; ++f) {
      if (c->bc[f] == 0) continue;
      int ib[2], jb[2], kb[2];
      region(f, ib, jb, kb);
      orc_prim_to_cons(c, geom, prim, cons, csig, ib, jb, kb);
    }
  }
}

// fluid.cpp:899-944
double orc_estimate_dt(const orc_cfg *c, const double *csig, const double *fsig, int has_face_speeds) {
  Lay L(c);
  const long N = L.ncell;
  const long st[3] = {1, L.ni, (long)L.ni * L.nj};
  Real min_dt = std::numeric_limits<Real>::max();
  for (int k = L.ks; k <= L.ke; ++k)
    for (int j = L.js; j <= L.je; ++j)
      for (int i = L.is; i <= L.ie; ++i) {
        const long o = L.idx(k, j, i);
        Real ldt = 0.0;
        for (int d = 0; d < L.ndim; d++) {
          Real max_s = csig[(size_t)d * N + o];
          if (has_face_speeds) max_s = std::max(max_s, std::max(fsig[(size_t)d * N + o], fsig[(size_t)d * N + o + st[d]]));
          ldt += max_s / c->dx[d];
        }
        min_dt = std::min(min_dt, 1.0 / (ldt + 1.e-50));
      }
  return min_dt;
}

} // extern "C"
