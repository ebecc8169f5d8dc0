// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's geometry providers used on
// the hot path: Analytic<Minkowski> and FMKS = Modified<SphericalKerrSchild,
// McKinneyGammieRyan>, tabulated the way Geometry::SetCachedCoordinateSystem does
// (src/geometry/cached_system.hpp:593-775). See oracle.h.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {
typedef double Real;
constexpr Real R_SMALL = 10 * std::numeric_limits<Real>::min();
inline Real ratio(Real a, Real b) {
  const Real sgn = b >= 0 ? 1 : -1;
  return a / (b + sgn * R_SMALL);
}

// src/geometry/spherical_kerr_schild.hpp:42-244
struct SKS {
  Real a_, a2_;
  explicit SKS(Real a) : a_(a), a2_(a * a) {}
  Real rho2_(Real r, Real cth) const { return r * r + a2_ * cth * cth; }
  void vars_(Real X1, Real X2, Real &r, Real &th, Real &cth, Real &sth, Real &sth2, Real &rho2) const {
    r = std::abs(X1);
    th = X2;
    cth = std::cos(th);
    sth = std::sin(th);
    sth2 = sth * sth;
    rho2 = rho2_(r, cth);
  }
  Real Lapse(Real X1, Real X2) const {
    const Real r = std::abs(X1);
    const Real cth = std::cos(X2);
    const Real rho2 = rho2_(r, cth);
    return std::sqrt(ratio(rho2, 2 * r + rho2));
  }
  void Shift(Real X1, Real X2, Real beta[3]) const {
    Real r = std::abs(X1), cth = std::cos(X2);
    const Real rho2 = rho2_(r, cth);
    beta[0] = beta[1] = beta[2] = 0;
    beta[0] = ratio(2 * r, 2 * r + rho2);
  }
  void SpacetimeMetric(Real X1, Real X2, Real g[4][4]) const {
    Real r, th, cth, sth, sth2, rho2;
    vars_(X1, X2, r, th, cth, sth, sth2, rho2);
    std::memset(g, 0, 16 * sizeof(Real));
    g[0][1] = g[1][0] = ratio(2 * r, rho2);
    g[0][0] = -1 + g[0][1];
    g[0][3] = g[3][0] = -ratio(2 * a_ * r * sth2, rho2);
    g[1][1] = 1 + g[0][1];
    g[1][3] = g[3][1] = -a_ * (1 + g[0][1]) * sth2;
    g[2][2] = rho2;
    g[3][3] = (rho2 + a2_ * (1 + g[0][1]) * sth2) * sth2;
  }
  void MetricInverse(Real X1, Real X2, Real g[3][3]) const {
    Real r, th, cth, sth, sth2, rho2;
    vars_(X1, X2, r, th, cth, sth, sth2, rho2);
    std::memset(g, 0, 9 * sizeof(Real));
    g[0][0] = ratio(rho2, 2 * r + rho2) + ratio(a2_ * sth2, rho2);
    g[0][2] = g[2][0] = ratio(a_, rho2);
    g[1][1] = ratio(1, rho2);
    g[2][2] = ratio(1, rho2 * sth2);
  }
  Real DetGamma(Real X1, Real X2) const {
    Real r, th, cth, sth, sth2, rho2;
    vars_(X1, X2, r, th, cth, sth, sth2, rho2);
    return std::sqrt(std::abs(rho2 * (2 * r + rho2) * sth2));
  }
};

// src/geometry/mckinney_gammie_ryan.hpp:44-210
struct MGR {
  bool derefine_poles_;
  Real h_, xt_, alpha_, x0_, smooth_, norm_, hexp_br_, hexp_nsq_, hexp_csq_;
  explicit MGR(const Real *p)
      : derefine_poles_(p[9] != 0), h_(p[1]), xt_(p[2]), alpha_(p[3]), x0_(p[4]), smooth_(p[5]), hexp_br_(p[6]),
        hexp_nsq_(p[7]), hexp_csq_(p[8]) {
    norm_ = 0.5 * M_PI * 1. / (1. + 1. / (alpha_ + 1.) * 1. / std::pow(xt_, alpha_));
  }
  Real r_(Real X1) const {
    const Real hexp_dr = X1 - hexp_br_;
    return std::exp(X1 + (hexp_dr > 0.) * hexp_csq_ * std::pow(hexp_dr, hexp_nsq_));
  }
  Real thG_(Real X2) const { return M_PI * X2 + ((1. - h_) / 2.) * std::sin(2. * M_PI * X2); }
  void thJ_(Real X2, Real &y, Real &thJ) const {
    y = 2. * X2 - 1.;
    thJ = norm_ * y * (1. + std::pow(y / xt_, alpha_) / (alpha_ + 1.)) + 0.5 * M_PI;
  }
  void th_(Real X1, Real X2, Real &y, Real &th) const {
    Real thG = thG_(X2);
    if (derefine_poles_) {
      Real thJ;
      thJ_(X2, y, thJ);
      const Real fac = x0_ - X1;
      th = thG + std::exp(smooth_ * fac) * (thJ - thG);
    } else {
      y = 0;
      th = thG;
    }
    constexpr Real SINGSMALL = 1.e-20;
    if (std::fabs(th) < SINGSMALL) {
      if (th >= 0.) th = SINGSMALL;
      else if (th < 0) th = -SINGSMALL;
    }
    if (std::fabs(M_PI - th) < SINGSMALL) {
      if (th >= M_PI) th = M_PI + SINGSMALL;
      else if (th < M_PI) th = M_PI - SINGSMALL;
    }
  }
  void operator()(Real X1, Real X2, Real X3, Real C[3], Real Jcov[3][3], Real Jcon[3][3]) const {
    Real y, th, thJ;
    const Real r = r_(X1);
    th_(X1, X2, y, th);
    C[0] = r;
    C[1] = th;
    C[2] = X3;
    const Real hexp_dr = X1 - hexp_br_;
    const Real drdX1 = std::exp(X1 + (hexp_dr > 0.) * hexp_csq_ * std::pow(hexp_dr, hexp_nsq_)) *
                       (1. + (hexp_dr > 0.) * hexp_csq_ * hexp_nsq_ * std::pow(hexp_dr, hexp_nsq_ - 1.));
    const Real dthGdX2 = M_PI + M_PI * (1 - h_) * std::cos(2 * M_PI * X2);
    Real dthdX1, dthdX2;
    if (derefine_poles_) {
      thJ_(X2, y, thJ);
      const Real thG = thG_(X2);
      const Real dydX2 = 2.;
      const Real dthJdy = norm_ * (1 + std::pow(y / xt_, alpha_));
      const Real dthJdX2 = dthJdy * dydX2;
      const Real fac = x0_ - X1;
      dthdX1 = -smooth_ * (thJ - thG) * std::exp(smooth_ * fac);
      dthdX2 = dthGdX2 + std::exp(smooth_ * fac) * (dthJdX2 - dthGdX2);
    } else {
      dthdX1 = 0.0;
      dthdX2 = dthGdX2;
    }
    std::memset(Jcov, 0, 9 * sizeof(Real));
    std::memset(Jcon, 0, 9 * sizeof(Real));
    Jcov[0][0] = drdX1;
    Jcov[1][0] = dthdX1;
    Jcov[1][1] = dthdX2;
    Jcov[2][2] = 1.;
    Jcon[0][0] = ratio(1., drdX1);
    Jcon[1][0] = -ratio(dthdX1, drdX1 * dthdX2);
    Jcon[1][1] = ratio(1., dthdX2);
    Jcon[2][2] = 1.;
  }
};

inline Real S2ST(const Real A[3][3], int mu, int nu) {
  if (mu == 0 || nu == 0) return (mu == nu) ? 1 : 0;
  return A[mu - 1][nu - 1];
}
// linear_algebra.hpp:48-54
inline Real Det3(const Real A[3][3]) {
  return (A[0][0] * A[1][1] * A[2][2] + A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] -
          A[0][2] * A[1][1] * A[2][0] - A[0][1] * A[1][0] * A[2][2] - A[0][0] * A[1][2] * A[2][1]);
}

struct PointGeom {
  Real alpha, detgam, beta[3], gcov[4][4], gamcon[3][3], dalpha[4], dg[4][4][4];
};

// src/geometry/modified_system.hpp:53-246
struct FMKS {
  SKS s_;
  MGR t_;
  explicit FMKS(const Real *p) : s_(p[0]), t_(p) {}
  Real Lapse(Real X1, Real X2, Real X3) const {
    Real C[3], Jcov[3][3], Jcon[3][3];
    t_(X1, X2, X3, C, Jcov, Jcon);
    return s_.Lapse(C[0], C[1]);
  }
  void Shift(Real X1, Real X2, Real X3, Real beta[3]) const {
    Real beta0[3], C[3], Jcov[3][3], Jcon[3][3];
    t_(X1, X2, X3, C, Jcov, Jcon);
    s_.Shift(C[0], C[1], beta0);
    for (int i = 0; i < 3; ++i) {
      beta[i] = 0;
      for (int j = 0; j < 3; ++j) beta[i] += Jcon[i][j] * beta0[j];
    }
  }
  void SpacetimeMetric(Real X1, Real X2, Real X3, Real g[4][4]) const {
    Real g0[4][4], C[3], Jcov[3][3], Jcon[3][3];
    t_(X1, X2, X3, C, Jcov, Jcon);
    s_.SpacetimeMetric(C[0], C[1], g0);
    std::memset(g, 0, 16 * sizeof(Real));
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = 0; nu < 4; ++nu)
        for (int lam = 0; lam < 4; ++lam)
          for (int kap = 0; kap < 4; ++kap) g[mu][nu] += g0[lam][kap] * S2ST(Jcov, lam, mu) * S2ST(Jcov, kap, nu);
  }
  void MetricInverse(Real X1, Real X2, Real X3, Real g[3][3]) const {
    Real g0[3][3], C[3], Jcov[3][3], Jcon[3][3];
    t_(X1, X2, X3, C, Jcov, Jcon);
    s_.MetricInverse(C[0], C[1], g0);
    std::memset(g, 0, 9 * sizeof(Real));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l) g[i][j] += g0[k][l] * Jcon[i][k] * Jcon[j][l];
  }
  Real DetGamma(Real X1, Real X2, Real X3) const {
    Real C[3], Jcov[3][3], Jcon[3][3];
    t_(X1, X2, X3, C, Jcov, Jcon);
    Real detJ = Det3(Jcov);
    return std::abs(s_.DetGamma(C[0], C[1]) * detJ);
  }
  // geometry_utils.hpp:150-185
  void MetricDerivative(Real X1, Real X2, Real X3, Real dg[4][4][4]) const {
    Real EPS = std::pow(std::numeric_limits<Real>::epsilon(), 1.0 / 3.0);
    std::memset(dg, 0, 64 * sizeof(Real));
    Real gl[4][4], gr[4][4];
    for (int d = 1; d < 4; ++d) {
      Real X1L = X1, X2L = X2, X3L = X3, X1R = X1, X2R = X2, X3R = X3, dx = 0;
      if (d == 1) { dx = std::max(std::abs(X1), 1.0) * EPS; X1L -= dx; X1R += dx; }
      if (d == 2) { dx = std::max(std::abs(X2), 1.0) * EPS; X2L -= dx; X2R += dx; }
      if (d == 3) { dx = std::max(std::abs(X3), 1.0) * EPS; X3L -= dx; X3R += dx; }
      SpacetimeMetric(X1R, X2R, X3R, gr);
      SpacetimeMetric(X1L, X2L, X3L, gl);
      for (int mu = 0; mu < 4; ++mu)
        for (int nu = 0; nu < 4; ++nu) dg[mu][nu][d] = ratio(gr[mu][nu] - gl[mu][nu], 2 * dx);
    }
  }
  // geometry_utils.hpp:91-121
  void GradLnAlpha(Real X1, Real X2, Real X3, Real da[4]) const {
    Real EPS = std::pow(std::numeric_limits<Real>::epsilon(), 1.0 / 3.0);
    da[0] = da[1] = da[2] = da[3] = 0;
    for (int d = 1; d < 4; ++d) {
      Real X1p = X1, X1m = X1, X2p = X2, X2m = X2, X3p = X3, X3m = X3, dx = 0;
      if (d == 1) { dx = std::max(std::abs(X1), 1.0) * EPS; X1p += dx; X1m -= dx; }
      else if (d == 2) { dx = std::max(std::abs(X2), 1.0) * EPS; X2p += dx; X2m -= dx; }
      else if (d == 3) { dx = std::max(std::abs(X3), 1.0) * EPS; X3p += dx; X3m -= dx; }
      Real alpham = Lapse(X1m, X2m, X3m);
      Real alphap = Lapse(X1p, X2p, X3p);
      da[d] = ratio(alphap - alpham, dx * (alpham + alphap));
    }
  }
};

void point(int kind, const Real *par, Real X1, Real X2, Real X3, bool derivs, PointGeom &p) {
  if (kind == 0) { // minkowski.hpp:35-115
    p.alpha = 1.;
    p.detgam = 1.;
    for (int i = 0; i < 3; ++i) p.beta[i] = 0;
    for (int m = 0; m < 4; ++m)
      for (int n = 0; n < 4; ++n) p.gcov[m][n] = (m == n) ? (m == 0 ? -1 : 1) : 0;
    for (int m = 0; m < 3; ++m)
      for (int n = 0; n < 3; ++n) p.gamcon[m][n] = (m == n ? 1 : 0);
    std::memset(p.dalpha, 0, sizeof(p.dalpha));
    std::memset(p.dg, 0, sizeof(p.dg));
  } else {
    FMKS s(par);
    p.alpha = s.Lapse(X1, X2, X3);
    p.detgam = s.DetGamma(X1, X2, X3);
    s.Shift(X1, X2, X3, p.beta);
    s.SpacetimeMetric(X1, X2, X3, p.gcov);
    s.MetricInverse(X1, X2, X3, p.gamcon);
    if (derivs) {
      s.GradLnAlpha(X1, X2, X3, p.dalpha);
      s.MetricDerivative(X1, X2, X3, p.dg);
    }
  }
}
const int FLAT[4][4] = {{0, 1, 3, 6}, {1, 2, 4, 7}, {3, 4, 5, 8}, {6, 7, 8, 9}};
} // namespace

extern "C" {

void orc_geometry_point(int kind, const double *par, double X1, double X2, double X3, double *out) {
  PointGeom p;
  point(kind, par, X1, X2, X3, true, p);
  int o = 0;
  out[o++] = p.alpha;
  out[o++] = p.detgam;
  for (int i = 0; i < 3; ++i) out[o++] = p.beta[i];
  for (int m = 0; m < 4; ++m)
    for (int n = 0; n < 4; ++n) out[o++] = p.gcov[m][n];
  for (int m = 0; m < 3; ++m)
    for (int n = 0; n < 3; ++n) out[o++] = p.gamcon[m][n];
  for (int m = 0; m < 4; ++m) out[o++] = p.dalpha[m];
  for (int m = 0; m < 4; ++m)
    for (int n = 0; n < 4; ++n)
      for (int s = 0; s < 4; ++s) out[o++] = p.dg[m][n][s];
}

// cached_system.hpp:640-775 (do_fd_on_grid = false). Table extents come from cfg (gk,gj,gi);
// an extent of 1 evaluates at index 0 of that dimension, as the axisymmetric cache does.
void orc_build_geometry(int kind, const double *par, const orc_cfg *c, double *geom) {
  const int gk = c->gk, gj = c->gj, gi = c->gi;
  const long gcell = (long)gk * gj * gi;
  for (int loc = 0; loc < 4; ++loc)
    for (int k = 0; k < gk; ++k)
      for (int j = 0; j < gj; ++j)
        for (int i = 0; i < gi; ++i) {
          // analytic_system.hpp:59-90 + uniform_cartesian.hpp:86-145
          Real X1 = c->xmin[0] + (i + 0.5) * c->dx[0];
          Real X2 = c->xmin[1] + (j + 0.5) * c->dx[1];
          Real X3 = c->xmin[2] + (k + 0.5) * c->dx[2];
          if (loc == 1) X1 = c->xmin[0] + i * c->dx[0];
          if (loc == 2) X2 = c->xmin[1] + j * c->dx[1];
          if (loc == 3) X3 = c->xmin[2] + k * c->dx[2];
          PointGeom p;
          point(kind, par, X1, X2, X3, loc == 0, p);
          const long cell = ((long)k * gj + j) * gi + i;
          double *t = geom + (long)loc * ORC_NGLOC * gcell + cell;
          t[0] = p.alpha;
          for (int d = 0; d < 3; ++d) t[(1 + d) * gcell] = p.beta[d];
          for (int m = 0; m < 4; ++m)
            for (int n = m; n < 4; ++n) t[(4 + FLAT[m][n]) * gcell] = p.gcov[m][n];
          for (int m = 0; m < 3; ++m)
            for (int n = m; n < 3; ++n) t[(14 + FLAT[m][n]) * gcell] = p.gamcon[m][n];
          t[20 * gcell] = p.detgam;
          if (loc == 0) {
            double *dtab = geom + (long)4 * ORC_NGLOC * gcell + cell;
            for (int d = 1; d < 4; ++d) dtab[(d - 1) * gcell] = p.dalpha[d];
            for (int s = 1; s < 4; ++s)
              for (int m = 0; m < 4; ++m)
                for (int n = m; n < 4; ++n) dtab[(3 + 3 * FLAT[m][n] + (s - 1)) * gcell] = p.dg[m][n][s];
          }
        }
}

// ------------------------------------------------------------------------------------------------
// Fishbone-Moncrief torus problem generator, src/pgen/torus.cpp:130-474 (IdealGas, no radiation): one MeshBlock,
// zone by zone as the reference's three kernels do. prim [nprim][k][j][i] and temp are written over the ENTIRE block;
// the caller continues with orc_prim_to_cons + orc_apply_floors (torus.cpp:471-474).
// tp = {rin, rmax, kappa, Gamma, nsub, u_jitter}; u_jitter must be 0 (the reference's RNG stream is not restated).
int orc_pgen_torus(const orc_cfg *c, const double *par, const double *tp, const double *geom, double *prim, double *temp) {
  const int ng = c->ng, ndim = c->ndim, mhd = c->mhd;
  const int ni = c->nx[0] + 2 * ng, nj = ndim >= 2 ? c->nx[1] + 2 * ng : 1, nk = ndim >= 3 ? c->nx[2] + 2 * ng : 1;
  const long N = (long)ni * nj * nk;
  const int nprim = 7 + 3 * mhd, prs = mhd ? 8 : 5, gm1i = mhd ? 9 : 6;
  const Real a = par[0], rin = tp[0], rmax = tp[1], kappa = tp[2], gam = tp[3], u_jitter = tp[5];
  const int nsub = (int)tp[4];
  if (u_jitter != 0.0 || ndim < 2) return -1;
  // geom.Lapse(0, x1, x2, x3) etc. on the CACHED system interpolate the cell-centred tables linearly:
  // cached_system.hpp:132-139,150-155,229-236 with phoebus_interpolation.hpp:122-219 (LCInterp)
  const long gcellI = (long)c->gk * c->gj * c->gi;
  auto GetWeights = [&](int d, const Real x, const int nx, int &ix, Real w[2]) {
    const Real min = c->xmin[d] + (0 + 0.5) * c->dx[d];
    const Real dx = c->dx[d];
    ix = std::min(std::max(0, static_cast<int>(ratio(x - min, dx))), nx - 2);
    const Real floor = min + ix * dx;
    w[1] = ratio(x - floor, dx);
    w[0] = 1. - w[1];
  };
  auto LCInterp = [&](int v, Real X1, Real X2, Real X3) {
    auto p = [&](int kk, int jj, int ii) { return geom[(long)v * gcellI + ((long)kk * c->gj + jj) * c->gi + ii]; };
    if (c->gk > 1) {
      int ix[3];
      Real w[3][2];
      GetWeights(0, X1, c->gi, ix[0], w[0]);
      GetWeights(1, X2, c->gj, ix[1], w[1]);
      GetWeights(2, X3, c->gk, ix[2], w[2]);
      return (w[2][0] * (w[1][0] * (w[0][0] * p(ix[2], ix[1], ix[0]) + w[0][1] * p(ix[2], ix[1], ix[0] + 1)) +
                         w[1][1] * (w[0][0] * p(ix[2], ix[1] + 1, ix[0]) + w[0][1] * p(ix[2], ix[1] + 1, ix[0] + 1))) +
              w[2][1] * (w[1][0] * (w[0][0] * p(ix[2] + 1, ix[1], ix[0]) + w[0][1] * p(ix[2] + 1, ix[1], ix[0] + 1)) +
                         w[1][1] * (w[0][0] * p(ix[2] + 1, ix[1] + 1, ix[0]) + w[0][1] * p(ix[2] + 1, ix[1] + 1, ix[0] + 1))));
    }
    int ix1, ix2;
    Real w1[2], w2[2];
    GetWeights(0, X1, c->gi, ix1, w1);
    GetWeights(1, X2, c->gj, ix2, w2);
    return (w2[0] * (w1[0] * p(0, ix2, ix1) + w1[1] * p(0, ix2, ix1 + 1)) +
            w2[1] * (w1[0] * p(0, ix2 + 1, ix1) + w1[1] * p(0, ix2 + 1, ix1 + 1)));
  };
  Real trp[10];
  std::memcpy(trp, par, sizeof(trp));
  trp[6] = 1000.; // McKinneyGammieRyan's six-argument constructor, mckinney_gammie_ryan.hpp:58-62 (torus.cpp:244)
  trp[7] = 4.0;
  trp[8] = 4.0;
  const MGR tr(trp);
  auto bl_metric = [&](Real X1, Real X2, Real g[4][4]) { // boyer_lindquist.hpp:79-95,262-280
    const Real a2_ = a * a;
    const Real r = std::abs(X1), th = X2;
    const Real sth = std::sin(th), cth = std::cos(th);
    const Real r2 = r * r, r3 = r2 * r;
    const Real DD = 1. - ratio(2., r) + ratio(a2_, r2);
    const Real mu = 1. + ratio(a2_ * cth * cth, r2);
    const Real sth2 = sth * sth;
    std::memset(g, 0, 16 * sizeof(Real));
    g[0][0] = -(1.0 - ratio(2.0, r * mu));
    g[0][3] = g[3][0] = -ratio(2.0 * a * sth2, r * mu);
    g[1][1] = mu / DD;
    g[2][2] = r2 * mu;
    g[3][3] = r2 * sth2 * (1.0 + ratio(a2_, r2) + ratio(2.0 * a2_ * sth2, r3 * mu));
  };
  auto lfish_calc = [](Real r, Real a) { // torus.cpp:745-757
    return (((std::pow(a, 2) - 2. * a * std::sqrt(r) + std::pow(r, 2)) *
             ((-2. * a * r * (std::pow(a, 2) - 2. * a * std::sqrt(r) + std::pow(r, 2))) /
                  std::sqrt(2. * a * std::sqrt(r) + (-3. + r) * r) +
              ((a + (-2. + r) * std::sqrt(r)) * (std::pow(r, 3) + std::pow(a, 2) * (2. + r))) /
                  std::sqrt(1 + (2. * a) / std::pow(r, 1.5) - 3. / r))) /
            (std::pow(r, 3) * std::sqrt(2. * a * std::sqrt(r) + (-3. + r) * r) * (std::pow(a, 2) + (-2. + r) * r)));
  };
  auto log_enthalpy = [](const Real r, const Real th, const Real a, const Real rin, const Real l, Real &uphi) { // :759-797
    const Real sth = sin(th);
    const Real cth = cos(th);
    const Real DD = r * r - 2. * r + a * a;
    const Real AA = (r * r + a * a) * (r * r + a * a) - DD * a * a * sth * sth;
    const Real SS = r * r + a * a * cth * cth;
    const Real thin = M_PI / 2.;
    const Real sthin = sin(thin);
    const Real cthin = cos(thin);
    const Real DDin = rin * rin - 2. * rin + a * a;
    const Real AAin = (rin * rin + a * a) * (rin * rin + a * a) - DDin * a * a * sthin * sthin;
    const Real SSin = rin * rin + a * a * cthin * cthin;
    uphi = 0.0;
    const Real lnh =
        0.5 * std::log((1. + std::sqrt(1. + 4. * (l * l * SS * SS) * DD / (AA * AA * sth * sth))) / (SS * DD / AA)) -
        0.5 * std::sqrt(1. + 4. * (l * l * SS * SS) * DD / (AA * AA * sth * sth)) - 2. * a * r * l / AA -
        (0.5 * std::log((1. + std::sqrt(1. + 4. * (l * l * SSin * SSin) * DDin / (AAin * AAin * sthin * sthin))) /
                        (SSin * DDin / AAin)) -
         0.5 * std::sqrt(1. + 4. * (l * l * SSin * SSin) * DDin / (AAin * AAin * sthin * sthin)) - 2. * a * rin * l / AAin);
    if (lnh > 0.0) {
      Real expm2chi = SS * SS * DD / (AA * AA * sth * sth);
      Real up1 = std::sqrt((-1. + std::sqrt(1. + 4. * l * l * expm2chi)) / 2.);
      uphi = 2. * a * r * std::sqrt(1. + up1 * up1) / std::sqrt(AA * SS * DD) + std::sqrt(SS / AA) * up1 / sth;
    }
    return lnh;
  };
  auto state = [&](Real hm1, Real rho_rmax, Real &rho_out, Real &u_out) { // GetStateFromEnthalpy, :803-812
    rho_out = std::pow(hm1 * (gam - 1.) / (kappa * gam), 1. / (gam - 1.));
    u_out = kappa * std::pow(rho_out, gam) / (gam - 1.) / rho_rmax;
    rho_out /= rho_rmax;
  };
  auto ucon_norm = [](Real ucon[4], Real gcov[4][4]) { // :834-847
    Real AA = gcov[0][0];
    Real BB = 2. * (gcov[0][1] * ucon[1] + gcov[0][2] * ucon[2] + gcov[0][3] * ucon[3]);
    Real CC = 1. + gcov[1][1] * ucon[1] * ucon[1] + gcov[2][2] * ucon[2] * ucon[2] + gcov[3][3] * ucon[3] * ucon[3] +
              2. * (gcov[1][2] * ucon[1] * ucon[2] + gcov[1][3] * ucon[1] * ucon[3] + gcov[2][3] * ucon[2] * ucon[3]);
    Real discr = BB * BB - 4. * AA * CC;
    return (-BB - std::sqrt(discr)) / (2. * AA);
  };
  auto floors = [&](Real x1, Real &rflr, Real &sflr) { // Floors::GetFloors, fixup.hpp:92-134 (x1-dependent types)
    switch (c->floor_type) {
    case 0: rflr = c->rho0; sflr = c->sie0; break;
    case 1: rflr = c->rho0 * std::exp(c->rexp * x1); sflr = c->sie0 * std::exp(c->sexp * x1); break;
    case 2: { Real scratch = c->rho0 * std::exp(c->rexp * x1); sflr = c->sie0 * std::exp(c->sexp * x1) / std::max(rflr, scratch); rflr = scratch; } break;
    default: rflr = c->rho0 * std::min(1.0, std::pow(x1, c->rexp)); sflr = c->sie0 * std::min(1.0, std::pow(x1, c->sexp)); break;
    }
    rflr = std::max(rflr, c->rho_min_eos);
    sflr = std::max(sflr, c->sie_min_eos);
  };
  const Real angular_mom = lfish_calc(rmax, a);
  Real uphi_rmax;
  const Real hm1_rmax = std::exp(log_enthalpy(rmax, 0.5 * M_PI, a, rin, angular_mom, uphi_rmax)) - 1.0;
  Real rho_rmax, u_rmax;
  state(hm1_rmax, 1.0, rho_rmax, u_rmax);
  auto V = [&](int n, int k, int j, int i) -> double & { return prim[(long)n * N + ((long)k * nj + j) * ni + i]; };
  for (long q = 0; q < (long)nprim * N; ++q) prim[q] = 0.0;
  for (int k = 0; k < nk; ++k)
    for (int j = 0; j < nj; ++j)
      for (int i = 0; i < ni; ++i) {
        const Real dx_sub = c->dx[0] / nsub, dy_sub = c->dx[1] / nsub;
        const Real x3 = c->xmin[2] + (k + 0.5) * c->dx[2];
        for (int m = 0; m < nsub; m++)
          for (int n = 0; n < nsub; n++) {
            const Real x1 = (c->xmin[0] + i * c->dx[0]) + (m + 0.5) * dx_sub;
            const Real x2 = (c->xmin[1] + j * c->dx[1]) + (n + 0.5) * dy_sub;
            Real y, th;
            const Real r = tr.r_(x1);
            tr.th_(x1, x2, y, th);
            Real lnh = -1.0, uphi;
            if (r > rin) lnh = log_enthalpy(r, th, a, rin, angular_mom, uphi);
            if (lnh > 0.0) {
              Real hm1 = std::exp(lnh) - 1.;
              Real rho, u;
              state(hm1, rho_rmax, rho, u);
              Real ucon_bl[] = {0.0, 0.0, 0.0, uphi};
              Real gcov[4][4];
              bl_metric(r, th, gcov);
              ucon_bl[0] = ucon_norm(ucon_bl, gcov);
              // bl_to_fmks, mckinney_gammie_ryan.hpp:119-151
              Real trans[4][4], ucon_ks[4] = {0, 0, 0, 0}, ucon[4] = {0, 0, 0, 0};
              std::memset(trans, 0, sizeof(trans));
              const Real idenom = 1.0 / (r * r - 2.0 * r + a * a);
              trans[0][0] = 1.0;
              trans[0][1] = 2.0 * r * idenom;
              trans[1][1] = 1.0;
              trans[2][2] = 1.0;
              trans[3][1] = a * idenom;
              trans[3][3] = 1.0;
              for (int mu = 0; mu < 4; ++mu)
                for (int nu = 0; nu < 4; ++nu) ucon_ks[mu] += trans[mu][nu] * ucon_bl[nu];
              Real C[3], Jcov[3][3], Jcon[3][3];
              tr(x1, x2, x3, C, Jcov, Jcon);
              std::memset(trans, 0, sizeof(trans));
              trans[0][0] = 1.0;
              for (int p = 0; p < 3; ++p)
                for (int qq = 0; qq < 3; ++qq) trans[p + 1][qq + 1] = Jcon[p][qq];
              for (int mu = 0; mu < 4; ++mu)
                for (int nu = 0; nu < 4; ++nu) ucon[mu] += trans[mu][nu] * ucon_ks[nu];
              const Real lapse = LCInterp(0, x1, x2, x3);
              Real beta[3];
              for (int d = 0; d < 3; ++d) beta[d] = LCInterp(1 + d, x1, x2, x3);
              for (int mu = 0; mu < 4; ++mu)
                for (int nu = 0; nu < 4; ++nu) gcov[mu][nu] = LCInterp(4 + FLAT[mu][nu], x1, x2, x3);
              ucon[0] = ucon_norm(ucon, gcov);
              const Real W = lapse * ucon[0];
              Real vcon[] = {ucon[1] / W + beta[0] / lapse, ucon[2] / W + beta[1] / lapse, ucon[3] / W + beta[2] / lapse};
              V(0, k, j, i) += rho / (nsub * nsub);
              V(4, k, j, i) += u / (nsub * nsub);
              for (int d = 0; d < 3; d++) V(1 + d, k, j, i) += W * vcon[d] / (nsub * nsub);
            }
          }
        const Real x1v = c->xmin[0] + (i + 0.5) * c->dx[0];
        V(4, k, j, i) *= (1. + u_jitter * (0.5 - 0.5));
        Real rhoflr = 0, epsflr;
        floors(x1v, rhoflr, epsflr);
        V(0, k, j, i) = V(0, k, j, i) < rhoflr ? rhoflr : V(0, k, j, i);
        V(4, k, j, i) = V(4, k, j, i) / V(0, k, j, i) < epsflr ? V(0, k, j, i) * epsflr : V(4, k, j, i);
        // IdealGas in scale-free units, eos_ideal.hpp:59-137
        const Real T = std::max(0.0, (V(4, k, j, i) / V(0, k, j, i)) / c->Cv);
        temp[((long)k * nj + j) * ni + i] = T;
        V(prs, k, j, i) = std::max(0.0, c->gm1 * V(0, k, j, i) * c->Cv * T);
        V(gm1i, k, j, i) = std::max(0.0, (c->gm1 + 1) * c->gm1 * V(0, k, j, i) * c->Cv * T) / V(prs, k, j, i);
      }
  if (!mhd) return 0;
  // vector potential at the corners, from the density of plane kb.s of the entire domain (torus.cpp:435-451)
  std::vector<Real> A((size_t)nj * ni, 0.0);
  for (int j = 1; j <= nj - 1; ++j)
    for (int i = 1; i <= ni - 1; ++i) {
      const Real rho_av = 0.25 * (V(0, 0, j, i) + V(0, 0, j, i - 1) + V(0, 0, j - 1, i) + V(0, 0, j - 1, i - 1));
      const Real q = rho_av - 0.2;
      A[(size_t)j * ni + i] = (q > 0 ? q : 0.0);
    }
  const long gcell = (long)c->gk * c->gj * c->gi;
  for (int k = 0; k < nk; ++k)
    for (int j = 0; j <= nj - 2; ++j)
      for (int i = 0; i <= ni - 2; ++i) {
        const long gc = ((long)(c->gk == 1 ? 0 : k) * c->gj + (c->gj == 1 ? 0 : j)) * c->gi + (c->gi == 1 ? 0 : i);
        const Real gamdet = geom[(long)20 * gcell + gc]; // loc 0 (Cent), component detgam
        auto Aa = [&](int jj, int ii) { return A[(size_t)jj * ni + ii]; };
        V(5, k, j, i) = -(Aa(j, i) - Aa(j + 1, i) + Aa(j, i + 1) - Aa(j + 1, i + 1)) / (2.0 * c->dx[1] * gamdet);
        V(6, k, j, i) = (Aa(j, i) + Aa(j + 1, i) - Aa(j, i + 1) - Aa(j + 1, i + 1)) / (2.0 * c->dx[0] * gamdet);
        V(7, k, j, i) = 0.0;
      }
  return 0;
}

// ComputeBetas, torus.cpp:849-930, one MeshBlock's interior: out = {beta_min, bsq_max, press_max}
void orc_pgen_betas(const orc_cfg *c, const double *geom, const double *prim, double rho_min_bnorm, double out[3]) {
  const int ng = c->ng, ndim = c->ndim;
  const int ni = c->nx[0] + 2 * ng, nj = ndim >= 2 ? c->nx[1] + 2 * ng : 1, nk = ndim >= 3 ? c->nx[2] + 2 * ng : 1;
  const long N = (long)ni * nj * nk;
  const long gcell = (long)c->gk * c->gj * c->gi;
  Real beta_min = std::numeric_limits<Real>::infinity(), press_max = -beta_min, bsq_max = -beta_min;
  for (int k = (ndim >= 3 ? ng : 0); k < (ndim >= 3 ? nk - ng : 1); ++k)
    for (int j = (ndim >= 2 ? ng : 0); j < (ndim >= 2 ? nj - ng : 1); ++j)
      for (int i = ng; i < ni - ng; ++i) {
        const long o = ((long)k * nj + j) * ni + i;
        const long gc = ((long)(c->gk == 1 ? 0 : k) * c->gj + (c->gj == 1 ? 0 : j)) * c->gi + (c->gi == 1 ? 0 : i);
        Real gcov[3][3];
        for (int m = 0; m < 3; ++m)
          for (int n = 0; n < 3; ++n) gcov[m][n] = geom[(long)(4 + FLAT[m + 1][n + 1]) * gcell + gc];
        const Real vp[3] = {prim[1 * N + o], prim[2 * N + o], prim[3 * N + o]};
        const Real Bp[3] = {prim[5 * N + o], prim[6 * N + o], prim[7 * N + o]};
        Real vsq = 0.;
        for (int m = 0; m < 3; ++m)
          for (int n = 0; n < 3; ++n) vsq += gcov[m][n] * vp[m] * vp[n];
        const Real W = std::sqrt(1. + vsq);
        Real Bsq = 0, Bdotv = 0; // relativity_utils.hpp:131-141
        for (int m = 0; m < 3; ++m)
          for (int n = 0; n < 3; ++n) {
            Bdotv += gcov[m][n] * ratio(vp[m], W) * Bp[n];
            Bsq += gcov[m][n] * Bp[m] * Bp[n];
          }
        const Real bsq = ratio(Bsq, W * W) + Bdotv * Bdotv;
        const Real P = prim[8 * N + o];
        const Real beta = ratio(P, 0.5 * bsq);
        if (prim[o] > rho_min_bnorm && beta < beta_min) beta_min = beta;
        bsq_max = std::max(bsq, bsq_max);
        press_max = std::max(P, press_max);
      }
  out[0] = beta_min;
  out[1] = bsq_max;
  out[2] = press_max;
}
}
