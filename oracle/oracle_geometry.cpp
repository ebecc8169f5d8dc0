// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's geometry providers used on
// the hot path: Analytic<Minkowski> and FMKS = Modified<SphericalKerrSchild,
// McKinneyGammieRyan>, tabulated the way Geometry::SetCachedCoordinateSystem does
// (src/geometry/cached_system.hpp:593-775). See oracle.h.
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <limits>

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
}
