// Host-side table builders: spacetime geometry and x1-dependent floors.
//
// The reference tabulates the metric once at initialisation (Geometry::SetCachedCoordinateSystem,
// ref: src/geometry/cached_system.hpp:593-775) and every CI/regression build uses those tables;
// FMKS metric derivatives are central finite differences of transcendental functions with a step of
// ~6e-6, which amplifies libm differences by ~1e5 (SURVEY.md §7 hard part 4). We therefore build the
// tables on the HOST with glibc, in the reference's evaluation order, and the kernels only read them.
// Compile with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/phoebus_b200.h"

namespace {
typedef double Real;
constexpr Real kSmall = 10 * std::numeric_limits<Real>::min();
inline Real ratio(Real a, Real b) { return a / (b + (b >= 0 ? 1 : -1) * kSmall); }
inline int sym4(int m, int n) { return m <= n ? n * (n + 1) / 2 + m : m * (m + 1) / 2 + n; }

struct Point {
  Real alpha, detgam, beta[3], g[4][4], gamcon[3][3], dlnalpha[4], dg[4][4][4];
};

// ---- Kerr-Schild in spherical coordinates, ref: src/geometry/spherical_kerr_schild.hpp:42-244
struct KerrSchild {
  Real a, a2;
  struct V {
    Real r, cth, sth, sth2, rho2;
  };
  V vars(Real X1, Real X2) const {
    V v;
    v.r = std::abs(X1);
    v.cth = std::cos(X2);
    v.sth = std::sin(X2);
    v.sth2 = v.sth * v.sth;
    v.rho2 = v.r * v.r + a2 * v.cth * v.cth;
    return v;
  }
  Real lapse(Real X1, Real X2) const {
    const Real r = std::abs(X1), cth = std::cos(X2);
    const Real rho2 = r * r + a2 * cth * cth;
    return std::sqrt(ratio(rho2, 2 * r + rho2));
  }
  void shift(Real X1, Real X2, Real b[3]) const {
    const Real r = std::abs(X1), cth = std::cos(X2);
    const Real rho2 = r * r + a2 * cth * cth;
    b[0] = ratio(2 * r, 2 * r + rho2);
    b[1] = b[2] = 0;
  }
  void metric4(Real X1, Real X2, Real g[4][4]) const {
    const V v = vars(X1, X2);
    std::memset(g, 0, 16 * sizeof(Real));
    g[0][1] = g[1][0] = ratio(2 * v.r, v.rho2);
    g[0][0] = -1 + g[0][1];
    g[0][3] = g[3][0] = -ratio(2 * a * v.r * v.sth2, v.rho2);
    g[1][1] = 1 + g[0][1];
    g[1][3] = g[3][1] = -a * (1 + g[0][1]) * v.sth2;
    g[2][2] = v.rho2;
    g[3][3] = (v.rho2 + a2 * (1 + g[0][1]) * v.sth2) * v.sth2;
  }
  void inverse3(Real X1, Real X2, Real g[3][3]) const {
    const V v = vars(X1, X2);
    std::memset(g, 0, 9 * sizeof(Real));
    g[0][0] = ratio(v.rho2, 2 * v.r + v.rho2) + ratio(a2 * v.sth2, v.rho2);
    g[0][2] = g[2][0] = ratio(a, v.rho2);
    g[1][1] = ratio(1, v.rho2);
    g[2][2] = ratio(1, v.rho2 * v.sth2);
  }
  Real detgamma(Real X1, Real X2) const {
    const V v = vars(X1, X2);
    return std::sqrt(std::abs(v.rho2 * (2 * v.r + v.rho2) * v.sth2));
  }
};

// ---- FMKS coordinate map, ref: src/geometry/mckinney_gammie_ryan.hpp:44-210
struct FmksMap {
  bool derefine;
  Real h, xt, al, x0, smooth, norm, br, nsq, csq;
  explicit FmksMap(const Real *p)
      : derefine(p[9] != 0), h(p[1]), xt(p[2]), al(p[3]), x0(p[4]), smooth(p[5]), br(p[6]), nsq(p[7]), csq(p[8]) {
    norm = 0.5 * M_PI * 1. / (1. + 1. / (al + 1.) * 1. / std::pow(xt, al));
  }
  Real thG(Real X2) const { return M_PI * X2 + ((1. - h) / 2.) * std::sin(2. * M_PI * X2); }
  void thJ(Real X2, Real &y, Real &t) const {
    y = 2. * X2 - 1.;
    t = norm * y * (1. + std::pow(y / xt, al) / (al + 1.)) + 0.5 * M_PI;
  }
  Real theta(Real X1, Real X2, Real &y) const {
    const Real tg = thG(X2);
    Real th;
    if (derefine) {
      Real tj;
      thJ(X2, y, tj);
      th = tg + std::exp(smooth * (x0 - X1)) * (tj - tg);
    } else {
      y = 0;
      th = tg;
    }
    constexpr Real tiny = 1.e-20; // pole clamps, ref: mckinney_gammie_ryan.hpp:172-185
    if (std::fabs(th) < tiny) th = (th >= 0.) ? tiny : -tiny;
    if (std::fabs(M_PI - th) < tiny) th = (th >= M_PI) ? M_PI + tiny : M_PI - tiny;
    return th;
  }
  // C = (r, theta, phi); J = dC/dX, Ji = dX/dC
  void eval(Real X1, Real X2, Real X3, Real C[3], Real J[3][3], Real Ji[3][3]) const {
    const Real dr = X1 - br;
    Real y;
    C[0] = std::exp(X1 + (dr > 0.) * csq * std::pow(dr, nsq));
    C[1] = theta(X1, X2, y);
    C[2] = X3;
    const Real drdX1 = std::exp(X1 + (dr > 0.) * csq * std::pow(dr, nsq)) * (1. + (dr > 0.) * csq * nsq * std::pow(dr, nsq - 1.));
    const Real dthGdX2 = M_PI + M_PI * (1 - h) * std::cos(2 * M_PI * X2);
    Real dthdX1 = 0.0, dthdX2 = dthGdX2;
    if (derefine) {
      Real tj;
      thJ(X2, y, tj);
      const Real tg = thG(X2);
      const Real dthJdX2 = norm * (1 + std::pow(y / xt, al)) * 2.;
      const Real fac = x0 - X1;
      dthdX1 = -smooth * (tj - tg) * std::exp(smooth * fac);
      dthdX2 = dthGdX2 + std::exp(smooth * fac) * (dthJdX2 - dthGdX2);
    }
    std::memset(J, 0, 9 * sizeof(Real));
    std::memset(Ji, 0, 9 * sizeof(Real));
    J[0][0] = drdX1;
    J[1][0] = dthdX1;
    J[1][1] = dthdX2;
    J[2][2] = 1.;
    Ji[0][0] = ratio(1., drdX1);
    Ji[1][0] = -ratio(dthdX1, drdX1 * dthdX2);
    Ji[1][1] = ratio(1., dthdX2);
    Ji[2][2] = 1.;
  }
};

inline Real pad4(const Real A[3][3], int m, int n) { return (m == 0 || n == 0) ? (m == n ? 1 : 0) : A[m - 1][n - 1]; }

// ---- Modified<KerrSchild, FmksMap>, ref: src/geometry/modified_system.hpp:53-246
struct Fmks {
  KerrSchild ks;
  FmksMap map;
  explicit Fmks(const Real *p) : ks{p[0], p[0] * p[0]}, map(p) {}
  Real lapse(Real X1, Real X2, Real X3) const {
    Real C[3], J[3][3], Ji[3][3];
    map.eval(X1, X2, X3, C, J, Ji);
    return ks.lapse(C[0], C[1]);
  }
  void shift(Real X1, Real X2, Real X3, Real b[3]) const {
    Real b0[3], C[3], J[3][3], Ji[3][3];
    map.eval(X1, X2, X3, C, J, Ji);
    ks.shift(C[0], C[1], b0);
    for (int i = 0; i < 3; ++i) {
      b[i] = 0;
      for (int j = 0; j < 3; ++j) b[i] += Ji[i][j] * b0[j];
    }
  }
  void metric4(Real X1, Real X2, Real X3, Real g[4][4]) const {
    Real g0[4][4], C[3], J[3][3], Ji[3][3];
    map.eval(X1, X2, X3, C, J, Ji);
    ks.metric4(C[0], C[1], g0);
    std::memset(g, 0, 16 * sizeof(Real));
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = 0; nu < 4; ++nu)
        for (int lam = 0; lam < 4; ++lam)
          for (int kap = 0; kap < 4; ++kap) g[mu][nu] += g0[lam][kap] * pad4(J, lam, mu) * pad4(J, kap, nu);
  }
  void inverse3(Real X1, Real X2, Real X3, Real g[3][3]) const {
    Real g0[3][3], C[3], J[3][3], Ji[3][3];
    map.eval(X1, X2, X3, C, J, Ji);
    ks.inverse3(C[0], C[1], g0);
    std::memset(g, 0, 9 * sizeof(Real));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l) g[i][j] += g0[k][l] * Ji[i][k] * Ji[j][l];
  }
  Real detgamma(Real X1, Real X2, Real X3) const {
    Real C[3], J[3][3], Ji[3][3];
    map.eval(X1, X2, X3, C, J, Ji);
    // linear_algebra.hpp:48-54
    const Real detJ = (J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
                       J[0][2] * J[1][1] * J[2][0] - J[0][1] * J[1][0] * J[2][2] - J[0][0] * J[1][2] * J[2][1]);
    return std::abs(ks.detgamma(C[0], C[1]) * detJ);
  }
  // central differences with step max(|X|,1)*cbrt(eps), ref: src/geometry/geometry_utils.hpp:91-185
  void derivatives(Real X1, Real X2, Real X3, Real dlna[4], Real dg[4][4][4]) const {
    const Real EPS = std::pow(std::numeric_limits<Real>::epsilon(), 1.0 / 3.0);
    std::memset(dg, 0, 64 * sizeof(Real));
    dlna[0] = 0;
    const Real X[3] = {X1, X2, X3};
    for (int d = 1; d < 4; ++d) {
      Real Xp[3] = {X1, X2, X3}, Xm[3] = {X1, X2, X3};
      const Real dx = std::max(std::abs(X[d - 1]), 1.0) * EPS;
      Xp[d - 1] += dx;
      Xm[d - 1] -= dx;
      Real gl[4][4], gr[4][4];
      metric4(Xp[0], Xp[1], Xp[2], gr);
      metric4(Xm[0], Xm[1], Xm[2], gl);
      for (int mu = 0; mu < 4; ++mu)
        for (int nu = 0; nu < 4; ++nu) dg[mu][nu][d] = ratio(gr[mu][nu] - gl[mu][nu], 2 * dx);
      const Real am = lapse(Xm[0], Xm[1], Xm[2]);
      const Real ap = lapse(Xp[0], Xp[1], Xp[2]);
      dlna[d] = ratio(ap - am, dx * (am + ap));
    }
  }
};

void eval_point(int kind, const Real *par, Real X1, Real X2, Real X3, bool derivs, Point &p) {
  if (kind == 0) { // ref: src/geometry/minkowski.hpp:35-115
    p.alpha = 1.;
    p.detgam = 1.;
    for (int i = 0; i < 3; ++i) p.beta[i] = 0;
    for (int m = 0; m < 4; ++m)
      for (int n = 0; n < 4; ++n) p.g[m][n] = (m == n) ? (m == 0 ? -1 : 1) : 0;
    for (int m = 0; m < 3; ++m)
      for (int n = 0; n < 3; ++n) p.gamcon[m][n] = (m == n ? 1 : 0);
    std::memset(p.dlnalpha, 0, sizeof(p.dlnalpha));
    std::memset(p.dg, 0, sizeof(p.dg));
    return;
  }
  const Fmks s(par);
  p.alpha = s.lapse(X1, X2, X3);
  p.detgam = s.detgamma(X1, X2, X3);
  s.shift(X1, X2, X3, p.beta);
  s.metric4(X1, X2, X3, p.g);
  s.inverse3(X1, X2, X3, p.gamcon);
  if (derivs) s.derivatives(X1, X2, X3, p.dlnalpha, p.dg);
}
} // namespace

extern "C" int pb_build_geometry_host(int kind, const double *par, const pb_cfg_t *cfg, const double xmin[3],
                                      const double dx[3], double *out) {
  if (!cfg || !out || (kind != 0 && kind != 1) || (kind == 1 && !par)) return PB_EINVAL;
  const int gk = cfg->gk, gj = cfg->gj, gi = cfg->gi;
  const long gcell = (long)gk * gj * gi;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < gk; ++k)
    for (int j = 0; j < gj; ++j)
      for (int i = 0; i < gi; ++i)
        for (int loc = 0; loc < 4; ++loc) {
          // cell centres / faces, ref: analytic_system.hpp:59-90 + uniform_cartesian.hpp:86-145
          Real X1 = xmin[0] + (i + 0.5) * dx[0], X2 = xmin[1] + (j + 0.5) * dx[1], X3 = xmin[2] + (k + 0.5) * dx[2];
          if (loc == 1) X1 = xmin[0] + i * dx[0];
          if (loc == 2) X2 = xmin[1] + j * dx[1];
          if (loc == 3) X3 = xmin[2] + k * dx[2];
          Point p;
          eval_point(kind, par, X1, X2, X3, loc == 0, p);
          const long cell = ((long)k * gj + j) * gi + i;
          double *t = out + (long)loc * PB_NGLOC * gcell + cell;
          t[0] = p.alpha;
          for (int d = 0; d < 3; ++d) t[(1 + d) * gcell] = p.beta[d];
          for (int m = 0; m < 4; ++m)
            for (int n = m; n < 4; ++n) t[(4 + sym4(m, n)) * gcell] = p.g[m][n];
          for (int m = 0; m < 3; ++m)
            for (int n = m; n < 3; ++n) t[(14 + sym4(m, n)) * gcell] = p.gamcon[m][n];
          t[20 * gcell] = p.detgam;
          if (loc == 0) {
            double *dt = out + (long)4 * PB_NGLOC * gcell + cell;
            for (int d = 1; d < 4; ++d) dt[(d - 1) * gcell] = p.dlnalpha[d];
            for (int s = 1; s < 4; ++s)
              for (int m = 0; m < 4; ++m)
                for (int n = m; n < 4; ++n) dt[(3 + 3 * sym4(m, n) + (s - 1)) * gcell] = p.dg[m][n][s];
          }
        }
  return PB_OK;
}

// Floors::GetFloors along x1, ref: src/fixup/fixup.hpp:99-147. out[loc][0][i] = density factor,
// out[loc][1][i] = sie (or u, for ExpX1RhoU) factor; loc 0 = cell centres, 1 = x1 faces.
extern "C" int pb_build_floor_table_host(const pb_cfg_t *c, const double xmin[3], const double dx[3], double *out) {
  if (!c || !out) return PB_EINVAL;
  const int ni = c->nx[0] + 2 * c->ng;
  for (int loc = 0; loc < 2; ++loc)
    for (int i = 0; i < ni; ++i) {
      const Real x1 = loc == 0 ? xmin[0] + (i + 0.5) * dx[0] : xmin[0] + i * dx[0];
      Real A, B;
      switch (c->floor_type) {
      case 0:
        A = c->rho0;
        B = c->sie0;
        break;
      case 1:
      case 2:
        A = c->rho0 * std::exp(c->rexp * x1);
        B = c->sie0 * std::exp(c->sexp * x1);
        break;
      case 3:
        A = c->rho0 * std::min(1.0, std::pow(x1, c->rexp));
        B = c->sie0 * std::min(1.0, std::pow(x1, c->sexp));
        break;
      case 4: // depends on all three coordinates: evaluated in the kernels
        A = B = 0.0;
        break;
      default:
        return PB_EINVAL;
      }
      out[(loc * 2 + 0) * ni + i] = A;
      out[(loc * 2 + 1) * ni + i] = B;
    }
  return PB_OK;
}

// ================================================================================================
// Fishbone-Moncrief torus, ref: src/pgen/torus.cpp:130-362,745-833 (IdealGas branch). The solution is
// axisymmetric and made of log / exp / pow / sin / cos, so — like the metric — it is tabulated on the HOST
// with glibc in the reference's evaluation order, one (j, i) table per MeshBlock column; the device side
// (pgen_kernels.cu) broadcasts it along x3 and does the rest (jitter, floors, EOS, vector potential, B).
namespace {
// ref: src/geometry/boyer_lindquist.hpp:79-95,262-280
void bl_metric4(Real a, Real X1, Real X2, Real g[4][4]) {
  const Real a2 = a * a;
  const Real r = std::abs(X1), th = X2;
  const Real sth = std::sin(th), cth = std::cos(th);
  const Real r2 = r * r, r3 = r2 * r;
  const Real DD = 1. - ratio(2., r) + ratio(a2, r2);
  const Real mu = 1. + ratio(a2 * cth * cth, r2);
  const Real sth2 = sth * sth;
  std::memset(g, 0, 16 * sizeof(Real));
  g[0][0] = -(1.0 - ratio(2.0, r * mu));
  g[0][3] = g[3][0] = -ratio(2.0 * a * sth2, r * mu);
  g[1][1] = mu / DD;
  g[2][2] = r2 * mu;
  g[3][3] = r2 * sth2 * (1.0 + ratio(a2, r2) + ratio(2.0 * a2 * sth2, r3 * mu));
}
// ref: torus.cpp:745-757
Real lfish_calc(Real r, Real a) {
  return (((std::pow(a, 2) - 2. * a * std::sqrt(r) + std::pow(r, 2)) *
           ((-2. * a * r * (std::pow(a, 2) - 2. * a * std::sqrt(r) + std::pow(r, 2))) / std::sqrt(2. * a * std::sqrt(r) + (-3. + r) * r) +
            ((a + (-2. + r) * std::sqrt(r)) * (std::pow(r, 3) + std::pow(a, 2) * (2. + r))) / std::sqrt(1 + (2. * a) / std::pow(r, 1.5) - 3. / r))) /
          (std::pow(r, 3) * std::sqrt(2. * a * std::sqrt(r) + (-3. + r) * r) * (std::pow(a, 2) + (-2. + r) * r)));
}
// ref: torus.cpp:759-797
Real log_enthalpy(const Real r, const Real th, const Real a, const Real rin, const Real l, Real &uphi) {
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
  const Real lnh = 0.5 * std::log((1. + std::sqrt(1. + 4. * (l * l * SS * SS) * DD / (AA * AA * sth * sth))) / (SS * DD / AA)) -
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
}
// ref: torus.cpp:803-812 (IdealGas)
void state_from_enthalpy(Real hm1, Real kappa, Real gam, Real rho_rmax, Real &rho_out, Real &u_out) {
  rho_out = std::pow(hm1 * (gam - 1.) / (kappa * gam), 1. / (gam - 1.));
  u_out = kappa * std::pow(rho_out, gam) / (gam - 1.) / rho_rmax;
  rho_out /= rho_rmax;
}
// ref: torus.cpp:834-847; returns false where the reference aborts (discr < 0)
bool ucon_norm(const Real ucon[4], const Real gcov[4][4], Real &u0) {
  Real AA = gcov[0][0];
  Real BB = 2. * (gcov[0][1] * ucon[1] + gcov[0][2] * ucon[2] + gcov[0][3] * ucon[3]);
  Real CC = 1. + gcov[1][1] * ucon[1] * ucon[1] + gcov[2][2] * ucon[2] * ucon[2] + gcov[3][3] * ucon[3] * ucon[3] +
            2. * (gcov[1][2] * ucon[1] * ucon[2] + gcov[1][3] * ucon[1] * ucon[3] + gcov[2][3] * ucon[2] * ucon[3]);
  Real discr = BB * BB - 4. * AA * CC;
  if (discr < 0) return false;
  u0 = (-BB - std::sqrt(discr)) / (2. * AA);
  return true;
}
} // namespace

namespace {
// The pgen reads the metric through the CACHED system's coordinate accessors (geom.Lapse(0, x1, x2, x3) ...), which
// interpolate the cell-centred tables linearly (ref: src/geometry/cached_system.hpp:132-139,150-155,229-236,
// src/phoebus_utils/phoebus_interpolation.hpp:122-219): at a cell centre the weights are (1, 0) up to the rounding of
// the coordinate arithmetic, which is where the last-bit structure of the reference's initial velocities comes from.
struct CentInterp {
  const double *tab; // one slot of the geometry table
  int gk, gj, gi;
  const double *xmin, *dx;
  void weights(int d, Real x, int nx, int &ix, Real w[2]) const {
    const Real mn = xmin[d] + (0 + 0.5) * dx[d]; // Xc<d>(0)
    ix = std::min(std::max(0, static_cast<int>(ratio(x - mn, dx[d]))), nx - 2);
    const Real fl = mn + ix * dx[d];
    w[1] = ratio(x - fl, dx[d]);
    w[0] = 1. - w[1];
  }
  Real operator()(int comp, Real X1, Real X2, Real X3) const { // comp: index into the Cent block of the table
    const long gcell = (long)gk * gj * gi;
    const double *p = tab + (long)comp * gcell;
    auto P = [&](int k, int j, int i) { return p[((long)k * gj + j) * gi + i]; };
    if (gk > 1) {
      int ix[3];
      Real w[3][2];
      weights(0, X1, gi, ix[0], w[0]);
      weights(1, X2, gj, ix[1], w[1]);
      weights(2, X3, gk, ix[2], w[2]);
      return (w[2][0] * (w[1][0] * (w[0][0] * P(ix[2], ix[1], ix[0]) + w[0][1] * P(ix[2], ix[1], ix[0] + 1)) +
                         w[1][1] * (w[0][0] * P(ix[2], ix[1] + 1, ix[0]) + w[0][1] * P(ix[2], ix[1] + 1, ix[0] + 1))) +
              w[2][1] * (w[1][0] * (w[0][0] * P(ix[2] + 1, ix[1], ix[0]) + w[0][1] * P(ix[2] + 1, ix[1], ix[0] + 1)) +
                         w[1][1] * (w[0][0] * P(ix[2] + 1, ix[1] + 1, ix[0]) + w[0][1] * P(ix[2] + 1, ix[1] + 1, ix[0] + 1))));
    }
    int ix1, ix2;
    Real w1[2], w2[2];
    weights(0, X1, gi, ix1, w1);
    weights(1, X2, gj, ix2, w2);
    return (w2[0] * (w1[0] * P(0, ix2, ix1) + w1[1] * P(0, ix2, ix1 + 1)) +
            w2[1] * (w1[0] * P(0, ix2 + 1, ix1) + w1[1] * P(0, ix2 + 1, ix1 + 1)));
  }
};
} // namespace

// torus_par = {rin, rmax, kappa, Gamma, nsub}; geom_tab = this MeshBlock column's geometry table (pb_build_geometry_host);
// out_host[5][tk][nj][ni] = rho, u, W v^1, W v^2, W v^3 of the torus interior (0 outside), before jitter and floors;
// tk = 1 for an axisymmetric (2-D) geometry table, else nk.
extern "C" int pb_pgen_torus_table_host(const double *geom_par, const double *torus_par, const pb_cfg_t *cfg,
                                        const double xmin[3], const double dx[3], const double *geom_tab, double *out) {
  if (!geom_par || !torus_par || !cfg || !out || !geom_tab) return PB_EINVAL;
  if (cfg->gj < 2 || cfg->gi < 2) return PB_EINVAL; // FMKS tables are at least 2-D
  const Real a = geom_par[0], rin = torus_par[0], rmax = torus_par[1], kappa = torus_par[2], gam = torus_par[3];
  const int nsub = (int)torus_par[4];
  if (nsub < 1) return PB_EINVAL;
  const int ni = cfg->nx[0] + 2 * cfg->ng, nj = cfg->ndim >= 2 ? cfg->nx[1] + 2 * cfg->ng : 1;
  const int tk = cfg->gk > 1 ? cfg->gk : 1;
  const CentInterp geom{geom_tab, cfg->gk, cfg->gj, cfg->gi, xmin, dx};
  // the pgen's own coordinate map: the six-argument constructor, i.e. hexp_br = 1000, nsq = csq = 4
  // (ref: torus.cpp:244, mckinney_gammie_ryan.hpp:57-62)
  Real tr_par[10];
  std::memcpy(tr_par, geom_par, sizeof(tr_par));
  tr_par[6] = 1000.;
  tr_par[7] = 4.0;
  tr_par[8] = 4.0;
  const FmksMap tr(tr_par);
  const Real angular_mom = lfish_calc(rmax, a);
  Real uphi_rmax;
  const Real hm1_rmax = std::exp(log_enthalpy(rmax, 0.5 * M_PI, a, rin, angular_mom, uphi_rmax)) - 1.0;
  Real rho_rmax, u_rmax;
  state_from_enthalpy(hm1_rmax, kappa, gam, 1.0, rho_rmax, u_rmax);
  const long N = (long)ni * nj * tk;
  int bad = 0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : bad)
  for (int k = 0; k < tk; ++k)
  for (int j = 0; j < nj; ++j)
    for (int i = 0; i < ni; ++i) {
      const Real dx_sub = dx[0] / nsub, dy_sub = dx[1] / nsub;
      Real acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      const Real x3 = xmin[2] + (k + 0.5) * dx[2];
      for (int m = 0; m < nsub; m++)
        for (int n = 0; n < nsub; n++) {
          const Real x1 = (xmin[0] + i * dx[0]) + (m + 0.5) * dx_sub;
          const Real x2 = (xmin[1] + j * dx[1]) + (n + 0.5) * dy_sub;
          Real C[3], J[3][3], Ji[3][3];
          tr.eval(x1, x2, x3, C, J, Ji);
          const Real r = C[0], th = C[1];
          Real lnh = -1.0, uphi;
          if (r > rin) lnh = log_enthalpy(r, th, a, rin, angular_mom, uphi);
          if (lnh > 0.0) {
            const Real hm1 = std::exp(lnh) - 1.;
            Real rho, u;
            state_from_enthalpy(hm1, kappa, gam, rho_rmax, rho, u);
            Real ucon_bl[4] = {0.0, 0.0, 0.0, uphi};
            Real gcov[4][4];
            bl_metric4(a, r, th, gcov);
            if (!ucon_norm(ucon_bl, gcov, ucon_bl[0])) bad++;
            // bl_to_ks then KS -> FMKS with the inverse Jacobian, ref: mckinney_gammie_ryan.hpp:119-151
            Real trans[4][4] = {{0}}, ucon_ks[4] = {0, 0, 0, 0}, ucon[4] = {0, 0, 0, 0};
            const Real idenom = 1.0 / (r * r - 2.0 * r + a * a);
            trans[0][0] = 1.0;
            trans[0][1] = 2.0 * r * idenom;
            trans[1][1] = 1.0;
            trans[2][2] = 1.0;
            trans[3][1] = a * idenom;
            trans[3][3] = 1.0;
            for (int mu = 0; mu < 4; ++mu)
              for (int nu = 0; nu < 4; ++nu) ucon_ks[mu] += trans[mu][nu] * ucon_bl[nu];
            Real t2[4][4] = {{0}};
            t2[0][0] = 1.0;
            for (int p = 0; p < 3; ++p)
              for (int q = 0; q < 3; ++q) t2[p + 1][q + 1] = Ji[p][q];
            for (int mu = 0; mu < 4; ++mu)
              for (int nu = 0; nu < 4; ++nu) ucon[mu] += t2[mu][nu] * ucon_ks[nu];
            const Real lapse = geom(0, x1, x2, x3);
            Real beta[3];
            for (int d = 0; d < 3; ++d) beta[d] = geom(1 + d, x1, x2, x3);
            for (int mu = 0; mu < 4; ++mu)
              for (int nu = 0; nu < 4; ++nu) gcov[mu][nu] = geom(4 + sym4(mu, nu), x1, x2, x3);
            if (!ucon_norm(ucon, gcov, ucon[0])) bad++;
            const Real W = lapse * ucon[0];
            const Real vcon[3] = {ucon[1] / W + beta[0] / lapse, ucon[2] / W + beta[1] / lapse, ucon[3] / W + beta[2] / lapse};
            acc[0] += rho / (nsub * nsub);
            acc[1] += u / (nsub * nsub);
            for (int d = 0; d < 3; d++) acc[2 + d] += W * vcon[d] / (nsub * nsub);
          }
        }
      for (int q = 0; q < 5; ++q) out[q * N + ((long)k * nj + j) * ni + i] = acc[q];
    }
  return bad ? PB_EINVAL : PB_OK; // discr < 0: the reference aborts ("discr < 0")
}
