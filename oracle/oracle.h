/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's GRMHD hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
 * The product (phoebus_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * Parity status: PINNED. Every function is checked against per-stage field dumps of the
 * unmodified reference (oracle/_ref/ref_dump_*, fixtures in tests/golden/ made by
 * oracle/make_golden.py) and against the geometry known-answer values of the reference's
 * tst/unit/geometry/test_geometry.cpp.
 *
 * All arrays are float64, per MeshBlock, C order [comp][k][j][i] (i fastest), with ghost
 * zones in every active dimension — the reference's per-block ParArrayND layout
 * (external/parthenon/src/interface/variable.hpp:138, fluid.cpp:300-304).
 */
#ifndef ORACLE_ORACLE_H_
#define ORACLE_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* geometry table: [loc(4)][ORC_NGLOC][gk][gj][gi] followed by [ORC_NGDER][gk][gj][gi];
 * loc order Cent, Face1, Face2, Face3 (src/phoebus_utils/cell_locations.hpp:21-37);
 * component order follows Geometry::Cached (src/geometry/cached_system.hpp:42-52, 640-690):
 * alpha, bcon[3], gcov[10 sym, Flatten2 order], gamcon[6 sym], detgam; then centre-only
 * dalpha[3], dg[10 sym][3]. gk/gj/gi are either the block extent or 1 (broadcast). */
#define ORC_NGLOC 21
#define ORC_NGDER 33

typedef struct {
  int ndim, nx[3], ng; /* interior zones per block; ghosts only in active dims */
  int mhd;             /* fluid/mhd */
  int recon;           /* 0 constant, 1 linear, 2 weno5z, 3 mp5  (reconstruction.hpp:23) */
  int riemann;         /* 0 llf, 1 hll (riemann.hpp:30) */
  double c2p_tol;
  int c2p_max_iter;
  double c2p_floor_scale_fac;
  int c2p_fail_on_floors, c2p_fail_on_ceilings;
  /* fixup::Bounds (fixup.hpp:72-327) */
  int floor_type; /* 0 ConstantRhoSie 1 ExpX1RhoSie 2 ExpX1RhoU 3 X1RhoSie 4 RRhoSie */
  double rho0, sie0, rexp, sexp, rho_min_eos, sie_min_eos;
  double gam_max, sie_max;
  double bsqorho_max, bsqou_max;
  int enable_floors, enable_mhd_floors, enable_c2p_fixup, enable_flux_fixup;
  int c2p_failure_strategy; /* 0 interpolate, 1 floors */
  /* IdealGas */
  double gm1, Cv;
  /* geometry table extents */
  int gk, gj, gi;
  /* block coordinates: xmin includes the ghost offset (uniform_cartesian.hpp:30-36) */
  double xmin[3], dx[3];
  /* boundary flags per face (ix1,ox1,ix2,ox2,ix3,ox3): 0 = not a physical boundary,
   * 1 gr_outflow, 2 reflect, 3 polar (only meaningful where parthenon flag == user) */
  int bc[6];
  int bc_prim; /* phoebus/mesh bc_vars == primitive */
  int ox1_inflow_check;
} orc_cfg;

/* number of primitive (recon) and conserved components */
int orc_nprim(const orc_cfg *c);
int orc_ncons(const orc_cfg *c);

/* fluid::CalculateFluxes (fluid.cpp:659-773): prim[nprim][cells] -> flux[ndim][ncons][cells],
 * fsig[ndim][cells]. ql/qr scratch may be NULL (allocated internally). */
void orc_calculate_fluxes(const orc_cfg *c, const double *geom, const double *prim,
                          double *flux, double *fsig);
/* fluid::FluxCT (fluid.cpp:775-841) */
void orc_flux_ct(const orc_cfg *c, double *flux, double *emf);
/* fixup::FixFluxes (fixup.cpp:728-1078) */
void orc_fix_fluxes(const orc_cfg *c, double *flux);
/* parthenon::Update::FluxDivergence (update.cpp:63-85): dudt interior only */
void orc_flux_divergence(const orc_cfg *c, const double *flux, double *dudt);
/* fluid::CalculateFluidSourceTerms (fluid.cpp:566-657): gsrc[4][cells] = S1,S2,S3,tau */
/* History reductions (the .hst columns). fluid package, fluid.cpp:325-347 + analysis/history.hpp:50-86 (plain sums:
 * the volume weighting never applies): out[6] = max p.density, sum c.density, sum c.energy, sum c.momentum[3].
 * geometry package, analysis/history.cpp:25-222 + history_utils.hpp:35-180: out[4] = mdot, jet energy flux,
 * jet momentum flux, 2 * magnetic Phi (the caller halves the total) for this block. */
void orc_history_fluid(const orc_cfg *c, const double *prim, const double *cons, double out[6]);
void orc_history_torus(const orc_cfg *c, const double *geom, const double *prim, const double *cons, double xh,
                       double sigma_cutoff, double out[4]);
/* fluid::CalculateDivB, fluid.cpp:843-897: corner-centred div B of the conserved field, interior zones */
void orc_divb(const orc_cfg *c, const double *cons, double *divb);
void orc_source_terms(const orc_cfg *c, const double *geom, const double *prim, double *gsrc);
/* SumData + AverageIndependentData + UpdateIndependentData (update.hpp:71-136) over entire
 * arrays: dudt[1..4] += gsrc; u0 = beta*u0 + (1-beta)*ubase (in place); uout = u0 + bdt*dudt */
void orc_rk_update(const orc_cfg *c, double *dudt, const double *gsrc, double *u0,
                   const double *ubase, double beta, double bdt, double *uout);
/* fluid::ConservedToPrimitiveRegion / ConToPrim::solve (fluid.cpp:477-517,
 * con2prim_robust.hpp:300-453) over [ib,jb,kb] inclusive ranges */
void orc_cons_to_prim(const orc_cfg *c, const double *geom, double *cons, double *prim,
                      double *temp, double *c2p_mu, double *fail, double *csig,
                      const int ib[2], const int jb[2], const int kb[2]);
/* fixup::ConservedToPrimitiveFixup (fixup_c2p.cpp:44-356), interior */
void orc_c2p_fixup(const orc_cfg *c, const double *geom, double *cons, double *prim,
                   double *temp, const double *fail, double *csig);
/* fixup::ApplyFloors (fixup.cpp:286-687), entire; counters[0] floors applied,
 * [1] ceilings applied, [2] in-kernel c2p failures */
void orc_apply_floors(const orc_cfg *c, const double *geom, double *cons, double *prim,
                      double *temp, double *c2p_mu, double *csig, long long *counters);
/* fluid::PrimitiveToConservedRegion (fluid.cpp:368-449) */
void orc_prim_to_cons(const orc_cfg *c, const double *geom, const double *prim, double *cons,
                      double *csig, const int ib[2], const int jb[2], const int kb[2]);
/* physical boundary conditions + ConvertBoundaryConditions
 * (phoebus_boundaries.cpp:44-448) for one block */
void orc_apply_bcs(const orc_cfg *c, const double *geom, double *cons, double *prim,
                   double *temp, double *c2p_mu, double *fail, double *csig);
/* the two halves of the above separately: what bit 0 = ghost fills, bit 1 = ConvertBoundaryConditions */
void orc_apply_bcs_ex(const orc_cfg *c, const double *geom, double *cons, double *prim, double *temp,
                      double *c2p_mu, double *fail, double *csig, int what);
/* Fishbone-Moncrief torus problem generator (src/pgen/torus.cpp:130-474, IdealGas, u_jitter = 0) on one MeshBlock:
 * p.* and temperature over the entire block; par as orc_build_geometry, tp = {rin, rmax, kappa, Gamma, nsub, u_jitter} */
int orc_pgen_torus(const orc_cfg *c, const double *par, const double *tp, const double *geom, double *prim, double *temp);
/* ComputeBetas (torus.cpp:849-930) over one block's interior: out = {beta_min, bsq_max, press_max} */
void orc_pgen_betas(const orc_cfg *c, const double *geom, const double *prim, double rho_min_bnorm, double out[3]);
/* fluid::EstimateTimestepBlock (fluid.cpp:899-944) without the cfl factor */
double orc_estimate_dt(const orc_cfg *c, const double *csig, const double *fsig,
                       int has_face_speeds);
/* single-stencil helpers for unit tests (reconstruction.hpp:111-244) */
void orc_recon5(int recon, const double q[5], double *ql, double *qr);

/* geometry tables: restatement of Analytic<Minkowski> (minkowski.hpp:35-115) and
 * FMKS = Modified<SphericalKerrSchild, McKinneyGammieRyan> (fmks.hpp:44) evaluated the way
 * SetCachedCoordinateSystem does (cached_system.hpp:593-775). kind 0 Minkowski, 1 FMKS.
 * par = {a, h, xt, alpha, x0, smooth, hexp_br, hexp_nsq, hexp_csq, derefine_poles} */
void orc_build_geometry(int kind, const double *par, const orc_cfg *c, double *geom);
/* pointwise accessors for known-answer tests: out = alpha, detgam, beta[3], gcov[16],
 * gamcon[9], dalpha[4], dg[64] */
void orc_geometry_point(int kind, const double *par, double X1, double X2, double X3,
                        double *out);

#ifdef __cplusplus
}
#endif
#endif
