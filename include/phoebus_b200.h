/* phoebus_b200 — C ABI of the B200-native GRMHD fluid-update path.
 *
 * Every entry point is one Parthenon/Phoebus task (or a fusion of adjacent tasks) applied to a
 * whole pack of MeshBlocks in one launch, asynchronously on a caller-supplied CUDA stream.
 * Plain pointers and sizes only; nothing here depends on torch, Kokkos or Parthenon.
 * "ref:" lines cite the reference interface (lanl/phoebus @ 830cb3ab, paths relative to the
 * reference root) that each entry point replaces.
 *
 * Conventions
 *  - All field data is FP64 in device memory, one slab per field per pack:
 *    [block][component][k][j][i], i fastest, ghost zones (ng) in every active dimension —
 *    per block this is exactly Parthenon's ParArrayND layout
 *    (ref: external/parthenon/src/interface/variable.hpp:138, src/fluid/fluid.cpp:300-304).
 *  - Component order (ref: src/fluid/fluid.cpp:272-286):
 *      prim: p.density, p.velocity[3], p.energy, [p.bfield[3]], pressure, gamma1   (7 + 3*mhd)
 *      cons: c.density, c.momentum[3], c.energy, [c.bfield[3]]                     (5 + 3*mhd)
 *  - Return value: 0 on success, negative PB_E* code otherwise; pb_last_error() gives text.
 *    Numerical failure is data (the `fail` field: 1.0 = success, 0.0 = failed zone,
 *    ref: src/fluid/fluid.cpp:506-514), never an error code.
 *  - Entry points never allocate field memory and never synchronise the stream, except the
 *    *_host convenience calls, which say so.
 *  - The library keeps small device work spaces of its own (redo queues, con2prim scratch, counters): they belong to
 *    ONE device per process and are used in stream order, i.e. one stage in flight at a time. A call from another
 *    device, or on another stream while the previous one still has work queued, returns PB_EINVAL.
 */
#ifndef PHOEBUS_B200_H_
#define PHOEBUS_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PB_OK 0
#define PB_EINVAL (-1)  /* bad argument / unsupported configuration */
#define PB_ECUDA (-2)   /* CUDA runtime error (launch failure, no device) */

/* geometry table: per slot [4 loc][PB_NGLOC][gk][gj][gi] then [PB_NGDER][gk][gj][gi].
 * loc: 0 Cent, 1 Face1, 2 Face2, 3 Face3 (ref: src/phoebus_utils/cell_locations.hpp:21-37).
 * components (ref: src/geometry/cached_system.hpp:42-52): alpha, bcon[3], gcov[10 sym],
 * gamcon[6 sym], detgam; centre only: dalpha[3], dg[10 sym][3]. */
#define PB_NGLOC 21
#define PB_NGDER 33
#define PB_GEOM_WORDS (4 * PB_NGLOC + PB_NGDER)

/* boundary codes in pb_block_t.bc */
#define PB_BC_BLOCK 0      /* neighbouring MeshBlock (filled by the halo exchange) */
#define PB_BC_PERIODIC (-1) /* periodic mesh face (filled by the halo exchange) */
#define PB_BC_OUTFLOW 1    /* "gr_outflow" (ref: src/phoebus_boundaries/phoebus_boundaries.cpp:118,222) */
#define PB_BC_REFLECT 2    /* "reflect" (ref: phoebus_boundaries.cpp:44-116) */
#define PB_BC_POLAR 3      /* "polar" (ref: phoebus_boundaries.cpp:153,187) */

/* neighbour codes in pb_block_t.nbr */
#define PB_NBR_NONE (-1)   /* physical boundary: no neighbour */
#define PB_NBR_REMOTE (-2) /* neighbour lives on another GPU: filled by pb_halo_unpack */

/* Per-MeshBlock metadata; the pack holds a device array of these. */
typedef struct pb_block {
  double xmin[3];  /* coordinate of cell index 0 (ghost offset included), ref: uniform_cartesian.hpp:30-36 */
  double dx[3];
  int bc[6];       /* ix1, ox1, ix2, ox2, ix3, ox3 */
  int geom_slot;   /* which geometry table this block reads */
  int floor_slot;  /* which floor table this block reads */
  int nbr[27];     /* block index in this pack of the neighbour at offset (o1+1)+3*(o2+1)+9*(o3+1) */
  int pad_;
} pb_block_t;

/* Flattened Params of the fluid, fixup, eos and geometry packages
 * (ref: src/fluid/fluid.cpp:53-137, src/fixup/fixup.cpp:53-277,
 *  src/microphysics/eos_phoebus/eos_phoebus.cpp:86-95). */
typedef struct pb_cfg {
  int ndim, nx[3], ng; /* interior zones per block; ghosts only in active dimensions */
  int mhd;             /* fluid/mhd */
  int recon;           /* 0 constant, 1 linear, 2 weno5z, 3 mp5 (ref: src/reconstruction.hpp:23) */
  int riemann;         /* 0 llf, 1 hll (ref: src/fluid/riemann.hpp:30) */
  double c2p_tol;
  int c2p_max_iter;
  double c2p_floor_scale_fac;
  int c2p_fail_on_floors, c2p_fail_on_ceilings;
  int floor_type;      /* 0 ConstantRhoSie 1 ExpX1RhoSie 2 ExpX1RhoU 3 X1RhoSie 4 RRhoSie (ref: fixup.hpp:59-147) */
  double rho0, sie0, rexp, sexp, rho_min_eos, sie_min_eos;
  double gam_max, sie_max;         /* Ceilings (ref: fixup.hpp:182-232) */
  double bsqorho_max, bsqou_max;   /* MHDCeilings (ref: fixup.hpp:234-280) */
  int enable_floors, enable_mhd_floors, enable_c2p_fixup, enable_flux_fixup;
  int c2p_failure_strategy;        /* 0 interpolate, 1 floors (ref: fixup_c2p.cpp:112-122) */
  double gm1, Cv;                  /* IdealGas (ref: singularity-eos eos_ideal.hpp:59-137) */
  int gk, gj, gi;                  /* geometry table extents: block extent or 1 (broadcast) */
  int bc_prim;                     /* phoebus/mesh bc_vars == primitive */
  int ox1_inflow_check;            /* fixup/enable_ox1_fmks_inflow_check */
} pb_cfg_t;

/* A pack of MeshBlocks: device pointers to the OneCopy fields shared by all stages
 * (ref: src/fluid/fluid.cpp:139-316). Conserved containers ("base", "1", ...) are passed
 * per call because they are per-stage (Metadata::Independent). */
typedef struct pb_pack {
  int nblocks;
  const pb_block_t *blocks; /* device [nblocks] */
  const double *geom;       /* device [nslots][PB_GEOM_WORDS][gk][gj][gi] */
  const double *floor_tab;  /* device [nfslots][2 loc (cell, face1)][2 (rho, sie/u)][ni] */
  double *prim;             /* [b][nprim][cells] */
  double *temp;             /* [b][cells]  temperature */
  double *c2p_mu;           /* [b][cells]  root-finder guess history */
  double *fail;             /* [b][cells] */
  double *csig;             /* [b][ndim][cells] cell_signal_speed */
  double *fsig;             /* [b][ndim][cells] face_signal_speed */
  double *flux;             /* [b][ndim][ncons][cells] */
  double *emf;              /* [b][3][cells] */
  /* optional accelerators (NULL / 0: the sequencer falls back to the exhaustive passes) */
  unsigned char *zflags;    /* [b][cells] zone flags kept by pb_rk_stage*: which zones the fixup / floor
                             * tasks have to touch; build once with pb_scan_zone_flags */
  const int *bc_blocks;     /* device [n_bc_blocks]: blocks with a physical or periodic mesh boundary */
  int n_bc_blocks;
  double *gsrc;             /* [b][4][cells] the "geometric source terms" container (ref: phoebus_driver.cpp:289-299);
                             * when present the sequencer runs the sources as their own launch */
  const int *inner_blocks;  /* device [n_inner_blocks]: the blocks NOT in bc_blocks (no mesh boundary on any face);
                             * optional, used by pb_rk_stage_post_early to start con2prim under an inter-GPU transfer */
  int n_inner_blocks;
  double *recon_lr;         /* [2][b][nprim][cells] left / right face states of ONE direction at a time (the
                             * reference's ql / qr fields, ref: src/fluid/fluid.cpp:288-299). When present,
                             * pb_calculate_fluxes runs reconstruction and Riemann solve as two launches per
                             * direction (higher occupancy for the reconstruction); NULL: one fused launch */
  /* optional ghost plan (pb_ghost_plan_host; all or none). The stage sequencer then solves con2prim only on the interior
   * zones and on the ghost zones listed in c2p_zones, and fills the ghost zones flagged in ghost_mask from the interior
   * zone of their same-pack neighbour AFTER con2prim: the reference solves those zones twice with identical inputs and
   * tables (ghost_plan_host.cpp). The task-level entry points (pb_cons_to_prim, pb_halo_exchange_local, ...) ignore it. */
  const unsigned char *ghost_mask; /* device [b][cells]: 1 = deferred ghost zone, 2 = twin (below) */
  const unsigned *c2p_zones;       /* device [n_c2p_zones]: ids b * cells + cell of the ghost zones that are solved */
  long long n_c2p_zones;
  /* twins (optional part of the plan): x1 ghost zones that would be deferred but for tables that differ from the
   * neighbour's in the last place. Their con2prim runs on the neighbour's inputs with this block's tables into twin_buf
   * (launched with the pre-con2prim ghost copy, before any con2prim of the stage overwrites those inputs) and the deferred
   * copy takes them from there. */
  const unsigned *twin_zones;      /* device [n_twin]: zone ids, ordered (block, x1 ghost column, k, j) */
  long long n_twin;
  double *twin_buf;                /* device [ncons + nprim + 3 + ndim][S], S = pb_twin_entries(cfg, nblocks) */
  unsigned char *twin_flags;       /* device [S] */
} pb_pack_t;

/* Per-MeshBlock allocations <-> the slabs of pb_pack_t. Parthenon allocates every variable per MeshBlock and variable
 * (ref: external/parthenon/src/interface/variable.cpp); a host that cannot allocate them from one arena per field gathers
 * them into the slab ([b][words_per_block], words_per_block = components * cells) before a cycle and scatters them back
 * after it. block_ptrs_dev: DEVICE array of nblocks device pointers; at most 65535 blocks per call. */
int pb_blocks_gather(double *slab, const double *const *block_ptrs_dev, int nblocks, long long words_per_block, void *stream);
int pb_blocks_scatter(const double *slab, double *const *block_ptrs_dev, int nblocks, long long words_per_block, void *stream);

/* Host-side construction of the ghost plan (ghost_plan_host.cpp states the conditions). blocks / geom_host / floor_host
 * are the HOST copies of what the pack holds on the device. mask_out: nblocks * cells bytes; zones_out: room for nblocks *
 * (cells - interior cells) entries, *nzones_out of them written; twins_out (optional, NULL: no twins — those zones are
 * listed in zones_out instead): room for nblocks * 2 ng * nk * nj entries. The caller uploads the arrays. Returns PB_OK; a configuration the
 * shortcut does not apply to (bc_vars = conserved, floor type RRhoSie) yields an all-zero mask and every ghost zone listed. */
int pb_ghost_plan_host(const pb_cfg_t *cfg, const pb_block_t *blocks, int nblocks, const double *geom_host,
                       const double *floor_host, unsigned char *mask_out, unsigned *zones_out, long long *nzones_out,
                       unsigned *twins_out, long long *ntwins_out);
/* entries of the twin side buffer: nblocks * 2 ng * nk * nj (every x1 ghost column of every block has a slot) */
long long pb_twin_entries(const pb_cfg_t *cfg, int nblocks);
/* Same-GPU ghost copy with the plan: phase 0 = every ghost zone (pb_halo_exchange_local); 1 = all but the deferred zones;
 * 2 = only the deferred zones, every field con2prim reads or writes (conserved, primitives, temperature, c2p_mu, fail,
 * cell signal speeds, zone flags), to be called after con2prim and before the fixup. */
int pb_halo_exchange_local_ex(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, int phase, void *stream);
/* 1 when pb_rk_stage / pb_rk_stage_post_early / pb_rk_stage_post2 use the pack's ghost plan (plan present, bc_vars =
 * primitive, option ghost_plan on): the phase a caller that issues the same-GPU ghost copy of a stage itself has to pass. */
int pb_ghost_plan_active(const pb_pack_t *pack, const pb_cfg_t *cfg);

const char *pb_version(void);
const char *pb_last_error(void);
int pb_nprim(const pb_cfg_t *cfg);
int pb_ncons(const pb_cfg_t *cfg);
long long pb_cells_per_block(const pb_cfg_t *cfg);
/* number of SMs of the current device (grid sizing); negative on error */
int pb_device_sm_count(void);

/* ---- host-side table builders (run on the CPU at initialisation, like the reference's
 * Geometry::SetCachedCoordinateSystem, ref: src/geometry/cached_system.hpp:593-775) ------- */
/* kind 0 Minkowski (ref: src/geometry/minkowski.hpp:35-115), 1 FMKS =
 * Modified<SphericalKerrSchild, McKinneyGammieRyan> (ref: src/geometry/fmks.hpp:44).
 * par = {a, h, xt, alpha, x0, smooth, hexp_br, hexp_nsq, hexp_csq, derefine_poles}.
 * out_host: PB_GEOM_WORDS*gk*gj*gi doubles. */
int pb_build_geometry_host(int kind, const double *par, const pb_cfg_t *cfg, const double xmin[3],
                           const double dx[3], double *out_host);
/* Floors::GetFloors tabulated along x1 (ref: src/fixup/fixup.hpp:99-147): out_host[2][2][ni] */
int pb_build_floor_table_host(const pb_cfg_t *cfg, const double xmin[3], const double dx[3],
                              double *out_host);

/* ---- problem generator: Fishbone-Moncrief torus (ref: src/pgen/torus.cpp) ------------------- */
/* Host table of the torus solution over one MeshBlock column (the solution is axisymmetric; log / exp / pow / sin /
 * cos evaluated with the host libm in the reference's order, ref: torus.cpp:285-333,745-833).
 * geom_par as pb_build_geometry_host; torus_par = {rin, rmax, kappa, Gamma, nsub}; geom_tab_host = the geometry table
 * of the same MeshBlock column (the pgen reads the metric through the cached system's interpolating accessors,
 * ref: src/geometry/cached_system.hpp:132-236). out_host[5][tk][nj][ni] = rho, u, W v^1, W v^2, W v^3 inside the torus,
 * 0 outside (before jitter and floors); tk = cfg->gk. PB_EINVAL where the reference aborts ("discr < 0"). */
int pb_pgen_torus_table_host(const double *geom_par, const double *torus_par, const pb_cfg_t *cfg, const double xmin[3],
                             const double dx[3], const double *geom_tab_host, double *out_host);
/* ref: torus.cpp:285-362 — p.density, p.velocity, p.energy (x (1 + u_jitter (rand - 0.5))), atmosphere at the floors,
 * temperature, pressure, gamma1 on every zone of the pack; p.bfield = 0. table_dev: [geom slot][5][gk][nj][ni]. */
int pb_pgen_torus_fill(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *table_dev, double u_jitter,
                       unsigned long long seed, void *stream);
/* ref: torus.cpp:435-470 — A_phi = max(<rho>_corner - 0.2, 0), p.bfield = curl A / sqrt(gamma) (unnormalised) */
int pb_pgen_torus_bfield(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream);
/* ref: torus.cpp:849-930 (ComputeBetas) over the interior zones of the pack: out_host = {min beta over zones with
 * rho > rho_min_bnorm, max b^2, max pressure}. partial_dev: 3 doubles per block. Synchronises the stream. */
int pb_pgen_betas(const pb_pack_t *pack, const pb_cfg_t *cfg, double rho_min_bnorm, double *partial_dev,
                  double out_host[3], void *stream);
/* ref: torus.cpp:721-728 — p.bfield *= factor on interior zones */
int pb_scale_bfield(const pb_pack_t *pack, const pb_cfg_t *cfg, double factor, void *stream);

/* ---- tasks ------------------------------------------------------------------------------ */
/* ref: TaskStatus fluid::CalculateFluxes(MeshBlockData<Real>*), src/fluid/fluid.cpp:659-773.
 * Fused reconstruction + Riemann solve per direction; writes pack->flux and pack->fsig. */
int pb_calculate_fluxes(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream);
/* ref: TaskStatus fluid::FluxCT(MeshBlockData<Real>*), src/fluid/fluid.cpp:775-841 */
int pb_flux_ct(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream);
/* ref: TaskStatus fixup::FixFluxes(MeshBlockData<Real>*), src/fixup/fixup.cpp:728-1078 */
int pb_fix_fluxes(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream);
/* ref: parthenon::Update::FluxDivergence<MeshData>, external/parthenon/src/interface/update.cpp:63-85 */
int pb_flux_divergence(const pb_pack_t *pack, const pb_cfg_t *cfg, double *dudt, void *stream);
/* ref: fluid::CalculateFluidSourceTerms(MeshData*, MeshData*), src/fluid/fluid.cpp:566-657;
 * gsrc [b][4][cells] = sources of S1,S2,S3,tau on interior zones */
int pb_source_terms(const pb_pack_t *pack, const pb_cfg_t *cfg, double *gsrc, void *stream);
/* ref: parthenon::Update::WeightedSumData (SumData / AverageIndependentData /
 * UpdateIndependentData), external/parthenon/src/interface/update.hpp:71-136:
 * z = w1*x + w2*y over n doubles (entire arrays, ghosts included) */
int pb_weighted_sum(const double *x, const double *y, double w1, double w2, double *z, long long n,
                    void *stream);
/* Fusion of FluxDivergence + CalculateFluidSourceTerms + SumData + AverageIndependentData +
 * UpdateIndependentData (ref: src/phoebus_driver.cpp:560-576):
 *   dudt = -div F (+ gsrc on momentum, energy);  uout = (beta*u0 + (1-beta)*ubase) + bdt*dudt
 * on interior zones and, if update_ghosts, uout = beta*u0 + (1-beta)*ubase on ghost zones (the
 * reference's WeightedSumData runs over entire arrays; dUdt is zero outside the interior).
 * dudt_out / gsrc_out may be NULL (not materialised). u0 is NOT averaged in place. */
int pb_update_fused(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *u0, const double *ubase,
                    double beta, double bdt, double *uout, double *dudt_out, double *gsrc_out,
                    int update_ghosts, void *stream);
/* The same update with the sources read from gsrc (written earlier by pb_source_terms). */
int pb_update_from_sources(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *u0, const double *ubase,
                           double beta, double bdt, double *uout, const double *gsrc, double *dudt_out,
                           int update_ghosts, void *stream);
/* ref: fluid::ConservedToPrimitiveRegion / ConservedToPrimitiveRobust, src/fluid/fluid.cpp:452-517,
 * src/fluid/con2prim_robust.hpp:300-453; inclusive index ranges */
int pb_cons_to_prim(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int ib[2],
                    const int jb[2], const int kb[2], void *stream);
/* ref: fixup::ConservedToPrimitiveFixup, src/fixup/fixup_c2p.cpp:44-393 */
int pb_c2p_fixup(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, void *stream);
/* ref: fixup::ApplyFloors, src/fixup/fixup.cpp:286-726. counters (device, 3 x uint64, may be
 * NULL): [0] floors applied, [1] ceilings applied, [2] in-kernel c2p failures; accumulated. */
int pb_apply_floors(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons,
                    unsigned long long *counters, void *stream);
/* (Re)build pack->zflags from the current p.* and fail fields: call once before the first
 * pb_rk_stage* on a pack, and again whenever something else has written p.* or fail. */
int pb_scan_zone_flags(const pb_pack_t *pack, const pb_cfg_t *cfg, void *stream);
/* ref: fluid::PrimitiveToConservedRegion, src/fluid/fluid.cpp:368-449 */
int pb_prim_to_cons(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int ib[2],
                    const int jb[2], const int kb[2], void *stream);
/* Ghost fill between blocks of this pack (same GPU). ref: parthenon SendBoundBufs/SetBounds,
 * external/parthenon/src/bvals/comms/boundary_communication.cpp:49-346. Ships the reference's
 * FillGhost variable set: cons, plus (bc_prim) prim, temperature, c2p_mu, fail. */
int pb_halo_exchange_local(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, void *stream);
/* Inter-GPU halos: a "halo plan" is a device int array of nseg segments, 10 ints each:
 * {block, i0, i1, j0, j1, k0, k1 (inclusive cell ranges), offset_lo, offset_hi (segment start in
 * the buffer, in zones; 64-bit split in two ints), reserved}. pack gathers the listed zones of
 * every shipped field into buf; unpack scatters buf into the listed (ghost) zones.
 * Buffer layout per segment: [field comp][k][j][i]. nvar = pb_halo_nvar(cfg). */
int pb_halo_nvar(const pb_cfg_t *cfg);
int pb_halo_pack(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, const int *plan,
                 int nseg, double *buf, void *stream);
int pb_halo_unpack(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const int *plan,
                   int nseg, const double *buf, void *stream);
/* ---- inter-GPU ghost exchange over NCCL (ref: parthenon SendBoundBufs / ReceiveBoundBufs / SetBounds over MPI,
 * external/parthenon/src/bvals/comms/boundary_communication.cpp:49-346). NCCL is resolved at run time (dlopen of
 * libnccl.so.2); none of this is needed on one GPU. One process per GPU: rank 0 creates the id, every rank receives it
 * out of band (MPI_Bcast, torch.distributed broadcast, a file ...) and calls pb_comm_init. */
int pb_comm_unique_id(char out[128]);
int pb_comm_init(void **comm_out, int rank, int world, const char id[128]);
int pb_comm_destroy(void *comm);
/* All segments of all peers in ONE send and ONE receive buffer (device); NCCL moves the per-peer slices. */
typedef struct pb_halo_plan {
  const int *send_plan; /* device [nsend][10], rows as pb_halo_pack */
  int nsend;
  const int *recv_plan; /* device [nrecv][10] */
  int nrecv;
  double *send_buf, *recv_buf; /* device */
  int npeers;
  const int *peer;                         /* host [npeers] rank of each peer */
  const long long *send_off, *send_cnt;    /* host [npeers] slice of send_buf, in doubles */
  const long long *recv_off, *recv_cnt;    /* host [npeers] slice of recv_buf */
} pb_halo_plan_t;
/* begin: pack on `stream`, then one grouped ncclSend / ncclRecv per peer on the communicator's own stream. Between begin and
 * end the caller issues the work that hides the transfer on `stream` (pb_halo_exchange_local, pb_rk_stage_post_early).
 * end: `stream` waits for the arrival and unpacks into the ghost zones. */
int pb_halo_exchange_begin(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, const pb_halo_plan_t *plan,
                           void *comm, void *stream);
int pb_halo_exchange_end(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const pb_halo_plan_t *plan, void *comm,
                         void *stream);
/* min over ranks of n device doubles, in place on `stream` (the dt reduction, ref: parthenon driver.cpp:237) */
int pb_comm_allreduce_min(void *comm, double *value_dev, int n, void *stream);
/* Time-step control on the device (ref: parthenon driver.cpp:216-240): dt <- min(2 dt, cfl * est), est = the value
 * pb_estimate_dt reduced into; no host round trip. */
int pb_dt_control(const double *est_dev, double cfl, double *dt_dev, void *stream);
/* One RK stage captured as a CUDA graph (launch-latency bound packs: strong scaling): the launches of pb_rk_stage
 * (comm == NULL) or of pb_rk_stage_pre / exchange with the interior con2prim under the transfer / pb_rk_stage_post2
 * (comm != NULL), with dt read from device memory at run time. Call after at least one ordinary stage on the same pack
 * (work spaces are sized then); per-task profiling must be off. */
int pb_stage_graph_create(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1, double beta,
                          const double *dt_dev, unsigned long long *counters, const pb_halo_plan_t *plan, void *comm,
                          void *stream, void **graph_out);
int pb_stage_graph_launch(void *graph, void *stream);
int pb_stage_graph_destroy(void *graph);

/* Physical boundary ghost fill + Boundaries::ConvertBoundaryConditions,
 * ref: src/phoebus_boundaries/phoebus_boundaries.cpp:44-448 */
int pb_apply_bcs(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, void *stream);
/* The two halves separately: what = 1 ghost fill only (parthenon::ApplyBoundaryConditions as Mesh::Initialize's
 * CommunicateBoundaries runs it, ref: external/parthenon/src/mesh/mesh.cpp:692-706), 2 conversion only, 3 both. */
int pb_apply_bcs_ex(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, int what, void *stream);
/* ref: TaskStatus fluid::CalculateDivB(MeshBlockData<Real>*), src/fluid/fluid.cpp:843-897 (diagnostic run with the
 * last stage of a cycle, ref: src/phoebus_driver.cpp:724-726). divb: device [b][cells], interior zones written. */
int pb_calculate_divb(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, double *divb, void *stream);
/* History reductions of the fluid package (the .hst columns, ref: src/fluid/fluid.cpp:325-347,
 * src/analysis/history.hpp:50-86) over the interior zones of the pack: out_host = {maximum density (p.density),
 * total baryon number (sum c.density), total conserved energy tau, total X1, X2, X3 momentum}. Plain sums, as the
 * reference's (its volume weighting never applies). partial_dev: device scratch, 6 doubles per block.
 * Synchronises the stream (it returns host values, like the reference's reduction). */
int pb_history_fluid(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, double *partial_dev,
                     double out_host[6], void *stream);
/* Horizon-flux history of a black-hole run (ref: src/analysis/history.cpp:25-222, registered in
 * src/geometry/geometry.cpp:67-99): out_host = {mdot, jet_energy_flux, jet_momentum_flux, magnetic_Phi} at x1 = xh
 * (FMKS: xh = log(1 + sqrt(1 - a^2))); jets count zones with magnetisation > sigma_cutoff (fluid/sigma_cutoff).
 * partial_dev: device scratch, 4 doubles per block. Synchronises the stream. */
int pb_history_torus(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, double xh, double sigma_cutoff,
                     double *partial_dev, double out_host[4], void *stream);
/* ref: fluid::EstimateTimestepBlock, src/fluid/fluid.cpp:899-944, min over the pack, without
 * the cfl factor. dt_out: device double, combined with atomic min (caller initialises). */
int pb_estimate_dt(const pb_pack_t *pack, const pb_cfg_t *cfg, int has_face_speeds, double *dt_out,
                   void *stream);

/* One Runge-Kutta stage in the order of PhoebusDriver::RungeKuttaStage's fluid sub-graph
 * (ref: src/phoebus_driver.cpp:204-738) for a single-GPU pack: sources, fluxes, CT, FixFluxes,
 * update, stale-prim fixup + floors, local halo, BCs, c2p, fixup, floors.
 * sc0/sc1/base: conserved containers (stage in, stage out, "base"). */
int pb_rk_stage(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1,
                double beta, double dt, unsigned long long *counters, void *stream);
/* first half (through the stale-prim floors) and second half (BCs onwards) of pb_rk_stage, for
 * callers that put an inter-GPU halo exchange in between */
int pb_rk_stage_pre(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1,
                    double beta, double dt, unsigned long long *counters, void *stream);
int pb_rk_stage_post(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc1,
                     unsigned long long *counters, void *stream);
/* The second half in two parts, to hide an inter-GPU halo transfer under interior work: _post_early (con2prim on the
 * interior zones of pack->inner_blocks) may be issued once the same-GPU halo copy has been launched and before the
 * remote ghosts have arrived; _post2(early_done = 1) then runs the BCs, con2prim on everything else, fixup and floors.
 * _post2(early_done = 0) == pb_rk_stage_post. Results are bit-identical either way. */
int pb_rk_stage_post_early(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc1, void *stream);
/* 1 if the early / late split of con2prim pays for this pack (at least a third of its MeshBlocks touch no mesh boundary), else 0:
 * callers then skip pb_rk_stage_post_early and pass early_done = 0 (fewer launches; what a small pack needs) */
int pb_early_c2p_pays(const pb_pack_t *pack);
int pb_rk_stage_post2(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc1, unsigned long long *counters,
                      int early_done, void *stream);

/* Optional per-task device timing of pb_rk_stage*: CUDA events on the launching stream around each
 * task group. pb_profile_read synchronises the recorded events, returns the accumulated milliseconds
 * and call counts per PB_TASK_* and resets the accumulators. */
#define PB_TASK_FLUX 0
#define PB_TASK_CT_FIX 1
#define PB_TASK_SOURCE 2
#define PB_TASK_UPDATE 3
#define PB_TASK_STALE_FIXUP_FLOORS 4
#define PB_TASK_HALO 5
#define PB_TASK_BCS 6
#define PB_TASK_C2P 7
#define PB_TASK_FIXUP_FLOORS 8
#define PB_TASK_RECON 9    /* nested in PB_TASK_FLUX: the recon_kernel launches of the two-launch flux path */
#define PB_TASK_RIEMANN 10 /* nested in PB_TASK_FLUX: the riemann_kernel launches */
#define PB_NTASK 11
int pb_profile_enable(int on);
/* Device self-test of the library's deferred-branch FP64 divide / reciprocal / square root against
 * IEEE a/b, 1.0/x, sqrt(x) on n pseudo-random operand pairs with exponents in +-exponent_span.
 * counts_dev (device, 5 x uint64, caller zeroes): mismatches of div, divz, rcp, root; [4] = cases
 * routed to the exact path. The hot path relies on [0..3] being zero (bit parity with the reference). */
int pb_selftest_arith(long long n, unsigned long long seed, int exponent_span, unsigned long long *counts_dev,
                      void *stream);
int pb_profile_read(double ms[PB_NTASK], int calls[PB_NTASK]);

/* Run-time options of the library (initial values from the environment variables in parentheses):
 *   "flux_split" 1|0  (PB_FLUX_MODE=split|fused)  two-launch flux path when the pack has recon_lr, else the fused kernel
 *   "riemann_blocks" 1|0 (PB_RIEMANN_BLOCKS=1|0) two Riemann launches per direction: MeshBlocks whose face states carry no
 *                magnetic field (flagged by the reconstruction) go through the exact unmagnetised variant at a higher
 *                occupancy, the others through the general kernel; 0: one launch, variant chosen per warp
 *   "flag_compact" 1|0 (PB_FLAG_PASS=compact|packed) the flag-driven fixup / floor passes compact the flagged zones into a
 *                queue and handle one zone per thread (default), or walk 16 flags per thread serially
 *   "flux_tma"   1|0  (PB_FLUX_STAGE=tma|cpasync) the fused flux kernel stages its stencil tile with one TMA tensor load
 *                (cp.async.bulk.tensor, 5-D box of the primitive slab) per CTA, or with per-thread cp.async copies
 *   "c2p_arith"  2|3|0|1 (PB_C2P_ARITH=queue|split|exact|inline) con2prim: one launch with warp-level parking of slowly
 *                converging zones + redo queue for the zones ArithFast rejects (default); root find and unwrap as separate
 *                launches (same parking); the compiler's divide / sqrt; ArithFast with an inline exact re-solve.
 *   "recon_cpt"  1|2  (PB_RECON_CPT=1|2) cells per thread in the reconstruction launch of the split flux path (2 was
 *                measured slower, profiles/r02g_variants.txt)
 *   "ghost_plan" 1|0  (PB_GHOST_PLAN=1|0) the stage sequencer uses the pack's ghost plan (pb_ghost_plan_host) when present
 *                All settings produce identical bits. */
int pb_set_option(const char *key, int value);
int pb_get_option(const char *key);
/* Faces / zones that the fast arithmetic rejected and riemann_redo_kernel / c2p_redo_kernel recomputed with the compiler's divide / sqrt,
 * accumulated since the library was loaded (synchronises the device). */
unsigned long long pb_arith_redo_count(void);
/* Unbracketed con2prim roots counted since the library was loaded or last reset. The reference aborts the run on the first
 * one (PARTHENON_FAIL, ref: src/phoebus_utils/root_find.hpp:115-118); here the zone is marked failed and the caller decides.
 * Negative on error. Synchronises the device. */
long long pb_c2p_unbracketed_count(int reset);
/* Con2prim statistics (the reference's CON2PRIM_STATISTICS histogram, ref: src/fluid/con2prim_statistics.hpp:22-41): enable != 0
 * starts counting (zeroing the counters), 0 stops; out_host (may be NULL) receives the 36 counters accumulated so far:
 * [0..31] regula-falsi iterations per solve (31 = 31 or more), [32] solves that used a floor / ceiling, [33] failed zones,
 * [34] zones parked by the warp vote, [35] zones unwrapped. Counted by the default (2) and the split (3) con2prim paths. Synchronises the device. */
int pb_c2p_stats(int enable, unsigned long long out_host[36]);
/* Same idea one level up: riemann.hpp's prim_to_flux (face state -> U, F, v+-) evaluated with the
 * deferred-branch arithmetic and with the compiler's, on n synthetic MHD face states (flat != 0:
 * Minkowski metric; else a random metric). counts_dev[0] = states whose outputs differ in any bit,
 * [1] = states routed to the exact path; first_dev (16 doubles): the first differing case. */
int pb_selftest_riemann(long long n, unsigned long long seed, const pb_cfg_t *cfg, int flat,
                        unsigned long long *counts_dev, double *first_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PHOEBUS_B200_H_ */
