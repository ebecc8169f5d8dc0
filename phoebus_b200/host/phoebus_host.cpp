// Host side of the B200 fluid path: deck reader, package Initialize functions, the Mesh (block
// table, coordinates, tables, device slabs) and the reference-named task functions, each a thin
// forward to the C ABI of libphoebus_b200.so. See phoebus_host.hpp for the reference citations.
#include <algorithm>
#include <chrono>
#include <fstream>
#include <thread>

#include "phoebus_host.hpp"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include <cuda_runtime_api.h>

namespace {

[[noreturn]] void Fail(const std::string &msg) { throw std::runtime_error(msg); }

void CheckPb(int rc, const char *what) {
  if (rc != PB_OK) Fail(std::string(what) + " failed (" + std::to_string(rc) + "): " + pb_last_error());
}
void CheckCuda(cudaError_t e, const char *what) {
  if (e != cudaSuccess) Fail(std::string(what) + ": " + cudaGetErrorString(e));
}

std::string Trim(const std::string &s) {
  const size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  const size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

int RangeArg(const IndexRange &r, int out[2]) {
  out[0] = r.s;
  out[1] = r.e;
  return 0;
}

const Params &Pkg(const parthenon::PackView *rc, const char *name) {
  return rc->GetMeshPointer()->packages.Get(name)->AllParams();
}
bool FluidActive(const parthenon::PackView *rc) { return Pkg(rc, "fluid").Get<bool>("active"); }
bool FixupFlag(const parthenon::PackView *rc, const char *key) { return Pkg(rc, "fixup").Get<bool>(key); }

} // namespace

// =================================================================================================
// ParameterInput
// =================================================================================================
namespace parthenon {

void ParameterInput::LoadFromFile(const std::string &filename) {
  std::ifstream f(filename);
  if (!f) Fail("ParameterInput: cannot open " + filename);
  std::stringstream ss;
  ss << f.rdbuf();
  LoadFromString(ss.str());
}

void ParameterInput::LoadFromString(const std::string &text) {
  std::istringstream in(text);
  std::string line, cur;
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    line = Trim(line);
    if (line.empty()) continue;
    if (line[0] == '<') {
      const size_t close = line.find('>');
      if (close == std::string::npos) Fail("ParameterInput: unterminated block name: " + line);
      cur = line.substr(1, close - 1);
      blocks_[cur];
      continue;
    }
    const size_t eq = line.find('=');
    if (eq == std::string::npos || cur.empty()) continue;
    blocks_[cur][Trim(line.substr(0, eq))] = Trim(line.substr(eq + 1));
  }
}

void ParameterInput::ModifyFromCmdline(int argc, char *argv[]) {
  for (int i = 0; i < argc; ++i) {
    const std::string a = argv[i];
    const size_t eq = a.find('='), slash = a.rfind('/', eq);
    if (eq == std::string::npos || slash == std::string::npos) Fail("bad override (block/key=value): " + a);
    blocks_[a.substr(0, slash)][a.substr(slash + 1, eq - slash - 1)] = a.substr(eq + 1);
  }
}

const std::string *ParameterInput::Find(const std::string &block, const std::string &name) const {
  auto b = blocks_.find(block);
  if (b == blocks_.end()) return nullptr;
  auto k = b->second.find(name);
  return k == b->second.end() ? nullptr : &k->second;
}
const std::string &ParameterInput::Require(const std::string &block, const std::string &name) const {
  const std::string *v = Find(block, name);
  if (!v) Fail("Parameter name '" + name + "' not found in block '" + block + "'");
  return *v;
}
bool ParameterInput::DoesParameterExist(const std::string &block, const std::string &name) const {
  return Find(block, name) != nullptr;
}
static bool ToBool(const std::string &v) {
  std::string s;
  for (char c : v) s += (char)std::tolower((unsigned char)c);
  return s == "true" || s == "1";
}
int ParameterInput::GetInteger(const std::string &b, const std::string &n) const { return std::stoi(Require(b, n)); }
Real ParameterInput::GetReal(const std::string &b, const std::string &n) const { return std::strtod(Require(b, n).c_str(), nullptr); }
bool ParameterInput::GetBoolean(const std::string &b, const std::string &n) const { return ToBool(Require(b, n)); }
std::string ParameterInput::GetString(const std::string &b, const std::string &n) const { return Require(b, n); }
int ParameterInput::GetOrAddInteger(const std::string &b, const std::string &n, int def) {
  if (const std::string *v = Find(b, n)) return std::stoi(*v);
  blocks_[b][n] = std::to_string(def);
  return def;
}
Real ParameterInput::GetOrAddReal(const std::string &b, const std::string &n, Real def) {
  if (const std::string *v = Find(b, n)) return std::strtod(v->c_str(), nullptr);
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.17g", def);
  blocks_[b][n] = buf;
  return def;
}
bool ParameterInput::GetOrAddBoolean(const std::string &b, const std::string &n, bool def) {
  if (const std::string *v = Find(b, n)) return ToBool(*v);
  blocks_[b][n] = def ? "true" : "false";
  return def;
}
std::string ParameterInput::GetOrAddString(const std::string &b, const std::string &n, const std::string &def) {
  if (const std::string *v = Find(b, n)) return *v;
  blocks_[b][n] = def;
  return def;
}

} // namespace parthenon

// =================================================================================================
// package Initialize functions: deck -> Params, defaults as in the reference
// =================================================================================================
namespace fluid {
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin) {
  auto physics = std::make_shared<StateDescriptor>("fluid");
  const bool active = pin->GetOrAddBoolean("physics", "hydro", true);
  physics->AddParam("active", active);
  physics->AddParam("zero_fluxes", pin->GetOrAddBoolean("fluid", "zero_fluxes", false));
  physics->AddParam("zero_sources", pin->GetOrAddBoolean("fluid", "zero_sources", false));
  physics->AddParam("cfl", pin->GetOrAddReal("fluid", "cfl", 0.8));
  const std::string c2p_method = pin->GetOrAddString("fluid", "c2p_method", "robust");
  if (c2p_method != "robust") Fail("only fluid/c2p_method = robust is on the accelerated path");
  physics->AddParam("c2p_method", c2p_method);
  physics->AddParam("c2p_tol", pin->GetOrAddReal("fluid", "c2p_tol", 1.e-8));
  physics->AddParam("c2p_max_iter", pin->GetOrAddInteger("fluid", "c2p_max_iter", 20));
  physics->AddParam("c2p_floor_scale_fac", pin->GetOrAddReal("fluid", "c2p_floor_scale_fac", 1.));
  physics->AddParam("c2p_fail_on_floors", pin->GetOrAddBoolean("fluid", "c2p_fail_on_floors", false));
  physics->AddParam("c2p_fail_on_ceilings", pin->GetOrAddBoolean("fluid", "c2p_fail_on_ceilings", false));
  const std::string recon = pin->GetOrAddString("fluid", "recon", "linear");
  const int nghost = pin->GetOrAddInteger("parthenon/mesh", "nghost", 2);
  int rt;
  if (recon == "weno5" || recon == "weno5z") {
    if (nghost < 4) Fail("weno5 requires 4+ ghost cells");
    rt = 2;
  } else if (recon == "mp5") {
    if (nghost < 4) Fail("mp5 requires 4+ ghost cells");
    rt = 3;
  } else if (recon == "linear") {
    rt = 1;
  } else if (recon == "constant") {
    rt = 0;
  } else {
    Fail("Reconstruction method invalid");
  }
  physics->AddParam("Recon", rt);
  const std::string solver = pin->GetOrAddString("fluid", "riemann", "hll");
  int rs;
  if (solver == "llf") rs = 0;
  else if (solver == "hll") rs = 1;
  else Fail("Riemann solver invalid");
  physics->AddParam("RiemannSolver", rs);
  if (pin->GetOrAddBoolean("fluid", "Ye", false)) Fail("fluid/Ye = true is not on the accelerated path");
  physics->AddParam("Ye", false);
  physics->AddParam("mhd", pin->GetOrAddBoolean("fluid", "mhd", false));
  physics->AddParam("bc_vars", pin->GetOrAddString("phoebus/mesh", "bc_vars", "conserved"));
  physics->AddParam("has_face_speeds", false, true);
  return physics;
}
} // namespace fluid

namespace fixup {
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin) {
  auto fix = std::make_shared<StateDescriptor>("fixup");
  const bool mhd = pin->GetOrAddBoolean("fluid", "mhd", false);
  if (pin->GetOrAddBoolean("physics", "rad", false)) Fail("radiation is not on the accelerated path");
  fix->AddParam("enable_flux_fixup", pin->GetOrAddBoolean("fixup", "enable_flux_fixup", false));
  fix->AddParam("enable_ox1_fmks_inflow_check", pin->GetOrAddBoolean("fixup", "enable_ox1_fmks_inflow_check", false));
  bool floors = pin->GetOrAddBoolean("fixup", "enable_floors", false);
  const bool c2p_fixup = pin->GetOrAddBoolean("fixup", "enable_c2p_fixup", false);
  const bool ceilings = pin->GetOrAddBoolean("fixup", "enable_ceilings", false);
  const bool mhd_floors = pin->GetOrAddBoolean("fixup", "enable_mhd_floors", false) && mhd;
  if (c2p_fixup || mhd_floors) floors = true; // a fixup needs somewhere to fall back to
  fix->AddParam("enable_floors", floors);
  fix->AddParam("enable_c2p_fixup", c2p_fixup);
  fix->AddParam("enable_ceilings", ceilings);
  fix->AddParam("enable_mhd_floors", mhd_floors);
  // disabled floors are -DBL_MAX, disabled ceilings +DBL_MAX (ref: fixup.hpp:74-76,182-185)
  int ftype = 0;
  Real rho0 = -DBL_MAX, sie0 = -DBL_MAX, rexp = 0.0, sexp = 0.0;
  if (floors) {
    const std::string t = pin->GetOrAddString("fixup", "floor_type", "ConstantRhoSie");
    static const char *names[] = {"ConstantRhoSie", "ExpX1RhoSie", "ExpX1RhoU", "X1RhoSie", "RRhoSie"};
    ftype = -1;
    for (int i = 0; i < 5; ++i)
      if (t == names[i]) ftype = i;
    if (ftype < 0) Fail("invalid floor_type");
    rho0 = pin->GetOrAddReal("fixup", "rho0_floor", 0.0);
    if (ftype >= 2) { // ExpX1RhoU, X1RhoSie, RRhoSie read u0_floor / u_exp_floor (ref: fixup.cpp:148-165)
      sie0 = pin->GetOrAddReal("fixup", "u0_floor", 0.0);
      rexp = pin->GetOrAddReal("fixup", "rho_exp_floor", -2.0);
      sexp = pin->GetOrAddReal("fixup", "u_exp_floor", -3.0);
    } else { // exponents are read (and ignored by ConstantRhoSie) so that every front end flattens alike
      sie0 = pin->GetOrAddReal("fixup", "sie0_floor", 0.0);
      rexp = pin->GetOrAddReal("fixup", "rho_exp_floor", -2.0);
      sexp = pin->GetOrAddReal("fixup", "sie_exp_floor", -1.0);
    }
  }
  fix->AddParam("floor_type", ftype);
  fix->AddParam("rho0_floor", rho0);
  fix->AddParam("sie0_floor", sie0);
  fix->AddParam("rho_exp_floor", rexp);
  fix->AddParam("sie_exp_floor", sexp);
  Real gam0 = DBL_MAX, siec = DBL_MAX;
  if (ceilings) {
    if (pin->GetOrAddString("fixup", "ceiling_type", "ConstantGamSie") != "ConstantGamSie") Fail("invalid ceiling type");
    gam0 = pin->GetOrAddReal("fixup", "gam0_ceiling", 1000.0);
    siec = pin->GetOrAddReal("fixup", "sie0_ceiling", 100.0);
  }
  fix->AddParam("gam0_ceiling", gam0);
  fix->AddParam("sie0_ceiling", siec);
  Real bsqorho0 = DBL_MAX, bsqou0 = DBL_MAX;
  if (mhd_floors) {
    if (pin->GetOrAddString("fixup", "mhd_ceiling_type", "ConstantBsqRat") != "ConstantBsqRat") Fail("invalid mhd ceiling type");
    bsqorho0 = pin->GetOrAddReal("fixup", "bsqorho0_ceiling", 50.);
    bsqou0 = pin->GetOrAddReal("fixup", "bsqou0_ceiling", 2500.);
  }
  fix->AddParam("bsqorho0_ceiling", bsqorho0);
  fix->AddParam("bsqou0_ceiling", bsqou0);
  const std::string strat = pin->GetOrAddString("fixup", "fluid_c2p_failure_strategy", "interpolate");
  if (strat != "interpolate" && strat != "floors") Fail("fixup/fluid_c2p_failure_strategy invalid");
  fix->AddParam("fluid_c2p_failure_strategy", strat == "interpolate" ? 0 : 1);
  return fix;
}
} // namespace fixup

namespace Microphysics {
namespace EOS {
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin) {
  auto eos = std::make_shared<StateDescriptor>("eos");
  const std::string type = pin->GetOrAddString("eos", "type", "IdealGas");
  if (type != "IdealGas") Fail("only eos/type = IdealGas is on the accelerated path");
  const Real gamma = pin->GetOrAddReal("eos", "Gamma", 5.0 / 3.0);
  eos->AddParam("type", type);
  eos->AddParam("gm1", gamma - 1.0);
  eos->AddParam("Cv", pin->GetOrAddReal("eos", "Cv", 1.0));
  eos->AddParam("rho_min", pin->GetOrAddReal("eos", "rho_min", 0.0));
  eos->AddParam("sie_min", pin->GetOrAddReal("eos", "sie_min", 0.0));
  return eos;
}
} // namespace EOS
} // namespace Microphysics

namespace Geometry {
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin) {
  auto geom = std::make_shared<StateDescriptor>("geometry");
  std::string kind = pin->GetOrAddString("geometry", "kind", "minkowski");
  for (char &c : kind) c = (char)std::tolower((unsigned char)c);
  if (kind != "minkowski" && kind != "fmks") Fail("geometry/kind must be minkowski or fmks");
  geom->AddParam("kind", kind == "fmks" ? 1 : 0);
  geom->AddParam("axisymmetric", pin->GetOrAddBoolean("geometry", "axisymmetric", false));
  // FMKS keys, ref: src/geometry/fmks.cpp:38-91
  const int ng = pin->GetOrAddInteger("parthenon/mesh", "nghost", 2);
  const Real x1min = pin->GetOrAddReal("parthenon/mesh", "x1min", 0.0);
  const Real x1max = pin->GetOrAddReal("parthenon/mesh", "x1max", 1.0);
  const int nx1 = pin->GetOrAddInteger("parthenon/mesh", "nx1", 1);
  const Real dx1 = (x1max - x1min) / nx1;
  const Real x0 = std::min(x1min, std::max(0.0, x1min - ng * dx1));
  std::vector<Real> par = {pin->GetOrAddReal("geometry", "a", 0.0),
                           pin->GetOrAddReal("coordinates", "hslope", 0.3),
                           pin->GetOrAddReal("coordinates", "poly_xt", 0.82),
                           pin->GetOrAddReal("coordinates", "poly_alpha", 14.0),
                           x0,
                           pin->GetOrAddReal("coordinates", "smooth", 0.5),
                           pin->GetOrAddReal("coordinates", "hexp_br", 1e30),
                           pin->GetOrAddReal("coordinates", "hexp_nsq", 4.0),
                           pin->GetOrAddReal("coordinates", "hexp_csq", 4.0),
                           pin->GetOrAddBoolean("coordinates", "derefine_poles", true) ? 1.0 : 0.0};
  geom->AddParam("par", par);
  return geom;
}
} // namespace Geometry

namespace phoebus {
parthenon::Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  parthenon::Packages_t packages;
  packages.Add(Microphysics::EOS::Initialize(pin.get()));
  packages.Add(Geometry::Initialize(pin.get()));
  packages.Add(fluid::Initialize(pin.get()));
  packages.Add(fixup::Initialize(pin.get()));
  return packages;
}
} // namespace phoebus

// =================================================================================================
// Mesh
// =================================================================================================
namespace parthenon {

namespace {
// LogicalLocation::IndexToSymmetrizedCoordinate (ref: logical_location.cpp:61-74)
double IndexToSym(long index, int bloc, long nrange) {
  const long noffset = index - nrange / 2;
  const long noffset_ceil = index - (nrange + 1) / 2;
  return (double)(noffset + noffset_ceil + bloc) / (2.0 * (double)nrange);
}
// RegionSize::SymmetrizedLogicalToActualPosition (ref: defs.hpp:98-101)
double SymToActual(double u, double xmin, double xmax) { return 0.5 * (xmin + xmax) + (u * xmax - u * xmin); }
int BcCode(const std::string &name) {
  if (name == "periodic") return PB_BC_PERIODIC;
  if (name == "gr_outflow" || name == "outflow") return PB_BC_OUTFLOW;
  if (name == "reflect") return PB_BC_REFLECT;
  if (name == "polar") return PB_BC_POLAR;
  Fail("boundary condition '" + name + "' is not on the accelerated path");
}
} // namespace

Mesh::Mesh(ParameterInput *pin, Packages_t pkgs) : packages(std::move(pkgs)) {
  const Params &fl = packages.Get("fluid")->AllParams();
  const Params &fx = packages.Get("fixup")->AllParams();
  const Params &eos = packages.Get("eos")->AllParams();
  int mesh_nx[3];
  for (int d = 0; d < 3; ++d) {
    const std::string n = std::to_string(d + 1);
    mesh_nx[d] = pin->GetOrAddInteger("parthenon/mesh", "nx" + n, 1);
    mesh_xmin_[d] = pin->GetOrAddReal("parthenon/mesh", "x" + n + "min", 0.0);
    mesh_xmax_[d] = pin->GetOrAddReal("parthenon/mesh", "x" + n + "max", 1.0);
  }
  ndim = mesh_nx[2] > 1 ? 3 : (mesh_nx[1] > 1 ? 2 : 1);
  ng = pin->GetOrAddInteger("parthenon/mesh", "nghost", 2);
  for (int d = 0; d < 3; ++d) {
    nx[d] = pin->GetOrAddInteger("parthenon/meshblock", "nx" + std::to_string(d + 1), mesh_nx[d]);
    if (d < ndim && mesh_nx[d] % nx[d]) Fail("Block size is not evenly divisible into the base mesh size.");
    nblk[d] = d < ndim ? std::max(1, mesh_nx[d] / nx[d]) : 1;
    ncell[d] = d < ndim ? nx[d] + 2 * ng : 1;
  }
  cells = (long long)ncell[0] * ncell[1] * ncell[2];
  // boundary names: parthenon's "user" defers to <phoebus> ix1_bc ... (ref: phoebus_boundaries.cpp:450-543)
  for (int d = 0; d < 3; ++d)
    for (int o = 0; o < 2; ++o) {
      const std::string key = std::string(o ? "o" : "i") + "x" + std::to_string(d + 1) + "_bc";
      const std::string parth = pin->GetOrAddString("parthenon/mesh", key, "periodic");
      mesh_bc_[2 * d + o] = parth == "periodic" ? parth : pin->GetOrAddString("phoebus", key, parth);
    }
  for (int d = 0; d < 3; ++d) periodic_[d] = mesh_bc_[2 * d] == "periodic";

  // ---- flatten the package Params into pb_cfg_t
  cfg.ndim = ndim;
  for (int d = 0; d < 3; ++d) cfg.nx[d] = nx[d];
  cfg.ng = ng;
  cfg.mhd = fl.Get<bool>("mhd");
  cfg.recon = fl.Get<int>("Recon");
  cfg.riemann = fl.Get<int>("RiemannSolver");
  cfg.c2p_tol = fl.Get<Real>("c2p_tol");
  cfg.c2p_max_iter = fl.Get<int>("c2p_max_iter");
  cfg.c2p_floor_scale_fac = fl.Get<Real>("c2p_floor_scale_fac");
  cfg.c2p_fail_on_floors = fl.Get<bool>("c2p_fail_on_floors");
  cfg.c2p_fail_on_ceilings = fl.Get<bool>("c2p_fail_on_ceilings");
  cfg.floor_type = fx.Get<int>("floor_type");
  cfg.rho0 = fx.Get<Real>("rho0_floor");
  cfg.sie0 = fx.Get<Real>("sie0_floor");
  cfg.rexp = fx.Get<Real>("rho_exp_floor");
  cfg.sexp = fx.Get<Real>("sie_exp_floor");
  cfg.rho_min_eos = eos.Get<Real>("rho_min");
  cfg.sie_min_eos = eos.Get<Real>("sie_min");
  cfg.gam_max = fx.Get<Real>("gam0_ceiling");
  cfg.sie_max = fx.Get<Real>("sie0_ceiling");
  cfg.bsqorho_max = fx.Get<Real>("bsqorho0_ceiling");
  cfg.bsqou_max = fx.Get<Real>("bsqou0_ceiling");
  cfg.enable_floors = fx.Get<bool>("enable_floors");
  cfg.enable_mhd_floors = fx.Get<bool>("enable_mhd_floors");
  cfg.enable_c2p_fixup = fx.Get<bool>("enable_c2p_fixup");
  cfg.enable_flux_fixup = fx.Get<bool>("enable_flux_fixup");
  cfg.c2p_failure_strategy = fx.Get<int>("fluid_c2p_failure_strategy");
  cfg.gm1 = eos.Get<Real>("gm1");
  cfg.Cv = eos.Get<Real>("Cv");
  cfg.bc_prim = fl.Get<std::string>("bc_vars") == "primitive";
  cfg.ox1_inflow_check = fx.Get<bool>("enable_ox1_fmks_inflow_check");
  nprim = pb_nprim(&cfg);
  ncons = pb_ncons(&cfg);

  // ---- which blocks are mine: compact bricks, largest factors of the rank count on the dimensions with the most blocks
  const char *er = std::getenv("PB_RANK"), *ew = std::getenv("PB_WORLD");
  rank = pin->GetOrAddInteger("phoebus_b200", "rank", er ? std::atoi(er) : 0);
  world = pin->GetOrAddInteger("phoebus_b200", "world", ew ? std::atoi(ew) : 1);
  if (world < 1 || rank < 0 || rank >= world) Fail("phoebus_b200/rank must be in [0, world)");
  {
    std::vector<int> factors;
    int w = world;
    for (int f = 2; w > 1; ++f)
      while (w % f == 0) {
        factors.push_back(f);
        w /= f;
      }
    std::sort(factors.rbegin(), factors.rend());
    for (int f : factors) {
      int best = -1;
      for (int d = 0; d < ndim; ++d)
        if ((nblk[d] / rgrid[d]) % f == 0 && (best < 0 || nblk[d] / rgrid[d] >= nblk[best] / rgrid[best])) best = d;
      if (best < 0) Fail("cannot decompose the MeshBlocks over " + std::to_string(world) + " ranks");
      rgrid[best] *= f;
    }
    for (int d = 0; d < 3; ++d) brick[d] = nblk[d] / rgrid[d];
  }

  BuildBlocks(pin);
  BuildTables(pin);
  if (world > 1) BuildHaloPlans();
  if (pin->GetOrAddBoolean("phoebus_b200", "allocate", true)) {
    if (world > 1) {
      int ndev = 0;
      CheckCuda(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
      const char *lr = std::getenv("PB_LOCAL_RANK");
      CheckCuda(cudaSetDevice((lr ? std::atoi(lr) : rank) % std::max(1, ndev)), "cudaSetDevice");
    }
    Allocate();
    if (world > 1) InitComm(pin);
  }
}

void Mesh::BuildBlocks(ParameterInput *) {
  nblocks = brick[0] * brick[1] * brick[2];
  lx.resize(nblocks);
  blocks_host.assign(nblocks, pb_block_t{});
  const int rc[3] = {rank % rgrid[0], (rank / rgrid[0]) % rgrid[1], rank / (rgrid[0] * rgrid[1])};
  for (int n = 0; n < nblocks; ++n) { // x1 fastest inside the brick
    lx[n] = {rc[0] * brick[0] + n % brick[0], rc[1] * brick[1] + (n / brick[0]) % brick[1], rc[2] * brick[2] + n / (brick[0] * brick[1])};
    local_index_[Gid(lx[n])] = n;
  }
  // the forest is made of trees of pow2^ndim blocks (ref: forest.cpp:70-145)
  int pow2 = 1 << 30;
  for (int d = 0; d < ndim; ++d) pow2 = std::min(pow2, nblk[d] & (~(nblk[d] - 1)));
  for (int n = 0; n < nblocks; ++n) {
    pb_block_t &b = blocks_host[n];
    for (int d = 0; d < 3; ++d) {
      const double xmin = mesh_xmin_[d], xmax = mesh_xmax_[d];
      if (d >= ndim) { // symmetry direction: one zone, no ghosts
        b.dx[d] = (xmax - xmin) / 1;
        b.xmin[d] = xmin - 0 * b.dx[d];
        continue;
      }
      const int ntree = nblk[d] / pow2;
      const int it = lx[n][d] / pow2, il = lx[n][d] % pow2;
      const double tmin = SymToActual(IndexToSym(it, 0, ntree), xmin, xmax);
      const double tmax = SymToActual(IndexToSym(it, 2, ntree), xmin, xmax);
      const double bmin = SymToActual(IndexToSym(il, 0, pow2), tmin, tmax);
      const double bmax = SymToActual(IndexToSym(il, 2, pow2), tmin, tmax);
      b.dx[d] = (bmax - bmin) / nx[d];
      b.xmin[d] = bmin - ng * b.dx[d];
    }
    for (int f = 0; f < 6; ++f) {
      const int d = f / 2, outer = f % 2;
      const bool at_edge = outer ? lx[n][d] == nblk[d] - 1 : lx[n][d] == 0;
      if (d >= ndim) b.bc[f] = periodic_[d] ? PB_BC_PERIODIC : PB_BC_BLOCK;
      else b.bc[f] = at_edge ? BcCode(mesh_bc_[f]) : PB_BC_BLOCK;
    }
    for (int o3 = -1; o3 <= 1; ++o3)
      for (int o2 = -1; o2 <= 1; ++o2)
        for (int o1 = -1; o1 <= 1; ++o1) {
          const int code = (o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1);
          const int off[3] = {o1, o2, o3};
          int nl[3];
          bool none = (o1 == 0 && o2 == 0 && o3 == 0);
          for (int d = 0; d < 3 && !none; ++d) {
            if (off[d] != 0 && d >= ndim) none = true;
            int l = lx[n][d] + off[d];
            if (l < 0 || l >= nblk[d]) {
              if (d < ndim && periodic_[d]) l = (l + nblk[d]) % nblk[d];
              else none = true;
            }
            nl[d] = l;
          }
          if (none) {
            b.nbr[code] = PB_NBR_NONE;
          } else {
            const std::array<int, 3> l = {nl[0], nl[1], nl[2]};
            b.nbr[code] = Owner(l) == rank ? local_index_.at(Gid(l)) : PB_NBR_REMOTE;
          }
        }
  }
}

void Mesh::BuildTables(ParameterInput *) {
  const Params &gp = packages.Get("geometry")->AllParams();
  const int kind = gp.Get<int>("kind");
  const bool axisym = gp.Get<bool>("axisymmetric");
  const std::vector<Real> &par = gp.Get<std::vector<Real>>("par");
  // geometry table extents: broadcast (Minkowski), 2-D (axisymmetric) or full
  int mode;
  if (kind == 0) { cfg.gk = cfg.gj = cfg.gi = 1; mode = 0; }
  else if (axisym || ndim < 3) { cfg.gk = 1; cfg.gj = ncell[1]; cfg.gi = ncell[0]; mode = 1; }
  else { cfg.gk = ncell[2]; cfg.gj = ncell[1]; cfg.gi = ncell[0]; mode = 2; }
  const long long gcell = (long long)cfg.gk * cfg.gj * cfg.gi;
  std::map<long long, int> slots, fslots;
  for (int n = 0; n < nblocks; ++n) {
    long long key = 0;
    if (mode == 1) key = lx[n][0] + (long long)nblk[0] * lx[n][1];
    if (mode == 2) key = n;
    auto it = slots.find(key);
    if (it == slots.end()) {
      const int slot = (int)slots.size();
      it = slots.emplace(key, slot).first;
      geom_host_.resize((size_t)(slot + 1) * PB_GEOM_WORDS * gcell);
      CheckPb(pb_build_geometry_host(kind, par.data(), &cfg, blocks_host[n].xmin, blocks_host[n].dx,
                                     geom_host_.data() + (size_t)slot * PB_GEOM_WORDS * gcell),
              "pb_build_geometry_host");
    }
    blocks_host[n].geom_slot = it->second;
    auto ft = fslots.find(lx[n][0]);
    if (ft == fslots.end()) {
      const int slot = (int)fslots.size();
      ft = fslots.emplace(lx[n][0], slot).first;
      floor_host_.resize((size_t)(slot + 1) * 4 * ncell[0]);
      CheckPb(pb_build_floor_table_host(&cfg, blocks_host[n].xmin, blocks_host[n].dx,
                                        floor_host_.data() + (size_t)slot * 4 * ncell[0]),
              "pb_build_floor_table_host");
    }
    blocks_host[n].floor_slot = ft->second;
  }
}

// Inter-GPU halo plans (mirror of phoebus_b200/mesh.py: halo_plans / merged_halo_plans): for every peer rank the slabs I
// send and the ghost slabs I receive, both sides ordered by (gid of the receiving block, offset code seen from the
// receiver), all peers concatenated into one send and one receive buffer.
void Mesh::BuildHaloPlans() {
  struct Seg {
    long long gid;
    int code, block, r[3][2];
  };
  const int nvar = ncons + (cfg.bc_prim ? nprim + 3 : 0);
  std::map<int, std::vector<Seg>> send, recv;
  auto ranges = [&](const int off[3], bool ghost, int r[3][2]) {
    for (int d = 0; d < 3; ++d) {
      const int n = d < ndim ? nx[d] : 1, g = d < ndim ? ng : 0, lo = d < ndim ? ng : 0;
      if (off[d] == 0) r[d][0] = lo, r[d][1] = lo + n - 1;
      else if (ghost) {
        if (off[d] < 0) r[d][0] = 0, r[d][1] = g - 1;
        else r[d][0] = lo + n, r[d][1] = lo + n + g - 1;
      } else {
        if (off[d] > 0) r[d][0] = lo + n - g, r[d][1] = lo + n - 1;
        else r[d][0] = lo, r[d][1] = lo + g - 1;
      }
    }
  };
  for (int n = 0; n < nblocks; ++n)
    for (int o3 = (ndim >= 3 ? -1 : 0); o3 <= (ndim >= 3 ? 1 : 0); ++o3)
      for (int o2 = (ndim >= 2 ? -1 : 0); o2 <= (ndim >= 2 ? 1 : 0); ++o2)
        for (int o1 = -1; o1 <= 1; ++o1) {
          if (o1 == 0 && o2 == 0 && o3 == 0) continue;
          const int off[3] = {o1, o2, o3}, moff[3] = {-o1, -o2, -o3};
          std::array<int, 3> nl;
          bool none = false;
          for (int d = 0; d < 3; ++d) {
            int l = lx[n][d] + off[d];
            if (l < 0 || l >= nblk[d]) {
              if (d < ndim && periodic_[d]) l = (l + nblk[d]) % nblk[d];
              else none = true;
            }
            nl[d] = l;
          }
          if (none) continue;
          const int peer = Owner(nl);
          if (peer == rank) continue;
          auto code = [](const int o[3]) { return (o[0] + 1) + 3 * (o[1] + 1) + 9 * (o[2] + 1); };
          Seg r{Gid(lx[n]), code(off), n, {}}, s{Gid(nl), code(moff), n, {}};
          ranges(off, true, r.r);
          ranges(off, false, s.r);
          recv[peer].push_back(r);
          send[peer].push_back(s);
        }
  std::vector<int> all;
  for (auto &kv : send) all.push_back(kv.first);
  for (auto &kv : recv)
    if (!send.count(kv.first)) all.push_back(kv.first);
  std::sort(all.begin(), all.end());
  long long so = 0, ro = 0;
  for (int peer : all) {
    long long cnt[2] = {0, 0};
    for (int side = 0; side < 2; ++side) {
      std::vector<Seg> &lst = side == 0 ? send[peer] : recv[peer];
      std::stable_sort(lst.begin(), lst.end(), [](const Seg &a, const Seg &b) { return a.gid != b.gid ? a.gid < b.gid : a.code < b.code; });
      std::vector<int> &rows = side == 0 ? send_plan : recv_plan;
      const long long base = side == 0 ? so : ro;
      for (const Seg &g : lst) {
        const long long nz = (long long)(g.r[0][1] - g.r[0][0] + 1) * (g.r[1][1] - g.r[1][0] + 1) * (g.r[2][1] - g.r[2][0] + 1);
        const unsigned long long off = (unsigned long long)(base + cnt[side]);
        const int row[10] = {g.block, g.r[0][0], g.r[0][1], g.r[1][0], g.r[1][1], g.r[2][0], g.r[2][1], (int)(unsigned)(off & 0xFFFFFFFFull),
                             (int)(unsigned)(off >> 32), 0};
        rows.insert(rows.end(), row, row + 10);
        cnt[side] += nz * nvar;
      }
    }
    peers.push_back(peer);
    send_off.push_back(so), send_cnt.push_back(cnt[0]);
    recv_off.push_back(ro), recv_cnt.push_back(cnt[1]);
    so += cnt[0];
    ro += cnt[1];
  }
}

// NCCL communicator owned by the C library. The 128-byte unique id travels through a rendezvous file
// (<phoebus_b200> rendezvous or PB_RENDEZVOUS; an MPI build would MPI_Bcast it): rank 0 writes, the others poll.
void Mesh::InitComm(ParameterInput *pin) {
  const char *env = std::getenv("PB_RENDEZVOUS");
  const std::string path = pin->GetOrAddString("phoebus_b200", "rendezvous", env ? env : "");
  if (path.empty()) Fail("world > 1 needs <phoebus_b200> rendezvous = <file> (or PB_RENDEZVOUS)");
  char id[128];
  if (rank == 0) {
    CheckPb(pb_comm_unique_id(id), "pb_comm_unique_id");
    const std::string tmp = path + ".tmp";
    std::ofstream f(tmp, std::ios::binary);
    f.write(id, 128);
    f.close();
    if (std::rename(tmp.c_str(), path.c_str()) != 0) Fail("cannot write the rendezvous file " + path);
  } else {
    bool ok = false;
    for (int tries = 0; tries < 6000 && !ok; ++tries) {
      std::ifstream f(path, std::ios::binary);
      if (f && f.read(id, 128) && f.gcount() == 128) ok = true;
      else std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    if (!ok) Fail("timed out waiting for the rendezvous file " + path);
  }
  CheckPb(pb_comm_init(&comm, rank, world, id), "pb_comm_init");
  // plans and buffers on the device
  auto dev = [&](size_t bytes) {
    void *p = nullptr;
    CheckCuda(cudaMalloc(&p, bytes ? bytes : 8), "cudaMalloc");
    owned_.push_back(p);
    return p;
  };
  int *sp = (int *)dev(send_plan.size() * sizeof(int)), *rp = (int *)dev(recv_plan.size() * sizeof(int));
  if (!send_plan.empty()) CheckCuda(cudaMemcpy(sp, send_plan.data(), send_plan.size() * sizeof(int), cudaMemcpyHostToDevice), "send plan H2D");
  if (!recv_plan.empty()) CheckCuda(cudaMemcpy(rp, recv_plan.data(), recv_plan.size() * sizeof(int), cudaMemcpyHostToDevice), "recv plan H2D");
  long long stot = 0, rtot = 0;
  for (size_t q = 0; q < peers.size(); ++q) stot += send_cnt[q], rtot += recv_cnt[q];
  plan.send_plan = sp;
  plan.nsend = (int)(send_plan.size() / 10);
  plan.recv_plan = rp;
  plan.nrecv = (int)(recv_plan.size() / 10);
  plan.send_buf = (double *)dev((size_t)stot * 8);
  plan.recv_buf = (double *)dev((size_t)rtot * 8);
  plan.npeers = (int)peers.size();
  plan.peer = peers.data();
  plan.send_off = send_off.data();
  plan.send_cnt = send_cnt.data();
  plan.recv_off = recv_off.data();
  plan.recv_cnt = recv_cnt.data();
}

void Mesh::Allocate() {
  auto dev = [&](size_t bytes) {
    void *p = nullptr;
    CheckCuda(cudaMalloc(&p, bytes ? bytes : 8), "cudaMalloc");
    CheckCuda(cudaMemset(p, 0, bytes ? bytes : 8), "cudaMemset"); // Kokkos views start at zero
    owned_.push_back(p);
    return p;
  };
  auto field = [&](const std::string &name, int comps) {
    const long long w = (long long)nblocks * comps * cells;
    field_[name] = (double *)dev((size_t)w * 8);
    words_[name] = w;
  };
  field("prim", nprim);
  field("temp", 1);
  field("c2p_mu", 1);
  field("fail", 1);
  field("csig", ndim);
  field("fsig", ndim);
  field("flux", ndim * ncons);
  field("emf", 3);
  field("gsrc", 4);
  field("divb", 1);
  pb_block_t *blocks_dev = (pb_block_t *)dev(sizeof(pb_block_t) * nblocks);
  CheckCuda(cudaMemcpy(blocks_dev, blocks_host.data(), sizeof(pb_block_t) * nblocks, cudaMemcpyHostToDevice), "blocks H2D");
  double *geom_dev = (double *)dev(geom_host_.size() * 8);
  CheckCuda(cudaMemcpy(geom_dev, geom_host_.data(), geom_host_.size() * 8, cudaMemcpyHostToDevice), "geom H2D");
  double *floor_dev = (double *)dev(floor_host_.size() * 8);
  CheckCuda(cudaMemcpy(floor_dev, floor_host_.data(), floor_host_.size() * 8, cudaMemcpyHostToDevice), "floor H2D");
  std::vector<int> bcb;
  for (int n = 0; n < nblocks; ++n) {
    bool any = false;
    for (int f = 0; f < 2 * ndim; ++f) any = any || blocks_host[n].bc[f] != PB_BC_BLOCK;
    if (any) bcb.push_back(n);
  }
  std::vector<int> inner;
  for (int n = 0, q = 0; n < nblocks; ++n) {
    if (q < (int)bcb.size() && bcb[q] == n) ++q;
    else inner.push_back(n);
  }
  int *inner_dev = (int *)dev(sizeof(int) * std::max<size_t>(1, inner.size()));
  if (!inner.empty()) CheckCuda(cudaMemcpy(inner_dev, inner.data(), sizeof(int) * inner.size(), cudaMemcpyHostToDevice), "inner list H2D");
  int *bcb_dev = (int *)dev(sizeof(int) * std::max<size_t>(1, bcb.size()));
  if (!bcb.empty()) CheckCuda(cudaMemcpy(bcb_dev, bcb.data(), sizeof(int) * bcb.size(), cudaMemcpyHostToDevice), "bc list H2D");
  counters = (unsigned long long *)dev(3 * sizeof(unsigned long long));
  dt_dev = (double *)dev(sizeof(double));
  pack_.nblocks = nblocks;
  pack_.blocks = blocks_dev;
  pack_.geom = geom_dev;
  pack_.floor_tab = floor_dev;
  pack_.prim = field_["prim"];
  pack_.temp = field_["temp"];
  pack_.c2p_mu = field_["c2p_mu"];
  pack_.fail = field_["fail"];
  pack_.csig = field_["csig"];
  pack_.fsig = field_["fsig"];
  pack_.flux = field_["flux"];
  pack_.emf = field_["emf"];
  pack_.zflags = (unsigned char *)dev((size_t)nblocks * cells);
  pack_.bc_blocks = bcb_dev;
  pack_.n_bc_blocks = (int)bcb.size();
  pack_.inner_blocks = inner_dev;
  pack_.n_inner_blocks = (int)inner.size();
  pack_.gsrc = field_["gsrc"];
  pack_.recon_lr = (double *)dev((size_t)2 * nblocks * nprim * cells * 8); // ql / qr of one direction
  // ghost plan: ghost zones whose con2prim result is copied from the same-GPU neighbour (include/phoebus_b200.h)
  pack_.ghost_mask = nullptr;
  pack_.c2p_zones = nullptr;
  pack_.n_c2p_zones = 0;
  pack_.twin_zones = nullptr;
  pack_.n_twin = 0;
  pack_.twin_buf = nullptr;
  pack_.twin_flags = nullptr;
  if (cfg.bc_prim) {
    std::vector<unsigned char> mask((size_t)nblocks * cells);
    std::vector<unsigned> zones((size_t)nblocks * cells);
    const long long S = pb_twin_entries(&cfg, nblocks);
    std::vector<unsigned> twins((size_t)std::max<long long>(1, S));
    long long nz = 0, nt = 0;
    CheckPb(pb_ghost_plan_host(&cfg, blocks_host.data(), nblocks, geom_host_.data(), floor_host_.data(), mask.data(), zones.data(), &nz,
                               twins.data(), &nt),
            "pb_ghost_plan_host");
    bool any = false;
    for (unsigned char m : mask) any = any || m;
    if (any) {
      unsigned char *mask_dev = (unsigned char *)dev(mask.size());
      CheckCuda(cudaMemcpy(mask_dev, mask.data(), mask.size(), cudaMemcpyHostToDevice), "ghost mask H2D");
      unsigned *zones_dev = (unsigned *)dev(sizeof(unsigned) * (size_t)std::max<long long>(1, nz));
      if (nz) CheckCuda(cudaMemcpy(zones_dev, zones.data(), sizeof(unsigned) * (size_t)nz, cudaMemcpyHostToDevice), "c2p zone list H2D");
      pack_.ghost_mask = mask_dev;
      pack_.c2p_zones = zones_dev;
      pack_.n_c2p_zones = nz;
      if (nt) {
        unsigned *tw_dev = (unsigned *)dev(sizeof(unsigned) * (size_t)nt);
        CheckCuda(cudaMemcpy(tw_dev, twins.data(), sizeof(unsigned) * (size_t)nt, cudaMemcpyHostToDevice), "twin list H2D");
        pack_.twin_zones = tw_dev;
        pack_.n_twin = nt;
        pack_.twin_buf = (double *)dev(sizeof(double) * (size_t)S * (size_t)(ncons + nprim + 3 + ndim));
        pack_.twin_flags = (unsigned char *)dev((size_t)S);
      }
    }
  }
  AddContainer("base");
  AddContainer("1");
  AddContainer("dUdt");
  // "geometric source terms" aliases the gsrc slab (ref: phoebus_driver.cpp:289-299)
  cons_["geometric source terms"] = field_["gsrc"];
  words_["geometric source terms"] = words_["gsrc"];
}

Mesh::~Mesh() {
  if (comm) pb_comm_destroy(comm);
  for (void *p : owned_) cudaFree(p);
}

void Mesh::AddContainer(const std::string &stage) {
  if (cons_.count(stage)) return;
  void *p = nullptr;
  const long long w = (long long)nblocks * ncons * cells;
  CheckCuda(cudaMalloc(&p, (size_t)w * 8), "cudaMalloc");
  CheckCuda(cudaMemset(p, 0, (size_t)w * 8), "cudaMemset");
  owned_.push_back(p);
  cons_[stage] = (double *)p;
  words_[stage] = w;
}

std::shared_ptr<MeshData<Real>> &Mesh::GetMeshData(const std::string &stage) {
  auto it = md_.find(stage);
  if (it != md_.end()) return it->second;
  if (!cons_.count(stage)) AddContainer(stage);
  auto md = std::make_shared<MeshData<Real>>();
  md->pmesh = this;
  md->stage = stage;
  md->pack = pack_;
  md->cons = cons_[stage];
  md->first = 0;
  md->nblocks = nblocks;
  return md_[stage] = md;
}

std::shared_ptr<MeshBlockData<Real>> Mesh::GetBlockData(int b, const std::string &stage) {
  if (b < 0 || b >= nblocks) Fail("GetBlockData: no such block");
  if (!cons_.count(stage)) AddContainer(stage);
  auto rc = std::make_shared<MeshBlockData<Real>>();
  rc->pmesh = this;
  rc->stage = stage;
  rc->first = b;
  rc->nblocks = 1;
  pb_pack_t p = pack_;
  const long long c = cells, o = (long long)b * c;
  p.nblocks = 1;
  p.blocks = pack_.blocks + b;
  p.prim += o * nprim;
  p.temp += o;
  p.c2p_mu += o;
  p.fail += o;
  p.csig += o * ndim;
  p.fsig += o * ndim;
  p.flux += o * ndim * ncons;
  p.emf += o * 3;
  p.zflags += o;
  p.gsrc += o * 4;
  p.recon_lr = nullptr; // per-block calls use the fused flux kernel
  p.bc_blocks = nullptr; // "every block of this pack"
  p.n_bc_blocks = 0;
  p.inner_blocks = nullptr;
  p.n_inner_blocks = 0;
  p.ghost_mask = nullptr; // the plan describes the whole pack
  p.c2p_zones = nullptr;
  p.twin_zones = nullptr;
  p.n_twin = 0;
  rc->pack = p;
  rc->cons = cons_[stage] + o * (stage == "geometric source terms" ? 4 : ncons);
  return rc;
}

double *Mesh::Slab(const std::string &name) const {
  auto f = field_.find(name);
  if (f != field_.end()) return f->second;
  auto c = cons_.find(name);
  if (c != cons_.end()) return c->second;
  Fail("Mesh: no field or container named '" + name + "'");
}
long long Mesh::FieldWords(const std::string &name) const {
  auto w = words_.find(name);
  if (w == words_.end()) Fail("Mesh: no field or container named '" + name + "'");
  return w->second;
}
void Mesh::Upload(const std::string &name, const double *host) {
  CheckCuda(cudaMemcpy(Slab(name), host, (size_t)FieldWords(name) * 8, cudaMemcpyHostToDevice), "Upload");
  flags_valid = false;
}
void Mesh::Download(const std::string &name, double *host) const {
  CheckCuda(cudaMemcpy(host, Slab(name), (size_t)FieldWords(name) * 8, cudaMemcpyDeviceToHost), "Download");
}
void Mesh::Synchronize() const { CheckCuda(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

} // namespace parthenon

// =================================================================================================
// tasks
// =================================================================================================
namespace {
using parthenon::PackView;
void *S(const PackView *rc) { return rc->GetMeshPointer()->stream; }
const pb_cfg_t *Cfg(const PackView *rc) { return &rc->GetMeshPointer()->cfg; }
// every-zone ("exhaustive") variant of the pack: the stand-alone tasks do what the reference's do
pb_pack_t Plain(const PackView *rc) {
  pb_pack_t p = rc->pack;
  p.zflags = nullptr;
  return p;
}

TaskStatus FluxesImpl(PackView *rc) {
  if (!FluidActive(rc) || Pkg(rc, "fluid").Get<bool>("zero_fluxes")) return TaskStatus::complete;
  CheckPb(pb_calculate_fluxes(&rc->pack, Cfg(rc), S(rc)), "pb_calculate_fluxes");
  return TaskStatus::complete;
}
TaskStatus FluxCTImpl(PackView *rc) {
  if (!FluidActive(rc) || !Pkg(rc, "fluid").Get<bool>("mhd") || Pkg(rc, "fluid").Get<bool>("zero_fluxes"))
    return TaskStatus::complete;
  CheckPb(pb_flux_ct(&rc->pack, Cfg(rc), S(rc)), "pb_flux_ct");
  return TaskStatus::complete;
}
TaskStatus FixFluxesImpl(PackView *rc) {
  if (!Pkg(rc, "fixup").Get<bool>("enable_flux_fixup")) return TaskStatus::complete;
  CheckPb(pb_fix_fluxes(&rc->pack, Cfg(rc), S(rc)), "pb_fix_fluxes");
  return TaskStatus::complete;
}
TaskStatus P2CImpl(PackView *rc, const IndexRange &ib, const IndexRange &jb, const IndexRange &kb) {
  int i[2], j[2], k[2];
  RangeArg(ib, i), RangeArg(jb, j), RangeArg(kb, k);
  CheckPb(pb_prim_to_cons(&rc->pack, Cfg(rc), rc->cons, i, j, k, S(rc)), "pb_prim_to_cons");
  return TaskStatus::complete;
}
void EntireRanges(const PackView *rc, IndexRange &ib, IndexRange &jb, IndexRange &kb) {
  const Mesh *pm = rc->GetMeshPointer();
  ib = {0, pm->ncell[0] - 1};
  jb = {0, pm->ncell[1] - 1};
  kb = {0, pm->ncell[2] - 1};
}
Real DtImpl(PackView *rc) {
  Mesh *pm = rc->GetMeshPointer();
  const double big = DBL_MAX;
  CheckCuda(cudaMemcpyAsync(pm->dt_dev, &big, sizeof big, cudaMemcpyHostToDevice, (cudaStream_t)pm->stream), "dt init");
  CheckPb(pb_estimate_dt(&rc->pack, Cfg(rc), pm->has_face_speeds ? 1 : 0, pm->dt_dev, S(rc)), "pb_estimate_dt");
  if (pm->comm && rc->nblocks == pm->nblocks) // the mesh-wide estimate: MPI_Allreduce(MIN) in the reference (parthenon driver.cpp:237)
    CheckPb(pb_comm_allreduce_min(pm->comm, pm->dt_dev, 1, pm->stream), "pb_comm_allreduce_min");
  double dt = 0;
  CheckCuda(cudaMemcpyAsync(&dt, pm->dt_dev, sizeof dt, cudaMemcpyDeviceToHost, (cudaStream_t)pm->stream), "dt D2H");
  pm->Synchronize();
  return Pkg(rc, "fluid").Get<Real>("cfl") * dt;
}
} // namespace

namespace fluid {
TaskStatus CalculateFluxes(MeshBlockData<Real> *rc) { return FluxesImpl(rc); }
TaskStatus CalculateFluxes(MeshData<Real> *rc) { return FluxesImpl(rc); }
TaskStatus FluxCT(MeshBlockData<Real> *rc) { return FluxCTImpl(rc); }
TaskStatus FluxCT(MeshData<Real> *rc) { return FluxCTImpl(rc); }

TaskStatus CalculateFluidSourceTerms(MeshData<Real> *rc, MeshData<Real> *rc_src) {
  if (!FluidActive(rc) || Pkg(rc, "fluid").Get<bool>("zero_sources")) return TaskStatus::complete;
  CheckPb(pb_source_terms(&rc->pack, Cfg(rc), rc_src->cons, S(rc)), "pb_source_terms");
  return TaskStatus::complete;
}

TaskStatus PrimitiveToConservedRegion(MeshBlockData<Real> *rc, const IndexRange &ib, const IndexRange &jb,
                                      const IndexRange &kb) {
  return P2CImpl(rc, ib, jb, kb);
}
TaskStatus PrimitiveToConserved(MeshBlockData<Real> *rc) {
  IndexRange ib, jb, kb;
  EntireRanges(rc, ib, jb, kb);
  return P2CImpl(rc, ib, jb, kb);
}
TaskStatus PrimitiveToConserved(MeshData<Real> *rc) {
  IndexRange ib, jb, kb;
  EntireRanges(rc, ib, jb, kb);
  return P2CImpl(rc, ib, jb, kb);
}

template <typename T>
TaskStatus ConservedToPrimitiveRobust(T *rc, const IndexRange &ib, const IndexRange &jb, const IndexRange &kb) {
  int i[2], j[2], k[2];
  RangeArg(ib, i), RangeArg(jb, j), RangeArg(kb, k);
  const pb_pack_t p = Plain(rc);
  CheckPb(pb_cons_to_prim(&p, Cfg(rc), rc->cons, i, j, k, S(rc)), "pb_cons_to_prim");
  rc->GetMeshPointer()->InvalidateZoneFlags();
  return TaskStatus::complete;
}
template <typename T>
TaskStatus ConservedToPrimitiveRegion(T *rc, const IndexRange &ib, const IndexRange &jb, const IndexRange &kb) {
  return ConservedToPrimitiveRobust(rc, ib, jb, kb); // Params "c2p_func" (ref: fluid.cpp:223-249)
}
template <typename T>
TaskStatus ConservedToPrimitive(T *rc) {
  IndexRange ib, jb, kb;
  EntireRanges(rc, ib, jb, kb);
  return ConservedToPrimitiveRegion(rc, ib, jb, kb);
}
template TaskStatus ConservedToPrimitive<MeshData<Real>>(MeshData<Real> *);
template TaskStatus ConservedToPrimitive<MeshBlockData<Real>>(MeshBlockData<Real> *);
template TaskStatus ConservedToPrimitiveRegion<MeshData<Real>>(MeshData<Real> *, const IndexRange &, const IndexRange &, const IndexRange &);
template TaskStatus ConservedToPrimitiveRegion<MeshBlockData<Real>>(MeshBlockData<Real> *, const IndexRange &, const IndexRange &, const IndexRange &);
template TaskStatus ConservedToPrimitiveRobust<MeshData<Real>>(MeshData<Real> *, const IndexRange &, const IndexRange &, const IndexRange &);
template TaskStatus ConservedToPrimitiveRobust<MeshBlockData<Real>>(MeshBlockData<Real> *, const IndexRange &, const IndexRange &, const IndexRange &);

static TaskStatus DivBImpl(parthenon::PackView *rc) {
  if (!FluidActive(rc) || !Pkg(rc, "fluid").Get<bool>("mhd")) return TaskStatus::complete;
  Mesh *pm = rc->GetMeshPointer();
  CheckPb(pb_calculate_divb(&rc->pack, Cfg(rc), rc->cons, pm->DivB() + (long long)rc->first * pm->cells, S(rc)), "pb_calculate_divb");
  return TaskStatus::complete;
}
TaskStatus CalculateDivB(MeshBlockData<Real> *rc) { return DivBImpl(rc); }
TaskStatus CalculateDivB(MeshData<Real> *rc) { return DivBImpl(rc); }
Real EstimateTimestepBlock(MeshBlockData<Real> *rc) { return DtImpl(rc); }
Real EstimateTimestepMesh(MeshData<Real> *rc) { return DtImpl(rc); }
} // namespace fluid

namespace fixup {
TaskStatus FixFluxes(MeshBlockData<Real> *rc) { return FixFluxesImpl(rc); }
TaskStatus FixFluxes(MeshData<Real> *rc) { return FixFluxesImpl(rc); }

template <typename T>
TaskStatus ApplyFloors(T *rc) {
  if (!FixupFlag(rc, "enable_floors")) return TaskStatus::complete;
  const pb_pack_t p = Plain(rc);
  CheckPb(pb_apply_floors(&p, Cfg(rc), rc->cons, rc->GetMeshPointer()->counters, S(rc)), "pb_apply_floors");
  rc->GetMeshPointer()->InvalidateZoneFlags();
  return TaskStatus::complete;
}
template <typename T>
TaskStatus ConservedToPrimitiveFixup(T *rc) {
  if (!FixupFlag(rc, "enable_c2p_fixup")) return TaskStatus::complete;
  const pb_pack_t p = Plain(rc);
  CheckPb(pb_c2p_fixup(&p, Cfg(rc), rc->cons, S(rc)), "pb_c2p_fixup");
  rc->GetMeshPointer()->InvalidateZoneFlags();
  return TaskStatus::complete;
}
template TaskStatus ApplyFloors<MeshData<Real>>(MeshData<Real> *);
template TaskStatus ApplyFloors<MeshBlockData<Real>>(MeshBlockData<Real> *);
template TaskStatus ConservedToPrimitiveFixup<MeshData<Real>>(MeshData<Real> *);
template TaskStatus ConservedToPrimitiveFixup<MeshBlockData<Real>>(MeshBlockData<Real> *);
} // namespace fixup

namespace Boundaries {
static TaskStatus BcImpl(PackView *rc) {
  CheckPb(pb_apply_bcs(&rc->pack, Cfg(rc), rc->cons, S(rc)), "pb_apply_bcs");
  return TaskStatus::complete;
}
TaskStatus ConvertBoundaryConditions(std::shared_ptr<MeshBlockData<Real>> &rc) { return BcImpl(rc.get()); }
TaskStatus ConvertBoundaryConditions(std::shared_ptr<MeshData<Real>> &rc) { return BcImpl(rc.get()); }
} // namespace Boundaries

namespace parthenon {
template <class T>
TaskStatus ApplyBoundaryConditions(std::shared_ptr<T> &) {
  return TaskStatus::complete; // the ghost fill is fused into Boundaries::ConvertBoundaryConditions
}
template TaskStatus ApplyBoundaryConditions<MeshBlockData<Real>>(std::shared_ptr<MeshBlockData<Real>> &);
template TaskStatus ApplyBoundaryConditions<MeshData<Real>>(std::shared_ptr<MeshData<Real>> &);

// ref: boundary_communication.cpp:49-158 (SendBoundBufs), :160-231 (ReceiveBoundBufs), :233-346 (SetBounds). Neighbours on
// other GPUs: pack + grouped ncclSend / ncclRecv inside pb_halo_exchange_begin, join + unpack in SetBounds.
TaskStatus SendBoundBufs(std::shared_ptr<MeshData<Real>> &md) {
  Mesh *pm = md->pmesh;
  if (pm->comm) CheckPb(pb_halo_exchange_begin(&md->pack, &pm->cfg, md->cons, &pm->plan, pm->comm, pm->stream), "pb_halo_exchange_begin");
  CheckPb(pb_halo_exchange_local(&md->pack, &pm->cfg, md->cons, pm->stream), "pb_halo_exchange_local");
  return TaskStatus::complete;
}
TaskStatus ReceiveBoundBufs(std::shared_ptr<MeshData<Real>> &) { return TaskStatus::complete; }
TaskStatus SetBounds(std::shared_ptr<MeshData<Real>> &md) {
  Mesh *pm = md->pmesh;
  if (pm->comm) CheckPb(pb_halo_exchange_end(&md->pack, &pm->cfg, md->cons, &pm->plan, pm->comm, pm->stream), "pb_halo_exchange_end");
  return TaskStatus::complete;
}

namespace Update {
template <class T>
TaskStatus FluxDivergence(T *in, T *dudt_obj) {
  CheckPb(pb_flux_divergence(&in->pack, Cfg(in), dudt_obj->cons, S(in)), "pb_flux_divergence");
  return TaskStatus::complete;
}
// z = w1*x + w2*y over the listed variables of every block (entire arrays)
static void WeightedSum(PackView *x, PackView *y, Real w1, Real w2, PackView *z, const std::vector<std::string> *vars) {
  const Mesh *pm = x->GetMeshPointer();
  const long long c = pm->cells;
  if (!vars) {
    CheckPb(pb_weighted_sum(x->cons, y->cons, w1, w2, z->cons, (long long)x->nblocks * pm->ncons * c, S(x)), "pb_weighted_sum");
    return;
  }
  // variable subset: the momentum + energy slots (1..4) are what "geometric source terms" holds
  for (const std::string &v : *vars)
    if (v != "c.momentum" && v != "c.energy") Fail("SumData: variable '" + v + "' is not in the source container");
  const bool ysrc = y->stage == "geometric source terms", xsrc = x->stage == "geometric source terms";
  for (int b = 0; b < x->nblocks; ++b) {
    const double *xp = x->cons + (xsrc ? (long long)b * 4 * c : ((long long)b * pm->ncons + 1) * c);
    const double *yp = y->cons + (ysrc ? (long long)b * 4 * c : ((long long)b * pm->ncons + 1) * c);
    double *zp = z->cons + ((long long)b * pm->ncons + 1) * c;
    CheckPb(pb_weighted_sum(xp, yp, w1, w2, zp, 4 * c, S(x)), "pb_weighted_sum");
  }
}
template <class F, class T>
TaskStatus SumData(const F &flags, T *in1, T *in2, T *out) {
  WeightedSum(in1, in2, 1.0, 1.0, out, &flags);
  return TaskStatus::complete;
}
template <class T>
TaskStatus AverageIndependentData(T *c1, T *c2, const Real wgt1) {
  WeightedSum(c1, c2, wgt1, 1.0 - wgt1, c1, nullptr);
  return TaskStatus::complete;
}
template <class T>
TaskStatus UpdateIndependentData(T *in, T *dudt, const Real dt, T *out) {
  WeightedSum(in, dudt, 1.0, dt, out, nullptr);
  return TaskStatus::complete;
}
template <class T>
TaskStatus FillDerived(T *rc) {
  return fluid::ConservedToPrimitive(rc);
}
template TaskStatus FluxDivergence<MeshData<Real>>(MeshData<Real> *, MeshData<Real> *);
template TaskStatus FluxDivergence<MeshBlockData<Real>>(MeshBlockData<Real> *, MeshBlockData<Real> *);
template TaskStatus SumData<std::vector<std::string>, MeshData<Real>>(const std::vector<std::string> &, MeshData<Real> *,
                                                                     MeshData<Real> *, MeshData<Real> *);
template TaskStatus AverageIndependentData<MeshData<Real>>(MeshData<Real> *, MeshData<Real> *, const Real);
template TaskStatus UpdateIndependentData<MeshData<Real>>(MeshData<Real> *, MeshData<Real> *, const Real, MeshData<Real> *);
template TaskStatus FillDerived<MeshData<Real>>(MeshData<Real> *);
template TaskStatus FillDerived<MeshBlockData<Real>>(MeshBlockData<Real> *);
} // namespace Update
} // namespace parthenon

// =================================================================================================
// driver
// =================================================================================================
namespace phoebus {

PhoebusDriver::PhoebusDriver(ParameterInput *pin, Mesh *pm, bool fused) : pmesh(pm), fused_(fused) {
  const std::string integ = pin->GetOrAddString("parthenon/time", "integrator", "rk2");
  if (integ == "rk1") beta = {1.0};
  else if (integ == "rk2") beta = {1.0, 0.5};
  else Fail("integrator " + integ + " is not on the accelerated path");
  // stage containers (ref: external/parthenon/src/time_integration/staged_integrator.cpp:23-29)
  stage_name.push_back("base");
  for (size_t i = 1; i < beta.size(); ++i) stage_name.push_back("1");
  stage_name.push_back("base");
  dt_init_fact_ = pin->GetOrAddReal("parthenon/time", "dt_init_fact", 1.0);
  tlim_ = pin->GetOrAddReal("parthenon/time", "tlim", DBL_MAX);
}

TaskStatus PhoebusDriver::RungeKuttaStage(int stage) {
  using namespace parthenon;
  using namespace parthenon::Update;
  Mesh *pm = pmesh;
  const Real b = beta[stage - 1];
  auto &base = pm->GetMeshData("base");
  auto &sc0 = pm->GetMeshData(stage_name[stage - 1]);
  auto &sc1 = pm->GetMeshData(stage_name[stage]);
  if (fused_) {
    if (!pm->flags_valid) {
      CheckPb(pb_scan_zone_flags(&base->pack, &pm->cfg, pm->stream), "pb_scan_zone_flags");
      pm->flags_valid = true;
    }
    if (!pm->comm) {
      CheckPb(pb_rk_stage(&sc0->pack, &pm->cfg, sc0->cons, base->cons, sc1->cons, b, dt, pm->counters, pm->stream), "pb_rk_stage");
      return TaskStatus::complete;
    }
    // N GPUs: the transfer is hidden under the con2prim of the MeshBlocks away from the mesh boundary
    CheckPb(pb_rk_stage_pre(&sc0->pack, &pm->cfg, sc0->cons, base->cons, sc1->cons, b, dt, pm->counters, pm->stream), "pb_rk_stage_pre");
    CheckPb(pb_halo_exchange_begin(&sc1->pack, &pm->cfg, sc1->cons, &pm->plan, pm->comm, pm->stream), "pb_halo_exchange_begin");
    CheckPb(pb_halo_exchange_local_ex(&sc1->pack, &pm->cfg, sc1->cons, pb_ghost_plan_active(&sc1->pack, &pm->cfg), pm->stream),
            "pb_halo_exchange_local_ex");
    const int early = pb_early_c2p_pays(&sc1->pack);
    if (early) CheckPb(pb_rk_stage_post_early(&sc1->pack, &pm->cfg, sc1->cons, pm->stream), "pb_rk_stage_post_early");
    SetBounds(sc1);
    CheckPb(pb_rk_stage_post2(&sc1->pack, &pm->cfg, sc1->cons, pm->counters, early, pm->stream), "pb_rk_stage_post2");
    return TaskStatus::complete;
  }
  // ---- the reference's task list, one task at a time (ref: phoebus_driver.cpp:316-365, 556-597, 655-676)
  auto &dudt = pm->GetMeshData("dUdt");
  auto &gsrc = pm->GetMeshData("geometric source terms");
  fluid::CalculateFluidSourceTerms(sc0.get(), gsrc.get());
  for (int n = 0; n < pm->nblocks; ++n) { // async region over blocks
    auto rc = pm->GetBlockData(n, stage_name[stage - 1]);
    fluid::CalculateFluxes(rc.get());
    fluid::FluxCT(rc.get());
    fixup::FixFluxes(rc.get());
  }
  FluxDivergence(sc0.get(), dudt.get());
  SumData(std::vector<std::string>({"c.momentum", "c.energy"}), dudt.get(), gsrc.get(), dudt.get());
  AverageIndependentData(sc0.get(), base.get(), b);
  UpdateIndependentData(sc0.get(), dudt.get(), b * dt, sc1.get());
  // FillDerived<MeshData> is a no-op in the reference: fixup + floors see the previous prims
  fixup::ConservedToPrimitiveFixup(sc1.get());
  fixup::ApplyFloors(sc1.get());
  SendBoundBufs(sc1);
  ReceiveBoundBufs(sc1);
  SetBounds(sc1);
  for (int n = 0; n < pm->nblocks; ++n) {
    auto rc = pm->GetBlockData(n, stage_name[stage]);
    ApplyBoundaryConditions(rc);
    Boundaries::ConvertBoundaryConditions(rc);
    FillDerived(rc.get());
    fixup::ConservedToPrimitiveFixup(rc.get());
    fixup::ApplyFloors(rc.get());
    if (stage == (int)beta.size()) fluid::CalculateDivB(rc.get()); // ref: phoebus_driver.cpp:719-726
  }
  return TaskStatus::complete;
}

void PhoebusDriver::PostInitializationCommunication() {
  using namespace parthenon;
  Mesh *pm = pmesh;
  auto &base = pm->GetMeshData("base");
  fluid::PrimitiveToConserved(base.get());
  SendBoundBufs(base);
  Boundaries::ConvertBoundaryConditions(base);
  fluid::ConservedToPrimitive(base.get());
  fixup::ConservedToPrimitiveFixup(base.get());
  fixup::ApplyFloors(base.get());
}

Real PhoebusDriver::EstimateTimestep() {
  Real v = fluid::EstimateTimestepMesh(pmesh->GetMeshData("base").get());
  pmesh->has_face_speeds = true;
  return v;
}

TaskStatus PhoebusDriver::Step() {
  Real step_dt = dt;
  if (first_step_) step_dt *= dt_init_fact_; // ref: phoebus_driver.cpp:121-132
  first_step_ = false;
  dt = step_dt;
  for (int s = 1; s <= (int)beta.size(); ++s) RungeKuttaStage(s);
  ncycle += 1;
  time += step_dt;
  dt = std::min(2.0 * step_dt, EstimateTimestep()); // ref: parthenon driver.cpp:216-240
  if (time < tlim_ && (tlim_ - time) < dt) dt = tlim_ - time; // ref: parthenon driver.cpp:236-239
  return TaskStatus::complete;
}

} // namespace phoebus
