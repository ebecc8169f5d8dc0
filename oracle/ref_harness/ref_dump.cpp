// TEST INFRASTRUCTURE ONLY — not part of the product.
//
// ref_dump: drives the UNMODIFIED reference (lanl/phoebus, linked from the reference's
// own static libraries libphoebuslib.a / libparthenon.a / libkokkos*.a) exactly as
// src/main.cpp:33-91 does, but replaces PhoebusDriver::Step (src/phoebus_driver.cpp:121-143)
// by a copy of its control flow that dumps every fluid/geometry field of every MeshBlock
// after each Runge-Kutta stage (RungeKuttaStage, src/phoebus_driver.cpp:204-738) as raw
// little-endian float64 plus a JSON manifest (HDF5 is not available in this image).
//
// The reference sources are compiled where they lie; nothing from the reference is copied
// into this repository. Build recipe: oracle/build_ref.sh. Outputs go to oracle/_ref/.
//
// usage: ref_dump -i deck.pin [block/key=value ...]   with env REF_DUMP_DIR=<dir>
//        REF_DUMP_QLQR=1 additionally dumps ql/qr.
//        REF_DUMP_EVERY=N dumps only the stages of cycles 0, N, 2N, ... and of the last cycle (long runs).
//        REF_DUMP_PERTURB=amp multiplies p.density, p.energy and p.velocity[2] of every zone (ghosts included) by
//        1 + amp * (h - 0.5) right after the problem generator and re-runs fluid::PrimitiveToConserved; h is an
//        integer hash of the zone's GLOBAL index (periodic in x3) in [0, 1), exactly representable, so that a test
//        can rebuild the same 3-D initial state from an axisymmetric pgen (phoebus_b200/pgen.py: hash_perturbation).
//        REF_DUMP_FIELDS=hash writes, instead of the raw fields of the stage dumps, one FNV-1a 64-bit hash per array
//        (large fixtures; the "init" dump keeps one k-plane per array for diagnostics).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <parthenon/driver.hpp>
#include <parthenon_manager.hpp>

#define private public // test harness needs integrator / dt_init_fact of PhoebusDriver
#include "phoebus_driver.hpp"
#undef private

#include "fluid/fluid.hpp"
#include "geometry/geometry.hpp"
#include "pgen/pgen.hpp"
#include "phoebus_boundaries/phoebus_boundaries.hpp"

using namespace parthenon::driver::prelude;

namespace {

// position-weighted sum of the canonicalised bit patterns (−0 -> +0, any NaN -> one pattern), modulo 2^64; the same
// three lines of numpy reproduce it (oracle/refdump.py: array_hash)
inline unsigned long long ArrayHash(const Real *p, size_t n) {
  unsigned long long h = 0;
  for (size_t i = 0; i < n; ++i) {
    const double y = p[i] + 0.0;
    unsigned long long b;
    std::memcpy(&b, &y, 8);
    if (y != y) b = 0x7ff8000000000000ULL;
    const unsigned long long w = ((unsigned long long)i * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL) | 1ULL;
    h += b * w;
  }
  return h;
}

struct Dumper {
  std::string dir;
  bool with_qlqr = false;
  bool hash_only = false;
  int ndump = 0;

  void AppendVar(std::ofstream &bin, std::ostringstream &js, bool &first,
                 const std::string &container, const std::string &name, int gid,
                 parthenon::Variable<Real> &v, size_t &offset) {
    if (!v.IsAllocated()) return;
    size_t n = 1;
    int dims[6];
    for (int d = 1; d <= 6; ++d) {
      dims[d - 1] = v.GetDim(d);
      n *= dims[d - 1];
    }
    auto h = v.data.GetHostMirrorAndCopy();
    if (!first) js << ",\n";
    first = false;
    if (hash_only) {
      js << "  {\"container\":\"" << container << "\",\"name\":\"" << name
         << "\",\"gid\":" << gid << ",\"dims\":[" << dims[5] << "," << dims[4] << ","
         << dims[3] << "," << dims[2] << "," << dims[1] << "," << dims[0]
         << "],\"count\":" << n << ",\"hash\":\"" << std::hex << ArrayHash(h.data(), n) << std::dec << "\"}";
      return;
    }
    bin.write(reinterpret_cast<const char *>(h.data()), n * sizeof(Real));
    js << "  {\"container\":\"" << container << "\",\"name\":\"" << name
       << "\",\"gid\":" << gid << ",\"dims\":[" << dims[5] << "," << dims[4] << ","
       << dims[3] << "," << dims[2] << "," << dims[1] << "," << dims[0]
       << "],\"offset\":" << offset << ",\"count\":" << n << "}";
    offset += n;
  }

  void Dump(Mesh *pmesh, const std::string &tag, int cycle, int stage, Real time, Real dt,
            Real beta, const std::string &sc0, const std::string &sc1, bool with_geom) {
    std::ofstream bin(dir + "/" + tag + ".bin", std::ios::binary);
    std::ostringstream js;
    js.precision(17);
    js << "{\"tag\":\"" << tag << "\",\"cycle\":" << cycle << ",\"stage\":" << stage
       << ",\"time\":" << time << ",\"dt\":" << dt << ",\"beta\":" << beta
       << ",\"sc0\":\"" << sc0 << "\",\"sc1\":\"" << sc1 << "\",\n";
    {
      // geometry parameters as resolved by the reference (fmks.cpp:38-91); absent for Minkowski
      auto &gp = pmesh->packages.Get("geometry")->AllParams();
      js << "\"geom_params\":{";
      bool fg = true;
      for (auto k : {"a", "h", "xt", "alpha", "x0", "smooth", "hexp_br", "hexp_nsq", "hexp_csq"}) {
        if (gp.hasKey(k)) {
          js << (fg ? "" : ",") << "\"" << k << "\":" << gp.Get<Real>(k);
          fg = false;
        }
      }
      if (gp.hasKey("derefine_poles"))
        js << (fg ? "" : ",") << "\"derefine_poles\":" << (gp.Get<bool>("derefine_poles") ? 1 : 0);
      js << "},\n\"mesh\":{\"x1min\":" << pmesh->mesh_size.xmin(X1DIR) << ",\"x1max\":"
         << pmesh->mesh_size.xmax(X1DIR) << ",\"x2min\":" << pmesh->mesh_size.xmin(X2DIR)
         << ",\"x2max\":" << pmesh->mesh_size.xmax(X2DIR) << ",\"x3min\":"
         << pmesh->mesh_size.xmin(X3DIR) << ",\"x3max\":" << pmesh->mesh_size.xmax(X3DIR)
         << "},\n";
    }
    js << "\"blocks\":[\n";
    bool firstb = true;
    for (auto &pmb : pmesh->block_list) {
      if (!firstb) js << ",\n";
      firstb = false;
      const auto &c = pmb->coords;
      js << "  {\"gid\":" << pmb->gid << ",\"lx\":[" << pmb->loc.lx1() << ","
         << pmb->loc.lx2() << "," << pmb->loc.lx3() << "],\"nx\":["
         << pmb->block_size.nx(X1DIR) << "," << pmb->block_size.nx(X2DIR) << ","
         << pmb->block_size.nx(X3DIR) << "],\"xmin\":[" << c.Xf<1>(0) << "," << c.Xf<2>(0)
         << "," << c.Xf<3>(0) << "],\"dx\":[" << c.Dxc<1>() << "," << c.Dxc<2>() << ","
         << c.Dxc<3>() << "],\"bflag\":[";
      for (int f = 0; f < 6; ++f)
        js << static_cast<int>(pmb->boundary_flag[f]) << (f < 5 ? "," : "");
      js << "]}";
    }
    js << "\n],\n\"arrays\":[\n";
    bool first = true;
    size_t offset = 0;
    const std::vector<std::string> onecopy = {
        "p.density",   "p.velocity", "p.energy",          "p.bfield",
        "pressure",    "temperature", "gamma1",           "c2p_mu",
        "fail",        "cell_signal_speed", "face_signal_speed", "emf", "divb"};
    const std::vector<std::string> cons = {"c.density", "c.momentum", "c.energy",
                                           "c.bfield"};
    const std::vector<std::string> geom = {
        "g.c.alpha",  "g.c.dalpha",  "g.c.bcon",   "g.c.gcov",   "g.c.gamcon",
        "g.c.detgam", "g.c.dg",      "g.f1.alpha", "g.f1.bcon",  "g.f1.gcov",
        "g.f1.gamcon", "g.f1.detgam", "g.f2.alpha", "g.f2.bcon", "g.f2.gcov",
        "g.f2.gamcon", "g.f2.detgam", "g.f3.alpha", "g.f3.bcon", "g.f3.gcov",
        "g.f3.gamcon", "g.f3.detgam"};
    for (auto &pmb : pmesh->block_list) {
      auto &base = pmb->meshblock_data.Get("base");
      for (auto &n : onecopy)
        if (base->HasVariable(n)) AppendVar(bin, js, first, "base", n, pmb->gid, base->Get(n), offset);
      if (with_qlqr)
        for (auto n : {"ql", "qr"})
          if (base->HasVariable(n)) AppendVar(bin, js, first, "base", n, pmb->gid, base->Get(n), offset);
      for (auto &n : cons) {
        std::string fn = "bnd_flux::" + n;
        if (base->HasVariable(fn)) AppendVar(bin, js, first, "base", fn, pmb->gid, base->Get(fn), offset);
      }
      if (with_geom)
        for (auto &n : geom)
          if (base->HasVariable(n)) AppendVar(bin, js, first, "base", n, pmb->gid, base->Get(n), offset);
      for (auto cn : {"base", "1", "dUdt", "geometric source terms"}) {
        std::shared_ptr<parthenon::MeshBlockData<Real>> rc;
        try {
          rc = pmb->meshblock_data.Get(cn);
        } catch (...) {
          continue;
        }
        for (auto &n : cons)
          if (rc->HasVariable(n)) AppendVar(bin, js, first, cn, n, pmb->gid, rc->Get(n), offset);
      }
    }
    js << "\n]}\n";
    std::ofstream jf(dir + "/" + tag + ".json");
    jf << js.str();
    ndump++;
  }
};

Dumper g_dumper;
int g_every = 1;

// REF_DUMP_PERTURB: see the header comment. Global zone index = (block origin - mesh origin) / dx + local index;
// x3 wraps (periodic), x1 / x2 ghosts beyond the mesh get indices < 0 or >= nx, shifted by +4096 to stay positive.
inline double HashUnit(long gi, long gj, long gk, long salt) {
  unsigned long long x = (unsigned long long)(gi + 4096) * 73856093ULL ^ (unsigned long long)(gj + 4096) * 19349663ULL ^
                         (unsigned long long)(gk + 4096) * 83492791ULL ^ (unsigned long long)salt * 2654435761ULL;
  x ^= x >> 15;
  x *= 0x2c1b3c6dULL;
  x ^= x >> 12;
  return (double)(x & 1023ULL) / 1024.0;
}
void PerturbState(Mesh *pmesh, double amp) {
  const int ng = parthenon::Globals::nghost;
  const long nx3 = pmesh->mesh_size.nx(X3DIR);
  for (auto &pmb : pmesh->block_list) {
    auto &rc = pmb->meshblock_data.Get();
    auto rho = rc->Get("p.density").data;
    auto eng = rc->Get("p.energy").data;
    auto vel = rc->Get("p.velocity").data;
    const auto &c = pmb->coords;
    auto ib = pmb->cellbounds.GetBoundsI(IndexDomain::entire);
    auto jb = pmb->cellbounds.GetBoundsJ(IndexDomain::entire);
    auto kb = pmb->cellbounds.GetBoundsK(IndexDomain::entire);
    const bool d2 = pmesh->ndim >= 2, d3 = pmesh->ndim >= 3;
    const long o1 = std::lround((c.Xf<1>(ng) - pmesh->mesh_size.xmin(X1DIR)) / c.Dxc<1>());
    const long o2 = d2 ? std::lround((c.Xf<2>(ng) - pmesh->mesh_size.xmin(X2DIR)) / c.Dxc<2>()) : 0;
    const long o3 = d3 ? std::lround((c.Xf<3>(ng) - pmesh->mesh_size.xmin(X3DIR)) / c.Dxc<3>()) : 0;
    for (int k = kb.s; k <= kb.e; ++k)
      for (int j = jb.s; j <= jb.e; ++j)
        for (int i = ib.s; i <= ib.e; ++i) {
          const long gi = o1 + i - ng, gj = d2 ? o2 + j - ng : 0;
          long gk = d3 ? o3 + k - ng : 0;
          if (d3) gk = ((gk % nx3) + nx3) % nx3;
          rho(k, j, i) *= 1.0 + amp * (HashUnit(gi, gj, gk, 1) - 0.5);
          eng(k, j, i) *= 1.0 + amp * (HashUnit(gi, gj, gk, 2) - 0.5);
          vel(2, k, j, i) *= 1.0 + amp * (HashUnit(gi, gj, gk, 3) - 0.5);
        }
    fluid::PrimitiveToConserved(rc.get());
  }
}

class DumpDriver : public phoebus::PhoebusDriver {
 public:
  using phoebus::PhoebusDriver::PhoebusDriver;
  // Same control flow as PhoebusDriver::Step (src/phoebus_driver.cpp:121-143) + dumps.
  TaskListStatus Step() override {
    TaskListStatus status;
    Real dt_trial = tm.dt;
    if (first_call_) dt_trial = std::min(dt_trial, dt_init);
    first_call_ = false;
    if (!is_restart_) {
      dt_trial *= dt_init_fact;
      dt_init_fact = 1.0;
    }
    if (tm.time + dt_trial > tm.tlim) dt_trial = tm.tlim - tm.time;
    tm.dt = dt_trial;
    integrator->dt = dt_trial;
    if (tm.ncycle == 0) {
      g_dumper.Dump(pmesh, "init", 0, 0, tm.time, tm.dt, 0.0, "base", "base", true);
    }
    const bool last_cycle = (tm.time + dt_trial >= tm.tlim) || (tm.nlim >= 0 && tm.ncycle + 1 >= tm.nlim);
    const bool dump_cycle = g_every <= 1 || tm.ncycle % g_every == 0 || last_cycle;
    for (int stage = 1; stage <= integrator->nstages; stage++) {
      TaskCollection tc = RungeKuttaStage(stage);
      status = tc.Execute();
      if (status != TaskListStatus::complete) break;
      if (!dump_cycle) continue;
      char tag[64];
      std::snprintf(tag, sizeof(tag), "c%04ds%d", tm.ncycle, stage);
      g_dumper.Dump(pmesh, tag, tm.ncycle, stage, tm.time, tm.dt,
                    integrator->beta[stage - 1], integrator->stage_name[stage - 1],
                    integrator->stage_name[stage], false);
    }
    status = RadiationPostStep();
    return status;
  }

 private:
  bool first_call_ = true;
};

} // namespace

int main(int argc, char *argv[]) {
  const char *d = std::getenv("REF_DUMP_DIR");
  g_dumper.dir = d ? d : ".";
  g_dumper.with_qlqr = std::getenv("REF_DUMP_QLQR") != nullptr;
  if (const char *e = std::getenv("REF_DUMP_EVERY")) g_every = std::atoi(e);
  if (const char *e = std::getenv("REF_DUMP_FIELDS")) g_dumper.hash_only = std::string(e) == "hash";

  parthenon::ParthenonManager pman;
  auto manager_status = pman.ParthenonInitEnv(argc, argv);
  if (manager_status == ParthenonStatus::complete) {
    pman.ParthenonFinalize();
    return 0;
  }
  if (manager_status == ParthenonStatus::error) {
    pman.ParthenonFinalize();
    return 1;
  }
  pman.app_input->ProcessPackages = phoebus::ProcessPackages;
  pman.app_input->ProblemGenerator = phoebus::ProblemGenerator;
  pman.app_input->InitMeshBlockUserData = Geometry::SetGeometryBlock;
  pman.app_input->MeshBlockUserWorkBeforeOutput = phoebus::UserWorkBeforeOutput;
  phoebus::ProblemModifier(pman.pinput.get());
  Boundaries::ProcessBoundaryConditions(pman);
  pman.ParthenonInitPackagesAndMesh();
  if (const char *e = std::getenv("REF_DUMP_PERTURB")) PerturbState(pman.pmesh.get(), std::atof(e));
  {
    if (!pman.IsRestart()) {
      phoebus::PostInitializationModifier(pman.pinput.get(), pman.pmesh.get());
    }
    DumpDriver driver(pman.pinput.get(), pman.app_input.get(), pman.pmesh.get(),
                      pman.IsRestart());
    driver.PostInitializationCommunication();
    driver.Execute();
  }
  pman.ParthenonFinalize();
  return 0;
}
