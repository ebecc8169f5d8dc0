// phoebus_host_stage: drives the C++ host layer from the command line (tests).
//
//   phoebus_host_stage describe <deck> <tables.bin> [block/key=value ...]
//       no GPU needed: runs the package Initialize functions and builds the Mesh metadata, prints
//       the flattened pb_cfg_t as "key value" lines and writes blocks + geometry + floor tables.
//   phoebus_host_stage run <deck> <state_in.bin> <state_out.bin> <fused|tasks> <stage:dt>[,<stage:dt>...] [overrides]
//       loads prim temp c2p_mu fail csig fsig cons(base) cons(1) (raw doubles, pack order), runs the
//       listed Runge-Kutta stages through PhoebusDriver::RungeKuttaStage and writes the same fields.
//   phoebus_host_stage evolve <deck> <state_in.bin> <state_out.bin> <fused|tasks> <ncycles> <dt0> [overrides]
//       full cycles with the reference's dt control (EstimateTimestep, 2x limiter, dt_init_fact).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "phoebus_host.hpp"

static const char *kFields[] = {"prim", "temp", "c2p_mu", "fail", "csig", "fsig", "base", "1"};

// The state file always describes the WHOLE mesh (blocks in global order, x1 fastest); with world > 1 every rank loads its
// own MeshBlocks from it and writes them, in its local order, to <path>.rank<r>.
static void LoadState(Mesh &m, const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  for (const char *name : kFields) {
    const size_t per_block = (size_t)m.FieldWords(name) / (size_t)m.nblocks;
    std::vector<double> all(per_block * (size_t)m.NumBlocksTotal()), buf((size_t)m.FieldWords(name));
    f.read((char *)all.data(), (std::streamsize)(all.size() * 8));
    if (!f) throw std::runtime_error("short read of field " + std::string(name));
    for (int n = 0; n < m.nblocks; ++n)
      std::memcpy(buf.data() + (size_t)n * per_block, all.data() + (size_t)m.Gid(m.lx[n]) * per_block, per_block * 8);
    m.Upload(name, buf.data());
  }
}
static void SaveState(const Mesh &m, const std::string &path0, const std::vector<double> &extra) {
  const std::string path = m.world > 1 ? path0 + ".rank" + std::to_string(m.rank) : path0;
  std::ofstream f(path, std::ios::binary);
  for (const char *name : kFields) {
    std::vector<double> buf((size_t)m.FieldWords(name));
    m.Download(name, buf.data());
    f.write((const char *)buf.data(), (std::streamsize)(buf.size() * 8));
  }
  f.write((const char *)extra.data(), (std::streamsize)(extra.size() * 8));
}

int main(int argc, char *argv[]) {
  try {
    if (argc < 4) throw std::runtime_error("usage: see the header of host_stage.cpp");
    const std::string cmd = argv[1];
    auto pin = std::make_unique<ParameterInput>(argv[2]);
    const int fixed = cmd == "describe" ? 4 : (cmd == "run" ? 7 : 8);
    if (argc < fixed) throw std::runtime_error("too few arguments for " + cmd);
    pin->ModifyFromCmdline(argc - fixed, argv + fixed);
    if (cmd == "describe") pin->SetString("phoebus_b200", "allocate", "false");
    auto packages = phoebus::ProcessPackages(pin);
    Mesh mesh(pin.get(), packages);
    if (cmd == "describe") {
      const pb_cfg_t &c = mesh.cfg;
      std::printf("ndim %d\nnx %d %d %d\nng %d\nmhd %d\nrecon %d\nriemann %d\n", c.ndim, c.nx[0], c.nx[1], c.nx[2], c.ng, c.mhd,
                  c.recon, c.riemann);
      std::printf("c2p_tol %.17g\nc2p_max_iter %d\nc2p_floor_scale_fac %.17g\nc2p_fail_on_floors %d\nc2p_fail_on_ceilings %d\n",
                  c.c2p_tol, c.c2p_max_iter, c.c2p_floor_scale_fac, c.c2p_fail_on_floors, c.c2p_fail_on_ceilings);
      std::printf("floor_type %d\nrho0 %.17g\nsie0 %.17g\nrexp %.17g\nsexp %.17g\nrho_min_eos %.17g\nsie_min_eos %.17g\n", c.floor_type,
                  c.rho0, c.sie0, c.rexp, c.sexp, c.rho_min_eos, c.sie_min_eos);
      std::printf("gam_max %.17g\nsie_max %.17g\nbsqorho_max %.17g\nbsqou_max %.17g\n", c.gam_max, c.sie_max, c.bsqorho_max, c.bsqou_max);
      std::printf("enable_floors %d\nenable_mhd_floors %d\nenable_c2p_fixup %d\nenable_flux_fixup %d\nc2p_failure_strategy %d\n",
                  c.enable_floors, c.enable_mhd_floors, c.enable_c2p_fixup, c.enable_flux_fixup, c.c2p_failure_strategy);
      std::printf("gm1 %.17g\nCv %.17g\ngk %d\ngj %d\ngi %d\nbc_prim %d\nox1_inflow_check %d\n", c.gm1, c.Cv, c.gk, c.gj, c.gi, c.bc_prim,
                  c.ox1_inflow_check);
      std::printf("nblocks %d\nnprim %d\nncons %d\ngeom_words %zu\nfloor_words %zu\n", mesh.nblocks, mesh.nprim, mesh.ncons,
                  mesh.geom_host().size(), mesh.floor_host().size());
      std::printf("rank %d\nworld %d\nrgrid %d %d %d\nnsend %zu\nnrecv %zu\nnpeers %zu\n", mesh.rank, mesh.world, mesh.rgrid[0], mesh.rgrid[1],
                  mesh.rgrid[2], mesh.send_plan.size() / 10, mesh.recv_plan.size() / 10, mesh.peers.size());
      std::ofstream f(argv[3], std::ios::binary);
      f.write((const char *)mesh.blocks_host.data(), (std::streamsize)(mesh.blocks_host.size() * sizeof(pb_block_t)));
      f.write((const char *)mesh.geom_host().data(), (std::streamsize)(mesh.geom_host().size() * 8));
      f.write((const char *)mesh.floor_host().data(), (std::streamsize)(mesh.floor_host().size() * 8));
      // inter-GPU halo plans (world > 1): send rows, receive rows, then per peer {rank, send offset, send count, recv offset, recv count}
      f.write((const char *)mesh.send_plan.data(), (std::streamsize)(mesh.send_plan.size() * sizeof(int)));
      f.write((const char *)mesh.recv_plan.data(), (std::streamsize)(mesh.recv_plan.size() * sizeof(int)));
      for (size_t q = 0; q < mesh.peers.size(); ++q) {
        const long long row[5] = {mesh.peers[q], mesh.send_off[q], mesh.send_cnt[q], mesh.recv_off[q], mesh.recv_cnt[q]};
        f.write((const char *)row, sizeof row);
      }
      return 0;
    }
    const bool fused = std::string(argv[5]) == "fused";
    LoadState(mesh, argv[3]);
    phoebus::PhoebusDriver driver(pin.get(), &mesh, fused);
    mesh.has_face_speeds = true; // the loaded state is past its first EstimateTimestep
    std::vector<double> extra;
    if (cmd == "run") {
      std::string list = argv[6];
      size_t pos = 0;
      while (pos < list.size()) {
        size_t comma = list.find(',', pos);
        if (comma == std::string::npos) comma = list.size();
        const std::string item = list.substr(pos, comma - pos);
        const size_t colon = item.find(':');
        driver.dt = std::strtod(item.c_str() + colon + 1, nullptr);
        driver.RungeKuttaStage(std::atoi(item.substr(0, colon).c_str()));
        pos = comma + 1;
      }
    } else if (cmd == "evolve") {
      const int ncycles = std::atoi(argv[6]);
      driver.dt = std::strtod(argv[7], nullptr);
      for (int n = 0; n < ncycles; ++n) {
        driver.Step();
        extra.push_back(driver.dt);
      }
    } else {
      throw std::runtime_error("unknown command " + cmd);
    }
    mesh.Synchronize();
    SaveState(mesh, argv[4], extra);
    std::printf("ok %s %d blocks\n", cmd.c_str(), mesh.nblocks);
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "phoebus_host_stage: %s\n", e.what());
    return 1;
  }
}
