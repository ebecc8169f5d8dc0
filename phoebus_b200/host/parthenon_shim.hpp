// Parthenon-shaped host types for the B200 fluid path — NOT Parthenon, and no Kokkos.
//
// The reference's task functions take parthenon::MeshBlockData<Real>* / MeshData<Real>* and read
// their configuration from StateDescriptor Params filled by <pkg>::Initialize(ParameterInput*).
// This header supplies the subset of those types the fluid hot path touches, with the same names,
// so that the task functions in phoebus_host.hpp keep the reference's signatures
// (ref: src/fluid/fluid.hpp:28-60, src/fixup/fixup.hpp:34-42,
//  external/parthenon/src/interface/update.hpp:41-136). A Phoebus maintainer replaces this header
// with <parthenon/package.hpp> and fills pb_pack_t from the real containers.
//
// Storage model: a Mesh owns one device slab per field for ALL MeshBlocks of this GPU
// (include/phoebus_b200.h). MeshData<Real> is a view of the whole partition in one stage container
// ("base", "1", "dUdt", "geometric source terms"); MeshBlockData<Real> is the same view narrowed to
// one block — both are just a pb_pack_t plus the container's conserved slab.
#ifndef PHOEBUS_B200_HOST_PARTHENON_SHIM_HPP_
#define PHOEBUS_B200_HOST_PARTHENON_SHIM_HPP_

#include <any>
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/phoebus_b200.h"

namespace parthenon {

using Real = double;

// ref: external/parthenon/src/basic_types.hpp:61
enum class TaskStatus { fail, complete, incomplete, iterate, skip, waiting };
enum class IndexDomain { entire, interior };
// ref: external/parthenon/src/mesh/domain.hpp:37-46 (inclusive bounds)
struct IndexRange {
  int s = 0, e = 0;
};

// ref: external/parthenon/src/interface/params.hpp — typed key/value store
class Params {
 public:
  template <class T>
  void Add(const std::string &key, T value, bool /*is_mutable*/ = false) {
    if (map_.count(key)) throw std::invalid_argument("Key " + key + " already exists");
    map_[key] = std::move(value);
  }
  template <class T>
  void Update(const std::string &key, T value) {
    map_[key] = std::move(value);
  }
  template <class T>
  const T &Get(const std::string &key) const {
    auto it = map_.find(key);
    if (it == map_.end()) throw std::invalid_argument("Key " + key + " doesn't exist");
    const T *p = std::any_cast<T>(&it->second);
    if (!p) throw std::invalid_argument("WRONG TYPE FOR KEY '" + key + "'");
    return *p;
  }
  bool hasKey(const std::string &key) const { return map_.count(key) > 0; }

 private:
  std::map<std::string, std::any> map_;
};

// ref: external/parthenon/src/parameter_input.hpp — "<block>" / "key = value" / "# comment" decks
class ParameterInput {
 public:
  ParameterInput() = default;
  explicit ParameterInput(const std::string &filename) { LoadFromFile(filename); }
  void LoadFromFile(const std::string &filename);
  void LoadFromString(const std::string &text);
  void ModifyFromCmdline(int argc, char *argv[]); // block/key=value overrides
  bool DoesParameterExist(const std::string &block, const std::string &name) const;
  int GetInteger(const std::string &block, const std::string &name) const;
  Real GetReal(const std::string &block, const std::string &name) const;
  bool GetBoolean(const std::string &block, const std::string &name) const;
  std::string GetString(const std::string &block, const std::string &name) const;
  int GetOrAddInteger(const std::string &block, const std::string &name, int def);
  Real GetOrAddReal(const std::string &block, const std::string &name, Real def);
  bool GetOrAddBoolean(const std::string &block, const std::string &name, bool def);
  std::string GetOrAddString(const std::string &block, const std::string &name, const std::string &def);
  void SetString(const std::string &block, const std::string &name, const std::string &v) { blocks_[block][name] = v; }

 private:
  const std::string *Find(const std::string &block, const std::string &name) const;
  const std::string &Require(const std::string &block, const std::string &name) const;
  std::map<std::string, std::map<std::string, std::string>> blocks_;
};

// ref: external/parthenon/src/interface/state_descriptor.hpp
class StateDescriptor {
 public:
  explicit StateDescriptor(std::string label) : label_(std::move(label)) {}
  const std::string &label() const { return label_; }
  template <class T>
  void AddParam(const std::string &key, T value, bool is_mutable = false) {
    params_.Add(key, std::move(value), is_mutable);
  }
  template <class T>
  void UpdateParam(const std::string &key, T value) {
    params_.Update(key, std::move(value));
  }
  template <class T>
  const T &Param(const std::string &key) const {
    return params_.Get<T>(key);
  }
  Params &AllParams() { return params_; }
  const Params &AllParams() const { return params_; }

 private:
  std::string label_;
  Params params_;
};

class Packages_t {
 public:
  void Add(const std::shared_ptr<StateDescriptor> &pkg) { pkgs_[pkg->label()] = pkg; }
  const std::shared_ptr<StateDescriptor> &Get(const std::string &name) const {
    auto it = pkgs_.find(name);
    if (it == pkgs_.end()) throw std::invalid_argument("package " + name + " is not registered");
    return it->second;
  }
  bool Has(const std::string &name) const { return pkgs_.count(name) > 0; }

 private:
  std::map<std::string, std::shared_ptr<StateDescriptor>> pkgs_;
};

class Mesh;

// What both container types are underneath: a (sub-)pack and the conserved slab of one stage.
struct PackView {
  Mesh *pmesh = nullptr;
  std::string stage;          // container name: "base", "1", "dUdt", "geometric source terms"
  pb_pack_t pack{};           // device pointers already offset to block `first`
  double *cons = nullptr;     // [nblocks][ncons][cells] of this container (gsrc: [nblocks][4][cells])
  int first = 0, nblocks = 0; // block range inside the Mesh
  Mesh *GetMeshPointer() const { return pmesh; }
  Mesh *GetParentPointer() const { return pmesh; }
  int NumBlocks() const { return nblocks; }
};
template <class T>
struct MeshBlockData : PackView {};
template <class T>
struct MeshData : PackView {};

// One GPU's share of the uniform mesh: block table, coordinates, geometry / floor tables, device
// slabs and the named stage containers. Mirrors the pieces of parthenon::Mesh the path reads:
// ref: external/parthenon/src/mesh/forest/forest.cpp:70-145, mesh/forest/logical_location.cpp:61-79,
// defs.hpp:98-101, coordinates/uniform_cartesian.hpp:30-40.
class Mesh {
 public:
  Mesh(ParameterInput *pin, Packages_t packages);
  ~Mesh();
  Mesh(const Mesh &) = delete;
  Mesh &operator=(const Mesh &) = delete;

  Packages_t packages;
  pb_cfg_t cfg{};
  int ndim = 1, ng = 0, nblocks = 0;
  // domain decomposition by MeshBlock (ref: external/parthenon/src/mesh/mesh-amr_loadbalance.cpp:362-389; compact
  // bricks here): one process per GPU, <phoebus_b200> rank / world (or PB_RANK / PB_WORLD in the environment)
  int rank = 0, world = 1;
  int rgrid[3] = {1, 1, 1}, brick[3] = {1, 1, 1};
  int Owner(const std::array<int, 3> &l) const {
    return (l[0] / brick[0]) + rgrid[0] * ((l[1] / brick[1]) + rgrid[1] * (l[2] / brick[2]));
  }
  long long Gid(const std::array<int, 3> &l) const { return l[0] + (long long)nblk[0] * (l[1] + (long long)nblk[1] * l[2]); }
  int NumBlocksTotal() const { return nblk[0] * nblk[1] * nblk[2]; }
  // inter-GPU halo: plans as pb_halo_pack expects them, all peers in one send / one receive buffer
  std::vector<int> send_plan, recv_plan;                 // rows of 10 ints
  std::vector<int> peers;
  std::vector<long long> send_off, send_cnt, recv_off, recv_cnt; // slices of the buffers, in doubles
  pb_halo_plan_t plan{};
  void *comm = nullptr;                                  // pb_comm_init handle (world > 1)
  int nx[3] = {1, 1, 1}, nblk[3] = {1, 1, 1}, ncell[3] = {1, 1, 1}; // per block, ghosts included
  long long cells = 0;                                               // ncell[0]*ncell[1]*ncell[2]
  int nprim = 0, ncons = 0;
  std::vector<std::array<int, 3>> lx;  // logical location of each block (x1 fastest)
  std::vector<pb_block_t> blocks_host;
  void *stream = nullptr;              // cudaStream_t every task launches on (default stream)

  // containers -------------------------------------------------------------------------------
  std::shared_ptr<MeshData<Real>> &GetMeshData(const std::string &stage);         // mesh_data.Get(stage)
  std::shared_ptr<MeshBlockData<Real>> GetBlockData(int b, const std::string &stage); // block_list[b]->meshblock_data.Get(stage)
  bool HasContainer(const std::string &stage) const { return cons_.count(stage) > 0; }
  void AddContainer(const std::string &stage); // conserved slab, zero-initialised like a Kokkos::View

  // host <-> device of whole fields (tests, problem generators); names: prim temp c2p_mu fail csig
  // fsig flux emf, or a container name for its conserved slab
  long long FieldWords(const std::string &name) const;
  void Upload(const std::string &name, const double *host);
  void Download(const std::string &name, double *host) const;
  void Synchronize() const;
  double *DivB() const { return field_.at("divb"); } // diagnostic field of fluid::CalculateDivB
  void InvalidateZoneFlags() { flags_valid = false; }
  bool flags_valid = false;
  unsigned long long *counters = nullptr; // device [3]: floors, ceilings, in-kernel c2p failures
  double *dt_dev = nullptr;               // device scalar for EstimateTimestep

  // time-step state (ref: parthenon driver.cpp:216-240, fluid.cpp:911,940)
  bool has_face_speeds = false;
  // host copies of the tables the kernels read (tests compare them with the Python mirror)
  const std::vector<double> &geom_host() const { return geom_host_; }
  const std::vector<double> &floor_host() const { return floor_host_; }

 private:
  void BuildBlocks(ParameterInput *pin);
  void BuildHaloPlans();
  void InitComm(ParameterInput *pin);
  std::map<long long, int> local_index_;
  void BuildTables(ParameterInput *pin);
  void Allocate();
  double *Slab(const std::string &name) const;
  std::map<std::string, double *> field_;
  std::map<std::string, long long> words_;
  std::map<std::string, double *> cons_;
  std::map<std::string, std::shared_ptr<MeshData<Real>>> md_;
  std::vector<void *> owned_;
  pb_pack_t pack_{};
  std::vector<double> geom_host_, floor_host_;
  double mesh_xmin_[3], mesh_xmax_[3];
  std::string mesh_bc_[6];
  bool periodic_[3];
};

} // namespace parthenon

#endif // PHOEBUS_B200_HOST_PARTHENON_SHIM_HPP_
