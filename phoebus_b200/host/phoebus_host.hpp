// The reference's fluid-path task functions, with the reference's names and signatures, each one
// forwarding to one C-ABI call of libphoebus_b200.so (include/phoebus_b200.h) on the whole pack
// the container describes. No Kokkos, no CPU fallback: if the library or the device is missing
// the C-ABI call returns PB_ECUDA and the task throws (the reference's PARTHENON_FAIL behaviour).
//
// ref: src/fluid/fluid.hpp:28-60, src/fixup/fixup.hpp:34-42,
//      src/phoebus_boundaries/phoebus_boundaries.hpp, external/parthenon/src/interface/update.hpp:41-136,
//      external/parthenon/src/bvals/comms/bvals_in_one.hpp, src/phoebus_driver.cpp:204-738.
#ifndef PHOEBUS_B200_HOST_PHOEBUS_HOST_HPP_
#define PHOEBUS_B200_HOST_PHOEBUS_HOST_HPP_

#include <memory>
#include <string>
#include <vector>

#include "parthenon_shim.hpp"

using parthenon::IndexDomain;
using parthenon::IndexRange;
using parthenon::Mesh;
using parthenon::MeshBlockData;
using parthenon::MeshData;
using parthenon::ParameterInput;
using parthenon::Params;
using parthenon::Real;
using parthenon::StateDescriptor;
using parthenon::TaskStatus;

namespace fluid {
// ref: src/fluid/fluid.cpp:53-366
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin);

TaskStatus PrimitiveToConserved(MeshBlockData<Real> *rc);
TaskStatus PrimitiveToConservedRegion(MeshBlockData<Real> *rc, const IndexRange &ib, const IndexRange &jb,
                                      const IndexRange &kb);
template <typename T>
TaskStatus ConservedToPrimitiveRegion(T *rc, const IndexRange &ib, const IndexRange &jb, const IndexRange &kb);
template <typename T>
TaskStatus ConservedToPrimitive(T *rc);
template <typename T>
TaskStatus ConservedToPrimitiveRobust(T *rc, const IndexRange &ib, const IndexRange &jb, const IndexRange &kb);
TaskStatus CalculateFluidSourceTerms(MeshData<Real> *rc, MeshData<Real> *rc_src);
TaskStatus CalculateFluxes(MeshBlockData<Real> *rc);
TaskStatus FluxCT(MeshBlockData<Real> *rc);
TaskStatus CalculateDivB(MeshBlockData<Real> *rc);
TaskStatus CalculateDivB(MeshData<Real> *rc);
Real EstimateTimestepBlock(MeshBlockData<Real> *rc);
// whole-partition overloads (one launch for every block of the GPU instead of one per block)
TaskStatus CalculateFluxes(MeshData<Real> *rc);
TaskStatus FluxCT(MeshData<Real> *rc);
TaskStatus PrimitiveToConserved(MeshData<Real> *rc);
Real EstimateTimestepMesh(MeshData<Real> *rc);
} // namespace fluid

namespace fixup {
// ref: src/fixup/fixup.cpp:53-277
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin);
TaskStatus FixFluxes(MeshBlockData<Real> *rc);
TaskStatus FixFluxes(MeshData<Real> *rc);
template <typename T>
TaskStatus ApplyFloors(T *rc);
template <typename T>
TaskStatus ConservedToPrimitiveFixup(T *rc);
} // namespace fixup

namespace Microphysics {
namespace EOS {
// ref: src/microphysics/eos_phoebus/eos_phoebus.cpp:46-95 (IdealGas only on this path)
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin);
} // namespace EOS
} // namespace Microphysics

namespace Geometry {
// ref: src/geometry/geometry.cpp:38-60; the coordinate system is a run-time choice here
// (<geometry> kind = minkowski | fmks) instead of -DPHOEBUS_GEOMETRY
std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin);
} // namespace Geometry

namespace Boundaries {
// ref: src/phoebus_boundaries/phoebus_boundaries.cpp:423-448; also performs the ghost fill of
// parthenon::ApplyBoundaryConditions (gr_outflow / reflect / polar, :44-272) that precedes it
TaskStatus ConvertBoundaryConditions(std::shared_ptr<MeshBlockData<Real>> &rc);
TaskStatus ConvertBoundaryConditions(std::shared_ptr<MeshData<Real>> &rc);
} // namespace Boundaries

namespace parthenon {
// ref: external/parthenon/src/bvals/boundary_conditions.cpp:36-80 — folded into ConvertBoundaryConditions
template <class T>
TaskStatus ApplyBoundaryConditions(std::shared_ptr<T> &rc);
// ref: external/parthenon/src/bvals/comms/boundary_communication.cpp:49-346. Same-GPU neighbours
// are copied by SendBoundBufs in one launch; Receive/Set complete immediately.
TaskStatus SendBoundBufs(std::shared_ptr<MeshData<Real>> &md);
TaskStatus ReceiveBoundBufs(std::shared_ptr<MeshData<Real>> &md);
TaskStatus SetBounds(std::shared_ptr<MeshData<Real>> &md);

namespace Update {
// ref: external/parthenon/src/interface/update.cpp:63-85
template <class T>
TaskStatus FluxDivergence(T *in, T *dudt_obj);
// ref: external/parthenon/src/interface/update.hpp:71-136
template <class F, class T>
TaskStatus SumData(const F &flags, T *in1, T *in2, T *out);
template <class T>
TaskStatus AverageIndependentData(T *c1, T *c2, const Real wgt1);
template <class T>
TaskStatus UpdateIndependentData(T *in, T *dudt, const Real dt, T *out);
// ref: update.hpp:139-160 — the fluid package's FillDerived is ConservedToPrimitive (fluid.cpp:347)
template <class T>
TaskStatus FillDerived(T *rc);
} // namespace Update
} // namespace parthenon

namespace phoebus {
// The fluid-only sub-graph of PhoebusDriver (ref: src/phoebus_driver.cpp:54-134, 204-738) over one
// GPU's Mesh. `fused` selects pb_rk_stage (one call per stage); otherwise every reference task is
// issued separately, in the reference's order, through the functions above.
class PhoebusDriver {
 public:
  PhoebusDriver(ParameterInput *pin, Mesh *pm, bool fused = true);
  TaskStatus RungeKuttaStage(int stage); // 1-based, uses dt
  TaskStatus Step();                     // all stages, then the next dt proposal
  void PostInitializationCommunication(); // ref: phoebus_driver.cpp:137-202
  Real EstimateTimestep();
  Real dt = 0.0, time = 0.0;
  int ncycle = 0;
  std::vector<Real> beta; // low-storage integrator weights (ref: low_storage_integrator.cpp:67-86)
  std::vector<std::string> stage_name;

 private:
  Mesh *pmesh;
  bool fused_;
  bool first_step_ = true;
  Real dt_init_fact_ = 1.0;
  Real tlim_ = 0.0;
};

// packages on the path, in the reference's order (ref: src/phoebus_package.cpp / main.cpp ProcessPackages)
parthenon::Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin);
} // namespace phoebus

#endif // PHOEBUS_B200_HOST_PHOEBUS_HOST_HPP_
