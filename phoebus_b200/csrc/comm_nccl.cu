// Inter-GPU ghost exchange behind the C ABI: what Parthenon's SendBoundBufs / ReceiveBoundBufs / SetBounds do over MPI
// (ref: external/parthenon/src/bvals/comms/boundary_communication.cpp:49-346) done with NCCL send/recv over NVLink:
//   pb_halo_exchange_begin  pack the per-peer send slices (one launch) on the caller's stream, then one grouped
//                           ncclSend / ncclRecv per peer on the communicator's own stream (fork by event)
//   ... the caller issues the same-GPU ghost copy and the interior con2prim on its stream: they overlap the transfer ...
//   pb_halo_exchange_end    join (event) and unpack the received slices into the ghost zones (one launch)
// plus the min all-reduce of the time step (ref: parthenon driver.cpp:237) and CUDA-graph capture of a whole RK stage.
//
// NCCL is resolved at run time with dlopen (libnccl.so.2: the copy already loaded by the host process — torch's — or the
// system one), so the library loads on machines without NCCL and nothing here is needed for single-GPU use.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "pb_device.cuh"
#include "pb_host.hpp"

namespace {
// the part of nccl.h this file needs (NCCL 2.x ABI)
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclMin = 3 };
struct Nccl {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  const char *err = nullptr;
};
Nccl &nccl() {
  static Nccl n = [] {
    Nccl x;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      x.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (x.h) break;
    }
    if (!x.h) {
      x.err = "libnccl.so.2 not found (dlopen)";
      return x;
    }
#define PB_SYM(field, sym)                                                   \
  x.field = reinterpret_cast<decltype(x.field)>(dlsym(x.h, sym));            \
  if (!x.field) x.err = "NCCL symbol missing: " sym;
    PB_SYM(GetUniqueId, "ncclGetUniqueId")
    PB_SYM(CommInitRank, "ncclCommInitRank")
    PB_SYM(CommDestroy, "ncclCommDestroy")
    PB_SYM(GroupStart, "ncclGroupStart")
    PB_SYM(GroupEnd, "ncclGroupEnd")
    PB_SYM(Send, "ncclSend")
    PB_SYM(Recv, "ncclRecv")
    PB_SYM(AllReduce, "ncclAllReduce")
    PB_SYM(GetErrorString, "ncclGetErrorString")
#undef PB_SYM
    return x;
  }();
  return n;
}
int nccl_error(ncclResult_t r, const char *where) {
  char msg[256];
  snprintf(msg, sizeof msg, "%s: %s", where, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
  return pb::set_error(PB_ECUDA, msg);
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;     // the transfers run here, next to the caller's compute stream
  cudaEvent_t packed = nullptr, arrived = nullptr;
};
} // namespace

using namespace pb;

extern "C" int pb_comm_unique_id(char out[128]) {
  if (!out) return set_error(PB_EINVAL, "pb_comm_unique_id: NULL");
  if (nccl().err) return set_error(PB_ECUDA, nccl().err);
  ncclUniqueId id;
  const ncclResult_t r = nccl().GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_error(r, "ncclGetUniqueId");
  memcpy(out, id.internal, 128);
  return PB_OK;
}

extern "C" int pb_comm_init(void **comm_out, int rank, int world, const char id_bytes[128]) {
  if (!comm_out || !id_bytes || world < 1 || rank < 0 || rank >= world) return set_error(PB_EINVAL, "pb_comm_init: bad argument");
  if (nccl().err) return set_error(PB_ECUDA, nccl().err);
  Comm *c = new Comm;
  c->rank = rank;
  c->world = world;
  ncclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  const ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return nccl_error(r, "ncclCommInitRank");
  }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->packed, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->arrived, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return set_cuda_error(cudaGetLastError(), "pb_comm_init");
  }
  *comm_out = c;
  return PB_OK;
}

extern "C" int pb_comm_destroy(void *comm) {
  Comm *c = static_cast<Comm *>(comm);
  if (!c) return PB_OK;
  if (c->comm) nccl().CommDestroy(c->comm);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->packed) cudaEventDestroy(c->packed);
  if (c->arrived) cudaEventDestroy(c->arrived);
  delete c;
  return PB_OK;
}

static int check_plan(const pb_halo_plan_t *plan, void *comm) {
  if (!plan || !comm) return set_error(PB_EINVAL, "NULL halo plan or communicator");
  if (plan->npeers < 0 || (plan->npeers > 0 && (!plan->peer || !plan->send_off || !plan->send_cnt || !plan->recv_off || !plan->recv_cnt)))
    return set_error(PB_EINVAL, "halo plan: NULL peer table");
  if ((plan->nsend > 0 && (!plan->send_plan || !plan->send_buf)) || (plan->nrecv > 0 && (!plan->recv_plan || !plan->recv_buf)))
    return set_error(PB_EINVAL, "halo plan: NULL segment list or buffer");
  return PB_OK;
}

extern "C" int pb_halo_exchange_begin(const pb_pack_t *pack, const pb_cfg_t *cfg, const double *cons, const pb_halo_plan_t *plan,
                                      void *comm, void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if ((rc = check_plan(plan, comm))) return rc;
  Comm *c = static_cast<Comm *>(comm);
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = pb_halo_pack(pack, cfg, cons, plan->send_plan, plan->nsend, plan->send_buf, stream))) return rc;
  cudaError_t e = cudaEventRecord(c->packed, s);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->packed, 0);
  if (e != cudaSuccess) return set_cuda_error(e, "pb_halo_exchange_begin");
  ncclResult_t r = nccl().GroupStart();
  for (int p = 0; p < plan->npeers && r == ncclSuccess; ++p) {
    if (plan->send_cnt[p] > 0) r = nccl().Send(plan->send_buf + plan->send_off[p], (size_t)plan->send_cnt[p], ncclFloat64, plan->peer[p], c->comm, c->stream);
    if (r == ncclSuccess && plan->recv_cnt[p] > 0)
      r = nccl().Recv(plan->recv_buf + plan->recv_off[p], (size_t)plan->recv_cnt[p], ncclFloat64, plan->peer[p], c->comm, c->stream);
  }
  const ncclResult_t r2 = nccl().GroupEnd();
  if (r != ncclSuccess) return nccl_error(r, "ncclSend/ncclRecv");
  if (r2 != ncclSuccess) return nccl_error(r2, "ncclGroupEnd");
  if ((e = cudaEventRecord(c->arrived, c->stream)) != cudaSuccess) return set_cuda_error(e, "pb_halo_exchange_begin");
  return PB_OK;
}

extern "C" int pb_halo_exchange_end(const pb_pack_t *pack, const pb_cfg_t *cfg, double *cons, const pb_halo_plan_t *plan, void *comm,
                                    void *stream) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if ((rc = check_plan(plan, comm))) return rc;
  Comm *c = static_cast<Comm *>(comm);
  const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, c->arrived, 0);
  if (e != cudaSuccess) return set_cuda_error(e, "pb_halo_exchange_end");
  return pb_halo_unpack(pack, cfg, cons, plan->recv_plan, plan->nrecv, plan->recv_buf, stream);
}

extern "C" int pb_comm_allreduce_min(void *comm, double *value_dev, int n, void *stream) {
  Comm *c = static_cast<Comm *>(comm);
  if (!c || !value_dev || n <= 0) return set_error(PB_EINVAL, "pb_comm_allreduce_min: bad argument");
  const ncclResult_t r = nccl().AllReduce(value_dev, value_dev, (size_t)n, ncclFloat64, ncclMin, c->comm, (cudaStream_t)stream);
  return r == ncclSuccess ? PB_OK : nccl_error(r, "ncclAllReduce");
}

// ------------------------------------------------------------------ one RK stage as a CUDA graph
// A stage is 20-30 short launches; on small packs (strong scaling: 16 MeshBlocks per GPU) the launch latency dominates.
// The graph captures exactly what pb_rk_stage (comm == NULL) or pb_rk_stage_pre / halo exchange with the interior con2prim
// under the transfer / pb_rk_stage_post2 (comm != NULL) would launch, with the time step read from device memory (dt_dev)
// at run time, so one instantiated graph per (stage, container pair) serves every cycle.
namespace pb {
void set_stage_dt_pointer(const double *dt_dev); // update_kernels.cu
int presize_workspaces(const pb_pack_t *pack, const pb_cfg_t *cfg, cudaStream_t stream); // zone_kernels.cu
}

struct StageGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

extern "C" int pb_stage_graph_create(const pb_pack_t *pack, const pb_cfg_t *cfg, double *sc0, double *base, double *sc1, double beta,
                                     const double *dt_dev, unsigned long long *counters, const pb_halo_plan_t *plan, void *comm,
                                     void *stream, void **graph_out) {
  int rc = check_args(pack, cfg);
  if (rc) return rc;
  if (!graph_out || !dt_dev || !sc0 || !base || !sc1) return set_error(PB_EINVAL, "pb_stage_graph_create: NULL argument");
  if (comm && (rc = check_plan(plan, comm))) return rc;
  if (!stream) return set_error(PB_EINVAL, "pb_stage_graph_create: stream capture needs a non-default stream");
  if ((rc = presize_workspaces(pack, cfg, (cudaStream_t)stream))) return rc; // no allocation may happen inside the capture
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamBeginCapture");
  set_stage_dt_pointer(dt_dev);
  if (!comm) {
    rc = pb_rk_stage(pack, cfg, sc0, base, sc1, beta, 0.0, counters, stream);
  } else {
    rc = pb_rk_stage_pre(pack, cfg, sc0, base, sc1, beta, 0.0, counters, stream);
    if (!rc) rc = pb_halo_exchange_begin(pack, cfg, sc1, plan, comm, stream);
    if (!rc) rc = pb_halo_exchange_local_ex(pack, cfg, sc1, pb_ghost_plan_active(pack, cfg), stream);
    const int early = pb_early_c2p_pays(pack);
    if (!rc && early) rc = pb_rk_stage_post_early(pack, cfg, sc1, stream);
    if (!rc) rc = pb_halo_exchange_end(pack, cfg, sc1, plan, comm, stream);
    if (!rc) rc = pb_rk_stage_post2(pack, cfg, sc1, counters, early, stream);
  }
  set_stage_dt_pointer(nullptr);
  StageGraph *g = new StageGraph;
  e = cudaStreamEndCapture(s, &g->graph);
  if (rc) {
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return rc;
  }
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return set_cuda_error(e, "pb_stage_graph_create");
  }
  *graph_out = g;
  return PB_OK;
}

extern "C" int pb_stage_graph_launch(void *graph, void *stream) {
  StageGraph *g = static_cast<StageGraph *>(graph);
  if (!g || !g->exec) return set_error(PB_EINVAL, "pb_stage_graph_launch: NULL graph");
  const cudaError_t e = cudaGraphLaunch(g->exec, (cudaStream_t)stream);
  return e == cudaSuccess ? PB_OK : set_cuda_error(e, "cudaGraphLaunch");
}

extern "C" int pb_stage_graph_destroy(void *graph) {
  StageGraph *g = static_cast<StageGraph *>(graph);
  if (!g) return PB_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return PB_OK;
}
