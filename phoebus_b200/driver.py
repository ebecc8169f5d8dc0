"""FluidDriver: sequences the fluid-only sub-graph of PhoebusDriver::RungeKuttaStage
(ref: src/phoebus_driver.cpp:204-738) over a MeshPack through the C ABI, and the
time-step control of Parthenon's EvolutionDriver (ref: external/parthenon/src/driver/driver.cpp:210-240,
src/phoebus_driver.cpp:121-132).

One process per GPU. With world > 1 the ghost exchange of region 8 goes through the C ABI's NCCL path
(pb_halo_exchange_begin / _end: pack -> grouped ncclSend / ncclRecv per peer on the communicator's stream -> unpack, with
the same-GPU ghost copy and the interior con2prim issued in between); torch.distributed only carries the NCCL unique id
to the ranks (and remains available as the transport with halo="torch"). use_graphs=True replays each RK stage as one
CUDA graph (pb_stage_graph_*).
"""
import ctypes as C

from . import _lib


class FluidDriver:
    def __init__(self, mesh, dist=None, halo="nccl", use_graphs=False):
        self.m = mesh
        self.halo_mode = halo
        self.use_graphs = use_graphs
        self.comm = None
        self._graphs = {}
        self.L = _lib.lib()
        self.torch = mesh.torch
        self.dist = dist if mesh.world > 1 else None
        self.betas = mesh.p.rk_betas()
        self.names = ["base"] + ["1"] * (len(self.betas) - 1) + ["base"]  # ref: staged_integrator.cpp:23-29
        self.has_face_speeds = False
        self.first_step = True
        self.dt = None
        self.ncycle = 0
        self.time = 0.0
        self.launches = 0
        self.timers = None  # dict -> per-task CUDA-event timing (bench.py)
        self.force_split = False  # single GPU: sequence the stage as the multi-GPU path does (pre / halo + early c2p / post2)
        if self.dist is not None:
            self._setup_halo_buffers()
            if self.halo_mode == "nccl":
                self._setup_comm()

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _pk(self):
        return C.byref(self.m.pack), C.byref(self.m.cfg)

    def _setup_halo_buffers(self):
        """One send and one receive buffer for ALL peers (per-peer slices are what NCCL moves), and one
        pack / unpack plan covering every segment: 2 copy launches per exchange instead of 2 per peer."""
        import numpy as np
        t, m = self.torch, self.m
        sp, stot, rp, rtot, peers = m.merged_halo_plans()
        self.sbuf = t.empty(max(stot, 1), dtype=t.float64, device=m.device)
        self.rbuf = t.empty(max(rtot, 1), dtype=t.float64, device=m.device)
        self.halo = [dict(peer=peer, sbuf=self.sbuf[ss], rbuf=self.rbuf[rs], soff=ss.start, scnt=ss.stop - ss.start,
                          roff=rs.start, rcnt=rs.stop - rs.start) for peer, ss, rs in peers]
        self.nsend, self.nrecv = len(sp), len(rp)
        self.splan = t.from_numpy(sp if len(sp) else np.zeros((1, 10), dtype=np.int32)).to(m.device)
        self.rplan = t.from_numpy(rp if len(rp) else np.zeros((1, 10), dtype=np.int32)).to(m.device)

    def _setup_comm(self):
        """NCCL communicator owned by the C library: rank 0 draws the unique id, torch.distributed broadcasts it"""
        t, m, L = self.torch, self.m, self.L
        buf = C.create_string_buffer(128)
        if m.rank == 0:
            _lib.check(L.pb_comm_unique_id(buf), "pb_comm_unique_id")
        idt = t.tensor(list(buf.raw), dtype=t.uint8, device=m.device)
        self.dist.broadcast(idt, 0)
        raw = bytes(idt.cpu().tolist())
        comm = C.c_void_p()
        _lib.check(L.pb_comm_init(C.byref(comm), m.rank, m.world, raw), "pb_comm_init")
        self.comm = comm
        n = len(self.halo)
        self._plan_arrays = ((C.c_int * n)(*[h["peer"] for h in self.halo]),
                             (C.c_longlong * n)(*[h["soff"] for h in self.halo]), (C.c_longlong * n)(*[h["scnt"] for h in self.halo]),
                             (C.c_longlong * n)(*[h["roff"] for h in self.halo]), (C.c_longlong * n)(*[h["rcnt"] for h in self.halo]))
        pl = _lib.PbHaloPlan()
        pl.send_plan, pl.nsend, pl.recv_plan, pl.nrecv = self.splan.data_ptr(), self.nsend, self.rplan.data_ptr(), self.nrecv
        pl.send_buf, pl.recv_buf, pl.npeers = self.sbuf.data_ptr(), self.rbuf.data_ptr(), n
        pl.peer, pl.send_off, pl.send_cnt, pl.recv_off, pl.recv_cnt = self._plan_arrays
        self.plan = pl

    def close(self):
        for g in self._graphs.values():
            self.L.pb_stage_graph_destroy(g)
        self._graphs = {}
        if self.comm is not None:
            self.L.pb_comm_destroy(self.comm)
            self.comm = None

    # ------------------------------------------------------------------ tasks
    def exchange(self, cons, under_transfer=None, in_stage=False):
        """region 8: SendBoundBufs / ReceiveBoundBufs / SetBounds. `under_transfer`, if given, is called after the
        same-GPU ghost copy has been launched and before the remote ghosts are waited for (work to hide the
        NVLink transfer under). in_stage: the call is followed by pb_rk_stage_post2, which fills the ghost zones the
        pack's ghost plan defers (include/phoebus_b200.h), so the same-GPU copy may leave them out."""
        m, L = self.m, self.L
        pk, cfg = self._pk()
        s = self._stream()
        phase = int(L.pb_ghost_plan_active(pk, cfg)) if in_stage else 0
        if self.dist is None:
            _lib.check(L.pb_halo_exchange_local_ex(pk, cfg, cons.data_ptr(), phase, s), "pb_halo_exchange_local_ex")
            if under_transfer is not None:
                under_transfer()
            return
        if self.comm is not None:
            _lib.check(L.pb_halo_exchange_begin(pk, cfg, cons.data_ptr(), C.byref(self.plan), self.comm, s), "pb_halo_exchange_begin")
            _lib.check(L.pb_halo_exchange_local_ex(pk, cfg, cons.data_ptr(), phase, s), "pb_halo_exchange_local_ex")
            if under_transfer is not None:
                under_transfer()
            _lib.check(L.pb_halo_exchange_end(pk, cfg, cons.data_ptr(), C.byref(self.plan), self.comm, s), "pb_halo_exchange_end")
            return
        dist = self.dist
        _lib.check(L.pb_halo_pack(pk, cfg, cons.data_ptr(), self.splan.data_ptr(), self.nsend, self.sbuf.data_ptr(), s), "pb_halo_pack")
        ops = []
        for h in self.halo:
            if h["sbuf"].numel():
                ops.append(dist.P2POp(dist.isend, h["sbuf"], h["peer"]))
            if h["rbuf"].numel():
                ops.append(dist.P2POp(dist.irecv, h["rbuf"], h["peer"]))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        # same-GPU neighbours, then interior work, while the NVLink transfers are in flight
        _lib.check(L.pb_halo_exchange_local_ex(pk, cfg, cons.data_ptr(), phase, s), "pb_halo_exchange_local_ex")
        if under_transfer is not None:
            under_transfer()
        for r in reqs:
            r.wait()
        _lib.check(L.pb_halo_unpack(pk, cfg, cons.data_ptr(), self.rplan.data_ptr(), self.nrecv, self.rbuf.data_ptr(), s), "pb_halo_unpack")

    def ensure_flags(self):
        """zone flags must describe the current p.* / fail fields before a flag-driven stage"""
        m = self.m
        if m.zflags is not None and not m.flags_valid:
            pk, cfg = self._pk()
            _lib.check(self.L.pb_scan_zone_flags(pk, cfg, self._stream()), "pb_scan_zone_flags")
            m.flags_valid = True
            self.launches += 1

    def stage(self, stage, dt):
        m, L = self.m, self.L
        pk, cfg = self._pk()
        s = self._stream()
        beta = self.betas[stage - 1]
        sc0, sc1 = m.cons[self.names[stage - 1]], m.cons[self.names[stage]]
        base = m.cons["base"]
        cnt = C.c_void_p(m.counters.data_ptr())
        self.ensure_flags()
        if self.use_graphs and self.timers is None and (self.dist is None or self.comm is not None) and not self.force_split:
            # one CUDA graph per stage (container pair, beta); dt travels through device memory. Capture needs a real
            # stream (not the legacy default one), so graph mode runs on its own stream, ordered after / before the caller's.
            t = self.torch
            if not hasattr(self, "dt_stage_dev"):
                self.dt_stage_dev = t.zeros(1, dtype=t.float64, device=m.device)
                self._gstream = t.cuda.Stream(device=m.device)
            cur = t.cuda.current_stream()
            self._gstream.wait_stream(cur)
            with t.cuda.stream(self._gstream):
                gs = C.c_void_p(self._gstream.cuda_stream)
                self.dt_stage_dev.fill_(dt)
                g = self._graphs.get(stage)
                if g is None:
                    g = C.c_void_p()
                    _lib.check(L.pb_stage_graph_create(pk, cfg, sc0.data_ptr(), base.data_ptr(), sc1.data_ptr(), beta,
                                                       self.dt_stage_dev.data_ptr(), cnt,
                                                       C.byref(self.plan) if self.comm is not None else None, self.comm, gs, C.byref(g)),
                               "pb_stage_graph_create")
                    self._graphs[stage] = g
                _lib.check(L.pb_stage_graph_launch(g, gs), "pb_stage_graph_launch")
            cur.wait_stream(self._gstream)
            self.launches += self.LAUNCHES_PER_STAGE
            return
        if self.dist is None and not self.force_split:
            _lib.check(L.pb_rk_stage(pk, cfg, sc0.data_ptr(), base.data_ptr(), sc1.data_ptr(), beta, dt, cnt, s), "pb_rk_stage")
        else:
            _lib.check(L.pb_rk_stage_pre(pk, cfg, sc0.data_ptr(), base.data_ptr(), sc1.data_ptr(), beta, dt, cnt, s),
                       "pb_rk_stage_pre")
            # con2prim on the interior of the MeshBlocks away from the mesh boundary runs under the transfer, if the pack
            # has enough of them for the split to pay (pb_early_c2p_pays); force_split always splits (tests)
            pays = bool(L.pb_early_c2p_pays(pk)) or self.force_split
            early = (lambda: _lib.check(L.pb_rk_stage_post_early(pk, cfg, sc1.data_ptr(), s), "pb_rk_stage_post_early")) if pays else None
            if self.timers is not None:
                t = self.torch
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record()
                self.exchange(sc1, early, in_stage=True)
                e1.record()
                self.timers.setdefault("halo+early_c2p", []).append((e0, e1))
            else:
                self.exchange(sc1, early, in_stage=True)
            _lib.check(L.pb_rk_stage_post2(pk, cfg, sc1.data_ptr(), cnt, int(pays), s), "pb_rk_stage_post2")
        self.launches += self.LAUNCHES_PER_STAGE

    # kernels launched per stage by pb_rk_stage: flux x ndim, emf + ct, fix, [source,] update
    # [+ ghosts], fixup, floors, halo, bcs, c2p, fixup, floors
    @property
    def LAUNCHES_PER_STAGE(self):
        m = self.m
        split = m.recon_lr is not None and self.L.pb_get_option(b"flux_split") == 1
        # flux: recon + riemann + riemann_redo per direction (split) or one fused launch; update; fixup; floors...
        n = m.ndim * (3 if split else 1) + 1 + 1 + 1 + 1
        kc = {0: 1, 1: 1, 2: 3, 3: 4}[self.L.pb_get_option(b"c2p_arith")]  # kernels per con2prim task (solve, slow, unwrap, redo)
        n += kc - 1
        if m.gsrc is not None:
            n += 1
        if m.zflags is None:
            n += 1
        if m.p.mhd and m.ndim > 1:
            n += 2
        if m.p.enable_flux_fixup:
            n += 1
        if m.p.enable_c2p_fixup:
            n += 2
        if m.p.enable_floors:
            n += 2
        plan = bool(self.L.pb_ghost_plan_active(*self._pk()))
        if plan:
            n += 1  # the deferred ghost zones are copied after con2prim
            if getattr(m, "n_twin", 0):
                n += 1  # the twin solve, launched with the pre-con2prim ghost copy
        if self.dist is not None:
            n += 2  # one pack and one unpack launch for all peers
            if self.L.pb_early_c2p_pays(C.byref(m.pack)):
                # c2p as three tasks (inner interiors, inner shells, boundary blocks) instead of one; two with the ghost plan
                n += (1 if plan else 2) * kc
        return n

    def enable_profile(self, on=True):
        """per-task CUDA-event timing inside the C-ABI sequencer (pb_profile_enable)"""
        self.L.pb_profile_enable(int(on))
        self.timers = {} if on else None

    def timer_summary(self):
        """task -> (calls, total ms) since the last call; synchronises the recorded events"""
        ms = (C.c_double * len(_lib.TASK_NAMES))()
        calls = (C.c_int * len(_lib.TASK_NAMES))()
        self.L.pb_profile_read(ms, calls)
        out = {n: (calls[i], ms[i]) for i, n in enumerate(_lib.TASK_NAMES) if calls[i]}
        for name, evs in (self.timers or {}).items():
            out[name] = (len(evs), sum(e0.elapsed_time(e1) for e0, e1 in evs))
        if self.timers is not None:
            self.timers = {}
        return out

    def estimate_dt(self):
        """cfl * min over blocks of fluid::EstimateTimestepBlock (ref: src/fluid/fluid.cpp:899-944),
        MPI_Allreduce(MIN) over ranks (ref: parthenon driver.cpp:237)."""
        m, L = self.m, self.L
        pk, cfg = self._pk()
        m.dt_dev.fill_(1.7976931348623157e308)
        _lib.check(L.pb_estimate_dt(pk, cfg, int(self.has_face_speeds), m.dt_dev.data_ptr(), self._stream()), "pb_estimate_dt")
        self.has_face_speeds = True
        self.launches += 1
        if self.comm is not None:
            _lib.check(L.pb_comm_allreduce_min(self.comm, m.dt_dev.data_ptr(), 1, self._stream()), "pb_comm_allreduce_min")
        elif self.dist is not None:
            self.dist.all_reduce(m.dt_dev, op=self.dist.ReduceOp.MIN)
        return m.p.cfl * float(m.dt_dev.item())

    def history(self, cons=None):
        """the fluid package's .hst columns (ref: src/fluid/fluid.cpp:325-347): dict of maximum density, total baryon
        number, total conserved energy tau, total X1/X2/X3 momentum; all-reduced over ranks"""
        m, L = self.m, self.L
        pk, cfg = self._pk()
        cons = m.cons["base"] if cons is None else cons
        if not hasattr(self, "_hist_partial"):
            self._hist_partial = self.torch.zeros(6 * m.nblocks, dtype=self.torch.float64, device=m.device)
        out = (C.c_double * 6)()
        _lib.check(L.pb_history_fluid(pk, cfg, cons.data_ptr(), self._hist_partial.data_ptr(), out, self._stream()), "pb_history_fluid")
        self.launches += 1
        v = list(out)
        if self.dist is not None:
            t = self.torch.tensor(v, dtype=self.torch.float64, device=m.device)
            mx = t[:1].clone()
            self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(t)
            v = [float(mx[0])] + [float(x) for x in t[1:]]
        keys = ("maximum density", "total baryon number", "total conserved energy tau", "total X1 momentum",
                "total X2 momentum", "total X3 momentum")
        return dict(zip(keys, v))

    def history_torus(self, cons=None, sigma_cutoff=1.0):
        """mdot, jet_energy_flux, jet_momentum_flux, magnetic_Phi at the horizon (ref: src/analysis/history.cpp:25-222)"""
        import math
        m, L = self.m, self.L
        pk, cfg = self._pk()
        cons = m.cons["base"] if cons is None else cons
        if not hasattr(self, "_hist_partial4"):
            self._hist_partial4 = self.torch.zeros(4 * m.nblocks, dtype=self.torch.float64, device=m.device)
        a = m.p.geom_par["a"]
        xh = math.log(1.0 + math.sqrt(1.0 - a * a))  # ref: src/geometry/fmks.cpp:55-57
        out = (C.c_double * 4)()
        _lib.check(L.pb_history_torus(pk, cfg, cons.data_ptr(), xh, sigma_cutoff, self._hist_partial4.data_ptr(), out, self._stream()),
                   "pb_history_torus")
        self.launches += 1
        v = list(out)
        if self.dist is not None:
            t = self.torch.tensor(v, dtype=self.torch.float64, device=m.device)
            self.dist.all_reduce(t)
            v = [float(x) for x in t]
        return dict(zip(("mdot", "jet_energy_flux", "jet_momentum_flux", "magnetic_Phi"), v))

    def step(self, dt=None):
        """One cycle: all RK stages, then the next dt proposal (2x growth limiter)."""
        if dt is None:
            dt = self.dt
            if self.first_step:  # ref: src/phoebus_driver.cpp:121-132
                dt *= self.m.p.dt_init_fact
        self.first_step = False
        for s in range(1, len(self.betas) + 1):
            self.stage(s, dt)
        self.ncycle += 1
        self.time += dt
        return dt

    def propose_dt(self, dt_prev):
        """ref: external/parthenon/src/driver/driver.cpp:216-240"""
        self.dt = min(2.0 * dt_prev, self.estimate_dt())
        return self.dt
