#!/usr/bin/env python3
"""bench.py — zone-cycles/s of the GRMHD fluid update on the 3-D torus configuration.

  python bench.py --gpus N --steps K --warmup W           (torchrun launches N ranks for N > 1)
  python bench.py --impl reference ...                    (the reference's CPU path on the host cores)

A step is one RK2 cycle (two stages + the dt estimate) of the weak-scaling torus of
BASELINE.json configs[4]: FMKS Kerr-Schild a=0.9375, ideal gas, weno5 + llf, MHD, floors /
ceilings / c2p fixup / flux fixup on, bc_vars = primitive, 256^3 zones per GPU in 32^3 MeshBlocks.
Initial data: the reference's own Fishbone-Moncrief torus (phoebus_b200/pgen.py: init_torus — atmosphere at the floors,
vector-potential B normalised to beta = 100, 5 % energy jitter), generated through the C ABI; the reference arm runs the
same deck on a smaller mesh.
Prints ONE JSON line (contract in the task statement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEAK_MESH = {1: (256, 256, 256), 2: (512, 256, 256), 4: (512, 512, 256), 8: (512, 512, 512)}
STRONG_MESH = (256, 128, 128)  # inputs/torus.pin in 3-D: 128 MeshBlocks of 32^3
# contract-minimum words per zone-stage, 3-D MHD (Nc=8, Nr=10, g=1.953)
A_TASK_BYTES_PER_ZONE_CYCLE = 2368.0
G1_WORDS_PER_ZONE_STAGE = 37.0  # CalculateFluxes: R 10 prims, W 3x8 fluxes, W 3 face speeds
DT_INIT_FACT = 1.0e-2  # both arms: first step = 1e-2 x the CFL step, doubling per cycle (the deck's 1e-6 needs 20 cycles to get there)


def ncu_kernel_metrics(kernel):
    """dram bytes per launch and FP64-pipe utilisation of `kernel` from the committed summary of the ncu --set full capture
    of this same command (profiles/ncu_kernel_metrics.json, written by scripts/ncu_to_json.py); None when there is none"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json")) as f:
            m = json.load(f)
        k = m["kernels"].get(kernel)
        if k:
            return dict(k, source=m.get("source"))
    except (OSError, ValueError, KeyError):
        pass
    return None


_STDOUT_FD = None


def emit(obj):
    """the one JSON line, on the process's real stdout"""
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, line)


def deck_text(name="torus3d"):
    with open(os.path.join(ROOT, "phoebus_b200", "decks", name + ".pin")) as f:
        return f.read()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except (OSError, ValueError):
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def build_mesh(args, rank, world, device):
    from phoebus_b200.mesh import MeshPack
    from phoebus_b200.params import FluidParams, ParameterInput
    from phoebus_b200.pgen import torus_problem_modifier
    nx = tuple(args.mesh) if args.mesh else (STRONG_MESH if args.strong else WEAK_MESH[world])
    ov = ["parthenon/mesh/nx%d=%d" % (d + 1, nx[d]) for d in range(3)]
    ov += ["parthenon/meshblock/nx%d=%d" % (d + 1, args.block) for d in range(3)]
    ov += ["parthenon/time/dt_init_fact=%r" % DT_INIT_FACT]
    pin = ParameterInput(deck_text(), ov)
    torus_problem_modifier(pin)
    fp = FluidParams(pin)
    return MeshPack(fp, device=device, rank=rank, world=world), fp, nx


def init_state(mesh, drv):
    """the reference's start-up sequence for `problem = torus` (Mesh::Initialize + PostInitializationModifier)"""
    from phoebus_b200.pgen import init_torus
    fac = init_torus(mesh, drv)
    drv.ensure_flags()
    mesh.counters.zero_()
    return fac


def e2e_cycles(torch, drv, host_t, dev_t, steps, dt, pipelined=True):
    """`steps` cycles with the state living in pinned HOST buffers between cycles: every cycle uploads its inputs
    (host_t -> dev_t) and downloads its results (dev_t -> host_t). pipelined: PCIe is full duplex, so the download of
    field f of cycle n and the upload of an already-downloaded field for cycle n+1 run on two copy streams, ordered per
    field by events; the cycle itself starts when all of its inputs have arrived. Same bytes, same results as the
    naive upload -> cycle -> download -> synchronize loop (tests/test_gpu_parity.py)."""
    main = torch.cuda.current_stream()
    if pipelined:  # finer pipeline grain: MeshBlock groups of the big fields, so the uploads trail the downloads closely
        hs, ds = [], []
        for h, d in zip(host_t, dev_t):
            k = 8 if d.numel() > (1 << 20) and d.shape[0] >= 8 else 1
            hs += list(h.chunk(k, dim=0))
            ds += list(d.chunk(k, dim=0))
        host_t, dev_t = hs, ds
    if not pipelined:
        for _ in range(steps):
            for h, d in zip(host_t, dev_t):
                d.copy_(h, non_blocking=True)
            drv.step(dt)
            for h, d in zip(host_t, dev_t):
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
        return
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in dev_t]
    ev_out = [torch.cuda.Event() for _ in dev_t]
    ev_done = torch.cuda.Event()
    s_in.wait_stream(main)
    with torch.cuda.stream(s_in):
        for h, d, ev in zip(host_t, dev_t, ev_in):
            d.copy_(h, non_blocking=True)
            ev.record(s_in)
    for n in range(steps):
        for ev in ev_in:
            main.wait_event(ev)
        drv.step(dt)
        ev_done.record(main)
        s_out.wait_event(ev_done)
        with torch.cuda.stream(s_out):
            for h, d, ev in zip(host_t, dev_t, ev_out):
                h.copy_(d, non_blocking=True)
                ev.record(s_out)
        if n + 1 < steps:
            with torch.cuda.stream(s_in):
                for h, d, evo, evi in zip(host_t, dev_t, ev_out, ev_in):
                    s_in.wait_event(evo)
                    d.copy_(h, non_blocking=True)
                    evi.record(s_in)
    main.wait_stream(s_out)
    main.wait_stream(s_in)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from phoebus_b200.driver import FluidDriver
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, "--gpus must equal WORLD_SIZE"
    mesh, fp, nx = build_mesh(args, rank, world, device)
    drv = FluidDriver(mesh, dist if world > 1 else None, halo=args.halo, use_graphs=args.graphs)
    init_state(mesh, drv)
    zones_total = nx[0] * nx[1] * nx[2]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # time-step control as the reference's driver: first step = dt_init_fact x the CFL proposal, then min(2 dt, CFL)
    dt = drv.estimate_dt()
    drv.dt = dt
    drv.first_step = False
    dt *= fp.dt_init_fact
    for _ in range(args.warmup):
        drv.step(dt)
        dt = drv.propose_dt(dt)
    barrier()
    # ---- timed region: exactly K cycles, per-task CUDA events on the launching stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    drv.enable_profile(not args.graphs)  # the per-task events cannot be recorded inside a graph replay
    launches0 = drv.launches
    redo0 = drv.L.pb_arith_redo_count()
    mesh.counters.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        drv.step(dt)
        dt = drv.propose_dt(dt)
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = drv.launches - launches0
    tsum = drv.timer_summary()
    drv.enable_profile(False)
    redo = float(drv.L.pb_arith_redo_count() - redo0)
    counters = mesh.counters.clone()
    # con2prim statistics: one more cycle, outside the timed region, with the counters on
    import ctypes as C
    c2p_stats = (C.c_ulonglong * 36)()
    drv.L.pb_c2p_stats(1, None)
    drv.step(dt)
    dt = drv.propose_dt(dt)
    torch.cuda.synchronize()
    drv.L.pb_c2p_stats(0, c2p_stats)
    c2p_stats = [int(x) for x in c2p_stats]
    nfail = (mesh.fail == 0).sum()
    finite = torch.isfinite(mesh.cons["base"]).all() & torch.isfinite(mesh.prim).all()
    stats = torch.stack([counters[0], counters[1], counters[2], nfail, (~finite).to(torch.int64)]).to(torch.float64)
    stats = torch.cat([stats, torch.tensor([redo], dtype=torch.float64, device=device)])
    if world > 1:
        dist.all_reduce(stats)
    value = zones_total * args.steps / (ms_total * 1e-3)

    # ---- e2e: same cycle through the public API with HOST state buffers (pinned): H2D of the
    # step's inputs and D2H of its results inside the timed region
    e2e = None
    if args.e2e_steps > 0:
        names = ["prim", "temp", "c2p_mu", "fail"]
        dev_t = [getattr(mesh, n) for n in names] + [mesh.cons["base"]]
        host_t = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in dev_t]
        for h, d in zip(host_t, dev_t):
            h.copy_(d)
        nbytes = sum(x.numel() * 8 for x in dev_t)
        barrier()
        e0.record()
        # the duplex pipeline pays on one GPU (6.6 -> 7.6-8.2e7); with several ranks sharing the host's memory and PCIe
        # root it measured slightly slower than the plain loop (N=2: 1.16e8 vs 1.25e8), so those keep the plain loop
        e2e_cycles(torch, drv, host_t, dev_t, args.e2e_steps, dt, pipelined=(world == 1))
        e1.record()
        barrier()
        ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = {"value": zones_total * args.e2e_steps / (float(ms2.item()) * 1e-3), "unit": "zone-cycles/s",
               "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world, "steps": args.e2e_steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, which = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    # dominant kernel. Two-launch flux path (default): recon_kernel, one launch per direction and stage, timed by
    # the in-library CUDA events around exactly those launches. Its algorithmic traffic is what the kernel must
    # move given its task: Nr primitives read, ql + qr (2 Nr words) written per reconstructed cell
    # ((nx+2)^3 cells per MeshBlock). Fused path (PB_FLUX_MODE=fused): flux_kernel, SURVEY 8(d) G1 bytes.
    zones_rank = mesh.zones()
    if "recon" in tsum:
        ncalls, k_ms = tsum["recon"]
        nlaunch = max(1, ncalls)
        cells = mesh.nblocks
        for d in range(3):
            cells *= (mesh.nx[d] + 2) if d < mesh.ndim else 1
        alg_bytes = 3.0 * mesh.nprim * 8.0 * cells
        kname = "recon_kernel<DIR,weno5z,10 prims> (reconstruction, one launch per direction; riemann_kernel follows)"
    else:
        ncalls, k_ms = tsum.get("flux", (0, 0.0))
        nlaunch = max(1, ncalls * mesh.ndim)
        alg_bytes = G1_WORDS_PER_ZONE_STAGE * 8.0 / mesh.ndim * zones_rank  # per launch (one direction)
        kname = "flux_kernel<DIR,weno5z,mhd> (fused reconstruction + Riemann, one launch per direction)"
    avg_ms = k_ms / nlaunch
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    ncu = ncu_kernel_metrics(kname.split("<")[0]) or {}
    out = {
        "metric": "zone-cycles/s", "value": value, "unit": "zone-cycles/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (Fishbone-Moncrief torus generated in HBM)",
        "config": {"workload": "3-D FMKS torus GRMHD, strong scaling, fixed 256x128x128 mesh (BASELINE.json configs[3])" if args.strong
                   else "3-D FMKS torus GRMHD, weak scaling (BASELINE.json configs[4])", "mesh": list(nx),
                   "meshblock": [args.block] * 3, "zones_per_gpu": zones_rank, "blocks_per_gpu": mesh.nblocks,
                   "recon": fp.recon, "riemann": fp.riemann, "integrator": "rk2", "geometry": "fmks a=0.9375 (host-tabulated, axisymmetric)",
                   "initial_data": "Fishbone-Moncrief torus rin=6 rmax=12, atmosphere at the floors, beta=100, u_jitter=0.05 (ref: src/pgen/torus.cpp)",
                   "dt_init_fact": fp.dt_init_fact,
                   "parallelism": "meshblock bricks %s" % "x".join(map(str, mesh.rgrid)),
                   "halo_transport": ("ncclSend/ncclRecv inside the C ABI" if args.halo == "nccl" else "torch.distributed") if world > 1 else None,
                   "cuda_graphs": bool(args.graphs),
                   # ghost zones whose con2prim result is copied from the same-GPU neighbour (identical inputs and tables)
                   "ghost_plan": {"deferred_zones": int(getattr(mesh, "n_deferred", 0)), "twin_zones": int(getattr(mesh, "n_twin", 0)),
                                  "con2prim_zones": int(mesh.n_c2p_zones),
                                  "all_zones": int(mesh.nblocks * mesh.ncell)} if mesh.ghost_mask is not None else None,
                   "l2": "working set %.1f GB per GPU >> 126 MB L2" % (mesh.bytes_allocated() / 1e9)},
        "roofline": {"bound": "hbm", "kernel": kname, "co_bound": "fp64 pipe (no FMA contraction: bit parity with the reference)",
                     "fp64_pipe_pct_ncu": ncu.get("fp64_pipe_pct"), "traffic_source": ncu.get("source"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": which,
                     "traffic": ncu.get("dram_bytes_per_launch"), "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                     "step_frac_of_hbm_roofline": value / world * A_TASK_BYTES_PER_ZONE_CYCLE / (peak * 1e9)},
        "breakdown_ms_per_step": {k: v[1] / args.steps for k, v in sorted(tsum.items())},
        "gpu_launches": launches,
        "zone_stats": {"floors": int(stats[0]), "ceilings": int(stats[1]), "floor_c2p_fail": int(stats[2]),
                       "failed_zones_now": int(stats[3]), "nonfinite": int(stats[4]),
                       "arith_redo_items": int(stats[5]),
                       # rank 0's con2prim statistics of ONE cycle run after the timed region (the reference's CON2PRIM_STATISTICS, con2prim_statistics.hpp)
                       "c2p_iteration_histogram": c2p_stats[:32], "c2p_zones_solved": c2p_stats[35],
                       "c2p_used_floor_or_ceiling": c2p_stats[32], "c2p_failed": c2p_stats[33],
                       "c2p_parked_by_warp_vote": c2p_stats[34],
                       "c2p_mean_iterations": (sum(i * n for i, n in enumerate(c2p_stats[:32])) / max(1, sum(c2p_stats[:32])))},
        "clocks": sampler.summary(),
    }
    if e2e:
        out["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def ref_binary():
    exe = os.path.join(ROOT, "oracle", "_ref", "phoebus_fmks")
    deck = os.path.join(ROOT, "oracle", "_ref", "decks", "torus.pin")
    return (exe, deck) if os.path.exists(exe) and os.path.exists(deck) else None


def run_reference_binary(nx, block, ncycles, threads):
    """the UNMODIFIED reference (Kokkos OpenMP build, oracle/_ref/phoebus_fmks) on a bounded torus"""
    exe, deck = ref_binary()
    ov = ["parthenon/mesh/nx%d=%d" % (d + 1, nx[d]) for d in range(3)]
    ov += ["parthenon/meshblock/nx%d=%d" % (d + 1, block) for d in range(3)]
    ov += ["parthenon/time/nlim=%d" % (ncycles + 2), "parthenon/time/perf_cycle_offset=2", "torus/seed=7",
           "parthenon/time/dt_init_fact=%r" % DT_INIT_FACT]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread", OMP_PLACES="threads")
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe, "-i", deck] + ov, cwd=tmp, env=env, capture_output=True, text=True, timeout=1500)
    for line in r.stdout.splitlines():
        if "zone-cycles/wallsecond" in line:
            return float(line.split("=")[1])
    raise RuntimeError("reference run produced no zone-cycles line:\n" + r.stdout[-2000:] + r.stderr[-2000:])


def run_oracle_port(nx, block, ncycles, threads):
    """the CPU oracle (port of the reference's algorithm), blocks spread over host threads"""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    import ctypes as C
    from oracle import pyoracle as po
    from phoebus_b200.mesh import MeshPack
    from phoebus_b200.params import FluidParams, ParameterInput
    from phoebus_b200.pgen import torus_problem_modifier
    ov = ["parthenon/mesh/nx%d=%d" % (d + 1, nx[d]) for d in range(3)] + ["parthenon/meshblock/nx%d=%d" % (d + 1, block) for d in range(3)]
    pin = ParameterInput(deck_text(), ov)
    torus_problem_modifier(pin)
    fp = FluidParams(pin)
    mesh = MeshPack(fp, alloc_fields=False)
    from oracle import deck as odeck
    p = odeck.params_from_deck(pin.blocks)
    L = po.lib()
    blocks = []
    rng = np.random.default_rng(0)
    for n, lx in enumerate(mesh.lx):
        bc = list(mesh.blocks_host[n].bc)
        geom = mesh.geom_host[mesh.geom_slot[n]].reshape((-1,) + mesh.gshape)
        b = po.Block(p, mesh.bxmin[n], mesh.bdx[n], bc, lx, geom, mesh.gshape)
        x1, x2, x3 = mesh.cell_centers(n)
        r = np.exp(x1)[None, None, :]
        s = np.abs(np.sin(np.pi * x2))[None, :, None]
        tor = np.exp(-((r - 12.0) / 4.0) ** 2) * s ** 8 * np.ones_like(x3)[:, None, None]
        b.prim[0] = 1e-5 * r ** -1.5 + tor
        b.prim[3] = 0.5 * tor * r ** -1.5
        b.prim[4] = 0.01 * b.prim[0] ** fp.gamma / (fp.gamma - 1) + 1e-7 * r ** -2.5
        b.prim[5] = 2e-3 * tor / r
        b.prim[8] = (fp.gamma - 1) * b.prim[4]
        b.prim[9] = fp.gamma
        L.orc_prim_to_cons(C.byref(b.cfg), b.geom, b.prim, b.cons["base"], b.csig, *b.entire())
        blocks.append(b)
    om = po.OracleMesh(p, blocks, mesh.nblk, mesh.periodic)
    pool = ThreadPoolExecutor(threads)

    def stage(stage, dt):  # OracleMesh.stage with the per-block loops spread over threads
        names, beta = ["base", "1", "base"], [1.0, 0.5][stage - 1]
        sc0, sc1 = names[stage - 1], names[stage]
        cnt = np.zeros(3, dtype=np.int64)

        def pre(b):
            c = C.byref(b.cfg)
            L.orc_source_terms(c, b.geom, b.prim, b.gsrc)
            L.orc_calculate_fluxes(c, b.geom, b.prim, b.flux, b.fsig)
            L.orc_flux_ct(c, b.flux, b.emf)
            L.orc_fix_fluxes(c, b.flux)
            L.orc_flux_divergence(c, b.flux, b.dudt)
            u0 = b.cons[sc0]
            L.orc_rk_update(c, b.dudt, b.gsrc.ctypes.data, u0, b.cons["base"] if sc0 != "base" else u0.copy(), beta, beta * dt, b.cons[sc1])
            L.orc_c2p_fixup(c, b.geom, b.cons[sc1], b.prim, b.temp, b.fail, b.csig)
            L.orc_apply_floors(c, b.geom, b.cons[sc1], b.prim, b.temp, b.c2p_mu, b.csig, None)

        def post(b):
            c = C.byref(b.cfg)
            out = b.cons[sc1]
            L.orc_apply_bcs(c, b.geom, out, b.prim, b.temp, b.c2p_mu, b.fail, b.csig)
            L.orc_cons_to_prim(c, b.geom, out, b.prim, b.temp, b.c2p_mu, b.fail, b.csig, *b.entire())
            L.orc_c2p_fixup(c, b.geom, out, b.prim, b.temp, b.fail, b.csig)
            L.orc_apply_floors(c, b.geom, out, b.prim, b.temp, b.c2p_mu, b.csig, None)
        list(pool.map(pre, blocks))
        om.exchange(sc1)
        list(pool.map(post, blocks))
        _ = cnt
    for b in blocks:  # initial FillDerived
        L.orc_cons_to_prim(C.byref(b.cfg), b.geom, b.cons["base"], b.prim, b.temp, b.c2p_mu, b.fail, b.csig, *b.entire())
    dt = 1e-3 * fp.cfl * min(L.orc_estimate_dt(C.byref(b.cfg), b.csig, b.fsig, 0) for b in blocks)
    stage(1, dt); stage(2, dt)  # warm-up cycle
    t0 = time.time()
    for _ in range(ncycles):
        stage(1, dt)
        stage(2, dt)
    wall = time.time() - t0
    _ = rng
    return nx[0] * nx[1] * nx[2] * ncycles / wall


def cpu_baseline(args, steps=None):
    threads = os.cpu_count() or 1
    nx, block = (128, 64, 64), 32
    ncycles = steps or 24  # ~0.33 s per cycle on 16 cores: a 10-20 s sample after the 2 untimed warm-up cycles
    sample = "3-D FMKS torus %dx%dx%d (%d MeshBlocks of %d^3), %d RK2 cycles after 2 warm-up cycles" % (
        nx[0], nx[1], nx[2], (nx[0] // block) * (nx[1] // block) * (nx[2] // block), block, ncycles)
    if ref_binary() and not args.force_port:
        try:
            v = run_reference_binary(nx, block, ncycles, threads)
            return {"value": v, "unit": "zone-cycles/s", "cores": threads, "kind": "reference",
                    "sample": sample + "; unmodified lanl/phoebus, Kokkos OpenMP, g++ -O3, Fishbone-Moncrief IC"}
        except Exception as e:  # fall through to the port, but say why
            sample += " [reference binary failed: %s]" % str(e)[:120]
    nx = (64, 64, 32)
    v = run_oracle_port(nx, block, max(1, min(ncycles, 3)), threads)
    return {"value": v, "unit": "zone-cycles/s", "cores": threads, "kind": "port",
            "sample": "oracle port, synthetic torus %dx%dx%d, 32^3 MeshBlocks spread over host threads" % nx}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on this box's host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t0 = time.time()
    cb = cpu_baseline(args, steps=max(8, min(8 * args.steps, 48)))  # one 'step' = 8 RK2 cycles of the bounded sample
    nx = (128, 64, 64) if cb["kind"] == "reference" else (64, 64, 32)
    out = {"impl": "reference", "metric": "zone-cycles/s", "value": cb["value"], "unit": "zone-cycles/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * nx[0] * nx[1] * nx[2] / cb["value"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic (Fishbone-Moncrief torus generated in HBM)",
           "config": {"workload": "3-D FMKS torus GRMHD (bounded CPU sample of BASELINE.json configs[4])", "mesh": list(nx),
                      "meshblock": [32] * 3},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "zone-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.time() - t0}
    emit(out)


def main():
    # exactly ONE line on stdout: libraries that write there on their own (NCCL prints its version banner at communicator
    # creation when NCCL_DEBUG is set) are sent to stderr until the JSON line is printed
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", type=int, nargs=3, default=None, help="override the mesh size (zones)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the fixed 256x128x128 torus of inputs/torus.pin (BASELINE.json configs[3]) on N GPUs")
    ap.add_argument("--block", type=int, default=32)
    ap.add_argument("--graphs", action="store_true", help="replay every RK stage as one CUDA graph (no per-task breakdown then)")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "torch"],
                    help="inter-GPU ghost transport: NCCL inside the C ABI (default) or torch.distributed")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-port", action="store_true", help="cpu baseline: use the oracle port even if oracle/_ref exists")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(_STDOUT_FD, 1)


if __name__ == "__main__":
    main()
