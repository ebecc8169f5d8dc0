"""Worker for the multi-rank tests (launched by torch.distributed.run).

GPU mode (nccl): every rank advances its brick of a small 3-D torus for a few RK2 cycles with NCCL
halos and ALSO advances the whole mesh alone on its own GPU; the brick's blocks must equal the
corresponding blocks of the single-GPU run bit for bit (domain decomposition does not change any
zone's arithmetic on a uniform mesh).

CPU mode (gloo): the same halo plans drive a numpy pack/unpack (mirror of pb_halo_pack/unpack, test
infrastructure) and the ghost zones must equal a single-process ghost fill.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def make_params(mesh_nx, block, periodic):
    import bench
    from phoebus_b200.params import FluidParams, ParameterInput
    from phoebus_b200.pgen import torus_problem_modifier
    ov = ["parthenon/mesh/nx%d=%d" % (d + 1, mesh_nx[d]) for d in range(3)]
    ov += ["parthenon/meshblock/nx%d=%d" % (d + 1, block) for d in range(3)]
    if periodic:
        for d in (1, 2):
            ov += ["parthenon/mesh/ix%d_bc=periodic" % d, "parthenon/mesh/ox%d_bc=periodic" % d]
        ov += ["fixup/enable_ox1_fmks_inflow_check=false", "fixup/enable_flux_fixup=false"]
    pin = ParameterInput(bench.deck_text(), ov)
    torus_problem_modifier(pin)
    return FluidParams(pin)


def gpu_mode(periodic):
    import bench
    from phoebus_b200.driver import FluidDriver
    from phoebus_b200.mesh import MeshPack
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    # 8x4x4 MeshBlocks of 16^3: every rank owns blocks away from the mesh boundary, so the con2prim that runs
    # under the NCCL transfer (pb_rk_stage_post_early) is exercised
    fp = make_params((128, 64, 64), 16, periodic)
    whole = MeshPack(fp, device=dev, rank=0, world=1)
    d_whole = FluidDriver(whole)
    bench.init_state(whole, d_whole)
    dt0 = 0.05 * d_whole.estimate_dt()
    dts = [dt0]
    for cyc in range(3):
        d_whole.step(dts[-1])
        dts.append(d_whole.propose_dt(dts[-1]))
    torch.cuda.synchronize()
    bad = 0
    # the brick of this rank, three ways: NCCL through the C ABI (pb_halo_exchange_begin / _end), the same with every stage
    # replayed as a CUDA graph, and torch.distributed as the transport
    for halo, graphs in (("nccl", False), ("nccl", True), ("torch", False)):
        part = MeshPack(fp, device=dev, rank=rank, world=world)
        d_part = FluidDriver(part, dist, halo=halo, use_graphs=graphs)
        bench.init_state(part, d_part)
        dt = 0.05 * d_part.estimate_dt()
        assert dt == dt0, (dt, dt0)
        for cyc in range(3):
            d_part.step(dt)
            dt = d_part.propose_dt(dt)
            assert dt == dts[cyc + 1], (halo, graphs, cyc, dt, dts[cyc + 1])
        torch.cuda.synchronize()
        for n, lx in enumerate(part.lx):
            m = whole.local_index[lx]
            for name in ("prim", "temp", "c2p_mu", "fail", "csig"):
                if not torch.equal(getattr(part, name)[n], getattr(whole, name)[m]):
                    bad += 1
                    print("rank", rank, halo, graphs, "block", lx, name, "differs", flush=True)
            if not torch.equal(part.cons["base"][n], whole.cons["base"][m]):
                bad += 1
                print("rank", rank, halo, graphs, "block", lx, "cons differs", flush=True)
        cnt = part.counters.clone()
        dist.all_reduce(cnt)
        if not torch.equal(cnt, whole.counters):
            bad += 1
            print("rank", rank, halo, graphs, "floor counters differ", cnt.tolist(), whole.counters.tolist(), flush=True)
        d_part.close()
        nblocks, rgrid = part.nblocks, part.rgrid
        del part, d_part
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("MP_RESULT bad=%d blocks_per_rank=%d grid=%s variants=nccl_c,nccl_c+graphs,torch" % (int(t.item()), nblocks, rgrid), flush=True)
    return int(t.item())


def np_pack(fields, plan, buf):
    for row in plan:
        b, i0, i1, j0, j1, k0, k1 = [int(x) for x in row[:7]]
        off = (int(row[7]) & 0xFFFFFFFF) | ((int(row[8]) & 0xFFFFFFFF) << 32)
        nz = (i1 - i0 + 1) * (j1 - j0 + 1) * (k1 - k0 + 1)
        v = 0
        for f in fields:
            for n in range(f.shape[1]):
                buf[off + v * nz: off + (v + 1) * nz] = f[b, n, k0:k1 + 1, j0:j1 + 1, i0:i1 + 1].reshape(-1)
                v += 1


def np_unpack(fields, plan, buf):
    for row in plan:
        b, i0, i1, j0, j1, k0, k1 = [int(x) for x in row[:7]]
        off = (int(row[7]) & 0xFFFFFFFF) | ((int(row[8]) & 0xFFFFFFFF) << 32)
        shape = (k1 - k0 + 1, j1 - j0 + 1, i1 - i0 + 1)
        nz = shape[0] * shape[1] * shape[2]
        v = 0
        for f in fields:
            for n in range(f.shape[1]):
                f[b, n, k0:k1 + 1, j0:j1 + 1, i0:i1 + 1] = buf[off + v * nz: off + (v + 1) * nz].reshape(shape)
                v += 1


def cpu_mode(periodic):
    """gloo: plan-driven exchange between ranks == single-process ghost fill"""
    from oracle.pyoracle import halo_exchange
    from phoebus_b200.mesh import MeshPack
    rank, world = dist.get_rank(), dist.get_world_size()
    fp = make_params((32, 16, 16), 8, periodic)
    part = MeshPack(fp, rank=rank, world=world, alloc_fields=False)
    whole = MeshPack(fp, rank=0, world=1, alloc_fields=False)
    rng = np.random.default_rng(1234)
    nvar = [fp.ncons, fp.nprim, 1, 1, 1] if fp.bc_prim else [fp.ncons]
    full = [rng.standard_normal((whole.nblocks, nv) + whole.shape) for nv in nvar]
    mine = [np.stack([f[whole.local_index[lx]] for lx in part.lx]).copy() for f in full]

    class B:  # minimal block view for the oracle's halo_exchange
        def __init__(self, lx, arrs, p):
            self.lx, self.arrs = lx, arrs
            self.p = dict(nx=p.block_nx)
    blocks = [B(lx, [f[n] for f in full], fp) for n, lx in enumerate(whole.lx)]
    halo_exchange(blocks, whole.nblk, whole.periodic, [lambda b, q=q: b.arrs[q] for q in range(len(full))], fp.ng, fp.ndim)
    # local neighbours of the partial mesh: same routine restricted to my blocks would need the
    # remote ones; emulate pb_halo_exchange_local with the neighbour table instead
    ng, nx = part.ng, part.nx
    for n, lx in enumerate(part.lx):
        for code in range(27):
            nb = part.blocks_host[n].nbr[code]
            if nb < 0:
                continue
            o = (code % 3 - 1, (code // 3) % 3 - 1, code // 9 - 1)
            dst, src = [], []
            for d in range(3):
                if o[d] == 0:
                    dst.append(slice(ng, ng + nx[d])); src.append(slice(ng, ng + nx[d]))
                elif o[d] < 0:
                    dst.append(slice(0, ng)); src.append(slice(nx[d], nx[d] + ng))
                else:
                    dst.append(slice(ng + nx[d], nx[d] + 2 * ng)); src.append(slice(ng, 2 * ng))
            for f in mine:
                f[n][:, dst[2], dst[1], dst[0]] = f[nb][:, src[2], src[1], src[0]]
    # one send / one receive buffer for all peers, one pack and one unpack (what FluidDriver does on the GPU)
    splan, stot, rplan, rtot, peers = part.merged_halo_plans()
    sbuf, rbuf = np.zeros(max(stot, 1)), np.zeros(max(rtot, 1))
    np_pack(mine, splan, sbuf)
    st, rt = torch.from_numpy(sbuf), torch.from_numpy(rbuf)
    ops = []
    for peer, ss, rs in peers:
        if ss.stop > ss.start:
            ops.append(dist.P2POp(dist.isend, st[ss], peer))
        if rs.stop > rs.start:
            ops.append(dist.P2POp(dist.irecv, rt[rs], peer))
    for r in (dist.batch_isend_irecv(ops) if ops else []):
        r.wait()
    np_unpack(mine, rplan, rbuf)
    bad = 0
    for n, lx in enumerate(part.lx):
        m = whole.local_index[lx]
        for f, g in zip(mine, full):
            if not np.array_equal(f[n], g[m]):
                bad += 1
    t = torch.tensor([bad])
    dist.all_reduce(t)
    if rank == 0:
        print("MP_RESULT bad=%d blocks_per_rank=%d grid=%s" % (int(t.item()), part.nblocks, part.rgrid), flush=True)
    return int(t.item())


if __name__ == "__main__":
    mode, periodic = sys.argv[1], sys.argv[2] == "periodic"
    if mode == "gpu":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        rc = gpu_mode(periodic)
    else:
        dist.init_process_group("gloo")
        rc = cpu_mode(periodic)
    dist.destroy_process_group()
    sys.exit(1 if rc else 0)
