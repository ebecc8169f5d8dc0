"""MeshPack: the MeshBlocks one GPU owns, stored as one device slab per field, plus everything the
C ABI needs to know about them (coordinates, boundary codes, neighbour table, geometry and floor
tables, inter-GPU halo plans).

Host-side mirror of the pieces of Parthenon's Mesh / MeshBlock / MeshData the hot path touches:
  - uniform block grid and block coordinates (ref: external/parthenon/src/mesh/forest/forest.cpp:70-145,
    mesh/forest/tree.cpp:298-320, mesh/forest/logical_location.cpp:61-79, defs.hpp:98-101,
    coordinates/uniform_cartesian.hpp:30-40)
  - domain decomposition by MeshBlock (ref: mesh/mesh-amr_loadbalance.cpp:362-389); here compact bricks
  - field registry of the fluid package (ref: src/fluid/fluid.cpp:139-316)
PyTorch is used for device memory only.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import PB_GEOM_WORDS, PbBlock, PbPack
from .params import BC


def _itsc(index, bloc, nrange):
    """LogicalLocation::IndexToSymmetrizedCoordinate (logical_location.cpp:61-74)"""
    noffset = index - nrange // 2
    noffset_ceil = index - (nrange + 1) // 2
    return float(noffset + noffset_ceil + bloc) / (2.0 * nrange)


def _s2a(u, xmin, xmax):
    """RegionSize::SymmetrizedLogicalToActualPosition (defs.hpp:98-101)"""
    return 0.5 * (xmin + xmax) + (u * xmax - u * xmin)


def rank_grid(world, nblk, ndim):
    """Brick decomposition: factor `world` over the active dimensions, largest factors on the
    dimensions with the most blocks (minimises the halo surface)."""
    grid = [1, 1, 1]
    w = world
    f = 2
    factors = []
    while w > 1:
        while w % f == 0:
            factors.append(f)
            w //= f
        f += 1
    for f in sorted(factors, reverse=True):
        cand = [d for d in range(ndim) if (nblk[d] // grid[d]) % f == 0]
        if not cand:
            raise ValueError("cannot decompose %s blocks over %d ranks" % (nblk, world))
        d = max(cand, key=lambda d: (nblk[d] // grid[d], d))
        grid[d] *= f
    return grid


class MeshPack:
    def __init__(self, params, device=None, rank=0, world=1, mesh_xmin=None, mesh_xmax=None, geom_par=None,
                 alloc_fields=True, use_zflags=True, use_gsrc=True, use_recon_lr=True, use_ghost_plan=True):
        import torch
        self.torch = torch
        self.p = params
        self.rank, self.world = rank, world
        self.device = device
        self.use_zflags, self.use_gsrc, self.use_recon_lr = use_zflags, use_gsrc, use_recon_lr
        self.use_ghost_plan = use_ghost_plan
        p = params
        self.ndim, self.ng = p.ndim, p.ng
        self.nx = tuple(p.block_nx)
        self.mesh_xmin = tuple(mesh_xmin if mesh_xmin is not None else p.xmin)
        self.mesh_xmax = tuple(mesh_xmax if mesh_xmax is not None else p.xmax)
        self.nblk = [max(1, p.mesh_nx[d] // p.block_nx[d]) if d < p.ndim else 1 for d in range(3)]
        for d in range(p.ndim):
            if p.mesh_nx[d] % p.block_nx[d]:
                raise ValueError("Block size is not evenly divisible into the base mesh size.")
        self.ni = self.nx[0] + 2 * self.ng
        self.nj = self.nx[1] + 2 * self.ng if self.ndim >= 2 else 1
        self.nk = self.nx[2] + 2 * self.ng if self.ndim >= 3 else 1
        self.shape = (self.nk, self.nj, self.ni)
        self.ncell = self.ni * self.nj * self.nk
        self.nprim, self.ncons = p.nprim, p.ncons
        self.periodic = [p.mesh_bc[2 * d] == "periodic" for d in range(3)]
        # ---- which blocks are mine
        self.rgrid = rank_grid(world, self.nblk, self.ndim)
        self.brick = [self.nblk[d] // self.rgrid[d] for d in range(3)]
        rc = (rank % self.rgrid[0], (rank // self.rgrid[0]) % self.rgrid[1], rank // (self.rgrid[0] * self.rgrid[1]))
        self.rcoord = rc
        self.lx = [(rc[0] * self.brick[0] + a, rc[1] * self.brick[1] + b, rc[2] * self.brick[2] + c)
                   for c in range(self.brick[2]) for b in range(self.brick[1]) for a in range(self.brick[0])]
        self.nblocks = len(self.lx)
        self.local_index = {l: n for n, l in enumerate(self.lx)}
        # ---- coordinates
        self._coords()
        # ---- metadata, tables
        self.geom_par = geom_par if geom_par is not None else p.geom_par_array()
        self.geom_kind = {"minkowski": 0, "fmks": 1}[p.geometry]
        self._tables()
        self._blocks_meta()
        self.cfg = p.make_cfg(self.gshape)
        if alloc_fields:
            self._alloc()

    # ------------------------------------------------------------------ coordinates
    def owner(self, lx):
        return (lx[0] // self.brick[0]) + self.rgrid[0] * ((lx[1] // self.brick[1]) + self.rgrid[1] * (lx[2] // self.brick[2]))

    def gid(self, lx):
        return lx[0] + self.nblk[0] * (lx[1] + self.nblk[1] * lx[2])

    def _coords(self):
        p = self.p
        active = range(self.ndim)
        pow2 = min((self.nblk[d] & (~(self.nblk[d] - 1))) for d in active)
        ntree = [self.nblk[d] // pow2 if d < self.ndim else 1 for d in range(3)]
        self.bxmin = np.zeros((self.nblocks, 3))
        self.bdx = np.zeros((self.nblocks, 3))
        for n, lx in enumerate(self.lx):
            for d in range(3):
                xmin, xmax = self.mesh_xmin[d], self.mesh_xmax[d]
                if d >= self.ndim:  # symmetry direction: one zone, no ghosts
                    self.bdx[n, d] = (xmax - xmin) / 1
                    self.bxmin[n, d] = xmin - 0 * self.bdx[n, d]
                    continue
                it, il = lx[d] // pow2, lx[d] % pow2
                tmin = _s2a(_itsc(it, 0, ntree[d]), xmin, xmax)
                tmax = _s2a(_itsc(it, 2, ntree[d]), xmin, xmax)
                bmin = _s2a(_itsc(il, 0, pow2), tmin, tmax)
                bmax = _s2a(_itsc(il, 2, pow2), tmin, tmax)
                dx = (bmax - bmin) / self.nx[d]
                self.bdx[n, d] = dx
                self.bxmin[n, d] = bmin - self.ng * dx
        _ = p

    def set_coords(self, n, xmin, dx):
        """override block coordinates (tests feed the reference's own values)"""
        self.bxmin[n] = xmin
        self.bdx[n] = dx

    # ------------------------------------------------------------------ geometry / floor tables
    def _tables(self):
        p = self.p
        L = _lib.lib()
        if self.geom_kind == 0:
            self.gshape = (1, 1, 1)
            key = lambda lx: 0
        elif p.axisymmetric or self.ndim < 3:
            self.gshape = (1, self.nj, self.ni)
            key = lambda lx: (lx[0], lx[1])
        else:
            self.gshape = (self.nk, self.nj, self.ni)
            key = lambda lx: lx
        cfg = p.make_cfg(self.gshape)
        slots, self.geom_slot = {}, []
        tabs = []
        gcell = self.gshape[0] * self.gshape[1] * self.gshape[2]
        par = np.ascontiguousarray(self.geom_par, dtype=np.float64)
        for n, lx in enumerate(self.lx):
            k = key(lx)
            if k not in slots:
                slots[k] = len(tabs)
                t = np.zeros(PB_GEOM_WORDS * gcell)
                xm = (C.c_double * 3)(*self.bxmin[n])
                dx = (C.c_double * 3)(*self.bdx[n])
                _lib.check(L.pb_build_geometry_host(self.geom_kind, par.ctypes.data_as(C.POINTER(C.c_double)),
                                                    C.byref(cfg), xm, dx, t.ctypes.data), "pb_build_geometry_host")
                tabs.append(t)
            self.geom_slot.append(slots[k])
        self.geom_host = np.stack(tabs)
        fslots, self.floor_slot, ftabs = {}, [], []
        for n, lx in enumerate(self.lx):
            k = lx[0]
            if k not in fslots:
                fslots[k] = len(ftabs)
                t = np.zeros(4 * self.ni)
                xm = (C.c_double * 3)(*self.bxmin[n])
                dx = (C.c_double * 3)(*self.bdx[n])
                _lib.check(L.pb_build_floor_table_host(C.byref(cfg), xm, dx, t.ctypes.data), "pb_build_floor_table_host")
                ftabs.append(t)
            self.floor_slot.append(fslots[k])
        self.floor_host = np.stack(ftabs)

    def rebuild_tables(self):
        self._tables()
        self._blocks_meta()
        if hasattr(self, "geom"):
            self._upload_meta()

    # ------------------------------------------------------------------ per-block metadata
    def neighbor(self, lx, off):
        """logical location of the neighbour at offset off, or None at a physical boundary"""
        out = []
        for d in range(3):
            l = lx[d] + off[d]
            if l < 0 or l >= self.nblk[d]:
                if d < self.ndim and self.periodic[d]:
                    l %= self.nblk[d]
                else:
                    return None
            out.append(l)
        return tuple(out)

    def _blocks_meta(self):
        p = self.p
        arr = (PbBlock * self.nblocks)()
        for n, lx in enumerate(self.lx):
            b = arr[n]
            for d in range(3):
                b.xmin[d] = self.bxmin[n, d]
                b.dx[d] = self.bdx[n, d]
            for f in range(6):
                d, outer = f // 2, f % 2
                at_edge = (lx[d] == self.nblk[d] - 1) if outer else (lx[d] == 0)
                if d >= self.ndim:
                    b.bc[f] = BC["periodic"] if self.periodic[d] else 0
                elif not at_edge:
                    b.bc[f] = 0
                else:
                    b.bc[f] = BC[p.mesh_bc[f]]
            b.geom_slot = self.geom_slot[n]
            b.floor_slot = self.floor_slot[n]
            for o3 in (-1, 0, 1):
                for o2 in (-1, 0, 1):
                    for o1 in (-1, 0, 1):
                        code = (o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)
                        off = (o1, o2, o3)
                        if any(off[d] != 0 and d >= self.ndim for d in range(3)) or off == (0, 0, 0):
                            b.nbr[code] = _lib.PB_NBR_NONE
                            continue
                        nl = self.neighbor(lx, off)
                        if nl is None:
                            b.nbr[code] = _lib.PB_NBR_NONE
                        elif self.owner(nl) == self.rank:
                            b.nbr[code] = self.local_index[nl]
                        else:
                            b.nbr[code] = _lib.PB_NBR_REMOTE
        self.blocks_host = arr

    def set_bc(self, n, bc):
        """override the boundary codes of block n (tests feed the reference's flags)"""
        for f in range(6):
            self.blocks_host[n].bc[f] = bc[f]

    def invalidate_flags(self):
        """call after writing p.* or fail from outside the sequencer (pgens, state loads)"""
        self.flags_valid = False

    # ------------------------------------------------------------------ device memory
    def _alloc(self):
        t = self.torch
        dev = self.device
        nb = self.nblocks
        z = lambda *s: t.zeros((nb,) + s + self.shape, dtype=t.float64, device=dev)
        self.prim = z(self.nprim)
        self.temp, self.c2p_mu, self.fail = z(), z(), z()
        self.csig, self.fsig = z(self.ndim), z(self.ndim)
        self.flux = z(self.ndim, self.ncons)
        self.emf = z(3)
        self.cons = {"base": z(self.ncons), "1": z(self.ncons)}
        # accelerators of the sequencer (include/phoebus_b200.h): zone flags, sources container
        self.zflags = t.zeros((nb,) + self.shape, dtype=t.uint8, device=dev) if self.use_zflags else None
        self.gsrc = z(4) if self.use_gsrc else None
        # ql / qr scratch of the two-launch flux path (one direction at a time)
        self.recon_lr = t.empty((2, nb, self.nprim) + self.shape, dtype=t.float64, device=dev) if self.use_recon_lr else None
        self.flags_valid = False
        self.counters = t.zeros(3, dtype=t.int64, device=dev)
        self.dt_dev = t.zeros(1, dtype=t.float64, device=dev)
        self._upload_meta()

    def _upload_meta(self):
        t = self.torch
        raw = np.frombuffer(bytes(self.blocks_host), dtype=np.uint8).copy()
        self.blocks_dev = t.from_numpy(raw).to(self.device)
        self.geom = t.from_numpy(self.geom_host).to(self.device)
        self.floor_tab = t.from_numpy(self.floor_host).to(self.device)
        bcb = [n for n in range(self.nblocks) if any(self.blocks_host[n].bc[f] != 0 for f in range(2 * self.ndim))]
        self.bc_blocks = t.tensor(bcb if bcb else [0], dtype=t.int32, device=self.device)
        self.n_bc_blocks = len(bcb)
        inner = [n for n in range(self.nblocks) if n not in set(bcb)]
        self.inner_blocks = t.tensor(inner if inner else [0], dtype=t.int32, device=self.device)
        self.n_inner_blocks = len(inner)
        pk = PbPack()
        pk.nblocks = self.nblocks
        pk.blocks = self.blocks_dev.data_ptr()
        pk.geom = self.geom.data_ptr()
        pk.floor_tab = self.floor_tab.data_ptr()
        for name in ("prim", "temp", "c2p_mu", "fail", "csig", "fsig", "flux", "emf"):
            setattr(pk, name, getattr(self, name).data_ptr())
        pk.zflags = self.zflags.data_ptr() if self.zflags is not None else None
        pk.gsrc = self.gsrc.data_ptr() if self.gsrc is not None else None
        pk.recon_lr = self.recon_lr.data_ptr() if self.recon_lr is not None else None
        pk.bc_blocks = self.bc_blocks.data_ptr()
        pk.n_bc_blocks = self.n_bc_blocks
        pk.inner_blocks = self.inner_blocks.data_ptr()
        pk.n_inner_blocks = self.n_inner_blocks
        self._ghost_plan(pk)
        self.pack = pk
        self.flags_valid = False

    def ghost_plan_host(self, twins=True):
        """(mask [nblocks*ncell] u8: 1 deferred, 2 twin; ids of the ghost zones that keep their own con2prim u32; ids of the
        twin zones u32) of pb_ghost_plan_host for this pack — host arrays only"""
        L = _lib.lib()
        nint = self.nx[0] * (self.nx[1] if self.ndim >= 2 else 1) * (self.nx[2] if self.ndim >= 3 else 1)
        mask = np.zeros(self.nblocks * self.ncell, dtype=np.uint8)
        zones = np.zeros(max(1, self.nblocks * (self.ncell - nint)), dtype=np.uint32)
        n, nt = C.c_longlong(0), C.c_longlong(0)
        cfg = self.p.make_cfg(self.gshape)
        tw = np.zeros(max(1, int(L.pb_twin_entries(C.byref(cfg), self.nblocks))), dtype=np.uint32)
        _lib.check(L.pb_ghost_plan_host(C.byref(cfg), C.addressof(self.blocks_host), self.nblocks, self.geom_host.ctypes.data,
                                        self.floor_host.ctypes.data, mask.ctypes.data, zones.ctypes.data, C.byref(n),
                                        tw.ctypes.data if twins else None, C.byref(nt) if twins else None),
                   "pb_ghost_plan_host")
        return mask, zones[:n.value].copy(), tw[:nt.value].copy()

    def _ghost_plan(self, pk):
        """which ghost zones take their con2prim result from the same-GPU neighbour instead of solving again, and which are
        solved as twins in a side buffer (pb_ghost_plan_host, csrc/ghost_plan_host.cpp); use_ghost_plan=False leaves the pack
        without a plan, use_ghost_plan="notwins" without the twin part"""
        t = self.torch
        self.ghost_mask = self.c2p_zones = self.twin_zones = None
        self.n_deferred, self.n_twin, self.n_c2p_zones = 0, 0, self.nblocks * self.ncell
        if not self.use_ghost_plan or not self.p.bc_prim:
            return
        import os
        mask, zones, tw = self.ghost_plan_host(twins=self.use_ghost_plan != "notwins" and os.environ.get("PB_GHOST_TWINS", "1") != "0")
        self.n_deferred = int((mask == 1).sum())
        self.n_twin = len(tw)
        if self.n_deferred == 0 and self.n_twin == 0:
            return
        self.n_c2p_zones = self.zones() + len(zones) + len(tw)  # interiors + the listed ghost zones + the twins
        self.ghost_mask = t.from_numpy(mask).to(self.device)
        self.c2p_zones = t.from_numpy(zones if len(zones) else np.zeros(1, dtype=np.uint32)).to(self.device)
        pk.ghost_mask = self.ghost_mask.data_ptr()
        pk.c2p_zones = self.c2p_zones.data_ptr()
        pk.n_c2p_zones = len(zones)
        if self.n_twin:
            S = int(_lib.lib().pb_twin_entries(C.byref(self.p.make_cfg(self.gshape)), self.nblocks))
            self.twin_zones = t.from_numpy(tw).to(self.device)
            self.twin_buf = t.zeros((self.ncons + self.nprim + 3 + self.ndim, S), dtype=t.float64, device=self.device)
            self.twin_flags = t.zeros(S, dtype=t.uint8, device=self.device)
            pk.twin_zones = self.twin_zones.data_ptr()
            pk.n_twin = self.n_twin
            pk.twin_buf = self.twin_buf.data_ptr()
            pk.twin_flags = self.twin_flags.data_ptr()

    def zones(self):
        """interior zones owned by this pack"""
        return self.nblocks * self.nx[0] * (self.nx[1] if self.ndim >= 2 else 1) * (self.nx[2] if self.ndim >= 3 else 1)

    def bytes_allocated(self):
        n = 0
        for name in ("prim", "temp", "c2p_mu", "fail", "csig", "fsig", "flux", "emf", "geom", "floor_tab", "gsrc", "recon_lr"):
            if getattr(self, name) is not None:
                n += getattr(self, name).numel() * 8
        if self.zflags is not None:
            n += self.zflags.numel()
        return n + sum(c.numel() * 8 for c in self.cons.values())

    # ------------------------------------------------------------------ cell-centre coordinates (pgens)
    def cell_centers(self, n):
        """(x1[ni], x2[nj], x3[nk]) of block n, ref: uniform_cartesian.hpp:86-90"""
        out = []
        for d, m in enumerate((self.ni, self.nj, self.nk)):
            out.append(self.bxmin[n, d] + (np.arange(m) + 0.5) * self.bdx[n, d])
        return out

    # ------------------------------------------------------------------ inter-GPU halo plans
    def halo_plans(self):
        """For every peer rank: (send plan, recv plan, doubles in the buffer). Plan rows follow
        include/phoebus_b200.h (pb_halo_pack). Both sides order segments by
        (gid of the receiving block, offset code as seen from the receiver)."""
        if hasattr(self, "_plans"):
            return self._plans
        nvar = self.ncons + (self.nprim + 3 if self.p.bc_prim else 0)
        ng, nx = self.ng, self.nx
        lo = [ng if d < self.ndim else 0 for d in range(3)]

        def ranges(off, ghost):
            r = []
            for d in range(3):
                n = nx[d] if d < self.ndim else 1
                g = ng if d < self.ndim else 0
                if off[d] == 0:
                    r.append((lo[d], lo[d] + n - 1))
                elif ghost:  # ghost slab of the receiver on side off[d]
                    r.append((0, g - 1) if off[d] < 0 else (lo[d] + n, lo[d] + n + g - 1))
                else:  # interior slab of the sender adjacent to the face on side off[d]
                    r.append((lo[d] + n - g, lo[d] + n - 1) if off[d] > 0 else (lo[d], lo[d] + g - 1))
            return r

        send, recv = {}, {}
        offs = [(a, b, c) for c in ((-1, 0, 1) if self.ndim >= 3 else (0,)) for b in ((-1, 0, 1) if self.ndim >= 2 else (0,))
                for a in (-1, 0, 1) if (a, b, c) != (0, 0, 0)]
        code = lambda o: (o[0] + 1) + 3 * (o[1] + 1) + 9 * (o[2] + 1)
        for n, lx in enumerate(self.lx):
            for off in offs:
                nl = self.neighbor(lx, off)
                if nl is None:
                    continue
                peer = self.owner(nl)
                if peer == self.rank:
                    continue
                # I receive into my ghost slab on side `off` from block nl
                recv.setdefault(peer, []).append(((self.gid(lx), code(off)), n, ranges(off, True)))
                # I send my interior slab on side `off`; it lands in nl's ghost slab on side -off
                moff = (-off[0], -off[1], -off[2])
                send.setdefault(peer, []).append(((self.gid(nl), code(moff)), n, ranges(off, False)))
        plans = {}
        for peer in sorted(set(send) | set(recv)):
            out = []
            for lst in (send.get(peer, []), recv.get(peer, [])):
                lst.sort(key=lambda e: e[0])
                rows, offset = [], 0
                for _, n, r in lst:
                    nz = (r[0][1] - r[0][0] + 1) * (r[1][1] - r[1][0] + 1) * (r[2][1] - r[2][0] + 1)
                    rows.append([n, r[0][0], r[0][1], r[1][0], r[1][1], r[2][0], r[2][1], offset & 0xFFFFFFFF, offset >> 32, 0])
                    offset += nz * nvar
                out.append((np.array(rows, dtype=np.int64).astype(np.uint32).view(np.int32).reshape(-1, 10), offset))
            plans[peer] = out
        self._plans = plans
        return plans

    def merged_halo_plans(self):
        """All peers in ONE send and ONE receive buffer: (send plan, send words, recv plan, recv words,
        [(peer, send slice, recv slice)]). Segment offsets of the per-peer plans are shifted by the peer's slice
        start, so that a single pb_halo_pack / pb_halo_unpack launch serves every peer and NCCL moves the slices."""
        def shifted(plan, base):
            rows = plan.view(np.uint32).astype(np.uint64)
            off = (rows[:, 7] | (rows[:, 8] << np.uint64(32))) + np.uint64(base)
            out = plan.copy().view(np.uint32)
            out[:, 7] = (off & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            out[:, 8] = (off >> np.uint64(32)).astype(np.uint32)
            return out.view(np.int32)

        sall, rall, peers, so, ro = [], [], [], 0, 0
        for peer, ((splan, sn), (rplan, rn)) in sorted(self.halo_plans().items()):
            peers.append((peer, slice(so, so + sn), slice(ro, ro + rn)))
            if len(splan):
                sall.append(shifted(splan, so))
            if len(rplan):
                rall.append(shifted(rplan, ro))
            so += sn
            ro += rn
        cat = lambda lst: np.ascontiguousarray(np.concatenate(lst, axis=0)) if lst else np.zeros((0, 10), dtype=np.int32)
        return cat(sall), so, cat(rall), ro, peers

