"""The C++ host layer (phoebus_b200/host): Parthenon-shaped task functions above the C ABI.

CPU part: the library exports the reference's task names, and its deck reader / package Initialize
functions / Mesh metadata agree bit for bit with the Python mirror (phoebus_b200.params, .mesh).
GPU part: PhoebusDriver::RungeKuttaStage — fused (pb_rk_stage) and task-by-task in the reference's
order, per-block tasks on MeshBlockData like the reference's driver — reproduces the reference's
per-stage dumps bit for bit.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import from_dump as fd
from tests.helpers import CASES, ROOT, getter, load_case, load_product_state, product_mesh, stage_tags

HOST = os.path.join(ROOT, "phoebus_b200", "host")
EXE = os.path.join(HOST, "phoebus_host_stage")
LIBHOST = os.path.join(ROOT, "phoebus_b200", "libphoebus_host.so")
STATE_FIELDS = ("prim", "temp", "c2p_mu", "fail", "csig", "fsig")

# the reference's task functions a Phoebus build would link against (demangled prefixes)
TASKS = [
    "fluid::Initialize(parthenon::ParameterInput*)",
    "fluid::CalculateFluxes(parthenon::MeshBlockData<double>*)",
    "fluid::FluxCT(parthenon::MeshBlockData<double>*)",
    "fluid::CalculateFluidSourceTerms(parthenon::MeshData<double>*, parthenon::MeshData<double>*)",
    "fluid::PrimitiveToConserved(parthenon::MeshBlockData<double>*)",
    "fluid::PrimitiveToConservedRegion(parthenon::MeshBlockData<double>*, parthenon::IndexRange const&, parthenon::IndexRange const&, parthenon::IndexRange const&)",
    "parthenon::TaskStatus fluid::ConservedToPrimitive<parthenon::MeshBlockData<double> >(parthenon::MeshBlockData<double>*)",
    "parthenon::TaskStatus fluid::ConservedToPrimitive<parthenon::MeshData<double> >(parthenon::MeshData<double>*)",
    "parthenon::TaskStatus fluid::ConservedToPrimitiveRegion<parthenon::MeshBlockData<double> >(",
    "parthenon::TaskStatus fluid::ConservedToPrimitiveRobust<parthenon::MeshData<double> >(",
    "fluid::EstimateTimestepBlock(parthenon::MeshBlockData<double>*)",
    "fluid::CalculateDivB(parthenon::MeshBlockData<double>*)",
    "fixup::Initialize(parthenon::ParameterInput*)",
    "fixup::FixFluxes(parthenon::MeshBlockData<double>*)",
    "parthenon::TaskStatus fixup::ApplyFloors<parthenon::MeshData<double> >(parthenon::MeshData<double>*)",
    "parthenon::TaskStatus fixup::ApplyFloors<parthenon::MeshBlockData<double> >(parthenon::MeshBlockData<double>*)",
    "parthenon::TaskStatus fixup::ConservedToPrimitiveFixup<parthenon::MeshData<double> >(parthenon::MeshData<double>*)",
    "parthenon::TaskStatus fixup::ConservedToPrimitiveFixup<parthenon::MeshBlockData<double> >(parthenon::MeshBlockData<double>*)",
    "Boundaries::ConvertBoundaryConditions(std::shared_ptr<parthenon::MeshBlockData<double> >&)",
    "parthenon::TaskStatus parthenon::Update::FluxDivergence<parthenon::MeshData<double> >(parthenon::MeshData<double>*, parthenon::MeshData<double>*)",
    "parthenon::TaskStatus parthenon::Update::SumData<",
    "parthenon::TaskStatus parthenon::Update::AverageIndependentData<parthenon::MeshData<double> >(",
    "parthenon::TaskStatus parthenon::Update::UpdateIndependentData<parthenon::MeshData<double> >(",
    "parthenon::TaskStatus parthenon::Update::FillDerived<parthenon::MeshBlockData<double> >(",
    "parthenon::SendBoundBufs(std::shared_ptr<parthenon::MeshData<double> >&)",
    "phoebus::PhoebusDriver::RungeKuttaStage(int)",
]


def _deck_text(pin):
    out = []
    for blk, kv in pin.blocks.items():
        out.append("<%s>" % blk)
        out += ["%s = %s" % (k, v) for k, v in kv.items()]
        out.append("")
    return "\n".join(out)


def _case_deck(name, tmp_path):
    from tests.helpers import product_params
    from phoebus_b200.params import ParameterInput
    g, _ = load_case(name)
    fp = product_params(g)
    # product_params mutates a ParameterInput built from the fixture; rebuild the same blocks here
    pin = ParameterInput.from_blocks(g.meta["deck_blocks"])
    man = g.manifest("init")
    mesh = pin.blocks.setdefault("parthenon/mesh", {})
    for k, v in man["mesh"].items():
        mesh[k] = repr(float(v))
    pin.blocks.setdefault("geometry", {})["kind"] = "fmks" if g.meta["geometry"] == 1 else "minkowski"
    path = os.path.join(str(tmp_path), name + ".pin")
    with open(path, "w") as f:
        f.write(_deck_text(pin))
    return g, fp, path


def test_host_library_is_built_and_exports_the_reference_task_names():
    assert os.path.exists(LIBHOST) and os.path.exists(EXE), "build with `make -C phoebus_b200/host`"
    syms = subprocess.run(["nm", "-DC", "--defined-only", LIBHOST], capture_output=True, text=True, check=True).stdout
    missing = [t for t in TASKS if t not in syms]
    assert not missing, missing


@pytest.mark.parametrize("name", CASES)
def test_host_cfg_and_mesh_metadata_match_python_mirror(name, tmp_path):
    from phoebus_b200._lib import PbBlock
    g, fp, deck = _case_deck(name, tmp_path)
    mesh, _ = product_mesh(g, alloc_fields=False)
    tab = os.path.join(str(tmp_path), "tables.bin")
    r = subprocess.run([EXE, "describe", deck, tab], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.splitlines() if ln.strip()}
    cfg = mesh.cfg
    for key, _ctype in cfg._fields_:
        v = getattr(cfg, key)
        if key == "nx":
            assert [int(x) for x in got["nx"]] == list(v)
        elif isinstance(v, float):
            assert float(got[key][0]) == v, key
        else:
            assert int(got[key][0]) == v, key
    assert int(got["nblocks"][0]) == mesh.nblocks
    raw = open(tab, "rb").read()
    nb = C.sizeof(PbBlock) * mesh.nblocks
    assert raw[:nb] == bytes(mesh.blocks_host), "block table (coordinates, bc codes, slots, neighbours)"
    rest = np.frombuffer(raw[nb:], dtype=np.float64)
    ng = mesh.geom_host.size
    assert np.array_equal(rest[:ng], mesh.geom_host.ravel(), equal_nan=True), "geometry tables"
    assert np.array_equal(rest[ng:], mesh.floor_host.ravel(), equal_nan=True), "floor tables"


def test_host_rejects_what_the_reference_rejects(tmp_path):
    deck = os.path.join(str(tmp_path), "bad.pin")
    open(deck, "w").write("<parthenon/mesh>\nnx1 = 64\nnghost = 2\n<fluid>\nrecon = weno5\n")
    r = subprocess.run([EXE, "describe", deck, os.devnull], capture_output=True, text=True)
    assert r.returncode != 0 and "requires 4+ ghost cells" in r.stderr
    r = subprocess.run([EXE, "describe", deck, os.devnull, "fluid/recon=ppm", "parthenon/mesh/nghost=4"], capture_output=True, text=True)
    assert r.returncode != 0 and "Reconstruction method invalid" in r.stderr
    r = subprocess.run([EXE, "describe", deck, os.devnull, "fluid/recon=linear", "parthenon/meshblock/nx1=48"], capture_output=True, text=True)
    assert r.returncode != 0 and "not evenly divisible" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "tasks"])
@pytest.mark.parametrize("name", CASES)
def test_host_driver_matches_reference_dumps(name, mode, tmp_path):
    import torch
    assert torch.cuda.is_available()
    g, fp, deck = _case_deck(name, tmp_path)
    mesh, _ = product_mesh(g, device=torch.device("cuda:0"))
    order = load_product_state(mesh, g, "init")
    fields = [getattr(mesh, k) for k in STATE_FIELDS] + [mesh.cons["base"], mesh.cons["1"]]
    sin, sout = os.path.join(str(tmp_path), "in.bin"), os.path.join(str(tmp_path), "out.bin")
    with open(sin, "wb") as f:
        for t in fields:
            f.write(t.cpu().numpy().tobytes())
    tags = stage_tags(g)
    for ntag in range(1, len(tags) + 1):
        stages = ",".join("%d:%r" % (g.manifest(t)["stage"], float(g.manifest(t)["dt"])) for t in tags[:ntag])
        r = subprocess.run([EXE, "run", deck, sin, sout, mode, stages], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        raw = np.fromfile(sout, dtype=np.float64)
        off, out = 0, {}
        for k, t in zip(STATE_FIELDS + ("base", "1"), fields):
            out[k] = raw[off:off + t.numel()].reshape(tuple(t.shape))
            off += t.numel()
        tag = tags[ntag - 1]
        man = g.manifest(tag)
        for mb in man["blocks"]:
            n = order[mb["gid"]]
            ref = fd.ref_state(getter(g, tag), mb["gid"], mesh.shape, man["sc1"])
            for k in ("prim", "temp", "c2p_mu", "fail", "csig"):
                assert np.array_equal(out[k][n].reshape(ref[k].shape), ref[k], equal_nan=True), (name, mode, tag, k)
            assert np.array_equal(out[man["sc1"]][n], ref["cons"], equal_nan=True), (name, mode, tag, "cons")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "tasks"])
def test_host_driver_long_run_with_its_own_dt_control(mode, tmp_path):
    """phoebus_host_stage evolve: PhoebusDriver::Step (EstimateTimestep, 2x limiter, tlim clip) for one full wave
    period from the reference's init dump must land on the reference's final dump bit for bit"""
    import torch
    name = "linmodes2d_long"
    g, fp, deck = _case_deck(name, tmp_path)
    mesh, _ = product_mesh(g, device=torch.device("cuda:0"))
    order = load_product_state(mesh, g, "init")
    fields = [getattr(mesh, k) for k in STATE_FIELDS] + [mesh.cons["base"], mesh.cons["1"]]
    sin, sout = os.path.join(str(tmp_path), "in.bin"), os.path.join(str(tmp_path), "out.bin")
    with open(sin, "wb") as f:
        for t in fields:
            f.write(t.cpu().numpy().tobytes())
    tags = stage_tags(g)
    last = g.manifest(tags[-1])
    ncycles = last["cycle"] + 1
    tlim = last["time"] + last["dt"]
    r = subprocess.run([EXE, "evolve", deck, sin, sout, mode, str(ncycles), repr(float(g.manifest("init")["dt"])),
                        "parthenon/time/dt_init_fact=1.0", "parthenon/time/tlim=%r" % float(tlim)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = np.fromfile(sout, dtype=np.float64)
    off, out = 0, {}
    for k, t in zip(STATE_FIELDS + ("base", "1"), fields):
        out[k] = raw[off:off + t.numel()].reshape(tuple(t.shape))
        off += t.numel()
    for mb in last["blocks"]:
        n = order[mb["gid"]]
        ref = fd.ref_state(getter(g, tags[-1]), mb["gid"], mesh.shape, last["sc1"])
        for k in ("prim", "temp", "c2p_mu", "fail", "csig"):
            assert np.array_equal(out[k][n].reshape(ref[k].shape), ref[k], equal_nan=True), (mode, k)
        assert np.array_equal(out[last["sc1"]][n], ref["cons"], equal_nan=True), (mode, "cons")


# ---------------------------------------------------------------------------------------------------------------
# N GPUs: MeshBlock bricks, inter-GPU halo plans and the NCCL exchange of the C library driven by the C++ host layer
def _brick_deck(tmp_path, mesh_nx=(64, 32, 32), block=16):
    import bench
    from phoebus_b200.params import FluidParams, ParameterInput
    from phoebus_b200.pgen import torus_problem_modifier
    ov = ["parthenon/mesh/nx%d=%d" % (d + 1, mesh_nx[d]) for d in range(3)]
    ov += ["parthenon/meshblock/nx%d=%d" % (d + 1, block) for d in range(3)]
    pin = ParameterInput(bench.deck_text(), ov)
    torus_problem_modifier(pin)
    fp = FluidParams(pin)
    path = os.path.join(str(tmp_path), "brick.pin")
    with open(path, "w") as f:
        f.write(_deck_text(pin))
    return fp, path


@pytest.mark.parametrize("world", [2, 4, 8])
def test_host_brick_decomposition_and_halo_plans_match_python_mirror(world, tmp_path):
    """every rank's block table (REMOTE neighbour markers included) and its merged inter-GPU halo plans, C++ vs Python"""
    from phoebus_b200._lib import PbBlock
    from phoebus_b200.mesh import MeshPack
    fp, deck = _brick_deck(tmp_path)
    for rank in range(world):
        mesh = MeshPack(fp, rank=rank, world=world, alloc_fields=False)
        tab = os.path.join(str(tmp_path), "tables_%d.bin" % rank)
        r = subprocess.run([EXE, "describe", deck, tab, "phoebus_b200/rank=%d" % rank, "phoebus_b200/world=%d" % world],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.splitlines() if ln.strip()}
        assert [int(x) for x in got["rgrid"]] == list(mesh.rgrid) and int(got["nblocks"][0]) == mesh.nblocks
        raw = open(tab, "rb").read()
        nb = C.sizeof(PbBlock) * mesh.nblocks
        assert raw[:nb] == bytes(mesh.blocks_host), "block table of rank %d" % rank
        sp, stot, rp, rtot, peers = mesh.merged_halo_plans()
        nd = mesh.geom_host.size + mesh.floor_host.size
        rest = raw[nb + 8 * nd:]
        ns, nr = int(got["nsend"][0]), int(got["nrecv"][0])
        assert (ns, nr, int(got["npeers"][0])) == (len(sp), len(rp), len(peers))
        plans = np.frombuffer(rest[:40 * (ns + nr)], dtype=np.int32).reshape(-1, 10)
        assert np.array_equal(plans[:ns], sp) and np.array_equal(plans[ns:], rp), "halo plans of rank %d" % rank
        tail = np.frombuffer(rest[40 * (ns + nr):], dtype=np.int64).reshape(-1, 5)
        want = [[p, ss.start, ss.stop - ss.start, rs.start, rs.stop - rs.start] for p, ss, rs in peers]
        assert tail.tolist() == want


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "tasks"])
def test_host_driver_on_two_gpus_matches_one_gpu(mode, tmp_path):
    """two phoebus_host_stage processes, one per GPU, each owning a brick of MeshBlocks (NCCL halo exchange inside the C
    library, rendezvous through a file) against the same deck on one GPU: every field of every block, bit for bit"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import bench
    from phoebus_b200.driver import FluidDriver
    from phoebus_b200.mesh import MeshPack
    fp, deck = _brick_deck(tmp_path)
    mesh = MeshPack(fp, device=torch.device("cuda:0"))
    drv = FluidDriver(mesh)
    bench.init_state(mesh, drv)
    dt = 0.05 * drv.estimate_dt()
    torch.cuda.synchronize()
    fields = [getattr(mesh, k) for k in STATE_FIELDS] + [mesh.cons["base"], mesh.cons["1"]]
    sin = os.path.join(str(tmp_path), "in.bin")
    with open(sin, "wb") as f:
        for t in fields:
            f.write(t.cpu().numpy().tobytes())
    stages = "1:%r,2:%r,1:%r,2:%r" % (dt, dt, dt, dt)
    one = os.path.join(str(tmp_path), "one.bin")
    r = subprocess.run([EXE, "run", deck, sin, one, mode, stages], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    two = os.path.join(str(tmp_path), "two.bin")
    rdv = os.path.join(str(tmp_path), "nccl_id")
    procs = [subprocess.Popen([EXE, "run", deck, sin, two, mode, stages],
                              env=dict(os.environ, PB_RANK=str(rk), PB_WORLD="2", PB_LOCAL_RANK=str(rk), PB_RENDEZVOUS=rdv),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for rk in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    ref = np.fromfile(one, dtype=np.float64)
    off = 0
    parts = [MeshPack(fp, rank=rk, world=2, alloc_fields=False) for rk in range(2)]
    raws = [np.fromfile(two + ".rank%d" % rk, dtype=np.float64) for rk in range(2)]
    offs = [0, 0]
    for k, t in zip(STATE_FIELDS + ("base", "1"), fields):
        per = t[0].numel()
        whole = ref[off:off + t.numel()].reshape(mesh.nblocks, per)
        off += t.numel()
        for rk, part in enumerate(parts):
            mine = raws[rk][offs[rk]:offs[rk] + part.nblocks * per].reshape(part.nblocks, per)
            offs[rk] += part.nblocks * per
            for n, lx in enumerate(part.lx):
                assert np.array_equal(mine[n], whole[mesh.local_index[lx]], equal_nan=True), (mode, k, rk, lx)
