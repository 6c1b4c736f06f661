"""World-size-2 tests of the multi-GPU host logic on CPU (gloo): the packed all-reduce, the
all-gather of the DF coefficients and the work split reproduce the single-process result when the
partial Fock pieces come from the oracle (the GPU kernels are covered by the -m gpu tests; the
sharding arithmetic around them is what runs here)."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = fn(rank, world)
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def _run(fn, tmp_path, world=2):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return torch.load(out)


def _packed_allreduce(rank, world):
    from dqc_b200.utils.dist import ParallelContext
    ctx = ParallelContext()
    assert (ctx.rank, ctx.world) == (rank, world)
    a = torch.full((3, 3), float(rank + 1), dtype=torch.float64)
    b = torch.arange(5, dtype=torch.float64) * (rank + 1)
    c = torch.tensor(2.0 * (rank + 1), dtype=torch.float64)
    ra, rb, rc = ctx.allreduce_packed([a, b, c])
    g = ctx.allgather_cat(torch.full((rank + 2,), float(rank), dtype=torch.float64), [2, 3])
    return ra, rb, rc, g


def test_packed_allreduce_and_allgather(tmp_path):
    ra, rb, rc, g = _run(_packed_allreduce, tmp_path)
    assert torch.equal(ra, torch.full((3, 3), 3.0, dtype=torch.float64))
    assert torch.equal(rb, torch.arange(5, dtype=torch.float64) * 3)
    assert float(rc) == 6.0 and rc.shape == ()
    assert g.tolist() == [0, 0, 1, 1, 1]


def _sharded_fock(rank, world):
    """Grid rows split with split_rows, aux functions split in two: partial Vxc and DF-J from the
    oracle, one packed all-reduce -- the data flow of HamiltonCGTO.get_fock_2e."""
    from dqc_b200.utils.dist import ParallelContext, split_rows
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    from oracle import fock_ref, cint
    ctx = ParallelContext()
    xcstr = "gga_x_pbe + gga_c_pbe"
    w, _ = util.make_wrapper(*util.H2O, "3-21g")
    aux, _ = util.make_wrapper(*util.H2O, "etb-jfit")
    bw, aw = LibcintWrapper.concatenate(w, aux)
    pts = util.random_points(1000, seed=4, span=2.5)
    wts = np.random.RandomState(1).uniform(0, 0.2, len(pts))
    nao = w.nao()
    dm = util.seeded_dm(nao, 5, seed=9)
    g0, g1 = split_rows(len(pts), world, rank, 128)
    h = fock_ref.RefHamilton(bw, auxwrapper=aw, orthozer=False).build_df()
    h.setup_grid(pts[g0:g1], wts[g0:g1], xcstr)
    vxc_part = h.get_vxc(dm) if g1 > g0 else torch.zeros(nao, nao, dtype=torch.float64)
    exc_part = h.get_e_xc(dm) if g1 > g0 else torch.zeros((), dtype=torch.float64)
    naux = h.j3c.shape[-1]
    a0, a1 = split_rows(naux, world, rank)
    temp = ctx.allgather_cat(torch.einsum("ij,ijl->l", dm, h.j3c[:, :, a0:a1]),
                             [split_rows(naux, world, r)[1] - split_rows(naux, world, r)[0] for r in range(world)])
    coef = temp @ h.inv_j2c
    j_part = torch.einsum("k,ijk->ij", coef[a0:a1], h.j3c[:, :, a0:a1])
    j, vxc, exc = ctx.allreduce_packed([j_part, vxc_part, exc_part])
    return j, vxc, exc


def test_sharded_fock_equals_single_process(tmp_path):
    from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
    from oracle import fock_ref
    j, vxc, exc = _run(_sharded_fock, tmp_path)
    w, _ = util.make_wrapper(*util.H2O, "3-21g")
    aux, _ = util.make_wrapper(*util.H2O, "etb-jfit")
    bw, aw = LibcintWrapper.concatenate(w, aux)
    pts = util.random_points(1000, seed=4, span=2.5)
    wts = np.random.RandomState(1).uniform(0, 0.2, len(pts))
    dm = util.seeded_dm(w.nao(), 5, seed=9)
    h = fock_ref.RefHamilton(bw, auxwrapper=aw, orthozer=False).build_df()
    h.setup_grid(pts, wts, "gga_x_pbe + gga_c_pbe")
    assert float((j - h.get_elrep(dm)).abs().max()) < 1e-12
    assert float((vxc - h.get_vxc(dm)).abs().max()) < 1e-12
    assert abs(float(exc) - float(h.get_e_xc(dm))) < 1e-12
