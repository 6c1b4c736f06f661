"""Launched by tests/test_gpu_multirank.py (torchrun or mp.spawn): every rank builds the sharded
Hamiltonian, rank 0 compares the all-reduced Fock pieces with an unsharded build on its own GPU."""
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class SoloContext(object):
    """world == 1 stand-in used for the unsharded reference build inside a multi-rank job."""
    rank, world, group = 0, 1, None

    def allreduce_(self, t):
        return t

    def allreduce_packed(self, ts):
        return ts

    def allgather_cat(self, t, sizes):
        return t


def main(rank, world, backend):
    from dqc_b200 import Mol, get_xc, SpinParam
    from dqc_b200.utils.dist import ParallelContext
    from tests import util
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    ctx = ParallelContext()
    assert ctx.world == world
    zs, pos = util.CH4ISH
    xc = "gga_x_pbe + gga_c_pbe"
    worst = worstk = 0.0
    for df in (False, True):
        def build(c):
            mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis="def2-svp" if df else "3-21g",
                      grid="sg2", device=dev, ctx=c)
            if df:
                mol.densityfit(auxbasis="etb-jfit")
            h = mol.get_hamiltonian()
            mol.setup_grid()
            h.setup_grid(mol.get_grid(), get_xc(xc))
            h.build()
            return h
        h = build(ctx)
        nao = h.nao
        dm = util.seeded_dm(nao, 6, seed=1).to(dev)
        du, dd = util.seeded_dm(nao, 6, seed=2).to(dev) * 0.5, util.seeded_dm(nao, 5, seed=3).to(dev) * 0.5
        res = [h.get_fock_2e(dm, exx=0.0 if df else 0.25).fullmatrix(), h.get_elrep(dm).fullmatrix(),
               h.get_vxc(dm).fullmatrix(), h.get_e_xc(dm).reshape(1)]
        sp = h.get_fock_2e(SpinParam(u=du, d=dd), exx=0.0 if df else 0.25)
        res += [sp.u.fullmatrix(), sp.d.fullmatrix()]
        # exact exchange: 4-centre work items dealt to ranks, or (density fitting) the aux functions of the
        # two tcgen05 GEMM stages sharded -- every rank whitens (ij|P) for its own slice of P
        resk = [h.get_exchange(dm).fullmatrix()]
        if df:
            resk.append(h.get_fock_2e(dm, exx=0.25).fullmatrix())
        if rank == 0:
            h1 = build(SoloContext())
            ref = [h1.get_fock_2e(dm, exx=0.0 if df else 0.25).fullmatrix(), h1.get_elrep(dm).fullmatrix(),
                   h1.get_vxc(dm).fullmatrix(), h1.get_e_xc(dm).reshape(1)]
            sp1 = h1.get_fock_2e(SpinParam(u=du, d=dd), exx=0.0 if df else 0.25)
            ref += [sp1.u.fullmatrix(), sp1.d.fullmatrix()]
            refk = [h1.get_exchange(dm).fullmatrix()]
            if df:
                refk.append(h1.get_fock_2e(dm, exx=0.25).fullmatrix())
            for a, b in zip(res, ref):
                worst = max(worst, float((a - b).abs().max()))
            for a, b in zip(resk, refk):
                worstk = max(worstk, float((a - b).abs().max()))
        dist.barrier()
    if rank == 0:
        print("MULTIRANK_MAXDIFF %.3e (exchange %.3e) world %d backend %s" % (worst, worstk, world, backend))
        assert worst < 1e-11, worst
        # the sliced-integer GEMMs of DF-K scale per K chunk, and the chunks follow the aux sharding
        assert worstk < 1e-9, worstk


def _spawned(rank, world, port, backend):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        main(rank, world, backend)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    # torchrun entry: one rank per GPU over NCCL
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    try:
        main(rank, world, "nccl")
    finally:
        dist.destroy_process_group()
