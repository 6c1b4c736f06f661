"""Generates tests/golden/h2o_def2svp_fock_pieces.npz with the CPU ORACLE (the reference itself cannot be imported in
the build container: dqclibs / pylibxc2 / xitorch are absent -- see DESIGN.md section 3).  Fock pieces of
H2O / def2-SVP at a seeded density: 4-centre J and K', density-fitted J and K', PBE Vxc and E_xc on a small
Becke-Lebedev grid (level-1 atomic grids, oracle Becke weights) that is stored in the file.
The CPU suite checks that the oracle still reproduces the file, the GPU suite checks the CUDA path against it.

    python tests/golden/make_fixtures.py
"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402
from oracle import fock_ref  # noqa: E402
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper  # noqa: E402
from dqc_b200.grid.factory import get_predefined_grid  # noqa: E402
from oracle import becke_ref  # noqa: E402

XC = "gga_x_pbe + gga_c_pbe"


def build():
    zs, pos = util.H2O
    w, _ = util.make_wrapper(zs, pos, "def2-svp")
    aux, _ = util.make_wrapper(zs, pos, "etb-jfit")
    bw, aw = LibcintWrapper.concatenate(w, aux)
    # atomic grids are host-side (single atom: no partitioning); the molecular weights come from the oracle's Becke
    posn = np.asarray(pos, dtype=np.float64)
    atoms = [get_predefined_grid(1, [z], torch.zeros(1, 3, dtype=torch.float64), device=torch.device("cpu")) for z in zs]
    rgrid = np.concatenate([g.get_rgrid().numpy() + p for g, p in zip(atoms, posn)])
    owner = np.concatenate([np.full(g.get_rgrid().shape[0], i) for i, g in enumerate(atoms)])
    dvol = np.concatenate([g.get_dvolume().numpy() for g in atoms]) * becke_ref.becke_weights(rgrid, owner, posn)
    ref4 = fock_ref.RefHamilton(w, orthozer=False).build_eri()
    refd = fock_ref.RefHamilton(bw, auxwrapper=aw, orthozer=False).build_df()
    refd.setup_grid(rgrid, dvol, XC)
    dm = util.seeded_dm(ref4.nao, 5, seed=0)
    return {"dm": dm.numpy(), "j_4c": ref4.get_elrep(dm).numpy(), "k_4c": ref4.get_exchange(dm).numpy(),
            "j_df": refd.get_elrep(dm).numpy(), "k_df": refd.get_exchange_df(dm).numpy(),
            "vxc_pbe": refd.get_vxc(dm).numpy(), "exc_pbe": np.array(float(refd.get_e_xc(dm))),
            "rgrid": rgrid, "dvolume": dvol}


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "h2o_def2svp_fock_pieces.npz")
    np.savez_compressed(out, **build())
    print("wrote", out)
