import sys, os, torch, json
sys.path.insert(0, os.getcwd())
from dqc_b200 import Mol
from dqc_b200.utils import systems
for name in ("c60", "taxol_like"):
    zs, pos = getattr(systems, name)()
    dev = torch.device("cuda:0")
    mol = Mol((torch.tensor(zs), torch.tensor(pos, dtype=torch.float64)), basis="def2-svp", grid="sg2", device=dev,
              orthogonalize_basis=False).densityfit(auxbasis="etb-jfit")
    h = mol.get_hamiltonian()
    h.build()
    j3c = h.df._j3c_packed
    rm = j3c.abs().amax(1)
    res = {"system": name, "npair": int(rm.numel()), "ld": int(j3c.shape[1])}
    for t in (1e-8, 1e-10, 1e-12, 1e-14, 1e-16):
        res["frac_rows_below_%g" % t] = float((rm < t).double().mean())
    print(json.dumps(res))
    del h, mol, j3c
    torch.cuda.empty_cache()
