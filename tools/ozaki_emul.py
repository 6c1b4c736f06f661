import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util
from oracle import cint, xc_ref, fock_ref
from dqc_b200.utils import systems
from dqc_b200.grid.factory import get_predefined_grid
np.random.seed(0)
zs,pos = systems.benzene()
w,_ = util.make_wrapper(zs,pos.tolist(),"def2-svp")
atm,bas,env = w.atm_bas_env
one = {z:get_predefined_grid("sg3",[z],torch.zeros(1,3,dtype=torch.float64),device=torch.device("cpu")) for z in set(zs)}
pts = np.concatenate([one[z].get_rgrid().numpy()+p for z,p in zip(zs,pos)])
dv = np.concatenate([one[z].get_dvolume().numpy() for z in zs])
# take every 4th block of 512 points
BS=512
sel = np.concatenate([np.arange(b*BS,(b+1)*BS) for b in range(0,len(pts)//BS,6)])
pts, dv = pts[sel], dv[sel]*0.5
ao = cint.eval_gto(atm,bas,env,pts,0)         # (ng, nao)
dao = cint.eval_gto(atm,bas,env,pts,1)
nao = ao.shape[1]
dm = util.seeded_dm(nao, 21, seed=1).numpy()
x = ao@dm
rho = (x*ao).sum(1); grad = 2*np.stack([(x*dao[d]).sum(1) for d in range(3)])
e,vr,vg = xc_ref.eval_unpol("gga_x_pbe + gga_c_pbe", torch.tensor(np.maximum(rho,1e-30)), torch.tensor(grad))
vr,vg = vr.numpy(), vg.numpy()
vb = dv[:,None]*(vr[:,None]*ao + 2*(vg[0][:,None]*dao[0]+vg[1][:,None]*dao[1]+vg[2][:,None]*dao[2]))
Mref = ao.T@vb
def slices(X, S, axis_scale):   # X (K, n): scale per column over the K block
    mx = np.abs(X).max(0); mx[mx==0]=1
    ex = np.ceil(np.log2(mx)); sc = 2.0**ex
    y = X/sc*64.0
    out=[]
    for s in range(S):
        q = np.rint(y); out.append(q.astype(np.int64)); y = (y-q)*128.0
    return out, sc
def ozaki(A,B,S,BS=512):
    M = np.zeros((A.shape[1],B.shape[1]))
    for b in range(0,A.shape[0],BS):
        a,sa = slices(A[b:b+BS],S,0); bb,sb = slices(B[b:b+BS],S,0)
        acc = np.zeros((A.shape[1],B.shape[1]))
        for d in range(S-1,-1,-1):     # small terms first
            t = np.zeros((A.shape[1],B.shape[1]),dtype=np.int64)
            for s in range(d+1):
                t += a[s].T@bb[d-s]
            acc += t*2.0**(-12-7*d)
        M += acc*sa[:,None]*sb[None,:]
    return M
for S in (4,5,6,7,8):
    M = ozaki(ao,vb,S)
    print("S",S,"products",S*(S+1)//2,"max abs err Vxc %.3e  (|M|max %.3e)"%(np.abs(M-Mref).max(), np.abs(Mref).max()))
# rho path: X = ao@dm per block rows scaling: A=ao rows(g) x K(mu): scale per row g over mu ; B = dm columns
def ozaki_rows(A,B,S):
    # A (M,K) scale per row; B (K,N) scale per column
    at,sa = slices(A.T,S,0); bt,sb = slices(B,S,0)
    acc=np.zeros((A.shape[0],B.shape[1]))
    for d in range(S-1,-1,-1):
        t=np.zeros_like(acc,dtype=np.int64)
        for s in range(d+1): t += at[s].T@bt[d-s]
        acc += t*2.0**(-12-7*d)
    return acc*sa[:,None]*sb[None,:]
for S in (5,6,7,8):
    X = ozaki_rows(ao,dm,S)
    r2 = (X*ao).sum(1)
    print("S",S,"rho max abs err %.3e rel-to-rho max %.3e ; Nel err %.3e"%(np.abs(r2-rho).max(), (np.abs(r2-rho)/np.maximum(np.abs(rho),1e-6)).max(), abs(((r2-rho)*dv).sum())))
