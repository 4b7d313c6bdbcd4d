"""Experiment (CPU, exact-product emulation of the tensor-core operand handling: tests/fake_ops.py; prints only): the
mixed-precision schedule (precision "tf32mix") against 3xTF32 in every pass over data regimes a slice can present, both
against the float64 oracle.  Printout: profiles/r02_emulated_regimes.txt."""
import sys; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from fake_ops import FakeOps
from dmd_era5_b200.rsvd import randomized_svd_device, draw_omega, PREC_TF32MIX, PREC_TF32X3
from oracle.compare import sigma_rel_err, vector_angles
from oracle.svd_ref import randomized_svd_ref
from oracle.synthetic_np import lowrank_field_np
from oracle.slice_tools_np import delay_embed_np
rng=np.random.RandomState(5)
def run(name,X,k,d=1):
    X=X.astype(np.float32)
    Xe=delay_embed_np(X.astype(np.float64),d)
    U0,s0,V0=randomized_svd_ref(Xe,k,1)
    n=Xe.shape[1]
    out=[]
    for nm,p in (("x3",PREC_TF32X3),("mix",PREC_TF32MIX)):
        U,s,Vt=randomized_svd_device(FakeOps(),torch.from_numpy(X),k,draw_omega(n,k,1,torch.float32),precision=p,delay=d)
        a=vector_angles(U.double().numpy(),U0)
        out.append(f"{nm}: sigma {sigma_rel_err(s.numpy(),s0):.1e} U[:k/2] {a[:k//2].max():.1e} Uall {a.max():.1e}")
    print(f"{name:40s}", " | ".join(out), flush=True)
base=lowrank_field_np(8000,300,r=80,rho=0.9,seed=3)
run("centred low rank", base-base.mean(axis=1,keepdims=True), 30)
Z=base-base.mean(axis=1,keepdims=True); Z=Z/Z.std(axis=1,keepdims=True)
run("centred + unit variance", Z, 30)
scales=np.concatenate([np.full(3000,1e4),np.full(3000,1.0),np.full(2000,1e-3)])[:,None]
run("three variables, scales 1e4 / 1 / 1e-3", (base-base.mean(axis=1,keepdims=True))*scales, 30)
w=np.sqrt(np.clip(np.cos(np.deg2rad(np.linspace(-90,90,8000))),0,None))[:,None]
run("cos-latitude weights (zero rows at poles)", (base-base.mean(axis=1,keepdims=True))*w, 30)
# autocorrelated series, delay 3
t=np.arange(300); B=np.stack([np.sin(2*np.pi*t/(12+3*i)+i) for i in range(20)])+0.05*rng.standard_normal((20,300))
A=rng.standard_normal((8000,20))*(0.85**np.arange(20))
run("oscillatory, delay 3", A@B, 20, d=3)
run("slow decay rho=0.985, k=60", (lambda Y: Y-Y.mean(axis=1,keepdims=True))(lowrank_field_np(8000,300,r=250,rho=0.985,seed=4)), 60)

# fields that still carry their time mean (what precision 'auto' now keeps on 3xTF32: DESIGN.md section 3)
run('noise + 250 K mean (uncentred)', 5.0 * rng.standard_normal((8000, 300)) + 250.0, 30)
