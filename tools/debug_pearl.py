import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "progressive-x_b200"))
from oracle import oracle as O
from pyprogressivex import _native, synthetic as syn
G = np.load(ROOT / "tests" / "golden" / "reference_scenes.npz")
def fit_F(c):
    x1=c[:,:2]; x2=c[:,2:]
    def norm(x):
        m=x.mean(0); d=np.sqrt(((x-m)**2).sum(1)).mean(); s=np.sqrt(2)/d
        T=np.array([[s,0,-s*m[0]],[0,s,-s*m[1]],[0,0,1]]); return (x-m)*s, T
    a,T1=norm(x1); b,T2=norm(x2)
    A=np.c_[b[:,0]*a[:,0], b[:,0]*a[:,1], b[:,0], b[:,1]*a[:,0], b[:,1]*a[:,1], b[:,1], a[:,0], a[:,1], np.ones(len(a))]
    _,_,vt=np.linalg.svd(A); F=vt[-1].reshape(3,3); u,s,v=np.linalg.svd(F); s[2]=0; F=u@np.diag(s)@v
    F=T2.T@F@T1; return F/F[2,2]
c=G["cubetoy_corrs"]; ref=G["cubetoy_labels"]; N=len(c)
rng=np.random.default_rng(0)
good=fit_F(c[ref==1]).reshape(-1)
bad=fit_F(c[rng.choice(N,8,replace=False)]).reshape(-1)
ctx=_native.Context(0); ctx.upload_points(1,c)
for models in (np.stack([good,bad]), np.stack([good]), np.stack([bad,good])):
    for lam in (0.5,0.3):
        D=O.pearl_datacost(1,c,models,0.75,lam)
        off,idx=ctx.knn_graph(50.0,5)
        lab_o,e_o,_=O.gco_pearl_label(D,lam,7.0,off,idx,None)
        lab,e=ctx.pearl_label(D,lam,7.0,off,idx,None)
        print("L",len(models),"lam",lam,"ref hist",np.bincount(lab_o,minlength=len(models)+1),"E",e_o,"| gpu hist",np.bincount(lab,minlength=len(models)+1),"E",e, "differ",int((lab!=lab_o).sum()))
