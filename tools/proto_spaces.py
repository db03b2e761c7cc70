"""Prototype: discrete gradient G (P2 Lagrange -> Nedelec-2 dofs) and Whitney prolongation P, built numerically
by local L2 projection in every tet (scratch / research code for the preconditioner design)."""
import numpy as np, scipy.sparse as sp
from proto_common import *

# degree-5 15-point Keast-like rule is overkill; use a 14-point degree-5 rule (Walkington)
def tet_quadrature():
    a1=0.31088591926330060980; a2=0.092735250310891226402; a3=0.045503704125649649492
    w1=0.11268792571801585080/6; w2=0.073493043116361949544/6; w3=0.042546020777081466438/6
    pts=[]; w=[]
    for a,ww in ((a1,w1),(a2,w2)):
        for i in range(4):
            l=[a]*4; l[i]=1-3*a; pts.append(l); w.append(ww)
    for i in range(4):
        for j in range(i+1,4):
            l=[a3]*4; l[i]=l[j]=0.5-a3; pts.append(l); w.append(w3)
    return np.array(pts), np.array(w)*6   # weights sum to 1 (multiply by V)

def local_funcs(t, tet_ids):
    """(s,l,X,P,Q) per tet in reference-local order + geometry. returns dict."""
    nE=t.edges.shape[1]
    tv=t.tets[:,tet_ids]; ttf=t.tet_to_field[:,tet_ids]
    lem=O.local_mapping(tv, t.edges[:,ttf[:6]]); ltm=O.local_mapping(tv, t.tris[:,ttf[6:10]-nE])
    p=t.nodes[:,tv].transpose(2,1,0); n=len(tet_ids); ar=np.arange(n)
    e1,e2,e3=p[:,1]-p[:,0],p[:,2]-p[:,0],p[:,3]-p[:,0]
    G1,G2,G3=np.cross(e2,e3),np.cross(e3,e1),np.cross(e1,e2); det=np.einsum('ij,ij->i',e1,G1)
    grad=np.stack([-(G1+G2+G3),G1,G2,G3],axis=1)/det[:,None,None]    # grad lambda
    D=np.linalg.norm(p[:,:,None,:]-p[:,None,:,:],axis=3)
    s=np.ones((n,20)); X=np.zeros((n,20),int); P=np.zeros((n,20),int); Q=np.zeros((n,20),int); ell=np.zeros((n,20))
    for e in range(6):
        A_,B_=lem[0,e],lem[1,e]
        for off,xv in ((0,A_),(10,B_)):
            X[:,e+off],P[:,e+off],Q[:,e+off]=xv,A_,B_; ell[:,e+off]=D[ar,A_,B_]
    for f in range(4):
        A_,B_,E_=ltm[0,f],ltm[1,f],ltm[2,f]
        X[:,6+f],P[:,6+f],Q[:,6+f]=B_,A_,E_; s[:,6+f]=-1; ell[:,6+f]=D[ar,A_,E_]
        X[:,16+f],P[:,16+f],Q[:,16+f]=E_,A_,B_; ell[:,16+f]=D[ar,A_,B_]
    return dict(s=s,X=X,P=P,Q=Q,ell=ell,grad=grad,V=np.abs(det)/6,lem=lem,ltm=ltm,tv=tv)

def eval_basis(F, lam):
    """lam (nq,4) -> N (n,20,nq,3)"""
    n=F['s'].shape[0]; ar=np.arange(n)[:,None]
    lX=lam[:,F['X']].transpose(1,2,0); lP=lam[:,F['P']].transpose(1,2,0); lQ=lam[:,F['Q']].transpose(1,2,0)  # (n,20,nq)
    gP=F['grad'][ar,F['P']]; gQ=F['grad'][ar,F['Q']]   # (n,20,3)
    w=lQ[...,None]*gP[:,:,None,:]-lP[...,None]*gQ[:,:,None,:]
    return (F['s']*F['ell'])[:,:,None,None]*lX[...,None]*w

def build_G_P(t, chunk=20000):
    nN=t.nodes.shape[1]; nE=t.edges.shape[1]; nT=t.tets.shape[1]; N=t.n_field
    qp,qw=tet_quadrature()
    Gr=[];Gc=[];Gv=[];Pr=[];Pc=[];Pv=[]
    for s0 in range(0,nT,chunk):
        ids=np.arange(s0,min(nT,s0+chunk)); n=len(ids); ar=np.arange(n)
        F=local_funcs(t,ids)
        Nq=eval_basis(F,qp)                                   # (n,20,nq,3)
        Mloc=np.einsum('q,naqx,nbqx->nab',qw,Nq,Nq)           # true symmetric mass / V
        # targets: grad of P2 functions (10) and Whitney (6)
        grad=F['grad']
        # vertex fn v: lam_v(2lam_v-1): grad=(4lam_v-1)grad_v ; edge fn (A,B) [local edge e]: 4(lam_A grad_B+lam_B grad_A)
        tg=[]
        for v in range(4):
            tg.append((4*qp[:,v]-1)[None,:,None]*grad[:,v][:,None,:])
        lem=F['lem']
        for e in range(6):
            A_,B_=lem[0,e],lem[1,e]
            tg.append(4*(qp[:,A_].T[...,None]*grad[ar,B_][:,None,:]+qp[:,B_].T[...,None]*grad[ar,A_][:,None,:]))
        for e in range(6):
            A_,B_=lem[0,e],lem[1,e]     # w_AB = lam_B grad_A - lam_A grad_B  (reference sign convention)
            tg.append(qp[:,B_].T[...,None]*grad[ar,A_][:,None,:]-qp[:,A_].T[...,None]*grad[ar,B_][:,None,:])
        tg=np.stack(tg,axis=1)                                # (n,16,nq,3)
        rhs=np.einsum('q,naqx,nkqx->nak',qw,Nq,tg)
        coef=np.linalg.solve(Mloc,rhs)                        # (n,20,16)
        dof=t.tet_to_field[:,ids].T                           # (n,20)
        tv=F['tv'].T                                          # (n,4)
        eid=t.tet_to_edge[:,ids].T                            # (n,6)
        cols=np.concatenate([tv, nN+eid],axis=1)              # (n,10)
        Gr.append(np.repeat(dof,10,axis=1).ravel()); Gc.append(np.tile(cols,(1,20)).ravel()); Gv.append(coef[:,:,:10].ravel())
        Pr.append(np.repeat(dof,6,axis=1).ravel()); Pc.append(np.tile(eid,(1,20)).ravel()); Pv.append(coef[:,:,10:].ravel())
    def mk(r,c,v,ncol):
        r=np.concatenate(r);c=np.concatenate(c);v=np.concatenate(v)
        keep=np.abs(v)>1e-9*np.abs(v).max()
        M=sp.coo_matrix((v[keep],(r[keep],c[keep])),shape=(N,ncol)).tocsr()
        cnt=sp.coo_matrix((np.ones(keep.sum()),(r[keep],c[keep])),shape=(N,ncol)).tocsr()
        M.data/=cnt.data      # same value from every tet -> average
        return M
    return mk(Gr,Gc,Gv,nN+nE), mk(Pr,Pc,Pv,nE)
