"""numpy model of the planned element kernel (mode-space sum factorisation), checked against the
oracle's literal element matrices.  Prototype only."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle.bloch_oracle import *

def tables(p):
    g,w = gauss_legendre(p); l = gauss_lobatto(p+1)
    I,D = lagrange(l,g)            # [p, p+1]
    xq,wq = gauss_legendre(p+1)
    V,_ = lagrange(l,xq)
    Lp = np.polynomial.legendre.Legendre.basis(p)(2*xq-1)
    alpha = (2*p+1)*(wq*Lp)@V
    TI = np.vstack([I,alpha])
    Dt = D@np.linalg.inv(TI)
    om = np.concatenate([w,[1/(2*p+1)]])
    return TI,Dt,om

def cyc_maps(p):
    """natural local index for cyclic (c,o,j1,j2) ND and (c,j,o1,o2) RT"""
    P1=p+1; nb=p*P1*P1; rb=p*p*P1
    nd = np.zeros((3,p,P1,P1),int); rt=np.zeros((3,P1,p,p),int)
    for c in range(3):
        dims=[P1]*3; dims[c]=p
        for o in range(p):
            for j1 in range(P1):
                for j2 in range(P1):
                    a=[0,0,0]; a[c]=o; a[(c+1)%3]=j1; a[(c+2)%3]=j2
                    nd[c,o,j1,j2]=c*nb+a[0]+dims[0]*(a[1]+dims[1]*a[2])
        dims=[p]*3; dims[c]=P1
        for j in range(P1):
            for o1 in range(p):
                for o2 in range(p):
                    a=[0,0,0]; a[c]=j; a[(c+1)%3]=o1; a[(c+2)%3]=o2
                    rt[c,j,o1,o2]=c*rb+a[0]+dims[0]*(a[1]+dims[1]*a[2])
    return nd,rt

def apply_local(p, J, kappa, eps, muinv, x_nat, ca, cm):
    TI,Dt,om = tables(p); P1=p+1
    ndm,rtm = cyc_maps(p)
    det=np.linalg.det(J); G=J.T@J/det; H=det*np.linalg.inv(J.T@J); kh=J.T@kappa
    F = x_nat[ndm].astype(complex)          # [3,p,P1,P1]
    F = np.einsum('rj,cojk->cork',TI,F); F=np.einsum('rk,cojk->cojr',TI,F)
    # curl
    R = np.zeros((3,P1,p,p),complex)
    for c in range(3):
        c1,c2=(c+1)%3,(c+2)%3
        A = F[c2]   # [o=o2, j1=j, j2=t]
        t1 = np.einsum('at,bjt->jab',Dt,A) - 1j*kh[c1]*np.transpose(A[:,:,:p],(1,2,0))   # [j,o1,o2]
        B = F[c1]   # [o=o1, j1=t, j2=j]
        t2 = np.einsum('bt,atj->jab',Dt,B) - 1j*kh[c2]*np.transpose(B[:,:p,:],(2,0,1))
        R[c]=t1-t2
    # pointwise M2 on grid points
    Y = np.zeros_like(R)
    for i0 in range(P1):
      for i1 in range(P1):
        for i2 in range(P1):
            i=[i0,i1,i2]; Om=om[i0]*om[i1]*om[i2]
            ex=[all(i[d]<p for d in range(3) if d!=c) for c in range(3)]
            for c in range(3):
                if not ex[c]: continue
                s=0
                for d in range(3):
                    if ex[d]: s+=G[c,d]*R[d,i[d],i[(d+1)%3],i[(d+2)%3]]
                Y[c,i[c],i[(c+1)%3],i[(c+2)%3]]=muinv*Om*s
    # pointwise M1
    MF = np.zeros_like(F)
    for i0 in range(P1):
      for i1 in range(P1):
        for i2 in range(P1):
            i=[i0,i1,i2]; Om=om[i0]*om[i1]*om[i2]
            ex=[i[c]<p for c in range(3)]
            for c in range(3):
                if not ex[c]: continue
                s=0
                for d in range(3):
                    if ex[d]: s+=H[c,d]*F[d,i[d],i[(d+1)%3],i[(d+2)%3]]
                MF[c,i[c],i[(c+1)%3],i[(c+2)%3]]=eps*Om*s
    # adjoint curl
    Fp = cm*MF
    for c in range(3):
        c1,c2=(c+1)%3,(c+2)%3
        Y1=Y[c1]  # [j=j1, o1, o2=o]
        Y2=Y[c2]  # [j=j2, o1=o, o2]
        t = np.einsum('at,jao->ojt',Dt,Y1)        # sum_o1 Dt[o1,j2] Y1[j1,o1,o] -> [o,j1,j2]
        t[:,:,:p] += 1j*kh[c2]*np.transpose(Y1,(2,0,1))   # [o, j1, j2<p] = Y1[j1, j2, o]
        u = np.einsum('bt,job->otj',Dt,Y2)        # sum_o2 Dt[o2,j1] Y2[j2,o,o2] -> [o,j1,j2]
        u[:,:p,:] += 1j*kh[c1]*np.transpose(Y2,(1,2,0))   # [o, j1<p, j2] = Y2[j2,o,j1]
        Fp[c] += ca*(t-u)
    Fp = np.einsum('rj,cork->cojk',TI,Fp); Fp=np.einsum('rk,cojr->cojk',TI,Fp)
    y = np.zeros(3*p*P1*P1,complex); y[ndm]=Fp
    return y

if __name__=="__main__":
    rng=np.random.default_rng(1)
    for p in [1,2,3]:
        J = rng.normal(size=(3,3)); 
        if np.linalg.det(J)<0: J[:,0]*=-1
        kappa = rng.normal(size=3); beta=np.linalg.norm(kappa); zeta=kappa/beta
        ref=RefElem(p); em=element_matrices(ref,J,zeta)
        C = em['T12']-1j*beta*em['Z12']
        eps,mu=2.5,0.7
        Ae = C.conj().T@(mu*em['M2'])@C; Me=eps*em['M1']
        n=ref.n_nd
        Apa=np.zeros((n,n),complex); Mpa=np.zeros((n,n),complex)
        for k in range(n):
            e=np.zeros(n); e[k]=1
            Apa[:,k]=apply_local(p,J,kappa,eps,mu,e,1.0,0.0)
            Mpa[:,k]=apply_local(p,J,kappa,eps,mu,e,0.0,1.0)
        print(p,"A err",abs(Apa-Ae).max()/abs(Ae).max(),"M err",abs(Mpa-Me).max()/abs(Me).max())
