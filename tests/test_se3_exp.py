"""se3Exp on batched jets (xs_se3_exp; KinectFusionReconstruction.h:176-219): the reference's small-angle branch at xi = 0 (first order
h G_i, second order only the omega^ v cross term) and, at a finite rotation, value / first / second order against finite
differences of the matrix exponential.  Host code: runs without a GPU."""
import numpy as np
from scipy.linalg import expm


def test_se3_exp_jets(xs):
    G=xs.se3_generators()
    def hat(x): return np.tensordot(x,G,1)
    h=1e-7
    # (1) xi = 0: Hessian batch seeds along the 6 axes -> first order h G_i; second order from the linear branch
    n=6; pairs=xs.all_pairs(n)
    xi=np.zeros((1+n+len(pairs),6),np.float32)
    for i in range(n): xi[1+i,i]=h
    T=xs.se3_exp(xi,comps=2,dirs=n)
    assert np.allclose(T[0],np.eye(4))
    for i in range(n): assert np.allclose(T[1+i],h*G[i],atol=1e-14)
    # reference quirk: the small-angle branch is linear in omega, so only the omega^ v cross term is second order
    k=pairs.index((0,4)); S=T[1+n+k]
    want=np.zeros((4,4)); want[:3,3]=h*h*(G[4][:3,:3]@np.array([1,0,0]))
    assert np.allclose(S,want,atol=1e-20), (S,want)
    # (2) a finite rotation: first order against finite differences of the matrix exponential, second order likewise
    x0=np.array([0.1,-0.2,0.05,0.3,-0.1,0.2])
    xi=np.zeros((1+n+len(pairs),6),np.float32); xi[0]=x0
    for i in range(n): xi[1+i,i]=h
    T=xs.se3_exp(xi,comps=2,dirs=n)
    f=lambda x: expm(hat(x))
    assert np.abs(T[0]-f(x0)).max()<2e-7
    d=1e-4
    for i in range(n):
        e=np.zeros(6); e[i]=d
        fd=(f(x0+e)-f(x0-e))/(2*d)
        assert np.abs(T[1+i]/h-fd).max()<2e-5,(i,np.abs(T[1+i]/h-fd).max())
    for k,(i,j) in enumerate(pairs):
        ei=np.zeros(6); ei[i]=d; ej=np.zeros(6); ej[j]=d
        fd=(f(x0+ei+ej)-f(x0+ei-ej)-f(x0-ei+ej)+f(x0-ei-ej))/(4*d*d)
        err=np.abs(T[1+n+k]/h/h-fd).max()
        assert err<2e-3*max(1,np.abs(fd).max()),(i,j,err)
    print('se3_exp ok')

