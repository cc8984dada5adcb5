import os, sys, math, time
import numpy as np, scipy.linalg as sla, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import gphm_oracle as O
from toeplitz_numerics_gs import toep_lower_apply, toep_upper_apply, gs_apply, refine, levinson

def schur_lattice(r):
    """Schur recursion for reflection coefficients (no dot products) + lattice for predictor."""
    n = len(r)
    al = r.copy(); be = r.copy(); be[0] = 0.0     # generator rows (unnormalised)
    a = np.zeros(n); a[0] = 1.0; b = np.zeros(n); b[0] = 1.0   # A_0 = 1, B_0 = 1
    E = r[0]; logdet = math.log(E)
    ks = np.zeros(n)
    for k in range(1, n):
        # shift alpha right by one: alpha_shift[j] = al[j-1]
        als = np.empty(n); als[0] = 0; als[1:] = al[:-1]
        kap = -be[k] / als[k]
        ks[k] = kap
        al_new = als + kap * be
        be_new = be + kap * als
        al, be = al_new, be_new
        # lattice: A_k = A_{k-1} + kap z B_{k-1};  B_k = z B_{k-1} + kap A_{k-1}
        zb = np.empty(n); zb[0] = 0; zb[1:] = b[:-1]
        a, b = a + kap * zb, zb + kap * a
        E = E * (1 - kap * kap)
        logdet += math.log(E)
    return a / E, logdet, ks

def toep_sym_apply(r, V):
    n = len(r); L = 2*n
    c = np.zeros(L); c[:n] = r; c[L-n+1:] = r[:0:-1]
    return np.fft.irfft(np.fft.rfft(c)[:,None]*np.fft.rfft(V, L, axis=0), L, axis=0)[:n]

if __name__ == "__main__":
    for N in [int(s) for s in sys.argv[1:]] or [1024]:
        p, _, _ = O.make_problem_2d("poisson_2d-sin_add_cos", "Matern52_Cos_1d", N, 2*math.pi)
        for name, params in [("S0", O.init_params_2d(N, N, 30, 20.0)), ("S1", O.state_S1(p))]:
            K, D = O._gram_pair(p.kernel, p.x, params["kernel_paras_1"], 2, p.jitter)
            K = K.numpy(); r = K[:,0].copy()
            cf = sla.cho_factor(K, lower=True)
            e1 = np.zeros((N,1)); e1[0]=1
            x_true = refine(K, cf, e1)[:,0]
            rel = lambda a,b: float(np.linalg.norm((a-b).astype(np.float64))/np.linalg.norm(b.astype(np.float64)))
            xl, ldl, kl = levinson(r)
            xs, lds, ksx = schur_lattice(r)
            ld_chol = 2*np.log(np.diag(cf[0])).sum()
            print(f"N={N} {name}: lev x err {rel(xl,x_true):.2e} ld {abs(ldl-ld_chol)/abs(ld_chol):.1e} | schur x err {rel(xs,x_true):.2e} ld {abs(lds-ld_chol)/abs(ld_chol):.1e}")
            for nm, x in [("lev", xl), ("schur", xs)]:
                for it in range(3):
                    res = e1[:,0] - toep_sym_apply(r, x[:,None])[:,0]
                    x = x + gs_apply(x, res[:,None])[:,0]
                    print(f"   {nm} refine {it+1}: x err {rel(x,x_true):.2e}")
