import os, sys, math, time
import numpy as np, scipy.linalg as sla, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gphm_oracle as O

def toep_lower_apply(c, V):
    # lower-triangular Toeplitz with first column c, times V (n x m), via FFT
    n = len(c); L = 2*n
    fc = np.fft.rfft(c, L)
    return np.fft.irfft(fc[:,None]*np.fft.rfft(V, L, axis=0), L, axis=0)[:n]
def toep_upper_apply(c, V):
    # (lower Toeplitz(c))^T times V
    n = len(c); L = 2*n
    fc = np.conj(np.fft.rfft(c, L))
    return np.fft.irfft(fc[:,None]*np.fft.rfft(V, L, axis=0), L, axis=0)[:n]
def gs_apply(x, V):
    n = len(x)
    c1 = x.copy()
    c2 = np.zeros(n); c2[1:] = x[:0:-1]
    y = toep_lower_apply(c1, toep_upper_apply(c1, V)) - toep_lower_apply(c2, toep_upper_apply(c2, V))
    return y / x[0]

def refine(K, cf, B, iters=4):
    # iterative refinement with longdouble residual
    Kl = K.astype(np.longdouble)
    X = sla.cho_solve(cf, B).astype(np.longdouble)
    Bl = B.astype(np.longdouble)
    for _ in range(iters):
        R = Bl - Kl @ X
        X = X + sla.cho_solve(cf, R.astype(np.float64)).astype(np.longdouble)
    return X

def levinson(r):
    # Levinson-Durbin for SPD Toeplitz first column r: returns x=K^-1 e1, logdet, refl coeffs
    n = len(r)
    a = np.zeros(n); a[0] = 1.0   # predictor polynomial
    E = r[0]; logdet = math.log(E)
    ks = np.zeros(n)
    for k in range(1, n):
        acc = np.dot(a[:k], r[k:0:-1])
        kap = -acc / E
        ks[k] = kap
        a[:k+1] = a[:k+1] + kap * a[:k+1][::-1].copy()
        E = E * (1 - kap*kap)
        logdet += math.log(E)
    # K a_rev = E e_n -> K^-1 e_n = a_rev/E; by persymmetry K^-1 e_1 = reverse = a / E ... a[0]=1 -> x[0]=1/E
    return a / E, logdet, ks

if __name__ == "__main__":
  for N in [int(s) for s in sys.argv[1:]] or [400, 1024]:
      p, _, _ = O.make_problem_2d("poisson_2d-sin_add_cos", "Matern52_Cos_1d", N, 2*math.pi)
      for name, params in [("S0", O.init_params_2d(N, N, 30, 20.0)), ("S1", O.state_S1(p))]:
          th = params["kernel_paras_1"]
          K, D = O._gram_pair(p.kernel, p.x, th, 2, p.jitter)
          K = K.numpy(); r = K[:,0].copy()
          U = O.state_S1(p)["U"].numpy()[:, :16].copy()
          ev = np.linalg.eigvalsh(K); print(f"N={N} {name} cond={ev[-1]/ev[0]:.3e} lmin={ev[0]:.3e}")
          cf = sla.cho_factor(K, lower=True)
          e1 = np.zeros((N,1)); e1[0]=1
          x_true = refine(K, cf, e1)[:,0]
          A_true = refine(K, cf, U)
          A_chol = sla.cho_solve(cf, U)
          x_chol = sla.cho_solve(cf, e1)[:,0]
          t=time.time(); x_lev, ld_lev, ks = levinson(r); tl=time.time()-t
          ld_chol = 2*np.log(np.diag(cf[0])).sum()
          rel = lambda a,b: float(np.linalg.norm((a-b).astype(np.float64))/np.linalg.norm(b.astype(np.float64)))
          print(f"  chol solve err {rel(A_chol, A_true):.2e}; x chol err {rel(x_chol,x_true):.2e}; x lev err {rel(x_lev,x_true):.2e}; max|k|={np.abs(ks).max():.6f} lev {tl:.1f}s")
          print(f"  logdet chol {ld_chol:.10f} lev {ld_lev:.10f} rel {abs(ld_lev-ld_chol)/abs(ld_chol):.2e}")
          for nm, x in [("true", x_true.astype(np.float64)), ("chol", x_chol), ("lev", x_lev)]:
              print(f"  GS[{nm}] err {rel(gs_apply(x, U), A_true):.2e}")
          print(f"  x0={x_true[0]:.3e} |x|={np.linalg.norm(x_true.astype(float)):.3e} |Kinv|~{1/ev[0]:.3e}")
