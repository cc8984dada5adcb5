import os, sys, math, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import gphm_oracle as O
from toeplitz_numerics_gs import gs_apply
from toeplitz_numerics_schur import schur_lattice

def toep_apply(tab, V, antisym=False):
    n = len(tab); L = 2*n
    c = np.zeros(L); c[:n] = tab
    c[L-n+1:] = (-tab[:0:-1] if antisym else tab[:0:-1])
    if antisym: c[0] = 0
    return np.fft.irfft(np.fft.rfft(c)[:,None]*np.fft.rfft(V, L, axis=0), L, axis=0)[:n]

def kinv_diag_sums(x):
    """s[d] = sum_i Kinv[i,i+d] (d>=0) from the GS generator; returns symmetric sums (both sides for d>0)."""
    n = len(x)
    y = np.zeros(n); y[1:] = x[:0:-1]
    L = 2*n
    def wac(v):
        # sum_p (n-d-p) v[p] v[p+d] = (n-d) R_vv(d) - R_{pv,v}(d)
        fv = np.fft.rfft(v, L); fpv = np.fft.rfft(np.arange(n)*v, L)
        Rvv = np.fft.irfft(np.conj(fv)*fv, L)[:n]
        Rpv = np.fft.irfft(np.conj(fpv)*fv, L)[:n]
        d = np.arange(n)
        return (n-d)*Rvv - Rpv
    s = (wac(x) - wac(y)) / x[0]
    s[1:] *= 2
    return s

def step(p, params):
    U = params["U"].numpy(); N1, N2 = U.shape
    tau, v = float(params["log_tau"]), float(params["log_v"])
    order, c1, lam, ld = p.deriv_order, p.c1, p.llk_weight, float(p.logdet)
    anti = order == 1
    ax = []
    for x, th in [(p.x, params["kernel_paras_1"]), (p.y, params["kernel_paras_2"])]:
        d = (x - x[0]).abs()
        tK = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], 0).sum(-1).numpy().copy()
        tD = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], order).sum(-1).numpy().copy()
        tK[0] += p.jitter
        xg, logdet, _ = schur_lattice(tK)
        ax.append(dict(tK=tK, tD=tD, x=xg, logdet=logdet))
    k1 = lambda V: gs_apply(ax[0]["x"], V)                 # K1^-1 V (columns)
    k2 = lambda V: gs_apply(ax[1]["x"], V.T).T             # V K2^-1
    A = k1(U); Bt = k2(U)
    Ux = toep_apply(ax[0]["tD"], A, anti)
    Uy = toep_apply(ax[1]["tD"], Bt.T, anti).T
    nl = U*(U*U-1) if p.eq_type.startswith("allencahn") else 0.0
    R = c1*Ux + Uy + nl - p.src.numpy()
    eqgap = (R*R).sum()
    eb = O.boundary_vector_2d(torch.from_numpy(U)).numpy() - p.bvals.reshape(-1).numpy()
    bgap = (eb*eb).sum(); quad = (A*Bt).sum()
    Nb, Nc = eb.size, N1*N2
    loss = (0.5*ld*(N2*ax[0]["logdet"] + N1*ax[1]["logdet"]) + 0.5*quad - lam*(0.5*Nb*tau - 0.5*math.exp(tau)*bgap)
            - (0.5*Nc*v - 0.5*math.exp(v)*eqgap))
    G = math.exp(v)*R
    W = k1(Bt)
    sg = -1.0 if anti else 1.0
    S1 = k1(c1*sg*toep_apply(ax[0]["tD"], G, anti))        # D1^T G = +-D1 G
    S2 = k2(sg*toep_apply(ax[1]["tD"], G.T, anti).T)       # G D2 = (D2^T G^T)^T
    gU = W + S1 + S2
    if p.eq_type.startswith("allencahn"): gU = gU + G*(3*U*U-1)
    s = lam*math.exp(tau)
    gU[0,:] += s*eb[:N2]; gU[-1,:] += s*eb[N2:2*N2]; gU[:,0] += s*eb[2*N2:2*N2+N1]; gU[:,-1] += s*eb[2*N2+N1:]
    T = torch.from_numpy
    sK1 = 0.5*ld*N2*kinv_diag_sums(ax[0]["x"]) - O._diag_sums(T((S1+0.5*W) @ A.T)).numpy()
    sD1 = O._diag_sums(T(c1*(G @ A.T)), antisym=anti).numpy()
    sK2 = 0.5*ld*N1*kinv_diag_sums(ax[1]["x"]) - O._diag_sums(T((S2+0.5*W).T @ Bt)).numpy()
    sD2 = O._diag_sums(T(G.T @ Bt), antisym=anti).numpy()
    grads = {"U": T(gU)}
    for key, x, sK, sD in [("kernel_paras_1", p.x, sK1, sD1), ("kernel_paras_2", p.y, sK2, sD2)]:
        th = params[key]; d = (x - x[0]).abs()
        _, pK = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], 0, True)
        _, pD = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], order, True)
        g = [T(sK) @ pk + T(sD) @ pd for pk, pd in zip(pK, pD)]
        grads[key] = {"log-w": g[0], "log-ls": g[1], "freq": g[2]}
    terms = {"loss": float(loss), "logdet1": ax[0]["logdet"], "logdet2": ax[1]["logdet"], "quad": float(quad), "bgap": float(bgap), "eqgap": float(eqgap)}
    return terms, grads

def relerr(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1); b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a-b).norm()/b.norm().clamp_min(1e-300))

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    cases = [("poisson_2d-sin_add_cos", k, 1.0) for k in ["Matern52_Cos_1d", "SE_Cos_1d", "Matern52_1d", "SE_1d"]]
    cases += [("allencahn_2d-mix-sincos", "SE_Cos_1d", 1.0), ("advection-sin", "SE_Cos_1d", 200.0)]
    for eq, kern, beta in cases:
        try:
            p, _, _ = O.make_problem_2d(eq, kern, N, 2*math.pi if not eq.startswith("adv") else 1.0, beta=beta, llk_weight=500.0 if eq.startswith("adv") else 200.0)
        except Exception as e:
            print("skip", eq, e); continue
        for nm, params in [("S0", O.init_params_2d(N, N, 30, 20.0)), ("S1", O.state_S1(p))]:
            t0, g0 = O.loss_and_grad_efficient(p, params)
            t1, g1 = step(p, params)
            te = {k: abs(t1[k]-t0[k])/max(abs(t0[k]),1e-300) for k in t0}
            ge = {"U": relerr(g1["U"], g0["U"])}
            for key in ["kernel_paras_1", "kernel_paras_2"]:
                for leaf in ["log-w", "log-ls", "freq"]:
                    ge[key[-1]+leaf] = relerr(g1[key][leaf], g0[key][leaf])
            print(eq, kern, nm, "terms max %.1e" % max(te.values()), {k: "%.1e" % v for k, v in ge.items()})
