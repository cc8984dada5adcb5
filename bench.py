#!/usr/bin/env python
"""Benchmark of the GP-HM log-joint + gradient + Adam iteration (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (libgphm, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU oracle port

Workload (N=1 and N>1): 2-D multi-scale Poisson `poisson_2d-sin_add_cos` on a 4096^2 collocation
grid, Matern52_Cos_1d, Q=30, FP64, reference initial state.  A "step" is one full
iteration: Gram tables + K^-1 (Schur/Levinson generator + Gohberg-Semencul FFT applications on the
uniform grid; blocked Cholesky + DMMA GEMMs on general grids) + Kronecker contractions + reductions
+ hand-derived backward + Adam on every leaf.  One JSON line is printed by rank 0.
N>1 shards the same problem (strong scaling): U row/column blocks per rank, NCCL all-to-all.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank; MKL then stays single-threaded whatever torch.set_num_threads says.
# Rank 0 runs the CPU oracle (parity block, cpu_baseline, --impl reference) and needs the host cores: undo it before
# torch is imported.  The other ranks keep one thread.
if os.environ.get("LOCAL_RANK", "0") == "0" and os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import torch

METRIC = "log-joint+grad iters/sec, 2D Poisson 4096^2 grid"
UNIT = "it/s"
EQUATION, KERNEL, Q, FREQ_SCALE, LLK, LR = "poisson_2d-sin_add_cos", "Matern52_Cos_1d", 30, 20.0, 200.0, 0.01


def flops_per_iter(n):
    return 28.0 * float(n) ** 3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="collocation points per axis")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="profiling runs only")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--no-peak", action="store_true", help="profiling runs only: skip the cuBLAS DGEMM denominator")
    ap.add_argument("--no-general", action="store_true", help="skip the short dense-path (Cholesky + DGEMM) measurement")
    ap.add_argument("--workload", default="grid", choices=["grid", "ensemble"],
                    help="grid: the 4096^2 step (BASELINE metric, the default); ensemble: BASELINE configs[4], independent "
                         "solves partitioned over the ranks (64 members per GPU, weak scaling)")
    ap.add_argument("--ensemble-equation", default="poisson_2d-sin_sin", help="any equation of the config table (1-D or 2-D)")
    ap.add_argument("--members-per-gpu", type=int, default=64)
    ap.add_argument("--streams", type=int, default=16)
    ap.add_argument("--no-graph", action="store_true", help="ensemble: plain launches instead of CUDA-graph replay")
    ap.add_argument("--no-batched", action="store_true", help="ensemble: one plan + stream-issued steps per member (round 1) "
                    "instead of the batched kernels")
    ap.add_argument("--no-parity", action="store_true", help="profiling runs only: skip the oracle parity block")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the ensemble sub-record of the grid line")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region: NVML polled from a thread every 5 ms
    (the timed region of the default run is ~0.15 s - an `nvidia-smi -lms` child does not even start in that time, and its
    start-up disturbs the first timed step); `nvidia-smi` is the fall-back when NVML cannot be loaded.
    Create it before the warm-up (NVML initialisation is slow), call start() right before the timed region, stop() after."""
    SMI_FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                  "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.nvml = None
        self.handle = None
        self.proc = None
        self.thread = None
        self.samples = []            # (sm_mhz, reasons bit mask)
        self.sm_max = None
        self.running = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = self._handle(pynvml, index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _handle(pynvml, index):
        """NVML enumerates all GPUs of the box, CUDA only the visible ones: map through the PCI bus id."""
        try:
            pr = torch.cuda.get_device_properties(index)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            return pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(index)

    def _poll(self):
        nv = self.nvml
        while self.running:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            import threading
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.SMI_FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            nv = self.nvml
            self.running = False
            if self.thread is not None:
                self.thread.join(timeout=2)
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            sm = sorted(x for x, _ in self.samples)
            reasons = sorted(name for name, bit in names if any(m & bit for _, m in self.samples))
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                    "reasons": reasons, "source": "NVML polled every 5 ms inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def measure_fp64_peak(device):
    """cuBLAS DGEMM 8192^3 through torch.matmul - the native-FP64 roofline denominator
    (MEASURED_PEAKS.json has only HBM and bf16; SURVEY 8d asks for this in-run measurement)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / reps
    fl = 2.0 * n ** 3
    return fl / best / 1e9, fl / sustained / 1e9


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def trick_paras(n):
    return {"equation": EQUATION, "kernel": KERNEL, "Q": Q, "freq_scale": FREQ_SCALE, "N_col": n, "llk_weight": LLK,
            "lr": LR, "logdet": True, "nepoch": 1, "tol": -1, "scale": 2 * math.pi, "num_fold": 1}


def build_inputs(n):
    import gphm_b200 as G
    tp = trick_paras(n)
    bvals, X_col, src, X_test, u_test = G.model_GP_solver_2d.build_problem(tp, M=300)
    return tp, bvals, X_col, src, X_test, u_test


# ------------------------------------------------------------------------------------------------
TERM_ORDER = ("loss", "logdet1", "logdet2", "quad", "bgap", "eqgap")      # first six entries of the C-ABI terms[8]
PARITY_BOUND = 1e-6          # north_star: loss and gradients within 1e-6 relative in FP64
REL_L2_STEPS = 5             # Adam steps from S0 before the rel-L2 comparison with the oracle


def small_from_params(params):
    """Packed small vector of include/gphm.h: [log-w1|log-ls1|freq1|log-w2|log-ls2|freq2|log_tau|log_v]."""
    s = torch.zeros(6 * Q + 2, dtype=torch.float64)
    for a, key in enumerate(("kernel_paras_1", "kernel_paras_2")):
        for j, leaf in enumerate(("log-w", "log-ls", "freq")):
            s[(3 * a + j) * Q:(3 * a + j + 1) * Q] = torch.as_tensor(params[key][leaf], dtype=torch.float64)
    s[6 * Q], s[6 * Q + 1] = float(params["log_tau"]), float(params["log_v"])
    return s


def cpu_reference_iteration(n, steps, warmup, threads):
    """Oracle port (efficient formulation = the stronger CPU baseline) on the host cores."""
    from oracle import gphm_oracle as O
    torch.set_num_threads(threads)
    p, _, _ = O.make_problem_2d(EQUATION, KERNEL, n, 2 * math.pi, llk_weight=LLK, M=8)
    params = O.init_params_2d(n, n, Q, FREQ_SCALE)
    st = O.adam_init(params)
    for _ in range(warmup):
        params, st, _ = O.step(p, params, st, LR, "efficient")
    t0 = time.perf_counter()
    for _ in range(steps):
        params, st, terms = O.step(p, params, st, LR, "efficient")
    dt = time.perf_counter() - t0
    return dt / max(steps, 1), terms["loss"]


def cpu_literal_iteration(n, threads):
    """BASELINE.md section 3 variant (i): the reference's own dataflow ((N,N,Q) tensors, autograd, LU solve + slogdet)."""
    from oracle import gphm_oracle as O
    torch.set_num_threads(threads)
    p, _, _ = O.make_problem_2d(EQUATION, KERNEL, n, 2 * math.pi, llk_weight=LLK, M=8)
    params = O.init_params_2d(n, n, Q, FREQ_SCALE)
    st = O.adam_init(params)
    params, st, _ = O.step(p, params, st, LR, "literal")
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        params, st, _ = O.step(p, params, st, LR, "literal")
    t_lit = (time.perf_counter() - t0) / reps
    t_eff, _ = cpu_reference_iteration(n, 3, 1, threads)
    return t_lit, t_eff


class OracleSide(object):
    """Rank 0's CPU checker for the bench: the oracle's efficient formulation on the bench workload itself."""

    def __init__(self, n):
        from oracle import gphm_oracle as O
        self.O = O
        torch.set_num_threads(os.cpu_count() or 1)
        self.p, (self.xt, self.yt), self.ut = O.make_problem_2d(EQUATION, KERNEL, n, 2 * math.pi, llk_weight=LLK, M=300)
        self.n = n

    def state(self, name):
        return self.O.state_S1(self.p, Q, FREQ_SCALE) if name == "S1" else self.O.init_params_2d(self.n, self.n, Q, FREQ_SCALE)

    def compare(self, name, params, terms8, gU, gsmall):
        """One value_and_grad at `params`: six loss terms + d/dlog_tau, d/dlog_v and every gradient leaf, relative."""
        te, ge = self.O.loss_and_grad_efficient(self.p, params)
        want_t = [te[k] for k in TERM_ORDER] + [float(ge["log_tau"]), float(ge["log_v"])]
        names = list(TERM_ORDER) + ["d_log_tau", "d_log_v"]
        rel_t = {}
        for k, got, want in zip(names, terms8.tolist(), want_t):
            rel_t[k] = abs(got - want) / abs(want) if want != 0.0 else abs(got)
        rel_l = {"U": float((gU.reshape(-1) - ge["U"].reshape(-1)).norm() / ge["U"].norm())}
        gs_want = small_from_params(ge)
        for a in (1, 2):
            for j, leaf in enumerate(("log-w", "log-ls", "freq")):
                sl = slice((3 * (a - 1) + j) * Q, (3 * (a - 1) + j + 1) * Q)
                den = float(gs_want[sl].norm())
                rel_l["kernel_paras_%d/%s" % (a, leaf)] = float((gsmall[sl] - gs_want[sl]).norm()) / (den if den > 0 else 1.0)
        return {"state": name, "max_rel_term": max(rel_t.values()), "max_rel_leaf": max(rel_l.values()),
                "worst_term": max(rel_t, key=rel_t.get), "worst_leaf": max(rel_l, key=rel_l.get), "loss": te["loss"]}

    def steps_and_rel_l2(self, steps):
        """`steps` Adam iterations from S0 (the first one untimed), then preds on the 300^2 test grid."""
        O = self.O
        params = self.state("S0")
        st = O.adam_init(params)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            params, st, terms = O.step(self.p, params, st, LR, "efficient")
            times.append(time.perf_counter() - t0)
        pred = O.preds_2d(self.p, params, self.xt, self.yt)
        timed = times[1:] if len(times) > 1 else times
        return {"rel_l2": O.rel_l2(pred, self.ut), "loss_last_step": terms["loss"], "U": params["U"],
                "s_per_iter": sum(timed) / len(timed), "timed_iters": len(timed)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.size
    # bounded sample: calibrate on N=1024 and pick the largest grid whose (W+K) iterations fit ~4 min
    t1024, _ = cpu_reference_iteration(1024, 1, 1, threads)
    budget, n_s = 240.0, n
    while n_s > 1024 and t1024 * (n_s / 1024.0) ** 3 * (args.steps + args.warmup) > budget:
        n_s //= 2
    t_iter, loss = cpu_reference_iteration(n_s, args.steps, args.warmup, threads)
    scale = (float(n_s) / n) ** 3                        # 28 N^3 FLOPs per iteration
    value = scale / t_iter
    sample = ("one full iteration per step at N=%d" % n_s) if n_s == n else \
        ("one full iteration per step at N=%d, it/s scaled by (N_s/N)^3 = %.4g to the N=%d workload" % (n_s, scale, n))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "%s %dx%d %s Q=%d S0" % (EQUATION, n, n, KERNEL, Q),
                       "reference": "torch-FP64 CPU port of the reference step (oracle/gphm_oracle.py, efficient "
                                    "formulation); the JAX reference cannot be installed here (no jax wheel, no network)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def measure_cufft(device, L=8192, batch=2048):
    """cuFFT Z2Z (through torch.fft.fft), `batch` transforms of length L: the library kernel the in-shared-memory
    transforms of gs_apply_fused_kernel are compared with (comparison only, never on the path)."""
    x = torch.randn(batch, L, dtype=torch.complex128, device=device)
    for _ in range(2):
        torch.fft.fft(x, dim=1)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.fft.fft(x, dim=1); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = 5.0 * L * math.log2(L) * batch
    return {"ms": best, "tflops": fl / best / 1e9, "batch": batch, "L": L,
            "hbm_gbs": 2.0 * 16 * L * batch / best / 1e6,
            "note": "out-of-place Z2Z, operands in HBM (one read + one write of the batch per transform)"}


def run_ours(args, rank, world, local):
    import gphm_b200 as G
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    lib = G._lib.load()
    n = args.size
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    tp, bvals, X_col, src, X_test, u_test = build_inputs(n)
    model_init = G.GP_solver_2d_single.init_params

    class _M:                                          # init_params only needs these attributes
        trick_paras, N1, N2 = tp, n, n
    s0_small = small_from_params(model_init(_M))

    if world == 1:
        core = G.solver_core.SolverCore(2, KERNEL, "poisson", X_col[0], X_col[1], src, bvals, None, LLK, 1.0, 1.0, 1e-6, Q)
        st = core.new_state(model_init(_M))
        step = lambda: core.step_inplace(st, LR)
        parallelism = "1 GPU"

        def set_state(U_full, small):
            st.U.copy_(U_full.reshape(-1)); st.small.copy_(small)
            for t in (st.mU, st.vU, st.msmall, st.vsmall, st.count):
                t.zero_()

        def value_and_grad():
            terms, gU, gs = core.value_and_grad(st)
            return terms.cpu(), gU.cpu(), gs.cpu()

        full_U = lambda: st.U.reshape(n, n)
        cur_small = lambda: st.small
        cur_loss = lambda: float(st.terms[0])
    else:
        from importlib import import_module
        distmod = import_module("gaussian-process-slover-for-high-freq-pde_b200.dist")
        solver = distmod.ShardedSolver2D(KERNEL, "poisson", X_col[0], X_col[1], src, bvals, LLK, 1.0, 1.0, 1e-6, Q, LR)
        solver.init_state(FREQ_SCALE)
        core = solver.ops.core
        step = solver.step
        st = None
        parallelism = "U row/column blocks over %d GPUs, %s" % (world, solver.exchange_name())
        set_state = lambda U_full, small: solver.set_state(U_full, small)

        def value_and_grad():
            terms, gU_r, gs = solver.value_and_grad()
            parts = [torch.empty_like(gU_r) for _ in range(world)]
            dist.all_gather(parts, gU_r.contiguous())
            return terms.cpu(), torch.cat(parts, 0).cpu(), gs.cpu()

        full_U = solver.gather_U
        cur_small = lambda: solver.small
        cur_loss = lambda: float(solver.last_loss())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bcast(t):
        t = t.to(device)
        if world > 1:
            dist.broadcast(t, src=0)
        return t

    def predict_rel_l2():
        class _S:
            pass
        ps = _S()
        ps.U, ps.small = full_U().reshape(-1).contiguous(), cur_small()
        pred = core.predict(ps, torch.as_tensor(X_test[0]), torch.as_tensor(X_test[1]))
        return float(core.rel_l2(pred, torch.as_tensor(u_test, dtype=torch.float64).to(device).contiguous()))

    # ---- parity at the bench workload's own size, BEFORE timing (rank 0 runs the CPU oracle) ----
    parity = None
    rel_l2_after = None
    oracle_run = None
    if not args.no_parity:
        side = OracleSide(n) if rank == 0 else None
        checks = []
        for name in ("S1", "S0"):
            params = side.state(name) if rank == 0 else None
            U_full = bcast(torch.as_tensor(params["U"], dtype=torch.float64) if rank == 0 else torch.empty(n, n, dtype=torch.float64))
            small = bcast(small_from_params(params) if rank == 0 else torch.empty(6 * Q + 2, dtype=torch.float64))
            set_state(U_full, small)
            terms8, gU, gs = value_and_grad()
            core.raise_on_bad_status()
            if rank == 0:
                checks.append(side.compare(name, params, terms8, gU.reshape(n, n), gs))
        # rel-L2 half of the metric: REL_L2_STEPS Adam steps from S0 on both sides, preds on the 300^2 test grid
        set_state(torch.zeros(n, n, dtype=torch.float64, device=device), s0_small.to(device))
        for _ in range(REL_L2_STEPS):
            step()
        ours_rel = predict_rel_l2()
        ours_loss = cur_loss()
        ours_U = full_U()                              # collective for N > 1: every rank calls it
        ours_U = ours_U.cpu() if rank == 0 else None
        if rank == 0:
            oracle_run = side.steps_and_rel_l2(REL_L2_STEPS)
            rel_l2_after = {"steps": REL_L2_STEPS, "ours": ours_rel, "oracle": oracle_run["rel_l2"],
                            "rel_diff": abs(ours_rel - oracle_run["rel_l2"]) / oracle_run["rel_l2"],
                            "loss_last_step_rel_diff": abs(ours_loss - oracle_run["loss_last_step"]) / abs(oracle_run["loss_last_step"]),
                            "max_abs_dU": float((ours_U - oracle_run["U"]).abs().max()),
                            "test_grid": "300x300", "bound": "final rel-L2 within 5% of the oracle's (north_star)"}
            worst = max(checks, key=lambda c: max(c["max_rel_term"], c["max_rel_leaf"]))
            parity = {"state": "+".join(c["state"] for c in checks), "max_rel_term": max(c["max_rel_term"] for c in checks),
                      "max_rel_leaf": max(c["max_rel_leaf"] for c in checks), "bound": PARITY_BOUND,
                      "worst": {k: worst[k] for k in ("state", "worst_term", "worst_leaf")},
                      "per_state": checks, "against": "oracle.loss_and_grad_efficient at N=%d on the host cores" % n,
                      "n_gpus": world}
            ok = parity["max_rel_term"] <= PARITY_BOUND and parity["max_rel_leaf"] <= PARITY_BOUND and rel_l2_after["rel_diff"] <= 0.05
            parity["ok"] = bool(ok)
            del side
        flag = bcast(torch.tensor([1.0 if (rank != 0 or parity["ok"]) else 0.0], dtype=torch.float64))
        if float(flag) == 0.0:
            if rank == 0:
                print(json.dumps({"parity": parity, "rel_l2_after": rel_l2_after}), flush=True)
            raise SystemExit("bench.py: parity against the oracle FAILED at N=%d" % n)
        set_state(torch.zeros(n, n, dtype=torch.float64, device=device), s0_small.to(device))

    # ---- timed region: profiling OFF ----
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = lib.gphm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    e0.record()
    marks = []
    for _ in range(args.steps):
        step()
        m = torch.cuda.Event(enable_timing=True)
        m.record()
        marks.append(m)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.gphm_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    per_step = sorted(a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks))      # this rank's steps
    step_spread = {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]}
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t)
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step
    loss = cur_loss()
    if not math.isfinite(loss):
        raise SystemExit("bench.py: non-finite loss")
    core.raise_on_bad_status()
    rel_l2_timed_state = predict_rel_l2()

    # ---- per-family breakdown: a SEPARATE pass with the event brackets on (not part of `value`) ----
    NF = 8
    psteps = min(args.steps, 10)
    barrier()
    lib.gphm_profile_start()
    for _ in range(psteps):
        step()
    cat_ms = (ctypes.c_double * NF)(); cat_fl = (ctypes.c_double * NF)(); cat_by = (ctypes.c_double * NF)()
    cat_n = (ctypes.c_longlong * NF)()
    lib.gphm_profile_stop(cat_ms, cat_fl, cat_by, cat_n)
    barrier()

    # ---- e2e: the same step through the host-buffer entry points (H2D + step + D2H every step) ----
    e2e = None
    if world == 1 and not args.no_e2e:
        nf, ns = n * n, 6 * Q + 2
        host = [torch.zeros(nf, dtype=torch.float64).pin_memory() for _ in range(3)]
        hs = [torch.zeros(ns, dtype=torch.float64).pin_memory() for _ in range(3)]
        hcount = torch.zeros(1, dtype=torch.int64).pin_memory()
        hterms = torch.zeros(8, dtype=torch.float64).pin_memory()
        host[0].copy_(st.U.cpu()); hs[0].copy_(st.small.cpu())
        e2e_steps = min(args.steps, 10)
        # (a) params in / params + loss out, opt_state resident on the device (what the reference's loop moves)
        core.step_host_params(host[0], hs[0], hterms, LR, reset_opt=True, hcount=hcount)       # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            core.step_host_params(host[0], hs[0], hterms, LR, hcount=hcount)
        dt_p = time.perf_counter() - t0
        # (b) params AND both Adam moments in and out every step (round 1's e2e definition)
        core.step_host(host[0], hs[0], host[1], host[2], hs[1], hs[2], hcount, hterms, LR)       # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            core.step_host(host[0], hs[0], host[1], host[2], hs[1], hs[2], hcount, hterms, LR)
        dt_f = time.perf_counter() - t0
        h2d_p = 8 * (nf + ns)
        h2d_f = 8 * (3 * nf + 3 * ns) + 8
        e2e = {"value": e2e_steps / dt_p, "unit": UNIT, "h2d_bytes_per_step": h2d_p, "d2h_bytes_per_step": h2d_p + 64 + 8,
               "steps": e2e_steps,
               "api": "gphm_step_host_params (pinned host params in, updated params + loss terms out every step; the Adam "
                      "state stays on the device like optax's state in the reference's loop)",
               "full_state": {"value": e2e_steps / dt_f, "unit": UNIT, "h2d_bytes_per_step": h2d_f,
                              "d2h_bytes_per_step": h2d_f + 64,
                              "api": "gphm_step_host (params AND Adam moments in and out every step)"}}
    elif world > 1 and not args.no_e2e:
        # every rank moves ITS row block of the params host -> device before, and device -> host after, each step
        blocks = [solver.U, solver.small]
        hostb = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in blocks]
        for h, t in zip(hostb, blocks):
            h.copy_(t.cpu())
        hloss = torch.zeros(1, dtype=torch.float64).pin_memory()

        def host_step():       # upload beside the factor stage, Adam(U) + download beside the theta-gradient tail (dist.py)
            solver.step_host(hostb[0], hostb[1], hloss)
        e2e_steps = min(args.steps, 10)
        host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        per_rank = sum(t.numel() * t.element_size() for t in blocks)
        e2e = {"value": e2e_steps / float(tt), "unit": UNIT, "h2d_bytes_per_step": per_rank * world,
               "d2h_bytes_per_step": per_rank * world + 8 * world, "steps": e2e_steps,
               "api": "ShardedSolver2D.step_host: every rank's row block of the params from / to pinned host memory every step "
                      "(Adam state resident on the device)"}

    # ---- ensemble sub-record (BASELINE configs[4]): this rank's share of the members, a few ensemble steps ----
    ensemble = None
    if not args.no_ensemble:
        ensemble = ensemble_record(args, rank, world, local, brief=True)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ----
    peaks = measured_peaks()
    burst = sustained = None
    if not args.no_peak:
        burst, sustained = measure_fp64_peak(device)
    fam = ("gram", "dgemm", "factor_serial", "reduce_elementwise", "adam", "fft_diag_sums", "gs_kinv_apply", "toeplitz_products")
    by_family = {name: cat_ms[i] / psteps for i, name in enumerate(fam)}
    peak_source = ("FP64 pipe: cuBLAS DGEMM 8192^3 measured in this run, sustained over %d back-to-back calls (burst %.1f); "
                   "MEASURED_PEAKS.json has no FP64 entry (bf16 %.0f TFLOP/s, HBM %.0f GB/s)"
                   % (40, burst or 0.0, peaks.get("bf16_tflops", 0.0), peaks.get("hbm_gbs", 0.0)))

    def family_roofline(idx, kernel, traffic_file, note):
        ms, fl = cat_ms[idx], cat_fl[idx]
        ach = fl / ms / 1e9 if ms > 0 else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", traffic_file)) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        return {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": sustained, "unit": "TFLOP/s",
                "frac": (ach / sustained) if (ach and sustained) else None, "traffic": traffic, "note": note,
                "peak_source": peak_source, "launches_per_step": cat_n[idx] / psteps, "ms_per_step": ms / psteps,
                "flops_per_step": fl / psteps,
                "algorithmic_bytes_per_launch": (cat_by[idx] / cat_n[idx]) if cat_n[idx] else None,
                "hbm_gbs_algorithmic": (cat_by[idx] / ms / 1e6) if ms > 0 else None,
                "timed": "CUDA events around every launch of the family in a separate %d-step pass after the timed region" % psteps}

    if cat_ms[6] >= cat_ms[1]:
        roofline = family_roofline(
            6, "gs_apply_fused_kernel (K^-1 rows by Gohberg-Semencul: six length-2N FP64 FFTs per row pair in shared memory)",
            "gs_apply_traffic.json",
            "compute-bound by the roofline model (24 FLOP/B algorithmic intensity vs 5.4 FLOP/B machine balance); FLOPs = "
            "5 L log2 L per complex transform + spectrum products; FP64 FMA pipe (same peak rate as the FP64 DMMA pipe; "
            "tcgen05 has no f64 kind); the binding on-chip resource is shared-memory bandwidth (ncu: see profiles/)")
    else:
        roofline = family_roofline(1, "dgemm_kernel (FP64 DMMA.8x8x4; tcgen05 has no f64 kind)", "dgemm_traffic.json",
                                   "FLOPs issued by the triangular / full GEMM launches of the step")
    roofline["step_tflops_of_28N3"] = flops_per_iter(n) / ms_step / 1e9
    roofline["step_frac_of_peak"] = (flops_per_iter(n) / ms_step / 1e9 / sustained) if sustained else None
    roofline["by_family_ms_per_step"] = by_family
    executed = sum(cat_fl[i] for i in range(NF)) / psteps
    roofline["step_executed_flops"] = executed
    roofline["step_executed_frac_of_peak"] = (executed / ms_step / 1e9 / sustained) if sustained else None
    if not args.no_peak and world == 1:
        roofline["cufft_z2z_comparison"] = measure_cufft(device)
    # the dense path (general grids, axes longer than 4096): same workload with the Toeplitz inverse generator and the
    # FFT products switched off, a few steps, to report the DGEMM kernel against the same FP64 peak
    if world == 1 and not args.no_general:
        del st
        core = None
        torch.cuda.empty_cache()
        core2 = G.solver_core.SolverCore(2, KERNEL, "poisson", X_col[0], X_col[1], src, bvals, None, LLK, 1.0, 1.0, 1e-6, Q,
                                         force_general=16 | 8)
        st2 = core2.new_state(model_init(_M))
        core2.step_inplace(st2, LR)
        torch.cuda.synchronize()
        lib.gphm_profile_start()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gsteps = 3
        for _ in range(gsteps):
            core2.step_inplace(st2, LR)
        g1.record()
        torch.cuda.synchronize()
        ms2 = (ctypes.c_double * NF)(); fl2 = (ctypes.c_double * NF)(); by2 = (ctypes.c_double * NF)(); n2_ = (ctypes.c_longlong * NF)()
        lib.gphm_profile_stop(ms2, fl2, by2, n2_)
        ach2 = fl2[1] / ms2[1] / 1e9 if ms2[1] > 0 else None
        roofline["general_path"] = {
            "what": "same workload, force_general = 16|8: blocked Cholesky + triangular DMMA GEMMs for K^-1, GEMMs for the "
                    "derivative-Gram contractions (the path of non-uniform grids)",
            "ms_per_step": g0.elapsed_time(g1) / gsteps, "dgemm_ms_per_step": ms2[1] / gsteps,
            "dgemm_tflops": ach2, "dgemm_frac_of_fp64_peak": (ach2 / sustained) if (ach2 and sustained) else None,
            "dgemm_launches_per_step": n2_[1] / gsteps}
        del st2, core2
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        if oracle_run is not None:
            t_iter, sample = oracle_run["s_per_iter"], ("%d full iterations at N=%d after 1 warm-up (oracle efficient formulation, "
                                                         "torch FP64/MKL) - the same run that produces rel_l2_after.oracle"
                                                         % (oracle_run["timed_iters"], n))
        else:
            t_iter, _ = cpu_reference_iteration(n, 2, 1, threads)
            sample = "2 full iterations at N=%d after 1 warm-up (oracle efficient formulation, torch FP64/MKL)" % n
        cpu = {"value": 1.0 / t_iter, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        try:
            t_lit, t_eff = cpu_literal_iteration(400, threads)
            cpu["literal_variant"] = {"n": 400, "it_per_s": 1.0 / t_lit, "efficient_it_per_s_same_n": 1.0 / t_eff,
                                      "what": "BASELINE.md section 3 variant (i): the reference's own dataflow ((N,N,Q) tensors, "
                                              "autograd, LU solve + slogdet) at N=400 (it cannot run at 4096: one (N^2,Q) "
                                              "intermediate is 4 GB); variant (ii) at the same N beside it"}
        except Exception as ex:                                  # context only; never fails the bench
            cpu["literal_variant"] = {"error": repr(ex)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s %dx%d %s Q=%d S0" % (EQUATION, n, n, KERNEL, Q), "parallelism": parallelism,
                       "l2": "working set per step ~%.1f GB >> 126 MB L2 (no flush needed)" % (28 * n * n * 8 / 1e9),
                       "loss_after": loss, "rel_l2_timed_state": rel_l2_timed_state, "step_ms_rank0": step_spread},
            "parity": parity, "rel_l2_after": rel_l2_after,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "ensemble": ensemble}
    print(json.dumps(line), flush=True)


def ensemble_record(args, rank, world, local, brief=False):
    """BASELINE configs[4]: 64 x 8 = 512 independent solves (seeds x frequency inits) over 8 GPUs = 64 members per
    GPU; every rank steps its own members, no data-path collective.  Returns the JSON record on rank 0 (None elsewhere).
    brief: the sub-record of the grid line (a few ensemble steps, no clock sampler)."""
    import gphm_b200 as G
    from importlib import import_module
    E = import_module("gaussian-process-slover-for-high-freq-pde_b200.ensemble")
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
    eq = args.ensemble_equation
    tp = G.configs.load_config(eq, "/nonexistent")
    tp.update(equation=eq, kernel=KERNEL, scale=2 * math.pi if tp["scale"] == "2pi" else 1.0)
    n_total = args.members_per_gpu * world
    seeds = (n_total + len(E.DEFAULT_FREQ_SCALES) - 1) // len(E.DEFAULT_FREQ_SCALES)
    members = E.ensemble_members(seeds)[:n_total]
    ens, mine = E.build_ensemble(tp, members, rank=rank, world=world, streams=args.streams, graph=not args.no_graph,
                                 batched=not args.no_batched)
    lib = G._lib.load()
    steps = 10 if brief else args.steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if (rank == 0 and not brief) else None
    ens.step(max(args.warmup, 3))
    barrier()
    launches0 = lib.gphm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    e0.record()
    ens.step(steps)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.gphm_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    res = torch.stack((ens.losses(), ens.errors()), dim=1)
    ens.raise_on_bad_status()
    full = E.gather_results(res, n_total, rank, world)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t)
    if rank != 0:
        return None
    if not bool(torch.isfinite(full).all()):
        raise SystemExit("bench.py: non-finite member loss")
    ms_step = ms_total / steps
    per_member = ens.models[0].core
    line = {"metric": "ensemble member-iterations/sec (log-joint+grad+Adam), %d independent solves" % n_total,
            "value": n_total * 1e3 / ms_step, "unit": "member-it/s", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ensemble %s N_col=%d %s Q=%d, %d members (%d per GPU): seeds x freq_scale %s"
                                   % (eq, tp["N_col"], KERNEL, tp["Q"], n_total, args.members_per_gpu, list(E.DEFAULT_FREQ_SCALES)),
                       "parallelism": "members partitioned over %d GPU(s), no collective on the data path" % world,
                       "issue": ens.issue_mode(),
                       "l2": "per-member working set %.1f MB x %d members resident" % (per_member.workspace.numel() / 1e6, len(ens)),
                       "loss_min_max": [float(full[:, 0].min()), float(full[:, 0].max())],
                       "rel_l2_min_max": [float(full[:, 1].min()), float(full[:, 1].max())]},
            "clocks": clocks, "e2e": None,
            # graph replay launches no kernel from the host: count the kernel nodes one replay executes
            "gpu_launches": int(launches) * world if not ens.use_graph else int(ens.kernels_per_step) * steps * world,
            "kernels_per_ensemble_step_per_gpu": ens.kernels_per_step}
    del ens
    torch.cuda.empty_cache()
    return line


def run_ensemble(args, rank, world, local):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = ensemble_record(args, rank, world, local)
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    if args.workload == "ensemble":
        run_ensemble(args, rank, world, local)
    else:
        run_ours(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
