"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores) prints one
well-formed JSON line; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--size", "256", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "it/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "256x256" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--size", "256",
                        "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_parity_block_helpers_on_cpu():
    """bench.py's parity checker (OracleSide) fed with the oracle's LITERAL formulation as the 'measured' side: the
    comparison code (term order, packed small layout, per-leaf norms) is exercised without a GPU."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import gphm_oracle as O
    side = bench.OracleSide(48)
    for name in ("S1", "S0"):
        params = side.state(name)
        tl, gl = O.loss_and_grad_literal(side.p, params)
        terms8 = torch.tensor([tl[k] for k in bench.TERM_ORDER] + [float(gl["log_tau"]), float(gl["log_v"])], dtype=torch.float64)
        rec = side.compare(name, params, terms8, gl["U"], bench.small_from_params(gl))
        assert rec["state"] == name and rec["max_rel_term"] <= 1e-9 and rec["max_rel_leaf"] <= 1e-7, rec
        bad = terms8.clone(); bad[3] *= 1.0 + 1e-3                     # a wrong quad term must show up
        assert side.compare(name, params, bad, gl["U"], bench.small_from_params(gl))["max_rel_term"] >= (9e-4 if name == "S1" else 0.0)
    run = side.steps_and_rel_l2(2)
    assert 0.0 < run["rel_l2"] < 2.0 and run["timed_iters"] == 1 and run["U"].shape == (48, 48)


def test_clock_sampler_contract_without_nvml():
    """The clocks record always has the contract's keys; NVML polling feeds it when the driver library loads, a fake NVML here
    (no GPU in this container): median SM clock, max clock, the reasons seen in ANY sample."""
    sys.path.insert(0, ROOT)
    import time
    import types
    import bench
    s = bench.ClockSampler(0)
    if s.nvml is None:                                   # no driver: the nvidia-smi fall-back reports that it is unavailable
        s.start()
        rec = s.stop()
        assert set(("sm_mhz", "sm_max_mhz", "samples", "reasons")) <= set(rec) and rec["samples"] == 0
    calls = {"n": 0}

    def clock(handle, kind):
        calls["n"] += 1
        return 1965 if calls["n"] != 2 else 1200         # one low sample must not move the median

    fake = types.SimpleNamespace(
        NVML_CLOCK_SM=0, nvmlDeviceGetClockInfo=clock,
        nvmlDeviceGetCurrentClocksEventReasons=lambda h: 0x4 if calls["n"] == 3 else 0,
        nvmlClocksThrottleReasonHwSlowdown=0x8, nvmlClocksThrottleReasonHwThermalSlowdown=0x40,
        nvmlClocksThrottleReasonSwThermalSlowdown=0x20, nvmlClocksThrottleReasonSwPowerCap=0x4)
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.index, s.nvml, s.handle, s.proc, s.thread, s.samples, s.sm_max, s.running = 0, fake, object(), None, None, [], 1965.0, False
    s.start()
    time.sleep(0.08)
    rec = s.stop()
    assert rec["samples"] >= 3 and rec["sm_mhz"] == 1965.0 and rec["sm_max_mhz"] == 1965.0 and rec["reasons"] == ["sw_power_cap"], rec
