"""Turn ncu captures (gpurun_out/) into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md [steps]
    python tools/summarize_ncu.py kernel   gpurun_out/prof_gemm.ncu-rep profiles/r01_dgemm_ncu.md
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys


def launches(src, dst, steps=1):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"<.*", "", r["Kernel Name"]).replace("void ", "").split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list: `%s`\n\n" % os.path.basename(src))
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prof_step.py 4096 %d` - the "
                "bench workload (poisson_2d-sin_add_cos 4096x4096 Matern52_Cos_1d Q=30), plan creation + %d full step(s), %d launches; "
                "run only after the same command exited 0 without ncu.\n" % (steps, steps, len(rows)))
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with the live CUDA-event "
                "numbers in the bench line, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.4f | %.1f%% |\n" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
        f.write("| **total** | %d | %.3f | | |\n" % (len(rows), tot))
    print(open(dst).read())


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full: %s\n\n" % os.path.basename(src))
        f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel -c 3 python "
                "tools/prof_gemm.py` (one N=4096 FP64 GEMM per launch: NN, NT, TN).\n\n")
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join("`%s`" % re.sub(r"\(gphm::GemmArgs\)|\(int\)|\(bool\)|gphm::|void ", "", r[idx["Kernel Name"]]) for r in data) + " |\n")
        for k in KEYS:
            if k in idx:
                f.write("| %s | %s | " % (k, units[idx[k]]) + " | ".join(r[idx[k]] for r in data) + " |\n")
    rd = [float(r[idx["dram__bytes_read.sum"]].replace(",", "")) for r in data]
    wr = [float(r[idx["dram__bytes_write.sum"]].replace(",", "")) for r in data]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    tr = [(a * scale[units[idx["dram__bytes_read.sum"]]] + b * scale[units[idx["dram__bytes_write.sum"]]]) for a, b in zip(rd, wr)]
    out = {"dram_bytes_per_launch": sum(tr) / len(tr), "launch": "N=4096 FP64 GEMM (algorithmic 3*N^2*8 = 402.7 MB)",
           "source": os.path.basename(dst)}
    json.dump(out, open(os.path.join(os.path.dirname(dst), "dgemm_traffic.json"), "w"), indent=1)
    print(open(dst).read())


KEYS2 = KEYS + ["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def kernels(src, dst, command, traffic_kernel=None, traffic_json=None, note=""):
    """Any capture: one column per launch; optionally the mean DRAM traffic of `traffic_kernel` launches -> json."""
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    short = lambda r: re.sub(r"<.*|\(.*|gphm::|void ", "", r[idx["Kernel Name"]])
    with open(dst, "w") as f:
        f.write("# ncu --set full: %s\n\nCommand: `%s`\n\n%s\n\n" % (os.path.basename(src), command, note))
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join("`%s`" % short(r) for r in data) + " |\n")
        for k in KEYS2:
            if k in idx and "dmma" not in k and "fp64.avg.pct_of_peak_sustained_elapsed" not in k:
                f.write("| %s | %s | " % (k, units[idx[k]]) + " | ".join(r[idx[k]] for r in data) + " |\n")
    if traffic_kernel:
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        sel = [r for r in data if traffic_kernel in r[idx["Kernel Name"]]]
        tr = [float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]] +
              float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]] for r in sel]
        json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "launches": len(tr), "kernel": traffic_kernel,
                   "launch": "one K^-1 application to 4096 rows of length 4096 (algorithmic 2*N^2*8 = 268.4 MB)",
                   "source": os.path.basename(dst)}, open(traffic_json, "w"), indent=1)
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "kernels":
        kernels(*sys.argv[2:])
        sys.exit(0)
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    else:
        kernel(sys.argv[2], sys.argv[3])
