"""Developer tool: turn the scratch outputs of scripts/gpu_round.sh (gpurun_out/) into the tracked summaries under
profiles/ (launch list shares, ncu --set full digests of the two hot kernels, the bench line)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def ncu_digest(rep, title, notes):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# {title}", "", notes, "", f"Source: `ncu --set full --clock-control none --import-source on` ({os.path.basename(rep)}, "
             f"{len(data)} launches captured; values of the first launch).", "", "| metric | value | unit |", "|---|---|---|"]
    r = data[0]
    lines.insert(2, f"Kernel: `{r[hdr.index('Kernel Name')]}`")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {r[i]} | {units[i]} |")
    return "\n".join(lines) + "\n", {k: r[hdr.index(k)] for k in KEYS if k in hdr}


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ix = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[start:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}[r[ix["Metric Unit"]]]
        name = r[ix["Kernel Name"]]
        tot[name] += v
        cnt[name] += 1
    return tot, cnt


def main():
    os.makedirs(OUT, exist_ok=True)
    tot, cnt = launch_shares(os.path.join(SRC, "launches.csv"))
    T = sum(tot.values())
    ours = sum(v for k, v in tot.items() if "i2v::" in k)
    with open(os.path.join(OUT, f"{TAG}_launches_step.md"), "w") as f:
        f.write(f"# Launch list of one denoise step ({TAG})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python bench.py --steps 1 "
                "--warmup 3 --no-cpu-baseline --no-e2e --eager --profiler-range` (one timed step between cudaProfilerStart/Stop).\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                f"Total {T:.2f} ms over {sum(cnt.values())} launches; library kernels (`i2v::*`) {ours:.2f} ms = "
                f"{ours / T * 100:.1f} % in {sum(c for k, c in cnt.items() if 'i2v::' in k)} launches.\n\n"
                "| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in tot.most_common(40):
            f.write(f"| {v:.3f} | {v / T * 100:.1f} % | {cnt[k]} | `{k[:150]}` |\n")
    t_rep = os.path.join(SRC, "prof_temporal.ncu-rep")
    d_txt, d = ncu_digest(os.path.join(SRC, "prof_dense.ncu-rep"), f"Dense attention kernel, level 0 ({TAG})",
                          "Fused spatial self-attention + I2V-Adapter cross-frame attention of one level-0 block at the C2 size "
                          "(32 frames x 8 heads x S=4096 x d=40, two problems, 1.374 TFLOP algorithmic per launch) on the "
                          "augmented operand layout, launched by `scripts/perf_aug.py`.  Expected before measuring: "
                          "exponential-bound (MUFU.EX2 16/clk/SM -> 1024 clk per 128x128 score tile vs 384 clk of MMA; 3 of 8 "
                          "column pairs on the FMA-pipe polynomial) -> XU and FMA pipes both busy, tensor pipe ~30 %, "
                          "DRAM << peak (K/V re-reads served by L2).")
    t_txt, t = (None, {}) if not os.path.exists(t_rep) else ncu_digest(t_rep, f"Temporal attention kernel, level 0 ({TAG})",
                          "Motion-module temporal self-attention at the C2 level-0 size (8192 positions x 16 frames x 8 heads x "
                          "d=40, 335.5 MB algorithmic bytes per launch), launched by `scripts/gpu_selftest.py --run perftemporal`.  "
                          "Expected before measuring: HBM-bound, DRAM traffic ~ algorithmic bytes, tensor/ALU pipes mostly idle.")
    ip_rep = os.path.join(SRC, "prof_ip.ncu-rep")
    if os.path.exists(ip_rep):
        i_txt, i = ncu_digest(ip_rep, f"IP-Adapter cross-attention kernel, level 0 ({TAG})",
                              "Decoupled text (77) + image (4) cross-attention at the C2 level-0 size (32 frames x 4096 "
                              "queries x 8 heads x d=40; reads Q and writes O once: 167.8 MB algorithmic bytes per launch), "
                              "launched by `scripts/perf_ip_one.py`: `ip_xattn_tc_kernel` (tcgen05, K / V of one (video, head) resident, "
                              "three query tiles in flight, Q through cp.async, output through a staging tile).  Expected before "
                              "measuring: HBM traffic ~ algorithmic (Q read once, O stays in L2), bound by the serial QK -> softmax "
                              "-> PV -> epilogue chain of a tile, not by a pipe.")
        open(os.path.join(OUT, f"{TAG}_ip_attn_l0.md"), "w").write(i_txt)
        print("ip:", i.get("gpu__time_duration.sum"), "dram r/w", i.get("dram__bytes_read.sum"), i.get("dram__bytes_write.sum"))
    ff_rep = os.path.join(SRC, "prof_ff.ncu-rep")
    if os.path.exists(ff_rep):
        f_txt, ff = ncu_digest(ff_rep, f"Feed-forward projection + GEGLU GEMM, level 0 ({TAG})",
                               "`ff_geglu_gemm_kernel<2>` (CTA pairs, tcgen05 cta_group::2) at the C2 level-0 shape: "
                               "131072 x 320 -> 1280 (+ ones column), 215 GFLOP and 84 MB in / 335 MB out algorithmic per "
                               "launch, launched by `scripts/perf_ff_one.py`.  Expected before measuring: tensor-pipe "
                               "bound in the mainloop (the epilogue-free kernel runs 130 us = 1650 TFLOP/s), the epilogue "
                               "(TMEM -> GELU -> staged TMA store) not fully hidden at K = 320.")
        open(os.path.join(OUT, f"{TAG}_ff_geglu_gemm_l0.md"), "w").write(f_txt)
        print("ff:", ff.get("gpu__time_duration.sum"), "dram r/w", ff.get("dram__bytes_read.sum"), ff.get("dram__bytes_write.sum"))
    open(os.path.join(OUT, f"{TAG}_dense_attn_l0.md"), "w").write(d_txt)
    if t_txt:
        open(os.path.join(OUT, f"{TAG}_temporal_attn_l0.md"), "w").write(t_txt)
    for n in (2, 4, 8):
        path = os.path.join(SRC, f"bench_n{n}.log")
        if os.path.exists(path):
            line = open(path).read().strip().splitlines()[-1]
            json.loads(line)
            open(os.path.join(OUT, f"{TAG}_bench_n{n}.json"), "w").write(line + "\n")
    bench = open(os.path.join(SRC, "bench.log")).read().strip().splitlines()[-1]
    json.loads(bench)
    open(os.path.join(OUT, f"{TAG}_bench_n1.json"), "w").write(bench + "\n")
    print("dense:", d.get("gpu__time_duration.sum"), "dram r/w", d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"))
    print("temporal:", t.get("gpu__time_duration.sum"), "dram r/w", t.get("dram__bytes_read.sum"), t.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main()
