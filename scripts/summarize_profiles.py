#!/usr/bin/env python
"""Turn the scratch artefacts a GPU session left in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01

  profiles/<round>_launches.txt        per-kernel device time and SHARE of the step (ncu launch list, cold + serialised)
  profiles/<round>_<kernel>_ncu.txt    selected metrics of the `ncu --set full` capture of a dominant kernel
  profiles/<round>_microbench.jsonl    instruction-rate microbenchmarks (POPC / LOP3 / dp4a / IMMA / HBM read)
  profiles/microbench.json             the same as a JSON list (bench.py reads the measured POPC rate from it)
  profiles/<round>_bench*.json         bench.py lines of that session
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "launch__cluster_size", "launch__cluster_max_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
]


def launches(tag, fname="launches.csv", suffix=""):
    src = os.path.join(OUT, fname)
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    tot = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r[vi].replace(",", "")) * scale[r[ui]]
    allms = sum(t[1] for t in tot.values())
    ours = sum(t[1] for k, t in tot.items() if "vdf::" in k)
    with open(os.path.join(PROF, f"{tag}_launches{suffix}.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): compare SHARES\n")
        f.write(f"# command: see scripts/gpu_round.sh (ncu step); {sum(t[0] for t in tot.values())} launches captured, "
                f"{allms:.3f} ms total, {100 * ours / allms:.2f}% in vdf:: kernels\n")
        f.write(f"{'ms':>12} {'share':>8} {'launches':>8}  kernel\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{ms:12.3f} {100 * ms / allms:7.2f}% {n:8d}  {k[:140]}\n")


def _bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    try:
        return float(v.replace(",", "")) * scale[unit]
    except Exception:
        return None


def ncu_summary(tag, rep, name):
    """name may carry the workload key after a double underscore (prof_<kernel>__<key>.ncu-rep): then the DRAM traffic of
    the captured launch goes to profiles/traffic.json[kernel][key] (bench.py's roofline.traffic)."""
    src = os.path.join(OUT, rep)
    if not os.path.exists(src):
        return
    key = None
    if "__" in name:
        name, key = name.split("__", 1)
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, f"{tag}_{name}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, {rep}; selected metrics per captured launch\n")
        for vals in rows[2:]:
            rec = dict(zip(hdr, vals))
            urec = dict(zip(hdr, units))
            if key:
                rd = _bytes(rec.get("dram__bytes_read.sum", ""), urec.get("dram__bytes_read.sum", ""))
                wr = _bytes(rec.get("dram__bytes_write.sum", ""), urec.get("dram__bytes_write.sum", ""))
                if rd is not None and wr is not None:
                    tf = os.path.join(PROF, "traffic.json")
                    t = json.load(open(tf)) if os.path.exists(tf) else {}
                    kname = re.sub(r"<.*", "", re.sub(r"\(.*", "", rec.get("Kernel Name", "?"))).split("::")[-1].strip().split(" ")[-1]
                    t.setdefault(kname, {})[key] = rd + wr
                    json.dump(t, open(tf, "w"), indent=1, sort_keys=True)
            f.write(f"\n## {rec.get('Kernel Name', '?')}  grid {rec.get('Grid Size', '?')} block {rec.get('Block Size', '?')}\n")
            for h, u, v in zip(hdr, units, vals):
                if any(h == k or (k in h and h.endswith(k)) for k in KEEP):
                    f.write(f"{h} [{u}] = {v}\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    for f in sorted(os.listdir(OUT)):
        if f.startswith("launches_") and f.endswith(".csv"):
            launches(tag, f, "_" + f[len("launches_"):-4])
    for rep in sorted(os.listdir(OUT)):
        if rep.endswith(".ncu-rep"):
            ncu_summary(tag, rep, rep[:-8].replace("prof_", ""))
    mb = os.path.join(OUT, "microbench.jsonl")
    if os.path.exists(mb):
        shutil.copy(mb, os.path.join(PROF, f"{tag}_microbench.jsonl"))
        recs = [json.loads(l) for l in open(mb) if l.strip().startswith("{")]
        json.dump(recs, open(os.path.join(PROF, "microbench.json"), "w"), indent=1)
    for f in sorted(os.listdir(OUT)):
        if f.startswith("bench") and f.endswith(".json") and os.path.getsize(os.path.join(OUT, f)) > 10 and "under_ncu" not in f:
            if f.startswith("bench_x"):  # kernel experiments (option sweeps): kept apart from the headline evidence
                os.makedirs(os.path.join(PROF, "experiments"), exist_ok=True)
                shutil.copy(os.path.join(OUT, f), os.path.join(PROF, "experiments", f"{tag}_{f}"))
                continue
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f"{tag}_{f}"))
    for f in sorted(os.listdir(OUT)):
        if f.startswith("dist_check") and f.endswith(".log"):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f"{tag}_{f}"))
    for f in ("gpu.csv", "host.txt", "pytest_gpu.log", "smoke.log", "sanitizer_memcheck.log"):
        if os.path.exists(os.path.join(OUT, f)):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f"{tag}_{f}"))
    print("profiles/:", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
