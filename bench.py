#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line.  Primary workload (`--workload search`, the default): all-pairs `search` over 1 M synthetic
hashes with equal durations (BASELINE.json configs[2], the configuration "Hamming pair-comparisons/s" is quoted
on); a step is one full pass: pair-matrix kernel -> edge all-gather -> sort -> greedy MatchGroups, inputs resident
in HBM.  The second headline metric (frame-stacks hashed/s, configs[1]: 1080p stacks resident in HBM) is measured
in the same run and reported under "secondary" with its own HBM roofline; `--workload hash` makes it primary.
`value` is device time (CUDA events on the stream the kernels run on, max over ranks); `e2e` goes through the
public API with host buffers; `cpu_baseline` times the oracle port on a bounded sample on this box's host cores.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIR_POPC32 = 32  # algorithmic POPC32 per pair: 16 x POPCNT64 over all 1024 stored bits (video_hash.rs:311-317)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="search", choices=["search", "hash", "refs", "e2e10m", "popc"])
    ap.add_argument("--n-query", type=int, default=100_000)
    ap.add_argument("--n-corpus", type=int, default=10_000_000)
    ap.add_argument("--n-e2e", type=int, default=10_000_000, help="hashes of the end-to-end dedup pipeline (BASELINE configs[4])")
    ap.add_argument("--n-popc", type=int, default=262_144, help="hashes of the XOR+POPC line (secondary_popc)")
    ap.add_argument("--n-hashes", "--n", dest="n", type=int, default=1_000_000, help="hashes in the all-pairs search")
    ap.add_argument("--tol", type=float, default=0.35)
    ap.add_argument("--tol-sweep", default="0.0,0.1,0.2,0.3,0.35,0.4,0.42", help="comma-separated tolerances of the sweep ('' = none)")
    ap.add_argument("--variant", type=int, default=-1, help="search kernel variant (-1: library default)")
    ap.add_argument("--hash-variant", type=int, default=-1, help="-1 (default) the fused persistent kernel; per-frame kernels: 0 IMMA 4 warps, 1 general, 2 IMMA 8 warps")
    ap.add_argument("--stacks", type=int, default=256, help="1080p stacks resident in HBM per GPU (8.5 GB at 256)")
    ap.add_argument("--hash-total", type=int, default=100_000, help="stack-hashes per hashing run (BASELINE configs[1]: 100 k stacks)")
    ap.add_argument("--mismatch-stacks", type=int, default=1024, help="stacks compared bit by bit with the CPU oracle")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU edge exchange: fused into the pair kernel over NVLink peer memory, or an NCCL all-gather")
    ap.add_argument("--plane", default="both", choices=["torch", "c", "both"],
                    help="N > 1: one process per GPU over torch.distributed, the in-library plane (rank 0 drives all GPUs through one "
                         "vdf_ctx_create_multi context, reported as e2e_cplane), or both")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=INT", help="extra vdf_ctx_set_option (kernel experiments)")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--secondary", default="hash,popc,refs,e2e10m", help="secondary records attached to the search line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-rows", type=int, default=96, help="rows brute-forced by the CPU oracle (untimed) per search record")
    ap.add_argument("--e2e-steps", type=int, default=3)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML from a
    thread of this process (nvidia_ml_py): a looping `nvidia-smi` child stalls kernel launches for milliseconds at a
    time, which shows up as idle gaps in a device-timed step; nvidia-smi remains the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = float(os.environ.get("VDF_CLOCK_PERIOD_S", "0.1"))):
        self.index, self.period, self.proc, self.lines = index, period_s, None, []
        self.sm, self.mx, self.reasons, self.stop, self.t, self.how = [], [], set(), threading.Event(), None, None

    def _nvml_loop(self, nv, h):
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self.stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(mx)
                r = int(reasons_fn(h))
                self.reasons.update(k for k, bit in names.items() if r & bit)
            except Exception:
                pass
            self.stop.wait(self.period)

    def __enter__(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            self.how = "nvml"
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": self.how}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "how": self.how}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


DEFAULT_SEARCH_VARIANT = 6   # vdf_ctx::search_variant in csrc/common.cuh
PAIR_MACS = 1024             # tensor-core variants: hamming = pc(a) + pc(b) - 2 <a, b>, one u8 MAC per stored bit
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from `ncu --set full` captures


def ncu_traffic(kernel: str, key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` at workload `key`, recorded from an
    `ncu --set full` capture (scripts/summarize_profiles.py writes profiles/traffic.json); None if not captured."""
    try:
        return json.load(open(TRAFFIC_FILE))[kernel][key]
    except Exception:
        return None


def search_roofline(variant: int, pairs_per_launch: float, k_ms: float, k_n: int, sm_count: int, sm_max_mhz: float, n_key=None):
    ms = k_ms / max(k_n, 1)
    if variant >= 3:
        # tensor-bound: 2 x 1024 operand-width integer ops per pair on tcgen05.mma (kind::i8, or kind::mxf4 for variant 6)
        p = os.path.join(ROOT, "MEASURED_PEAKS.json")
        bf16 = json.load(open(p)).get("bf16_tflops") if os.path.exists(p) else None
        rate = 4.0 if variant == 6 else 2.0   # 4-bit operands issue at four times the bf16 rate, 8-bit at twice
        macs = 16384 if variant == 6 else 8192  # MAC/clk/SM: 128 x 256 x {64 | 32} per 128 clk
        what = "4-bit" if variant == 6 else "8-bit"
        bf16_peak, bf16_src = (rate * bf16, f"{rate:g} x measured dense bf16 TFLOP/s (MEASURED_PEAKS.json, burst): {what} operands run "
                               f"at {rate:g}x the bf16 rate") if bf16 else (2250.0 * rate, f"fallback: nominal dense {what} peak (B200_PROFILING.md)")
        pipe = sm_count * macs * 2 * sm_max_mhz * 1e6 / 1e12
        # peak = the tcgen05 issue rate measured on this pool's B200 by csrc/microbench.cu (same instruction shape, same operand
        # encoding, resident operands, nothing else running; sustained run) -- the way the POPC roofline is defined
        peak, src = umma_peak_tops("tcgen05_mxf4_2cta_m256n192k64" if variant == 6 else "tcgen05_i8_2cta_m256n256k32")
        if peak is None:
            peak, src = bf16_peak, bf16_src
        achieved = 2.0 * PAIR_MACS * pairs_per_launch / (ms * 1e-3) / 1e12 if k_n else None
        kname = {3: "hamming_tc_kernel", 4: "hamming_tc2_kernel", 5: "hamming_tc5_kernel", 6: "hamming_tc6_kernel"}[variant]
        return {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(kname, n_key) if n_key else None,
                "peak_source": src, "pipe_peak": pipe, "pipe_frac": (achieved / pipe) if achieved else None,
                "pipe_peak_source": f"148 SMs x {macs} {what} MAC/clk/SM x 2 x max SM clock (nominal tcgen05 issue floor, B300_MICROARCH.md)",
                "cublas_scaled_peak": bf16_peak, "cublas_scaled_frac": (achieved / bf16_peak) if achieved else None,
                "cublas_scaled_peak_source": bf16_src,
                "kernel_ms_per_launch": ms, "kernel_launches_timed": k_n, "algorithmic_ops_per_pair": 2 * PAIR_MACS,
                "popc_equivalent": {"algorithmic_popc32_per_pair": PAIR_POPC32,
                                    "frac_of_popc_peak": PAIR_POPC32 * pairs_per_launch / (ms * 1e-3) / 1e9 /
                                    popc_peak_gpopc(sm_count, sm_max_mhz)[0] if k_n else None}}
    popc_peak, popc_src = popc_peak_gpopc(sm_count, sm_max_mhz)
    achieved = PAIR_POPC32 * pairs_per_launch / (ms * 1e-3) / 1e9 if k_n else None
    kname = "hamming_tiles_csa4_kernel" if variant == 2 else "hamming_tiles_kernel"
    return {"bound": "int_popc", "kernel": kname, "achieved": achieved, "peak": popc_peak,
            "unit": "GPOPC32/s", "frac": (achieved / popc_peak) if achieved else None,
            "traffic": ncu_traffic(kname, n_key) if n_key else None,
            "peak_source": popc_src, "kernel_ms_per_launch": ms, "kernel_launches_timed": k_n,
            "algorithmic_popc32_per_pair": PAIR_POPC32}


def umma_peak_tops(op: str):
    """Measured tcgen05.mma issue rate (2 ops per MAC) from profiles/microbench.json; (None, None) if not recorded."""
    p = os.path.join(ROOT, "profiles", "microbench.json")
    try:
        recs = [r for r in json.load(open(p)) if r.get("op") == op]
        best = [r for r in recs if r.get("run") == "sustained"] or recs
        return best[0]["ops_per_s"] / 1e12, f"measured {op} issue rate, sustained run (csrc/microbench.cu -> profiles/microbench.json)"
    except Exception:
        return None, None


def popc_peak_gpopc(sm_count: int, sm_max_mhz: float):
    """POPC32 results per second.  Prefer the rate measured on this pool's B200 by the microbenchmark
    (profiles/microbench.json, op popc_add); else the documented 16 results/clk/SM."""
    p = os.path.join(ROOT, "profiles", "microbench.json")
    if os.path.exists(p):
        try:
            for rec in json.load(open(p)):
                if rec.get("op") == "popc_add":
                    return rec["ops_per_s"] / 1e9, "measured POPC rate (profiles/microbench.json)"
        except Exception:
            pass
    return sm_count * 16 * sm_max_mhz * 1e6 / 1e9, "documented 16 POPC/clk/SM x SMs x max SM clock"


# ------------------------------------------------------------------------------------------------ shared by both arms
def workload_name(kind: str, args) -> str:
    """config.workload, identical in our arm and in the reference arm"""
    if kind == "search":
        return f"all-pairs search (find_all_matches), {args.n} synthetic hashes, equal durations, tolerance {args.tol}"
    if kind == "popc":
        return f"all-pairs search (find_all_matches), {args.n_popc} synthetic hashes, equal durations, tolerance {args.tol}"
    if kind == "refs":
        return (f"search_with_references, {args.n_query} queries x {args.n_corpus} sorted table entries, equal durations, "
                f"tolerance {args.tol}")
    if kind == "e2e10m":
        return (f"end-to-end dedup, {args.n_e2e} frame stacks ({args.width}x{args.height}x16 u8, letterbox cropdetect) -> hashes -> "
                f"all-pairs search -> MatchGroups, tolerance {args.tol}")
    return f"frame-stack hashing, {args.width}x{args.height}x16 u8 stacks resident in HBM, letterbox cropdetect"


def digest(*arrays) -> str:
    """one short hex string over result arrays (sorted edge keys, group CSR, ...): equal at every GPU count or the runs differ"""
    import hashlib

    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode())
        h.update(a.view(np.uint8).reshape(-1).data)
    return h.hexdigest()[:16]


class FixedWidthPaths:
    """"v/%08d" % (start + perm[i]) for i in range(n) without ten million str objects: a sequence for MatchGroups plus the
    (blob, offsets) the C ABI takes"""

    def __init__(self, ids: np.ndarray):
        self.ids = np.asarray(ids, dtype=np.int64)

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        return "v/%08d" % int(self.ids[i])

    def blob(self):
        n = len(self.ids)
        b = np.empty((n, 10), dtype=np.uint8)
        b[:, 0], b[:, 1] = ord("v"), ord("/")
        v = self.ids.copy()
        for k in range(8):
            b[:, 9 - k] = ord("0") + (v % 10)
            v //= 10
        return b.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(10)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path = the oracle port (the reference is Rust; no cargo here), on a
    bounded SAMPLE of our arm's workload (same config.workload string; `sample` says what was timed).
    search: Search::search_self's inner loop is single-threaded (no rayon in vid_dup_finder_lib) -> 1 thread;
    hash:   the app hashes files on a rayon pool -> all host cores."""
    from oracle import vdf_oracle as o
    from tests import synth

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = args.workload
    if kind in ("search", "popc"):
        n = args.n if kind == "search" else args.n_popc
        H, _ = synth.planted_hashes(n)
        dur = np.full(n, 600, np.uint32)
        tol = o.tolerance_int(args.tol)
        rows_per_step = max(8, int(2.0e9 // n))  # ~2e9 pairs per step: a few seconds of one core
        rng = np.random.default_rng(1)

        def step():
            rows = rng.integers(0, n, rows_per_step)
            o.search_refs(H, dur, H[rows], dur[rows], tol)
            return rows_per_step * n

        metric, unit, used = "hamming_pair_comparisons_per_s", "pairs/s", 1
        sample = f"sampled: {rows_per_step} random rows x {n} candidates per step (same table, tol {tol}), 1 thread"
    elif kind == "refs":
        # search_with_references: every query walks its whole duration slice of the sorted table (search_algorithm.rs:40-53)
        nc = args.n_corpus
        n_tab = min(nc, 1_000_000)  # a bounded slice of the table: the cost per (query, entry) pair does not depend on its length
        H, _ = synth.planted_hashes(n_tab)
        dur = np.full(n_tab, 600, np.uint32)
        Q = synth.random_hashes(max(8, int(2.0e9 // n_tab)), seed=synth.SEED + 1)
        qd = np.full(len(Q), 600, np.uint32)
        tol = o.tolerance_int(args.tol)

        def step():
            o.search_refs(H, dur, Q, qd, tol)
            return len(Q) * n_tab

        metric, unit, used = "hamming_pair_comparisons_per_s", "pairs/s", 1
        sample = f"sampled: {len(Q)} queries x {n_tab} table entries per step (tol {tol}), 1 thread"
    else:  # hash (and the hashing half of e2e10m, which the reference arm does not extrapolate)
        from concurrent.futures import ThreadPoolExecutor

        kind = "hash" if kind != "e2e10m" else kind
        per = max(cores, 8)
        st = synth.frame_stacks(per, args.width, args.height).numpy()

        def one(s):
            return o.hash_stack(st[s], 1)[0]

        def step():
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(one, range(per)))
            return per

        metric, unit, used = "frame_stacks_hashed_per_s", "stacks/s", cores
        sample = f"sampled: {per} synthetic {args.width}x{args.height} stacks per step, {cores} threads"
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    units = sum(step() for _ in range(args.steps))
    dt = time.perf_counter() - t0
    v = units / dt
    emit({
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": {"workload": workload_name(kind, args), "sampled": sample},
        "cpu_baseline": {"value": v, "unit": unit, "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------ our arm
def emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything libraries print (NCCL banner, torchrun notes) was
    re-routed to stderr by main()"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints from native libraries must not pollute the JSON line
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import vid_dup_finder_lib_b200 as vdf
    from tests import synth
    from vid_dup_finder_lib_b200 import _ffi
    from vid_dup_finder_lib_b200 import dist as vdist
    from vid_dup_finder_lib_b200.definitions import tolerance_to_int

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers while rank 0 drives every GPU (--plane c)
    ctx = _ffi.default_context()
    if args.variant >= 0:
        ctx.set_option("search_variant", args.variant)
    if args.hash_variant >= 0:
        ctx.set_option("hash_fused", 0)
        ctx.set_option("hash_variant", args.hash_variant)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    fused = world > 1 and args.exchange == "peer" and (args.variant in (-1, 6))
    if fused:
        try:
            vdist.enable_peer_exchange(ctx, capacity=1 << 22)
        except Exception as e:  # e.g. CUDA IPC unavailable in this container: all ranks fail alike and use the collective
            print(f"[bench] peer exchange unavailable ({e!r}); falling back to the NCCL all-gather", file=sys.stderr)
            fused = False
    exchange_note = ("edges appended to every rank's buffer by the pair kernel over NVLink peer memory" if fused else
                     "NCCL all-gather of per-rank edge lists") if world > 1 else "single GPU"
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    L2_BYTES = int(torch.cuda.get_device_properties(local).L2_cache_size)
    variant = args.variant if args.variant >= 0 else DEFAULT_SEARCH_VARIANT
    dtype_of = lambda v: "e2m1 x e2m1 -> f32 (exact)" if v == 6 else "u8 x u8 -> s32" if v >= 3 else "u32"  # noqa: E731

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, steps, warmup, flush_l2):
        """W warm-up steps, then K steps between two CUDA events on the kernels' stream, barrier + synchronize on
        both sides, max over ranks.  -> seconds.  flush_l2 (workloads whose inputs fit the L2): a 512 MiB fill, ~0.15 ms,
        INSIDE the timed region - the reported time is conservative by that much per step."""
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step()
            if flush_l2:
                flush.fill_(1)  # the first launch of torch's fill kernel loads its module (5 ms .. 1 s): not inside the timing
            gc.collect()
            gc.disable()  # a generation-2 collection between two launches would show up as GPU idle time
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                if flush_l2:
                    flush.fill_(1)
                step()
            e1.record(stream)
            barrier()
            gc.enable()
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    def parity_rows_self(H, dur, keys_np, tol_int, k_rows, seed=9):
        """K random rows of the pair matrix brute-forced by the CPU oracle against the edges the GPU returned (untimed)"""
        from oracle import vdf_oracle as o

        n = len(dur)
        rows = np.unique(np.random.default_rng(seed).integers(0, n, k_rows))
        rp, ci = o.search_refs(H, dur, H[rows], dur[rows], tol_int)
        ki, kj = (keys_np >> np.uint64(32)).astype(np.int64), (keys_np & np.uint64(0xFFFFFFFF)).astype(np.int64)
        bad = 0
        for k, r in enumerate(rows.tolist()):
            want = set(int(c) for c in ci[rp[k]:rp[k + 1]] if c != r)
            got = set(kj[ki == r].tolist()) | set(ki[kj == r].tolist())
            bad += got != want
        return {"rows": int(len(rows)), "candidates_per_row": int(n), "mismatching_rows": int(bad), "ok": bad == 0,
                "how": "oracle brute force of sampled rows vs the GPU's edge list, untimed"}

    # ---------------------------------------------------------------- search workload
    def bench_search(steps, warmup, n=None, use_variant=None, light=False):
        """all-pairs search on a resident, prepared table (`Search::from` done once, `search_self` per step).
        light: the XOR+POPC line (no e2e, no sweep)."""
        n = n or args.n
        tol_int = tolerance_to_int(args.tol)
        v = variant if use_variant is None else use_variant
        if use_variant is not None:
            ctx.set_option("search_variant", use_variant)
        use_exchange = fused and v == 6
        H, _ = synth.planted_hashes(n)
        dur = np.full(n, 600, np.uint32)
        pairs = n * (n - 1) // 2
        with torch.cuda.stream(stream):
            d_hash = torch.from_numpy(H.view(np.int64)).to(dev)
            d_dur = torch.from_numpy(dur.view(np.int32)).to(dev)
            torch.cuda.current_stream().synchronize()
            tbl = ctx.table_create_device(d_hash.data_ptr(), d_dur.data_ptr(), n, keepalive=(d_hash, d_dur))
        # what one launch reads on this rank: the 4-bit operand tiles (512 B per hash; its share of the rows and every column)
        # for the tensor-core kernels, the raw hashes (128 B) for the XOR+POPC ones.  Larger than twice the L2: nothing of a
        # step survives to the next one and no flush is needed; else a 512 MiB write sits between the steps, inside the timing.
        operand_bytes = (n // world + n) * 512 if v >= 5 else n * 128
        flush_l2 = operand_bytes < 2 * L2_BYTES
        l2_note = (f"operands read per launch {operand_bytes / 1e6:.0f} MB > 2 x L2 ({L2_BYTES / 1e6:.0f} MB): no flush" if not flush_l2 else
                   f"operands {operand_bytes / 1e6:.0f} MB fit L2: 512 MiB write between timed steps, inside the timed region")
        result = {}
        saved_world = ctx.peer_world
        if not use_exchange:
            ctx.peer_world = 0  # search_self_keys then merges with the NCCL all-gather

        def step(t=tol_int):
            keys = vdist.search_self_keys(ctx, tbl, None, t, device=dev)
            torch.cuda.current_stream().synchronize()
            gp, mm = ctx.group_greedy_device(n, keys.data_ptr(), keys.numel())
            result["keys"], result["gp"], result["mm"] = keys, gp, mm

        try:
            with ClockSampler(local) as cs:
                with torch.cuda.stream(stream):
                    for _ in range(warmup):
                        step()
                ctx.kernel_time(0, reset=True)
                l0 = ctx.counters()[0]
                secs = timed(step, steps, 0, flush_l2)
            launches = ctx.counters()[0] - l0
            k_ms, k_n = ctx.kernel_time(0, reset=True)
            value = pairs * steps / secs
            roof = search_roofline(v, pairs / world, k_ms, k_n, sm_count, sm_max_mhz, n_key=f"self_{n}_x{world}")
            keys_np = result["keys"].cpu().numpy().view(np.uint64)
            out = {"metric": "hamming_pair_comparisons_per_s", "value": value, "unit": "pairs/s", "ms_per_step": secs / steps * 1e3,
                   "scaling": "strong", "dtype": dtype_of(v), "roofline": roof, "gpu_launches": int(launches), "clocks": cs.summary(),
                   "result_digest": digest(keys_np, result["gp"], result["mm"]),
                   "config": {"workload": workload_name("search" if n == args.n else "popc", args), "n_hashes": n, "tol_int": tol_int,
                              "pairs_per_step": pairs, "edges": int(len(keys_np)), "groups": int(len(result["gp"]) - 1),
                              "parallelism": f"tile-block shard x{world}", "exchange": exchange_note if v == 6 else "NCCL all-gather" if world > 1 else "single GPU",
                              "table": "prepared once (vdf_table: packed tiles, windows, work units), searched per step",
                              "l2": l2_note,
                              "search_variant": v}}
            if rank == 0 and not args.no_cpu_baseline and args.parity_rows > 0:
                out["parity_sample"] = parity_rows_self(H, dur, keys_np, tol_int, args.parity_rows if not light else min(args.parity_rows, 32))
            if light:
                return out
            # ---- tolerance sweep (BASELINE configs[2] is a sweep): one warm + two timed steps per tolerance
            sweep = []
            for ts in [x for x in args.tol_sweep.split(",") if x.strip()]:
                t_int = tolerance_to_int(float(ts))
                with torch.cuda.stream(stream):
                    step(t_int)
                ctx.kernel_time(0, reset=True)
                # three steps, timed one by one: the median is the line's figure, all three are kept (a box that hiccups once -- one
                # 1.5 s stall was seen in one of ~40 sweeps -- shows up in ms_each, not in the median)
                each = sorted(timed(lambda: step(t_int), 1, 0, flush_l2) * 1e3 for _ in range(3))
                sk_ms, sk_n = ctx.kernel_time(0, reset=True)
                r = search_roofline(v, pairs / world, sk_ms, sk_n, sm_count, sm_max_mhz)
                sweep.append({"tolerance": float(ts), "ms_per_step": each[1], "ms_each": each, "kernel_ms": r["kernel_ms_per_launch"],
                              "frac": r["frac"], "edges": int(result["keys"].numel()), "groups": int(len(result["gp"]) - 1)})
            out["tol_sweep"] = sweep
        finally:
            ctx.peer_world = saved_world
            if use_variant is not None:
                ctx.set_option("search_variant", variant)
            tbl.close()
        # ---- e2e: host arrays -> public API (keys cut + GPU sort, H2D, kernels, D2H) -> MatchGroups
        e_steps = max(1, args.e2e_steps)
        # the caller's order is arbitrary: a fixed shuffle of the table, so that the (duration, Path) sort has real work
        paths = synth.paths(n)
        perm = np.random.default_rng(7).permutation(n)
        table = vdf.HashTable(np.ascontiguousarray(H[perm]), dur[perm], [paths[i] for i in perm.tolist()])
        table.path_blob()  # the table owns its struct-of-arrays buffers (hashes, durations, path blob) before the call
        vdist.search(table, args.tol, ctx=ctx)  # warm-up: pinned staging buffers get allocated
        c0 = ctx.counters()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            groups = vdist.search(table, args.tol, ctx=ctx)
        barrier()
        e_secs = max_over_ranks(time.perf_counter() - t0)
        c1 = ctx.counters()
        ph = ctx.last_phases()
        t0 = time.perf_counter()
        n_members = sum(g.len() for g in groups)  # every MatchGroup object, with its path strings, built once
        mat_ms = (time.perf_counter() - t0) * 1e3
        h2d = torch.tensor([c1[1] - c0[1], c1[2] - c0[2]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(h2d, op=dist.ReduceOp.MAX)  # rank 0 stages the table; the others receive it over NVLink
        out["e2e"] = {"value": pairs * e_steps / e_secs, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d[0].item() // e_steps),
                      "d2h_bytes_per_step": int(h2d[1].item() // e_steps), "steps": e_steps, "groups": len(groups), "ms_per_call": e_secs / e_steps * 1e3,
                      "api": "vid_dup_finder_lib_b200.dist.search(HashTable, tolerance) -> MatchGroups (lazy: objects are built on access)",
                      "input_order": "shuffled (fixed permutation of the synthetic table)",
                      "phases_ms_last_call": {"sort_keys_host": ph[0], "pinned_copy_h2d_gpu_sort": ph[1], "device_incl_d2h": ph[2],
                                              "note": "rank 0's vdf_search / vdf_stage_sorted (csrc/host.cu); at N > 1 the search itself follows the broadcast"},
                      "materialize_all_groups_ms": mat_ms, "members": int(n_members),
                      "result_digest": digest(*groups.csr)}
        # ---- the same call through ONE multi-device context: rank 0 drives all GPUs, the other ranks sit at a host barrier
        if world > 1 and args.plane in ("c", "both"):
            barrier()
            cp = None
            if rank == 0:
                try:
                    m = _ffi.Context(list(range(world)))
                    vdf.search(table, args.tol, ctx=m)
                    t0 = time.perf_counter()
                    for _ in range(e_steps):
                        g2 = vdf.search(table, args.tol, ctx=m)
                    dt = time.perf_counter() - t0
                    cp = {"value": pairs * e_steps / dt, "unit": "pairs/s", "ms_per_call": dt / e_steps * 1e3, "groups": len(g2),
                          "result_digest": digest(*g2.csr), "devices": m.device_count,
                          "api": "vdf_search on a vdf_ctx_create_multi context: one process, one C call, all GPUs"}
                    m.close()
                except Exception as e:
                    cp = {"error": repr(e)}
            dist.barrier(group=cpu_group)
            if cp is not None:
                out["e2e_cplane"] = cp
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import vdf_oracle as o

            rows = max(8, int(4.0e9 // n))
            sel = np.random.default_rng(1).integers(0, n, rows)
            t0 = time.perf_counter()
            o.search_refs(H, dur, H[sel], dur[sel], tol_int)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": rows * n / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                                   "sample": f"{rows} random rows x {n} candidates of the same table, oracle port, 1 thread "
                                             "(the reference's search is single-threaded)"}
        return out

    # ---------------------------------------------------------------- query-vs-reference workload (BASELINE configs[3])
    def bench_refs(steps, warmup):
        """search_with_references: --n-query hashes against a sorted table of --n-corpus hashes whose contiguous slices
        are spread over the ranks (each rank: all queries x its prepared slice); keys reach every rank through the exchange."""
        nq, nc, tol_int = args.n_query, args.n_corpus, tolerance_to_int(args.tol)
        with torch.cuda.stream(stream):
            corpus, _ = synth.planted_hashes_torch(nc, device=dev)
            d_q, _ = synth.planted_hashes_torch(nq, seed=synth.SEED + 1, device=dev, dup_frac_den=1 << 40)  # random queries ...
            hit = torch.arange(0, nq, 20, device=dev)  # ... 5 % of them perturbed copies of corpus entries
            pick = (synth._t_stream(synth.SEED + 2, hit, 0) & 0x7FFFFFFFFFFFFFFF) % nc
            flip = synth._t_stream(synth.SEED + 2, hit[:, None] * 16 + torch.arange(16, device=dev)[None, :], 1)
            flip = flip & synth._t_stream(synth.SEED + 2, hit[:, None] * 16 + torch.arange(16, device=dev)[None, :], 2)
            flip = flip & synth._t_stream(synth.SEED + 2, hit[:, None] * 16 + torch.arange(16, device=dev)[None, :], 3)
            flip[:, 15] &= (1 << 40) - 1
            d_q[hit] = corpus[pick] ^ flip
            d_cd = torch.full((nc,), 600, dtype=torch.int32, device=dev)
            d_qd = torch.full((nq,), 600, dtype=torch.int32, device=dev)
            b, e = vdist.shard_range(nc, rank, world)
            d_c = corpus[b:e].contiguous()
            torch.cuda.current_stream().synchronize()
            tbl = ctx.table_create_device(d_c.data_ptr(), d_cd[b:e].data_ptr(), e - b, keepalive=(d_c, d_cd))
        result = {}

        def step():
            result["keys"] = vdist.search_refs_keys(ctx, tbl, None, b, d_q, d_qd, tol_int)

        with ClockSampler(local) as cs:
            with torch.cuda.stream(stream):
                for _ in range(warmup):
                    step()
            ctx.kernel_time(0, reset=True)
            l0 = ctx.counters()[0]
            secs = timed(step, steps, 0, False)
        launches = ctx.counters()[0] - l0
        k_ms, k_n = ctx.kernel_time(0, reset=True)
        pairs = nq * nc
        keys_np = result["keys"].cpu().numpy().view(np.uint64)
        out = {"metric": "hamming_pair_comparisons_per_s", "value": pairs * steps / secs, "unit": "pairs/s",
               "ms_per_step": secs / steps * 1e3, "steps": steps, "scaling": "strong", "dtype": dtype_of(variant),
               "roofline": search_roofline(variant, pairs / world, k_ms, k_n, sm_count, sm_max_mhz),
               "gpu_launches": int(launches), "clocks": cs.summary(), "result_digest": digest(keys_np),
               "config": {"workload": workload_name("refs", args), "matches": int(len(keys_np)),
                          "parallelism": f"table slice x{world}", "exchange": exchange_note,
                          "table": "each rank's slice prepared once (vdf_table), queries packed per step",
                          "l2": f"table slice operands {(e - b) * 512 / 1e6:.0f} MB per rank > 2 x L2: no flush"}}
        tbl.close()
        Hc = corpus.cpu().numpy().view(np.uint64)
        cdur = np.full(nc, 600, np.uint32)
        Q = d_q.cpu().numpy().view(np.uint64)
        qdur = np.full(nq, 600, np.uint32)
        del corpus, d_c
        if rank == 0 and not args.no_cpu_baseline and args.parity_rows > 0:
            from oracle import vdf_oracle as o

            k_rows = max(4, min(args.parity_rows, int(8.0e8 // nc)))
            rows = np.unique(np.concatenate([np.random.default_rng(3).integers(0, nq, k_rows), np.arange(0, 20 * (k_rows // 2), 20) % nq]))
            rp, ci = o.search_refs(Hc, cdur, Q[rows], qdur[rows], tol_int)
            kr, kc = (keys_np >> np.uint64(32)).astype(np.int64), keys_np & np.uint64(0xFFFFFFFF)
            bad = sum(not np.array_equal(kc[kr == r], ci[rp[k]:rp[k + 1]]) for k, r in enumerate(rows.tolist()))
            out["parity_sample"] = {"rows": int(len(rows)), "candidates_per_row": int(nc), "mismatching_rows": int(bad), "ok": bad == 0,
                                    "how": "oracle brute force of sampled query rows (half of them planted hits) vs the GPU's match lists, untimed"}
        # e2e: host tables through the public API (sort keys + GPU sort of the 10 M table, uploads, search, CSR back)
        ids = np.random.default_rng(11).permutation(nc)
        cands = vdf.HashTable(Hc[ids], cdur, FixedWidthPaths(ids))
        cands._blob = cands.paths.blob()
        refs = vdf.HashTable(Q, qdur, FixedWidthPaths(np.arange(nq) + 10**7 * 5))
        vdist.search_with_references(refs, cands, args.tol, ctx=ctx)
        c0 = ctx.counters()
        barrier()
        t0 = time.perf_counter()
        g = vdist.search_with_references(refs, cands, args.tol, ctx=ctx)
        barrier()
        e_secs = max_over_ranks(time.perf_counter() - t0)
        c1 = ctx.counters()
        out["e2e"] = {"value": pairs / e_secs, "unit": "pairs/s", "ms_per_call": e_secs * 1e3, "groups": len(g),
                      "h2d_bytes_per_step": int(c1[1] - c0[1]), "d2h_bytes_per_step": int(c1[2] - c0[2]), "steps": 1,
                      "result_digest": digest(*g.csr), "input_order": "shuffled candidate table",
                      "api": "vid_dup_finder_lib_b200.dist.search_with_references(refs, candidates, tolerance) -> MatchGroups"}
        return out

    # ---------------------------------------------------------------- end-to-end dedup (BASELINE configs[4])
    def bench_e2e10m():
        """frame stacks -> hashes -> all-gather -> (duration, Path) sort -> all-pairs search -> MatchGroups, on all ranks.
        Stacks come from a resident pool of 1080p stacks cycled to --n-e2e hash operations (one pass over distinct content
        would need 330 TB of synthetic pixels); the hashes never leave HBM.  So that the search sees distinct entries, row r
        of the gathered table is XORed with salt[r] = synthetic[r] ^ H(pool[r % pool]), which turns it into the synthetic
        planted table EXACTLY IF every one of the hashes computed in the timed region is right."""
        n, ns, tol_int = args.n_e2e, args.stacks, tolerance_to_int(args.tol)
        w, h = args.width, args.height
        b, e = vdist.shard_range(n, rank, world)
        with torch.cuda.stream(stream):
            pool = synth.frame_stacks(ns, w, h, device=dev, first_id=0)  # the same pool on every rank: row r is stack r % ns
            descs = _ffi.make_descs(ns, w, h)
            pool_hash = torch.zeros((ns, 16), dtype=torch.int64, device=dev)
            torch.cuda.current_stream().synchronize()
            ctx.hash_stacks_device(pool.data_ptr(), descs, _ffi.CROPDETECT_LETTERBOX, pool_hash.data_ptr())  # untimed: the salt
            synth_tab, _ = synth.planted_hashes_torch(n, device=dev)
            salt = synth_tab ^ pool_hash[torch.arange(n, device=dev) % ns]
            del synth_tab
            n_calls = (e - b + ns - 1) // ns
            mine = torch.zeros((n_calls * ns, 16), dtype=torch.int64, device=dev)
            ids = np.random.default_rng(13).permutation(n)  # path of row r = "v/%08d" % ids[r]: the caller's order is arbitrary
            paths = FixedWidthPaths(ids)
            blob, off = paths.blob()
            durs = np.full(n, 600, np.uint32)
            d_order = torch.empty(n, dtype=torch.int32, device=dev)
            d_dur = torch.empty(n, dtype=torch.int32, device=dev)
            torch.cuda.current_stream().synchronize()
        # rows [b, e) of the table are stacks (b + k) % ns: rotate the descriptors so that every call hashes the right stacks
        rot = np.roll(descs, -(b % ns))
        ph = {}
        ctx.kernel_time(0, reset=True)
        l0 = ctx.counters()[0]
        with ClockSampler(local) as cs:
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(stream):
                for c in range(n_calls):  # H1-H5 on this rank's share
                    ctx.hash_stacks_device(pool.data_ptr(), rot, _ffi.CROPDETECT_LETTERBOX, mine[c * ns:].data_ptr())
                torch.cuda.current_stream().synchronize()
                ph["hash_s"] = time.perf_counter() - t0
                if world > 1:  # G1: one all-gather of the 128-byte hashes (ragged last share: gather padded, cut)
                    per = (n + world - 1) // world
                    buf = torch.zeros((world, per, 16), dtype=torch.int64, device=dev)
                    dist.all_gather_into_tensor(buf.view(-1), torch.nn.functional.pad(mine[:e - b], (0, 0, 0, per - (e - b))).contiguous().view(-1))
                    table = torch.cat([buf[r, :vdist.shard_range(n, r, world)[1] - vdist.shard_range(n, r, world)[0]] for r in range(world)])
                else:
                    table = mine[:n]
                table = table ^ salt
                torch.cuda.current_stream().synchronize()
                ph["allgather_s"] = time.perf_counter() - t0 - ph["hash_s"]
                t1 = time.perf_counter()
                if rank == 0:  # S1: keys cut by host threads, sorted on the GPU; hashes stay where they are
                    ctx.sort_order_device(durs, blob, off, d_order.data_ptr(), d_dur.data_ptr())
                if world > 1:
                    dist.broadcast(d_order, 0)
                    dist.broadcast(d_dur, 0)
                d_sorted = table[d_order.long()]
                torch.cuda.current_stream().synchronize()
                ph["sort_s"] = time.perf_counter() - t1
                t2 = time.perf_counter()
                keys = vdist.search_self_keys(ctx, d_sorted, d_dur, tol_int)  # S2-S3, G2
                torch.cuda.current_stream().synchronize()
                ph["search_s"] = time.perf_counter() - t2
                t3 = time.perf_counter()
                gp, mm = ctx.group_greedy_device(n, keys.data_ptr(), keys.numel(), d_remap=d_order.data_ptr())  # S4
                groups = vdf.MatchGroup.from_csr(paths, gp, mm)
                ph["group_s"] = time.perf_counter() - t3
            barrier()
            secs = max_over_ranks(time.perf_counter() - t0)
        launches = ctx.counters()[0] - l0
        k_ms, k_n = ctx.kernel_time(0, reset=True)
        pairs = n * (n - 1) // 2
        keys_np = keys.cpu().numpy().view(np.uint64)
        out = {"metric": "e2e_dedup_hashes_per_s", "value": n / secs, "unit": "hashes/s", "seconds": secs, "steps": 1, "scaling": "strong",
               "phases_s_rank0": ph, "search_pairs_per_s": pairs / max_over_ranks(ph["search_s"]),
               "hash_stacks_per_s": n / max_over_ranks(ph["hash_s"]),
               "roofline": search_roofline(variant, pairs / world, k_ms, k_n, sm_count, sm_max_mhz),
               "gpu_launches": int(launches), "clocks": cs.summary(), "edges": int(len(keys_np)), "groups": len(groups),
               "result_digest": digest(keys_np, gp, mm),
               "config": {"workload": workload_name("e2e10m", args), "pool_stacks": ns,
                          "data_flow": "pool stacks (HBM) -> vdf_hash_stacks_device -> all-gather -> xor salt -> vdf_sort_order_device -> "
                                       "gather -> sharded search + fused exchange -> greedy groups with device remap -> CSR to host",
                          "parallelism": f"stack shard x{world}, then tile-block shard x{world}", "exchange": exchange_note}}
        if rank == 0 and not args.no_cpu_baseline and args.parity_rows > 0:
            # positions are sorted positions: bring the sorted table to the host for the oracle
            Hs = d_sorted.cpu().numpy().view(np.uint64)
            out["parity_sample"] = parity_rows_self(Hs, durs, keys_np, tol_int, max(4, min(args.parity_rows, int(6.0e8 // n))))
        return out

    # ---------------------------------------------------------------- hashing workload (BASELINE configs[1])
    def bench_hash(steps, warmup):
        """--hash-total stack-hashes per GPU (configs[1]: 100 k stacks) from a resident pool of --stacks 1080p stacks"""
        w, h, ns = args.width, args.height, args.stacks
        steps = max(steps, (args.hash_total + ns - 1) // ns)
        stack_bytes = 16 * w * h
        with torch.cuda.stream(stream):
            pool = synth.frame_stacks(ns, w, h, device=dev, first_id=rank * ns)
            out = torch.zeros((ns, 16), dtype=torch.int64, device=dev)
        descs = _ffi.make_descs(ns, w, h)
        torch.cuda.synchronize()
        crops = {}
        cropdetect = int(os.environ.get("VDF_BENCH_CROPDETECT", _ffi.CROPDETECT_LETTERBOX))  # timing studies: 0 = Cropdetect::None

        def step():
            crops["c"] = ctx.hash_stacks_device(pool.data_ptr(), descs, cropdetect, out.data_ptr())[1]

        # a short run first: 20 calls, before the GPU reaches its power cap (the long run below does, and slows with the SM clock)
        with ClockSampler(local) as cs_b:
            with torch.cuda.stream(stream):
                for _ in range(warmup):
                    step()
            ctx.kernel_time(1, reset=True)
            b_secs = timed(step, 20, 0, False)
        b_ms, b_n = ctx.kernel_time(1, reset=True)
        burst = {"steps": 20, "ms_per_step": b_secs / 20 * 1e3, "kernel_ms_per_launch": b_ms / max(b_n, 1),
                 "value": ns * world * 20 / b_secs, "unit": "stacks/s",
                 "frac": (ns * (stack_bytes + 128) / (b_ms / max(b_n, 1) * 1e-3) / 1e9 / hbm_peak) if b_n else None,
                 "frac_whole_step": ns * (stack_bytes + 128) / (b_secs / 20) / 1e9 / hbm_peak, "clocks": cs_b.summary()}
        with ClockSampler(local) as cs:
            for k in range(4):
                ctx.kernel_time(k, reset=True)
            l0 = ctx.counters()[0]
            secs = timed(step, steps, 0, False)
        launches = ctx.counters()[0] - l0
        kt = [ctx.kernel_time(k, reset=True) for k in range(4)]
        value = ns * world * steps / secs
        k_ms = kt[1][0] / max(kt[1][1], 1)
        alg_bytes = ns * (stack_bytes + 128)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if kt[1][1] else None
        # bytes the kernel actually requests: the cropped rows in whole k-chunks (256 bytes in the fused kernel, 128 in the per-frame
        # kernels) from the 16-byte aligned start, never beyond the row; + strip 0 of the 8 scanned sides and the bars the scan walks
        c = crops["c"].astype(np.int64)
        cw, chh = w - c[:, 0] - c[:, 1], h - c[:, 2] - c[:, 3]
        fused = args.hash_variant < 0
        chunk = 256 if fused else 128
        x0 = c[:, 0] & ~15
        row_bytes = np.minimum((((c[:, 0] & 15) + cw + chunk - 1) // chunk) * chunk, w - x0)
        scan = 2 * (2 * h + 2 * w + (c[:, 0] + c[:, 1] + 64 * (c[:, 0] + c[:, 1] > 0)) * h + (c[:, 2] + c[:, 3] + 64 * (c[:, 2] + c[:, 3] > 0)) * w)
        moved = int((16 * chh * row_bytes + scan).sum()) + ns * 128
        kname = "hash_fused_kernel" if fused else "resize_mma_kernel"
        roof = {"bound": "hbm",
                "kernel": ("hash_fused_kernel (ONE persistent launch per call: letterbox scan of frames 0 and 8, crop, resize job, Lanczos3 -> 16x16 per "
                           "frame with both passes on the tensor path, 16^3 DCT, threshold, pack)") if fused else
                          "resize_mma_kernel (crop window + Lanczos3 -> 16x16 per frame, DCT + threshold + pack in the stack's last CTA), behind the letterbox scan kernels",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                "frac_bytes_moved": (moved / (k_ms * 1e-3) / 1e9 / hbm_peak) if kt[1][1] else None, "bytes_moved_per_launch": moved,
                "frac_whole_step": alg_bytes / (secs / steps) / 1e9 / hbm_peak,
                "traffic": ncu_traffic(kname, f"stacks_{ns}_{w}x{h}"), "peak_source": peak_src,
                "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_stack": stack_bytes + 128,
                "note": ("kernel_ms = CUDA events around the one launch of a call; frac = 16*W*H + 128 bytes per stack over that time (SURVEY M3: cropped "
                         "stacks read fewer rows, frac_bytes_moved counts what is requested); frac_whole_step = the same bytes over ms_per_step, "
                         "which also holds the descriptor upload, two memsets, the read-back of the crops and the host's wake-up") if fused else
                        "kernel_ms = CUDA events around the resize launches of one call (includes the job-build kernels); the letterbox scan runs before it",
                "step_share": {"fused_kernel" if fused else "resize_dct_pack": kt[1][0], "letterbox_kernels": kt[2][0], "unit": "ms over timed steps"}}
        # e2e: host frames (pinned) -> vdf_hash_stacks -> host hashes, bounded to a few stacks (PCIe-bound)
        ne = min(ns, 32)
        host = pool[:ne].cpu().pin_memory()
        hd = _ffi.make_descs(ne, w, h)
        ctx.hash_stacks(host.numpy().reshape(-1), hd, _ffi.CROPDETECT_LETTERBOX)  # warm-up: staging buffers get allocated
        barrier()
        t0 = time.perf_counter()
        got, st, _ = ctx.hash_stacks(host.numpy().reshape(-1), hd, _ffi.CROPDETECT_LETTERBOX)
        barrier()
        e_secs = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": ne * world / e_secs, "unit": "stacks/s", "h2d_bytes_per_step": int(ne * stack_bytes),
               "d2h_bytes_per_step": int(ne * 128), "stacks": ne, "api": "vdf_hash_stacks (host frames, pinned async staging)"}
        res = {"metric": "frame_stacks_hashed_per_s", "value": value, "unit": "stacks/s", "ms_per_step": secs / steps * 1e3, "steps": steps,
               "stack_hashes_per_gpu": ns * steps, "scaling": "weak", "dtype": "u8/i32 resize, f64 DCT", "roofline": roof, "burst": burst, "e2e": e2e,
               "gpu_launches": int(launches), "clocks": cs.summary(), "result_digest": digest(out.cpu().numpy()),
               "config": {"workload": workload_name("hash", args), "stacks_per_gpu_per_step": ns, "parallelism": f"stack shard x{world}",
                          "l2": f"pool of {ns * stack_bytes / 1e9:.1f} GB per GPU >> L2, no flush",
                          "hash_variant": args.hash_variant}}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from concurrent.futures import ThreadPoolExecutor

            from oracle import vdf_oracle as o

            cores = os.cpu_count() or 1
            total, diff, dt_cpu, done = max(ns, args.mismatch_stacks), 0, 0.0, 0
            for first in range(0, total, ns):  # bit mismatches over >= 1024 distinct stacks
                if first:
                    with torch.cuda.stream(stream):
                        pool.copy_(synth.frame_stacks(ns, w, h, device=dev, first_id=10_000 + first))
                        torch.cuda.current_stream().synchronize()
                    step()
                hs = pool.cpu().numpy()
                t0 = time.perf_counter()
                with ThreadPoolExecutor(cores) as ex:
                    ref = list(ex.map(lambda s: o.hash_stack(hs[s], 1)[1], range(ns)))
                dt_cpu += time.perf_counter() - t0
                gw = out.cpu().numpy().view(np.uint64)
                diff += int(np.unpackbits((np.stack(ref) ^ gw).view(np.uint8)).sum())
                done += ns
            res["cpu_baseline"] = {"value": done / dt_cpu, "unit": "stacks/s", "cores": cores, "kind": "port",
                                   "sample": f"{done} synthetic stacks (the bench pool and {done // ns - 1} more pools), oracle port on {cores} threads"}
            res["bit_mismatch"] = {"stacks": done, "bits_differing": diff, "rate": diff / (done * 1000.0),
                                   "note": "f64 DCT in the oracle's operation order: no epsilon band"}
        return res

    def guarded(fn, *a, **k):
        try:
            return fn(*a, **k)
        except Exception as e:  # a secondary record must not take the headline down with it
            import traceback

            traceback.print_exc()
            return {"error": repr(e)}

    if args.workload == "search":
        line = bench_search(args.steps, args.warmup)
        wanted = [] if args.no_secondary else [x.strip() for x in args.secondary.split(",") if x.strip()]
        if "hash" in wanted:
            line["secondary"] = guarded(bench_hash, max(2, args.steps), max(3, args.warmup))
        if "popc" in wanted:  # the north-star's graded kernel: XOR + POPC, its own roofline (integer XU pipe)
            line["secondary_popc"] = guarded(bench_search, 2, 1, n=args.n_popc, use_variant=0, light=True)
        if "refs" in wanted:
            line["secondary_refs"] = guarded(bench_refs, min(max(args.steps, 1), 3), 1)
        if "e2e10m" in wanted:
            line["secondary_e2e"] = guarded(bench_e2e10m)
    elif args.workload == "popc":
        line = bench_search(args.steps, args.warmup, n=args.n_popc, use_variant=0, light=True)
    elif args.workload == "refs":
        line = bench_refs(args.steps, args.warmup)
    elif args.workload == "e2e10m":
        line = bench_e2e10m()
    else:
        line = bench_hash(args.steps, args.warmup)
    line.update({"n_gpus": world, "steps": line.get("steps", args.steps) if args.workload != "search" else args.steps, "warmup": args.warmup,
                 "higher_is_better": True, "vs_baseline": None, "data": "synthetic", "impl": "ours"})
    if rank == 0:
        emit(line)
    if world > 1:
        if fused:
            vdist.disable_peer_exchange(ctx)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
