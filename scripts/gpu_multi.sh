#!/bin/bash
# Multi-GPU session (gpurun --gpus N): C-plane tests, torch-plane parity (dist_check), the full default bench under torchrun.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_x$N.csv 2>&1
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -q -k "multi_device or two_contexts" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/pytest_multi_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W" gpurun_out/dist_check_$N.log | tail -9
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps ${2:-5} --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench x$N rc=$?"; tail -4 gpurun_out/bench_g$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_g$N.json"))
    def show(k,v):
        if isinstance(v,dict) and "error" in v: print(k,"ERROR",v["error"]); return
        print(k, v.get("metric"), "%.4g"%v.get("value",0), "ms/step", v.get("ms_per_step"), "frac", (v.get("roofline") or {}).get("frac"), "digest", v.get("result_digest"), "parity", (v.get("parity_sample") or {}).get("ok"), "e2e ms", (v.get("e2e") or {}).get("ms_per_call"), (v.get("e2e") or {}).get("result_digest"))
    show("main",d)
    for k in d:
        if k.startswith("secondary"): show(k,d[k])
    print("cplane", d.get("e2e_cplane"))
    print("e2e phases", d["e2e"].get("phases_ms_last_call"))
    print("e2e10m", (d.get("secondary_e2e") or {}).get("phases_s_rank0"), (d.get("secondary_e2e") or {}).get("seconds"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_g$N.json 2> gpurun_out/bench_ref_g$N.err; echo "bench ref x$N rc=$?"; head -c 600 gpurun_out/bench_ref_g$N.json
