#!/bin/bash
# One GPU-box session: microbench, parity tests, smoke, bench, ncu launch list + full captures.
# Run under gpurun from the repo root; everything worth keeping lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/host.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/host.txt
STEP=${1:-all}
if [[ $STEP == all || $STEP == micro ]]; then
  timeout 180 ./vid_dup_finder_lib_b200/vdf_microbench > gpurun_out/microbench.jsonl 2>&1; echo "microbench rc=$?"
fi
if [[ $STEP == all || $STEP == test ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
fi
if [[ $STEP == all || $STEP == smoke ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
fi
if [[ $STEP == all || $STEP == bench ]]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 600 python bench.py --steps 2 --warmup 1 --variant 0 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err; echo "bench v0 rc=$?"; tail -c 1500 gpurun_out/bench_v0.json
fi
if [[ $STEP == hash ]]; then
  for v in 0 3 2; do
    timeout 600 python bench.py --workload hash --steps 5 --warmup 3 --hash-variant $v --no-cpu-baseline > gpurun_out/bench_hash_v$v.json 2> gpurun_out/bench_hash_v$v.err; echo "bench hash v$v rc=$?"; tail -c 1800 gpurun_out/bench_hash_v$v.json; tail -3 gpurun_out/bench_hash_v$v.err
  done
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:resize_mma -c 1 -f -o gpurun_out/prof_resize_mma \
      python bench.py --workload hash --steps 1 --warmup 0 --stacks 32 --no-cpu-baseline > gpurun_out/ncu_resize_mma.log 2>&1; echo "ncu resize_mma rc=$?"
fi
if [[ $STEP == all || $STEP == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --stacks 64 > gpurun_out/bench_under_ncu.json 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_tiles -c 1 -f -o gpurun_out/prof_hamming \
      python bench.py --steps 1 --warmup 0 --n-hashes 262144 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_hamming.log 2>&1; echo "ncu hamming rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:resize -c 1 -f -o gpurun_out/prof_resize \
      python bench.py --workload hash --steps 1 --warmup 0 --stacks 32 --no-cpu-baseline > gpurun_out/ncu_resize.log 2>&1; echo "ncu resize rc=$?"
fi
ls -la gpurun_out | head -30
if [[ $STEP == exp ]]; then
  timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "edges_and_groups" > gpurun_out/pytest_variants.log 2>&1; echo "pytest variants rc=$?"; tail -3 gpurun_out/pytest_variants.log
  for v in 2 1 0; do
    timeout 600 python bench.py --steps 3 --warmup 2 --variant $v --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_sv$v.json 2> gpurun_out/bench_sv$v.err; echo "bench search v$v rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_sv$v.json'));print(d['value'], d['roofline']['kernel_ms_per_launch'])"
  done
  for v in 1 2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_tiles -c 1 -f -o gpurun_out/prof_hamming_v$v \
      python bench.py --steps 1 --warmup 0 --n-hashes 262144 --variant $v --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_hamming_v$v.log 2>&1; echo "ncu hamming v$v rc=$?"
  done
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:letterbox_side -c 1 -f -o gpurun_out/prof_letterbox \
      python bench.py --workload hash --steps 1 --warmup 0 --stacks 256 --no-cpu-baseline > gpurun_out/ncu_letterbox.log 2>&1; echo "ncu letterbox rc=$?"
fi
if [[ $STEP == multi ]]; then
  N=${2:-2}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"; tail -6 gpurun_out/dist_check_$N.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench x$N rc=$?"; tail -c 1500 gpurun_out/bench_g$N.json; tail -3 gpurun_out/bench_g$N.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --n-hashes 100000 > gpurun_out/bench_ref_g$N.json 2>&1; echo "bench ref x$N rc=$?"
fi
if [[ $STEP == tc2 ]]; then
  timeout 300 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "2cta" > gpurun_out/pytest_tc2.log 2>&1; echo "pytest tc2 rc=$?"; tail -25 gpurun_out/pytest_tc2.log
  for v in 5 4; do
    timeout 300 python bench.py --steps 3 --warmup 2 --variant $v --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_sv$v.json 2> gpurun_out/bench_sv$v.err; echo "bench search v$v rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_sv$v.json'));print(d['value'], d['roofline']['kernel_ms_per_launch'], d['config']['edges'], d['config']['groups'])"; tail -2 gpurun_out/bench_sv$v.err
  done
  for v in 5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_tc -c 1 -f -o gpurun_out/prof_hamming_v$v \
      python bench.py --steps 1 --warmup 0 --n-hashes 262144 --variant $v --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_hamming_v$v.log 2>&1; echo "ncu hamming v$v rc=$?"
  done
fi
if [[ $STEP == list5 ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv \
      python bench.py --steps 1 --warmup 1 --variant 5 --no-cpu-baseline --no-secondary --e2e-steps 1 > gpurun_out/bench_under_ncu_v5.json 2>&1; echo "ncu list rc=$?"
  for i in 1 2; do
  timeout 300 python bench.py --steps 3 --warmup 2 --variant 5 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_sv5_$i.json 2> gpurun_out/bench_sv5_$i.err; echo "bench search v5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_sv5_$i.json'));print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['config']['edges'], d['config']['groups'])"
  done
fi
if [[ $STEP == ncu2 ]]; then
  # launch list of the default bench command (shares of the step), then full captures of the dominant kernels at bench size
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vdf|DeviceRadixSort|DeviceScan" -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --e2e-steps 1 > gpurun_out/bench_under_ncu.json 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vdf -c 400 --csv --log-file gpurun_out/launches_hash.csv \
      python bench.py --workload hash --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_hash.json 2>&1; echo "ncu list hash rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_tc5 -c 1 -f -o gpurun_out/prof_hamming_tc5__self_1000000_x1 \
      python bench.py --steps 1 --warmup 0 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_hamming_tc5.log 2>&1; echo "ncu tc5 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:resize_mma -c 1 -f -o gpurun_out/prof_resize_mma__stacks_256_1920x1080 \
      python bench.py --workload hash --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_resize_mma.log 2>&1; echo "ncu resize rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:letterbox_side -c 1 -f -o gpurun_out/prof_letterbox__stacks_256_1920x1080 \
      python bench.py --workload hash --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_letterbox.log 2>&1; echo "ncu letterbox rc=$?"
fi
if [[ $STEP == r6 ]]; then
  # refresh of the tracked evidence with search variant 6 (kind::mxf4) as the library default
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"; tail -c 1500 gpurun_out/bench_reference.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --e2e-steps 1 > gpurun_out/bench_under_ncu.json 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_tc6 -c 1 -f -o gpurun_out/prof_hamming_tc6__self_1000000_x1 \
      python bench.py --steps 1 --warmup 0 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_hamming_tc6.log 2>&1; echo "ncu tc6 rc=$?"
fi
# (the x7-x12 option sweeps whose bench lines are in profiles/experiments/ used kernel options that were removed again once
# measured: expander warp counts and teams, fence batching, timing-only debug forms - see the comments in csrc/search_tc.cu)
if [[ $STEP == x10 ]]; then
  for c in 64 256 512 1024 2048; do
    timeout 300 python bench.py --steps 3 --warmup 2 --opt tc_chunk=$c --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_x10_c${c}.json 2> gpurun_out/bench_x10_c${c}.err; echo "bench chunk=$c rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_x10_c${c}.json'));print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['config']['edges'], d['config']['groups'])"; tail -2 gpurun_out/bench_x10_c${c}.err
  done
fi
if [[ $STEP == x13 ]]; then
  timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -x -q > gpurun_out/pytest_search.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_search.log
  for i in 1 2; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_x13_$i.json 2> gpurun_out/bench_x13_$i.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_x13_$i.json'));print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['config']['edges'], d['config']['groups'])"; tail -2 gpurun_out/bench_x13_$i.err
  done
  VDF_BENCH_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_x13_dbg.json 2> gpurun_out/bench_x13_dbg.err; tail -12 gpurun_out/bench_x13_dbg.err
fi
if [[ $STEP == xg ]]; then
  # multi-GPU: parity of both exchange modes, then the bench with each (same box, back to back)
  N=${2:-2}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -v "^W1\|OMP_NUM\|^\*\*\*" gpurun_out/dist_check_$N.log | tail -9
  for ex in peer; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --exchange $ex --no-secondary --no-cpu-baseline > gpurun_out/bench_xg${N}_$ex.json 2> gpurun_out/bench_xg${N}_$ex.err; echo "bench x$N $ex rc=$?"; python -c "import json;d=json.loads(open('gpurun_out/bench_xg${N}_$ex.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['config']['edges'], d['config']['groups'], d['e2e']['ms_per_call'])"
  done
fi
if [[ $STEP == sanitize ]]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_search.py tests/test_gpu_hashing.py -m gpu -x -q \
      -k "variant6_kernel_options or (edges_and_groups and tcgen05_mxf4 and (129 or 1000)) or find_with_refs or known_group or components_mode or letterbox_kat or cropdetect_none or dct_threshold" \
      > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
fi
if [[ $STEP == sweep ]]; then
  # BASELINE config 2: tolerance sweep of the 1 M all-pairs search (0.42 stresses the edge buffer)
  for t in 0.0 0.1 0.2 0.3 0.4 0.42; do
    timeout 300 python bench.py --steps 2 --warmup 1 --tol $t --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_tol_$t.json 2> gpurun_out/bench_tol_$t.err; echo "bench tol=$t rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_tol_$t.json'));print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['config']['edges'], d['config']['groups'], d['e2e']['ms_per_call'])"
  done
fi
if [[ $STEP == big ]]; then
  # BASELINE configs[3] and [4] sizes on one GPU: 100k queries x 10M-entry table; all-pairs over a 10M-hash corpus
  timeout 900 python bench.py --workload refs --steps 3 --warmup 1 > gpurun_out/bench_refs_10m.json 2> gpurun_out/bench_refs_10m.err; echo "bench refs rc=$?"; tail -c 1200 gpurun_out/bench_refs_10m.json; tail -3 gpurun_out/bench_refs_10m.err
  timeout 1200 python bench.py --n-hashes 10000000 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_self_10m.json 2> gpurun_out/bench_self_10m.err; echo "bench 10M rc=$?"; tail -c 2500 gpurun_out/bench_self_10m.json; tail -3 gpurun_out/bench_self_10m.err
fi
if [[ $STEP == tc ]]; then
  timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "tcgen05" > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -25 gpurun_out/pytest_tc.log
  for v in 3 2; do
    timeout 600 python bench.py --steps 3 --warmup 2 --variant $v --no-secondary --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_sv$v.json 2> gpurun_out/bench_sv$v.err; echo "bench search v$v rc=$?"; python -c "import json;d=json.load(open('gpurun_out/bench_sv$v.json'));print(d['value'], d['roofline']['kernel_ms_per_launch'], d['config']['edges'], d['config']['groups'])"; tail -2 gpurun_out/bench_sv$v.err
  done
fi
