#!/bin/bash
# Round-2 evidence session on one B200 (run under gpurun from the repo root; everything lands in gpurun_out/,
# scripts/summarize_profiles.py r02 turns it into profiles/r02_*).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/host.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/host.txt
STEP=${1:-all}
if [[ $STEP == all || $STEP == micro ]]; then
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 250 > gpurun_out/microbench_clocks.csv 2>&1 &
  SMI=$!
  timeout 300 ./vid_dup_finder_lib_b200/vdf_microbench > gpurun_out/microbench.jsonl 2>&1; echo "microbench rc=$?"
  kill $SMI
  grep cublaslt gpurun_out/microbench.jsonl
fi
if [[ $STEP == all || $STEP == ncu ]]; then
  # launch list of one default step of both paths (search on the prepared table, hashing): per-launch times are cold and serialised
  # launch list of one default step of both paths (search on the prepared table, hashing): per-launch times are cold and serialised;
  # only this library's kernels and cub's (torch's synthetic-data kernels are not part of a step)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
      -k 'regex:^(void )?(vdf::|cub::|<unnamed>::)' -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --tol-sweep "" --secondary hash --hash-total 512 --parity-rows 0 > gpurun_out/bench_under_ncu.json 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_tc6 -c 1 -f -o gpurun_out/prof_hamming_tc6__self_1000000_x1 \
      python bench.py --steps 1 --warmup 0 --no-secondary --no-cpu-baseline --e2e-steps 1 --tol-sweep "" --parity-rows 0 > gpurun_out/ncu_tc6.log 2>&1; echo "ncu tc6 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_tiles -c 1 -f -o gpurun_out/prof_hamming_popc__self_262144_x1 \
      python bench.py --workload popc --steps 1 --warmup 0 --no-cpu-baseline --parity-rows 0 > gpurun_out/ncu_popc.log 2>&1; echo "ncu popc rc=$?"
  # the third launch: the first call meets crop sizes for the first time (a pass for the known sizes, a pass for the misses), the next call is warm
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hash_fused -s 2 -c 1 -f -o gpurun_out/prof_hash_fused__stacks_256_1920x1080 \
      python bench.py --workload hash --steps 1 --warmup 1 --hash-total 256 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
  # the per-frame kernels it replaced (context option hash_fused = 0), for the comparison in profiles/README.md
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:resize_mma -s 2 -c 1 -f -o gpurun_out/prof_resize_mma__stacks_256_1920x1080 \
      python bench.py --workload hash --hash-variant 0 --steps 1 --warmup 1 --hash-total 256 --no-cpu-baseline > gpurun_out/ncu_resize.log 2>&1; echo "ncu resize rc=$?"
fi
if [[ $STEP == all || $STEP == sanitizer ]]; then
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_search.py tests/test_gpu_hashing.py -m gpu -q -x \
      -k "fold_is_exact and clusters or prepared_table or many_matches or sort_ties or chunked_overlapped or fused_kernel_machinery or random_edge_lists or long_dependency or find_with_refs or cropdetect_none" \
      > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
fi
if [[ $STEP == all || $STEP == bench ]]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
  timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
  python scripts/show_bench.py gpurun_out/bench.json
fi
if [[ $STEP == all || $STEP == smoke ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
fi
ls -la gpurun_out | head -50
