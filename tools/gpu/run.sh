#!/bin/bash
# One parametrised GPU-box script (replaces the per-experiment scripts of round 1).  Usage, under gpurun:
#   bash tools/gpu/run.sh <stage> [<stage> ...]
# Stages write into gpurun_out/ (merged back by gpurun); each runs under its own timeout so that a hang costs
# minutes, not the box.  TAG (env, default r02) prefixes the output files.
#   tests            python -m pytest tests -m gpu
#   smoke            __graft_entry__.smoke()
#   bench            python bench.py (defaults) -> ${TAG}_bench_c2.json
#   bench:<args>     python bench.py <args with ',' for ' '> -> ${TAG}_bench_<args>.json
#   mbench:<N>:<args> torchrun, N ranks
#   probe[:c3,c5]    tools/gpu/probe.py (gather floor microbenchmarks, hot-table variants, conversion phases)
#   sigma            tools/gpu/sweep_sigma.py (sigma rule sweep)
#   launches:<wl>    ncu launch list of one bench run of workload <wl>
#   ncu:<wl>:<regex>[:<bench args>] ncu --set full of the kernels matching <regex>
#   mtests:<N>       tests/multigpu_worker.py on N ranks
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
export CUDA_DEVICE_MAX_CONNECTIONS=32
PORT=29517
for stage in "$@"; do
  name=${stage%%:*}; rest=${stage#*:}; [ "$rest" == "$stage" ] && rest=""
  echo "=== stage $stage ($(date +%T))"
  case $name in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q ${rest//,/ } 2>&1 | tail -25 | tee gpurun_out/${TAG}_tests.txt ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt ;;
    bench)
      args=${rest//,/ }; label=$(echo "${rest:-c2}" | tr -c 'A-Za-z0-9\n' '_')
      timeout 1200 python bench.py $args > gpurun_out/${TAG}_bench_${label}.json 2> gpurun_out/${TAG}_bench_${label}.log
      echo "rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_${label}.json; tail -3 gpurun_out/${TAG}_bench_${label}.log ;;
    mbench)
      n=${rest%%:*}; a=${rest#*:}; [ "$a" == "$rest" ] && a=""
      args=${a//,/ }; label=$(echo "n${n}_${a}" | tr -c 'A-Za-z0-9\n' '_')
      PORT=$((PORT+1))
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $n $args > gpurun_out/${TAG}_bench_${label}.json 2> gpurun_out/${TAG}_bench_${label}.log
      echo "rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench_${label}.json; grep -v "^W\|^\*\*\*" gpurun_out/${TAG}_bench_${label}.log | tail -8 ;;
    mtests)
      PORT=$((PORT+1))
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${rest:-2} --master-addr 127.0.0.1 --master-port $PORT \
        tests/multigpu_worker.py > gpurun_out/${TAG}_mtests_n${rest:-2}.txt 2>&1
      echo "rc=$?"; grep -E "rank [0-9]+:|Error|error|assert|Traceback|File " gpurun_out/${TAG}_mtests_n${rest:-2}.txt | head -30 ;;
    probe)
      timeout 1500 python tools/gpu/probe.py ${rest:-c3,c5,c2} 2>&1 | tee gpurun_out/${TAG}_probe.txt | tail -60 ;;
    sigma)
      timeout 1500 python tools/gpu/sweep_sigma.py 2>&1 | tee gpurun_out/${TAG}_sweep_sigma.txt | tail -40 ;;
    launches)
      wl=${rest:-c2}
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_${wl}.csv \
        python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${TAG}_launches_${wl}.log 2>&1
      echo "rc=$?"; tail -3 gpurun_out/${TAG}_launches_${wl}.csv ;;
    ncu)
      wl=${rest%%:*}; r2=${rest#*:}; rx=${r2%%:*}; extra=${r2#*:}; [ "$extra" == "$r2" ] && extra=""
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -s 10 -c 2 -f -o gpurun_out/${TAG}_ncu_${wl}_${rx} \
        python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-extra ${extra//,/ } > gpurun_out/${TAG}_ncu_${wl}_${rx}.log 2>&1
      echo "rc=$?"; rep=gpurun_out/${TAG}_ncu_${wl}_${rx}
      # gpurun brings back at most 64 MiB: keep the small exports, drop the 40 MB report unless KEEP_REP=1
      python tools/ncu_summary.py $rep.ncu-rep > $rep.md 2>/dev/null
      ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null | gzip > ${rep}_raw.csv.gz
      ncu -i $rep.ncu-rep --page details 2>/dev/null > ${rep}_details.txt
      ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > ${rep}_source.csv.gz
      [ "${KEEP_REP:-0}" == "1" ] || rm -f $rep.ncu-rep
      ls -la ${rep}*; head -12 $rep.md ;;
    sanitize)
      # compute-sanitizer over a small, adversarial subset of the GPU tests (memcheck, then racecheck on the shared-memory kernels)
      sel='long_empty or huge_span or device_coo_to_csr_matches or (chunks_bit_identical and example_c1) or (axpby and hub_row) or (native_sharded_on_one_gpu and two_packet_s26 and push)'
      # test_gpu_refcuda.py runs the REFERENCE's own kernels (libref_cuda.so), whose rounded-up grids read out of bounds
      #  -- memcheck flags them, so they are left out here
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$sel" \
          --ignore tests/test_gpu_refcuda.py > gpurun_out/${TAG}_sanitize_memcheck_full.txt 2>&1
      tail -8 gpurun_out/${TAG}_sanitize_memcheck_full.txt | tee gpurun_out/${TAG}_sanitize_memcheck.txt
      timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "device_coo_to_csr_matches or (metadata_word_for_word and long_empty)" 2>&1 | tail -8 | tee gpurun_out/${TAG}_sanitize_racecheck.txt ;;
    cppsharded)
      # the C++ host API on real peers: examples/sharded_spmv.cpp with one shard per visible GPU (or <rest> shards)
      g++ -O2 -std=c++17 -Iinclude examples/sharded_spmv.cpp -Lbenchmark_spmv_using_csr5_b200 -lcsr5_b200 \
          -Wl,-rpath,$PWD/benchmark_spmv_using_csr5_b200 -o /tmp/sharded_spmv || echo "compile failed"
      n=${rest:-$(nvidia-smi -L | wc -l)}
      for tr in 0 1 2 4; do
        timeout 300 /tmp/sharded_spmv $n 10000000 30 $tr 0 2>&1 | tail -3 | sed "s/^/[transport $tr] /" | tee -a gpurun_out/${TAG}_cppsharded_n${n}.txt
      done ;;
    tma)
      # closing A/B of the TMA-staged kernel with x prefetch against the direct-load kernel: ring-geometry sweep
      wl=${rest:-c2}
      : > gpurun_out/${TAG}_tma_sweep_${wl}.txt
      run1() { timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e --no-extra --steps 60 --sigma-rule 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-40s %-28s kernel_ms %.4f (min %.4f) frac %.3f  GFLOPS %.1f' % ('$*', r['kernel'], r['kernel_ms_avg'], r['kernel_ms_min'], r['frac'], d['value']))
" | tee -a gpurun_out/${TAG}_tma_sweep_${wl}.txt; }
      run1 --kernel 1
      for st in 2 3 4; do for w in 8 11 16; do run1 --kernel 4 --stages $st --warps $w --ctas-per-sm 1; done; done
      run1 --kernel 4 --stages 2 --warps 8 --ctas-per-sm 2
      run1 --kernel 4 --stages 3 --warps 5 --ctas-per-sm 2
      run1 --kernel 2 --stages 3 --warps 11 --ctas-per-sm 1
      run1 --kernel 1 ;;
    ab)
      # A/B of two builds of the library on the same box, interleaved: ab:<other .so>:<bench args>
      other=${rest%%:*}; a=${rest#*:}; [ "$a" == "$rest" ] && a=""
      for rep in 1 2; do for lib in libcsr5_b200.so $other; do
        CSR5B200_LIB=$lib timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 100 ${a//,/ } 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ex=d.get('extra_workloads') or {}
print('$lib rep $rep:', d['config']['workload'][:24], 'sigma', d['config']['sigma'], 'kernel_ms %.4f frac %.3f' % (d['roofline']['kernel_ms_avg'], d['roofline']['frac']),
      ' | '.join('%s sigma %s kernel_ms %.4f frac %.3f' % (k, v.get('sigma'), v.get('kernel_ms', 0), v.get('roofline_frac', 0)) for k, v in ex.items()))
" | tee -a gpurun_out/${TAG}_ab.txt
      done; done ;;
    *) echo "unknown stage $stage" ;;
  esac
done
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_nvidia_smi.csv 2>&1
echo "=== done ($(date +%T))"
