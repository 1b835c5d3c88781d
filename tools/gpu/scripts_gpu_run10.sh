# 1 GPU: default benches, launch lists, ncu --set full captures of the default kernels.  The .ncu-rep files are
# summarised ON THE BOX (raw page + source page as CSV, tools/ncu_summary.py) and deleted: gpurun brings back
# at most 64 MiB.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --steps 1000 --warmup 50 > gpurun_out/r10_bench_c2.json 2> gpurun_out/r10_bench_c2.err; tail -2 gpurun_out/r10_bench_c2.err
timeout 400 python bench.py --workload c3 --steps 1000 --warmup 50 > gpurun_out/r10_bench_c3.json 2> gpurun_out/r10_bench_c3.err; tail -2 gpurun_out/r10_bench_c3.err
timeout 400 python bench.py --workload c4 --steps 200 --warmup 20 > gpurun_out/r10_bench_c4.json 2> gpurun_out/r10_bench_c4.err; tail -2 gpurun_out/r10_bench_c4.err
timeout 300 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r10_bench_ref.json 2> gpurun_out/r10_bench_ref.err
K='regex:spmv_|calibrate|tile_|scan_|transpose|desc_offset|hot_|zero_rows'
prof() { tag=$1; kre=$2; shift 2
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -o /tmp/prof_$tag python bench.py "$@" --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r10_ncu_full_$tag.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$tag.ncu-rep > gpurun_out/r10_ncu_summary_$tag.md 2>> gpurun_out/r10_ncu_full_$tag.log
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r10_ncu_raw_$tag.csv.gz
  ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r10_ncu_source_$tag.csv.gz
  ncu -i /tmp/prof_$tag.ncu-rep --page details 2>/dev/null | head -400 > gpurun_out/r10_ncu_details_$tag.txt
}
for w in c2 c3 c4; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r10_launches_$w.csv python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r10_ncu_launch_$w.log 2>&1
prof ${w}_direct spmv_direct --workload $w
done
prof c3_hot spmv_hot --workload c3 --hot 12288 --hot-threads 1024 --nch 3
prof c2_tma spmv_tma --workload c2 --kernel 2
ls -la gpurun_out | grep r10_; du -sh gpurun_out
