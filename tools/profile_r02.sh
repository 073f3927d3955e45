set -x
mkdir -p gpurun_out/r02c/prof
cd /root/repo
# 1. plain run first (must exit 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02c/prof/bench_plain.json 2> gpurun_out/r02c/prof/bench_plain.err || exit 1
# 2. launch list
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02c/prof/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02c/prof/ncu_list.log 2>&1
# 3. full captures (one launch each) on the developer benchmark (same sampler, c3)
python tools/dev_bench.py c3 2 > gpurun_out/r02c/prof/dev_plain.log 2>&1 || exit 1
for spec in "k_diag128:40:r02_diag128" "k_panel_update:40:r02_panel_update" "k_dgemm_i8:9:r02_dgemm_i8" "k_ess_persist:2:r02_ess_persist"; do
  k=$(echo $spec | cut -d: -f1); s=$(echo $spec | cut -d: -f2); o=$(echo $spec | cut -d: -f3)
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/r02c/prof/$o -f python tools/dev_bench.py c3 2 > gpurun_out/r02c/prof/ncu_$o.log 2>&1
done
# the bulk rank-128 update of the factorisation (gemm_f64_kernel 128x128, NT): first big one
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 200 -c 1 -o gpurun_out/r02c/prof/r02_gemm_f64_bulk -f python tools/dev_bench.py c3 2 > gpurun_out/r02c/prof/ncu_gemm.log 2>&1
ls -la gpurun_out/r02c/prof
