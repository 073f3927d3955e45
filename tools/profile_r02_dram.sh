# DRAM bytes and duration of every k_dgemm_i8 launch of one un-pipelined and one pipelined sweep (dominant kernel of the
# bench: nu = L Z, the two K* products through L^-1, the f* product) -> profiles/r02_dgemm_i8_dram_bytes.csv
set -x
mkdir -p gpurun_out/r02c/prof
python tools/dev_bench.py c3 2 > gpurun_out/r02c/prof/dev_plain2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_dgemm_i8 --csv \
    --log-file gpurun_out/r02c/prof/r02_dgemm_i8_dram_bytes.csv python tools/dev_bench.py c3 2 > gpurun_out/r02c/prof/ncu_dram.log 2>&1
