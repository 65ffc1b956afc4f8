set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_learning.py -x -q) > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
timeout 120 python tools/stats_bench.py 2>gpurun_out/r2c_stats.err | tee gpurun_out/r2c_stats_rows.json; tail -3 gpurun_out/r2c_stats.err
for w in 12 16 17 20 24; do BN_STATS_WARPS=$w timeout 120 python tools/stats_bench.py 2>/dev/null; done | tee gpurun_out/r2c_stats_warps.txt
BN_STATS_NO_TMA=1 timeout 120 python tools/stats_bench.py 2>/dev/null | tee gpurun_out/r2c_stats_notma.json
timeout 120 python tools/stats_bench.py > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:statistics_rows -s 3 -c 1 -f -o gpurun_out/r2c_stats_rows python tools/stats_bench.py > gpurun_out/r2c_ncu_stats.log 2>&1
ls -la gpurun_out/r2c*
