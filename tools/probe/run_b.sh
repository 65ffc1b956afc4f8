set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -4 gpurun_out/r2b_tests.log
python tools/stats_bench.py > gpurun_out/r2b_stats_rows.json 2> gpurun_out/r2b_stats_rows.err; cat gpurun_out/r2b_stats_rows.json; tail -3 gpurun_out/r2b_stats_rows.err
BN_STATS_NODELANES=1 python tools/stats_bench.py > gpurun_out/r2b_stats_nodelanes.json 2>/dev/null; cat gpurun_out/r2b_stats_nodelanes.json
for w in 20 25 32; do BN_STATS_WARPS=$w python tools/stats_bench.py 2>/dev/null; done > gpurun_out/r2b_stats_warps.txt; cat gpurun_out/r2b_stats_warps.txt
# ncu: top kernels, each after its plain run
python tools/stats_bench.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:statistics_rows -s 3 -c 1 -f -o gpurun_out/r2b_stats_rows python tools/stats_bench.py > gpurun_out/r2b_ncu_stats.log 2>&1
python bench.py --steps 2 --warmup 3 --no-configs --no-cold --no-factors > gpurun_out/r2b_bench_short.json 2>/dev/null && ncu --set full --clock-control none --import-source on -k regex:lw_spec -s 4 -c 1 -f -o gpurun_out/r2b_lw_spec python bench.py --steps 2 --warmup 3 --no-configs --no-cold --no-factors > gpurun_out/r2b_ncu_lw.log 2>&1
python tools/gibbs_bench.py > gpurun_out/r2b_gibbs_bench.json 2>/dev/null && ncu --set full --clock-control none --import-source on -k regex:gibbs_spec -s 3 -c 1 -f -o gpurun_out/r2b_gibbs_spec python tools/gibbs_bench.py > gpurun_out/r2b_ncu_gibbs.log 2>&1
cat gpurun_out/r2b_gibbs_bench.json
ls -la gpurun_out/*.ncu-rep
