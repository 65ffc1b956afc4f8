# full round-2 verification on one GPU: all GPU tests, smoke, bench (+ ncu launch list after the plain run)
set -x
mkdir -p gpurun_out
TAG=${1:-r2f}
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/${TAG}_smoke.log
(time timeout 900 python bench.py) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
# ncu launch list of the same bench command (kernel shares of the step), after the plain run exited 0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cold > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
