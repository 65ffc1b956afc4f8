set -x
mkdir -p gpurun_out
nvidia-smi -L
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python __graft_entry__.py --smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2a_smoke.log
python tools/multi_gpu_check.py --all 0,1 --json > gpurun_out/r2a_multi_all.json 2> gpurun_out/r2a_multi_all.err; echo "multi all rc=$?"
cat gpurun_out/r2a_multi_all.json; tail -3 gpurun_out/r2a_multi_all.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py --json > gpurun_out/r2a_multi_rank.json 2> gpurun_out/r2a_multi_rank.err; echo "multi rank rc=$?"
cat gpurun_out/r2a_multi_rank.json; tail -3 gpurun_out/r2a_multi_rank.err
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench_2gpu.json 2> gpurun_out/r2a_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r2a_bench_2gpu.json; tail -5 gpurun_out/r2a_bench_2gpu.err
