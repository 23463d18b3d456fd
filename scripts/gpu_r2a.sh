# usage: bash scripts/gpu_r2a.sh TAG -- round-2 first check: new parity tests, small bench (all code paths), full bench
set -x
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt
nproc >> gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests/test_headline_gpu.py tests/test_drivers_gpu.py tests/test_fullsize_gpu.py -x -q -s > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_new.log
tail -5 gpurun_out/${TAG}_pytest_new.log
timeout 600 python bench.py --lanes 1200 --micro-lanes 2048 --sim-steps 100 --steps 1 --warmup 1 --net-replicas 64 > gpurun_out/${TAG}_bench_small.json 2> gpurun_out/${TAG}_bench_small.err; echo "rc=$?"
tail -3 gpurun_out/${TAG}_bench_small.err; cut -c1-1500 gpurun_out/${TAG}_bench_small.json
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_f64.json 2> gpurun_out/${TAG}_bench_f64.err; echo "rc=$?"
tail -3 gpurun_out/${TAG}_bench_f64.err; cut -c1-600 gpurun_out/${TAG}_bench_f64.json
