# usage: bash scripts/gpu_round_check.sh TAG   -- parity tests, full bench (f64, f32), launch list, ncu full capture
set -x
TAG=${1:-r1x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_f64.json 2> gpurun_out/${TAG}_bench_f64.err
tail -2 gpurun_out/${TAG}_bench_f64.err; cut -c1-700 gpurun_out/${TAG}_bench_f64.json
timeout 600 python bench.py --steps 3 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/${TAG}_bench_f32.json 2> gpurun_out/${TAG}_bench_f32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-net > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"arz_rollout|idm_rollout" -c 4 -o gpurun_out/${TAG}_full python bench.py --lanes 2368 --micro-lanes 16384 --sim-steps 64 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-net > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_full_sass.csv 2>/dev/null
ls -la gpurun_out | tail -12
