# usage: bash scripts/gpu_round2_check.sh TAG -- the full check of round 2: pytest -m gpu, bench.py (all sections), ncu launch list of the
# bench command, ncu --set full of a bench-shaped lane chunk (lane kernels) and of the network kernels (592 replicas)
set -x
TAG=${1:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_f64.json 2> gpurun_out/${TAG}_bench_f64.err
tail -2 gpurun_out/${TAG}_bench_f64.err; cut -c1-600 gpurun_out/${TAG}_bench_f64.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-net --no-drivers --no-parity > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"arz_rollout|idm_rollout" -c 4 -o gpurun_out/${TAG}_full python bench.py --lanes 6560 --micro-lanes 65536 --sim-steps 256 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-net --no-drivers --no-parity > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_full_sass.csv 2>/dev/null
rm -f gpurun_out/${TAG}_full.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:"net_rollout|hyb_rollout" -c 4 -o gpurun_out/${TAG}_net python scripts/net_prof.py > gpurun_out/${TAG}_ncu_net.log 2>&1
ncu -i gpurun_out/${TAG}_net.ncu-rep --page raw --csv > gpurun_out/${TAG}_net_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_net.ncu-rep
ls -la gpurun_out | tail -12
