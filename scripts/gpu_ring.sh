# usage: bash scripts/gpu_ring.sh TAG -- ARZ parity tests, quick ARZ timing at one lane chunk, ncu full of both ARZ rollout kernels
TAG=${1:-r1g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_arz_gpu.py tests/test_fullsize_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -15
run() { timeout 300 python bench.py --lanes 6560 --micro-lanes 1024 --ckpt-every 1 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('RESULT', '$1', '%.3e' % d['value'], d['phase_ms_per_step']['arz_fwd'], d['phase_ms_per_step']['arz_bwd'])"; }
run default
for v in $SWEEP; do env $v bash -c "$(declare -f run); run $v"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"arz_rollout" -c 2 -o gpurun_out/${TAG}_full python bench.py --lanes 2368 --micro-lanes 1024 --sim-steps 64 --ckpt-every 1 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_full_sass.csv 2>/dev/null
ls -la gpurun_out | tail -6
