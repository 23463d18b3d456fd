# usage: bash scripts/gpu_r2c.sh TAG -- rest of the GPU suite, A/B of the 80-register forward kernel, bench-shaped ncu capture
set -x
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fullsize_gpu.py tests/test_headline_gpu.py tests/test_hyb_gpu.py tests/test_idm_gpu.py tests/test_inverse_gpu.py tests/test_itscp_env_gpu.py tests/test_net_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-net --no-drivers --no-parity"
$B > gpurun_out/${TAG}_ab_base.json 2>/dev/null
DHTS_ARZ_MB_FWD=3 $B > gpurun_out/${TAG}_ab_mb3.json 2>/dev/null
for f in base mb3; do python -c "
import json,sys; d=json.load(open('gpurun_out/${TAG}_ab_$f.json')); print('$f', d['value'], d['phase_ms_per_step'])"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"arz_rollout|idm_rollout" -c 4 -o gpurun_out/${TAG}_full python bench.py --lanes 6560 --micro-lanes 65536 --sim-steps 256 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-net --no-drivers --no-parity > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_full_sass.csv 2>/dev/null
ls -la gpurun_out | tail -6
