# usage: bash scripts/gpu_try.sh TAG -- ARZ parity tests, then the device-resident ARZ/IDM bench only
TAG=${1:-try}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_arz_gpu.py tests/test_fullsize_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-net --no-cpu-baseline --no-e2e > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
echo "rc=$?"; tail -2 gpurun_out/${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}.json').read().strip().splitlines()[-1])
print('${TAG} VALUE %.3e'%d['value'], {k: round(v,1) for k,v in d['phase_ms_per_step'].items()}, 'losses', d['losses'], 'roof', round(d['roofline']['frac'],3))
PY
