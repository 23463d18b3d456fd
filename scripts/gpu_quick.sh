# usage: bash scripts/gpu_quick.sh TAG  -- parity tests + the default bench (fp64)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_f64.json 2> gpurun_out/${TAG}_bench_f64.err
tail -3 gpurun_out/${TAG}_bench_f64.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_f64.json').read().strip().splitlines()[-1])
print('VALUE %.3e'%d['value'], d['phase_ms_per_step'], 'roof', round(d['roofline']['frac'],3), 'e2e %.3e'%d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'])
PY
