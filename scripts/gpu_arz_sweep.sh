# usage: bash scripts/gpu_arz_sweep.sh  -- parity tests, then launch-shape sweep of the ARZ rollout kernels (every state stored)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
K=1
run() { timeout 300 python bench.py --lanes 8192 --micro-lanes 1024 --ckpt-every $K --steps 2 --warmup 1 --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('RESULT K=$K', '$1', '%.3e' % d['value'], d['phase_ms_per_step']['arz_fwd'], d['phase_ms_per_step']['arz_bwd'])"; }
run default
DHTS_ARZ_C_FWD=8 DHTS_ARZ_MB_FWD=2 run fwd_c8_mb2
DHTS_ARZ_C_FWD=8 DHTS_ARZ_MB_FWD=1 run fwd_c8_mb1
DHTS_ARZ_C_BWD=8 DHTS_ARZ_MB_BWD=2 run bwd_c8_mb2
DHTS_ARZ_C_BWD=8 DHTS_ARZ_MB_BWD=1 run bwd_c8_mb1
DHTS_ARZ_C_BWD=4 DHTS_ARZ_MB_BWD=1 run bwd_c4_mb1
DHTS_ARZ_C_BWD=2 DHTS_ARZ_MB_BWD=1 run bwd_c2_mb1
