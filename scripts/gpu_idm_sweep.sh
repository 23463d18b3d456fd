for K in 32 16 8 4 2 1; do
  timeout 300 python bench.py --lanes 296 --sim-steps 1000 --idm-ckpt-every $K --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-net 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT K=$K idm %.3e' % d['idm']['value'], round(d['phase_ms_per_step']['idm_fwd'],1), round(d['phase_ms_per_step']['idm_bwd'],1))"
done
