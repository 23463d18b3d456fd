# launch-shape sweep of the ARZ rollout kernels in the bench configuration (every state stored, TMA staging / ring)
run() { timeout 300 python bench.py --lanes 13120 --micro-lanes 1024 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-net 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT', '$1', '%.3e' % d['value'], round(d['phase_ms_per_step']['arz_fwd'],1), round(d['phase_ms_per_step']['arz_bwd'],1))"; }
run default
DHTS_ARZ_C_FWD=2 run fwd_c2
DHTS_ARZ_C_FWD=8 run fwd_c8
DHTS_ARZ_C_BWD=2 run bwd_c2
DHTS_ARZ_C_BWD=8 run bwd_c8
DHTS_ARZ_RING=2 run ring2
DHTS_ARZ_RING=3 run ring3
DHTS_ARZ_RING=6 run ring6
DHTS_ARZ_RING=0 run ring0_regprefetch
DHTS_ARZ_STAGE=0 run fwd_plain_stores
