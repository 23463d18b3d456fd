# usage: bash scripts/gpu_sanitize5.sh -- round 2: compute-sanitizer over the kernels this round changed (interface-parallel network
# kernels with their shared-memory phases, waiting-list sources, stored interface outcomes / ballots / uniform geometry of the ARZ
# rollouts, IDM EXT instantiation)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; tag=$2; shift 2
  timeout 420 $SAN --tool $tool --error-exitcode 77 --log-file gpurun_out/san5_${tool}_${tag}.log python -m pytest -m gpu -q -x "$@" > gpurun_out/san5_${tool}_${tag}.out 2>&1
  echo "$tool $tag rc=$? | $(grep -E 'SUMMARY' gpurun_out/san5_${tool}_${tag}.log | tail -1) | $(tail -1 gpurun_out/san5_${tool}_${tag}.out)"
}
run racecheck arz "tests/test_arz_gpu.py::test_rollout_vs_oracle_shapes" "tests/test_arz_gpu.py::test_rollout_with_vacuum_vs_oracle" "tests/test_arz_gpu.py::test_rollout_tma_paths_equal_plain_paths"
run memcheck arz tests/test_arz_gpu.py
run racecheck idm tests/test_idm_gpu.py
run racecheck net tests/test_net_gpu.py::test_itscp_macro_matches_live_reference
run racecheck hyb "tests/test_hyb_gpu.py::test_hybrid_itscp_matches_live_reference" tests/test_hyb_gpu.py::test_per_vehicle_idm_parameters_match_live_reference
run racecheck micro "tests/test_itscp_micro_gpu.py::test_micro_episode_matches_live_reference"
run memcheck nethyb tests/test_net_gpu.py tests/test_hyb_gpu.py "tests/test_itscp_micro_gpu.py::test_micro_episode_matches_live_reference"
run synccheck all tests/test_arz_gpu.py::test_rollout_vs_oracle_shapes tests/test_net_gpu.py::test_itscp_macro_matches_live_reference "tests/test_hyb_gpu.py::test_hybrid_itscp_matches_live_reference"
