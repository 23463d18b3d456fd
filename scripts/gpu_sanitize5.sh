# usage: bash scripts/gpu_sanitize5.sh -- round 2: compute-sanitizer over the kernels this round changed (interface-parallel network
# kernels with their shared-memory phases, waiting-list sources, stored interface outcomes of the ARZ rollouts)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; tag=$2; shift 2
  timeout 700 $SAN --tool $tool --error-exitcode 77 --log-file gpurun_out/san5_${tool}_${tag}.log python -m pytest -m gpu -q "$@" > gpurun_out/san5_${tool}_${tag}.out 2>&1
  echo "$tool $tag rc=$? | $(grep -E 'SUMMARY' gpurun_out/san5_${tool}_${tag}.log | tail -1) | $(tail -1 gpurun_out/san5_${tool}_${tag}.out)"
}
run racecheck net tests/test_net_gpu.py::test_itscp_macro_matches_live_reference "tests/test_net_gpu.py::test_random_networks_match_checker"
run racecheck hyb "tests/test_hyb_gpu.py::test_hybrid_itscp_matches_live_reference" tests/test_hyb_gpu.py::test_plain_mode_chain_matches_live_reference tests/test_hyb_gpu.py::test_per_vehicle_idm_parameters_match_live_reference
run racecheck micro "tests/test_itscp_micro_gpu.py::test_micro_episode_matches_live_reference"
run racecheck arz tests/test_arz_gpu.py
run memcheck nethyb tests/test_net_gpu.py tests/test_hyb_gpu.py "tests/test_itscp_micro_gpu.py::test_micro_episode_matches_live_reference"
run synccheck nethyb tests/test_net_gpu.py::test_itscp_macro_matches_live_reference "tests/test_hyb_gpu.py::test_hybrid_itscp_matches_live_reference"
