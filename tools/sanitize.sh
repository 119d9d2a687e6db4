#!/bin/bash
# compute-sanitizer over a slice of the GPU tests and one bench step (development aid; run under gpurun).
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer.log \
    python -m pytest -x -q \
    "tests/test_native_planner_gpu.py::test_replayed_native_chain_matches_oracle[g64_s1]" \
    "tests/test_native_planner_gpu.py::test_replayed_native_brainid_matches_oracle" \
    "tests/test_native_planner_gpu.py::test_native_mode_is_deterministic_and_sane" \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_s5_lowres]" \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_left_s9]" \
    "tests/test_interpol_gpu.py::test_tiled_prefilter_matches_line_kernel_and_add_identity" \
    "tests/test_interpol_gpu.py::test_fast_pull_paths_match_oracle" \
    "tests/test_shapeid_gpu.py" > gpurun_out/sanitizer_pytest.log 2>&1
echo "memcheck tests exit $?"; tail -2 gpurun_out/sanitizer_pytest.log; tail -2 gpurun_out/sanitizer.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_bench.log \
    python bench.py --steps 2 --warmup 3 --quick > gpurun_out/sanitizer_bench.out 2>&1
echo "memcheck bench exit $?"; tail -2 gpurun_out/sanitizer_bench.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
    python -m pytest -x -q \
    "tests/test_native_planner_gpu.py::test_native_mode_is_deterministic_and_sane" \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_s5_lowres]" \
    "tests/test_interpol_gpu.py::test_tiled_prefilter_matches_line_kernel_and_add_identity" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -2 gpurun_out/racecheck_pytest.log; tail -3 gpurun_out/racecheck.log
