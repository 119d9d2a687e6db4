#!/bin/bash
# compute-sanitizer over the kernels added / rewritten in round 2 (development aid; run under gpurun).
mkdir -p gpurun_out
python -c 'import torch' >/dev/null 2>&1
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitizer.log \
    python -m pytest -x -q -p no:cacheprovider \
    "tests/test_pathology_ops_gpu.py" \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_s1]" \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_pathol_s7]" \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_full_s4]" \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_svf_s31]" \
    "tests/test_gen_parity_gpu.py::test_pair_mode_is_identical_to_the_unpaired_gather" \
    "tests/test_native_planner_gpu.py::test_native_mode_is_deterministic_and_sane" \
    "tests/test_misc_gpu.py" "tests/test_solvers_gpu.py" > gpurun_out/r2_sanitizer_pytest.log 2>&1
echo "memcheck tests exit $?"; tail -3 gpurun_out/r2_sanitizer_pytest.log; tail -2 gpurun_out/r2_sanitizer.log
BFM_SLAB_ONE_GPU=1 BFM_SLAB_PLANNER=native timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_slab.log \
    python tests/_slab_noise_worker.py > gpurun_out/r2_sanitizer_slab.out 2>&1
echo "memcheck slab exit $?"; tail -2 gpurun_out/r2_sanitizer_slab.out; tail -2 gpurun_out/r2_sanitizer_slab.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_bench.log \
    python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2_sanitizer_bench.out 2>&1
echo "memcheck bench exit $?"; tail -2 gpurun_out/r2_sanitizer_bench.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2_racecheck.log \
    python -m pytest -x -q -p no:cacheprovider \
    "tests/test_gen_parity_gpu.py::test_chain_matches_oracle[g64_s5_lowres]" \
    "tests/test_gen_parity_gpu.py::test_pair_mode_is_identical_to_the_unpaired_gather" \
    "tests/test_pathology_ops_gpu.py" > gpurun_out/r2_racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -2 gpurun_out/r2_racecheck_pytest.log; tail -3 gpurun_out/r2_racecheck.log
