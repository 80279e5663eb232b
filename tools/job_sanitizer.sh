# GPU job: compute-sanitizer memcheck + racecheck over a small slice of the GPU suite (every library kernel runs at least once).
# usage: bash tools/job_sanitizer.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
SEL='tests/test_gpu_parity.py::test_intermediate_state_bit_exact tests/test_gpu_parity.py::test_edge_cases tests/test_gpu_parity.py::test_camera_gradients tests/test_gpu_parity.py::test_grad_targets_accumulate tests/test_gpu_parity.py::test_fused_rgb_depth_equals_two_passes tests/test_gpu_parity.py::test_huge_splats_and_duplicates tests/test_gpu_slam_ops.py::test_flat_adam_surgery_matches_reference_optimizer_surgery'
for tool in memcheck racecheck; do
  GSR_SANITIZER_SMALL=1 timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL -q -x -m gpu -k "not 1000000 and not 100003 and not 100000 and not 50000" > gpurun_out/${tag}_${tool}.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/${tag}_${tool}.log | tail -5
done
