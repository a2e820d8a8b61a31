#!/bin/bash
# compute-sanitizer passes over the operator tests and one small whole-model test (memory errors, then
# shared-memory races).  The tcgen05/TMA kernels are exercised through the same tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDNET_B200_CUDA_GRAPH=0
echo "== memcheck: ops"; timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 1 python -m pytest tests/test_ops_gpu.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/sanitize_memcheck_ops.log
echo "== memcheck: smallest golden model case"; timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 1 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "td2_r34_80x112-tc or forward_labels or forward_u8" 2>&1 | tail -15 | tee gpurun_out/sanitize_memcheck_model.log
echo "== racecheck: ops (pointwise + simt + stem)"; timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "not tc_conv and not attention" 2>&1 | tail -15 | tee gpurun_out/sanitize_racecheck_ops.log
