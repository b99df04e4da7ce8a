#!/bin/bash
# compute-sanitizer over the gather / blend / pointwise kernels (race + memory checks), under gpurun.
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the blend kernels' per-row patch lists, the
# final-conv halo tile, the depthwise conv has none).  The tcgen05 kernels (GEMM, attention) synchronise through mbarriers and
# the async proxy, which racecheck does not model; they are covered by memcheck only, at small shapes.
# Usage: bash scripts/gpu_sanitize.sh [tag]     -> gpurun_out/sanitizer_<tool>_<tag>.log
tag=${1:-r02}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
GEO='crop_resize or roi_gather or blend_fast_paths or blend_ragged or blend_rejects or sharded_blend'
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --error-exitcode 86 --print-limit 20 python -m pytest tests/test_geometry_gpu.py -q -m gpu -x -k "$GEO" -p no:cacheprovider \
    > gpurun_out/sanitizer_${tool}_geometry_${tag}.log 2>&1
  echo "$tool geometry rc=$?"; tail -n 4 gpurun_out/sanitizer_${tool}_geometry_${tag}.log
done
timeout 900 $CS --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_mnv4.py tests/test_network_gpu.py -q -m gpu -x \
  -k "dwconv or encoder_features or (test_attention and 257) or layernorm or resize" -p no:cacheprovider > gpurun_out/sanitizer_memcheck_network_${tag}.log 2>&1
echo "memcheck network rc=$?"; tail -n 4 gpurun_out/sanitizer_memcheck_network_${tag}.log
