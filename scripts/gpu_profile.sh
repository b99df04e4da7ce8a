#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command (cold-cache, serialised: compare SHARES),
# (2) --set full captures of the hot kernels at ViT-L shapes through scripts/prof_kernels.py.
# Usage (under gpurun): bash scripts/gpu_profile.sh [tag]
tag=${1:-r01}
mkdir -p gpurun_out
# one frame = ~2034 launches at ViT-L r32 with patch_batch 12; skip the 3 warm-up frames
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6102 -c 2100 --csv \
  --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
  > gpurun_out/bench_under_ncu_${tag}.log 2>&1
echo "launch list rc=$?"
for k in qkv proj attn conv blend; do
  case $k in attn) rx=attention;; blend) rx=blend;; *) rx=umma_gemm;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 3 -f -o gpurun_out/prof_${k}_${tag} \
    python scripts/prof_kernels.py $k > gpurun_out/prof_${k}_${tag}.log 2>&1
  echo "prof $k rc=$?"
done
ls -la gpurun_out
