#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command with device time + DRAM bytes per launch (cold-cache,
# serialised: compare SHARES), (2) --set full captures of the hot kernels at ViT-L shapes through scripts/prof_kernels.py.
# Usage (under gpurun): bash scripts/gpu_profile.sh [tag]
tag=${1:-r01}
mkdir -p gpurun_out
RX='regex:umma_gemm|final_conv|attention_|layernorm_kernel|patchify|assemble_tokens|resize_act|depth_taps|tap_stencil|phase_split|crop_resize|roi_gather|blend_|dwconv|encoder_input'
# one frame = 976 library calls = 1072 kernels at ViT-L r32 with patch_batch 27 (every attention call is the tile kernel + the
# one-row tail kernel); skip the warm-up frame
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$RX" -s 1072 -c 1072 --csv \
  --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --profile-run --no-cpu-baseline --no-e2e --no-fp32-mode --no-parity \
  > gpurun_out/bench_under_ncu_${tag}.log 2>&1
echo "launch list rc=$?"
for k in qkv fc1 fc2 proj conv attn blend ln; do
  case $k in attn) rx=attention_v6;; blend) rx=blend_;; resize) rx=resize_act;; finalconv) rx=final_conv;; ln) rx=layernorm_kernel;; *) rx=umma_gemm;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 3 -f -o gpurun_out/prof_${k}_${tag} \
    python scripts/prof_kernels.py $k > gpurun_out/prof_${k}_${tag}.log 2>&1
  echo "prof $k rc=$?"
  # the raw page as CSV travels even when a large .ncu-rep does not
  ncu -i gpurun_out/prof_${k}_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${k}_${tag}_raw.csv 2>/dev/null
  # gpurun copies back at most 64 MiB: keep the full report only for the small non-GEMM kernels, the top stall sites for all
  python scripts/ncu_top_stalls.py gpurun_out/prof_${k}_${tag}.ncu-rep 30 > gpurun_out/prof_${k}_${tag}_stalls.txt 2>&1
  case $k in attn|blend|finalconv) ;; *) rm -f gpurun_out/prof_${k}_${tag}.ncu-rep;; esac
done
du -sh gpurun_out
ls -la gpurun_out | head -40
