#!/bin/bash
# One `ncu --set full` capture per hot kernel of the current build (run on the GPU box under gpurun):
#   bash scripts/ncu_capture.sh            -> gpurun_out/r02_{fused,tail}_{fp32,bf16}.ncu-rep
# One 256x256 / 64-sample scene = 4 chunks of 16384 rays; the 2nd launch of each kernel is captured.
# Summaries (scripts/ncu_summary.py, keyed by build id) go to profiles/r02_ncu_*.json.
set -u
mkdir -p gpurun_out
for prec in ${PRECS:-fp32 bf16}; do
  ncu --set full --clock-control none --import-source on -k regex:k_fused_encode -s 1 -c 1 -f \
      -o gpurun_out/r02_fused_${prec} python scripts/render_once.py ${prec} 1 > gpurun_out/ncu_fused_${prec}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_tail -s ${TAIL_SKIP:-2} -c ${TAIL_COUNT:-2} -f \
      -o gpurun_out/r02_tail_${prec} python scripts/render_once.py ${prec} 1 > gpurun_out/ncu_tail_${prec}.log 2>&1
done
for prec in ${PRECS:-fp32 bf16}; do
  for k in fused tail; do
    [ -f gpurun_out/r02_${k}_${prec}.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/r02_${k}_${prec}.ncu-rep > gpurun_out/r02_ncu_${k}_${prec}.json
  done
done
# gpurun merges at most 64 MiB back: keep the fused-kernel reports (source page, stall reasons), drop the tail's
rm -f gpurun_out/r02_tail_*.ncu-rep
ls -la gpurun_out/*.ncu-rep
