#!/bin/bash
# One `ncu --set full` capture per hot kernel of the measurement update on the benchmark shape (1024 filters x 256
# features x 30 observations, 30-pose window = one front-end sub-batch of the 4096-filter step).  Run under gpurun;
# reports land in gpurun_out/<tag>_<kernel>.ncu-rep, raw pages in gpurun_out/<tag>_<kernel>.csv.
tag=${1:-r02}
mkdir -p gpurun_out
for k in front_geo_kernel front_asm_ps_kernel front_chol_kernel gram_ws_kernel update_smem_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${tag}_$k \
      python tools/front_probe.py 1024 1 exact update > gpurun_out/${tag}_$k.log 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_$k.csv 2>/dev/null
done
