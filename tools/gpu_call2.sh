#!/bin/bash
mkdir -p gpurun_out
# timing experiments on the new K2: variants and B-cache sizes (mode = nbc << 12)
timeout 600 python tools/prof_sb.py c60 3 rho 0:0,1:0,2:0,3:0,0:32768,0:49152,0:24576,1:32768 > gpurun_out/c2_prof_rho.log 2>&1
echo "prof rc=$?"
# ncu full-set capture of the new K2 (one launch) and of the fused vb slicer
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rho_i8_ps -s 1 -c 1 -o gpurun_out/c2_rho_ps python tools/prof_sb.py c60 2 rho > gpurun_out/c2_ncu_rho.log 2>&1
echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vxc_vbslice -s 1 -c 1 -o gpurun_out/c2_vbslice python tools/prof_sb.py c60 2 vxc > gpurun_out/c2_ncu_vb.log 2>&1
echo "ncu2 rc=$?"
tail -30 gpurun_out/c2_prof_rho.log
