#!/bin/bash
# ncu full-set captures of the round-2 kernels: K2 point-stationary, fused vb slicer, K4b
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rho_i8_ps -s 1 -c 1 -f -o gpurun_out/c6_rho_ps python tools/prof_sb.py c60 2 rho > gpurun_out/c6_ncu_rho.log 2>&1
echo "ncu rho rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vxc_vbslice -s 2 -c 1 -f -o gpurun_out/c6_vbslice python tools/prof_sb.py c60 2 vxc > gpurun_out/c6_ncu_vb.log 2>&1
echo "ncu vb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vxc_i8_gemm -s 1 -c 1 -f -o gpurun_out/c6_vxcgemm python tools/prof_sb.py c60 2 vxc > gpurun_out/c6_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
