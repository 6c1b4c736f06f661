#!/bin/bash
# round-2 final single-GPU run: tests, bench lines of every workload, reference arm, ncu launch list + full-set captures
mkdir -p gpurun_out
O=gpurun_out
if [ "$SKIP_TESTS" != "1" ]; then timeout 1500 python -m pytest tests -m gpu -q > $O/f1_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/f1_tests.log; fi
timeout 600 python bench.py --steps 10 --warmup 3 > $O/f1_c60.json 2> $O/f1_c60.err; echo "bench c60 rc=$?"
for w in c60-pbe0-df taxol-like-b3lyp-df taxol-like-pbe-df benzene-lda-4c benzene-scan-4c cluster36-pbe-df cluster72-pbe-df; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $O/f1_$w.json 2> $O/f1_$w.err; echo "bench $w rc=$?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/f1_ref.json 2> $O/f1_ref.err; echo "ref rc=$?"
python tools/show_bench.py $O/f1_c60.json $O/f1_c60-pbe0-df.json $O/f1_taxol-like-b3lyp-df.json $O/f1_taxol-like-pbe-df.json $O/f1_benzene-lda-4c.json $O/f1_benzene-scan-4c.json $O/f1_cluster36-pbe-df.json $O/f1_cluster72-pbe-df.json
# launch list of the same command as the bench (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/f1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/f1_launches.log 2>&1; echo "launch list rc=$?"
# full-set captures, one launch each
cap() { # name regex skip script...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/f1_$name "$@" > $O/f1_ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  # the reports are ~14 MB each and gpurun brings back 64 MiB: summarise on the box, keep only the text
  python tools/summarize_ncu.py full $O/f1_$name.ncu-rep $O/f1_full_$name.txt "$name: ncu --set full, one launch ($*)" > /dev/null 2>&1
  if [ "$name" != "rho_ps" ]; then rm -f $O/f1_$name.ncu-rep; fi
  rm -f $O/f1_ncu_$name.log
}
cap rho_ps rho_i8_ps 1 python tools/prof_sb.py c60 2 rho
cap gather sb_gather_slice_dm 1 python tools/prof_sb.py c60 2 rho
cap vbslice vxc_vbslice 2 python tools/prof_sb.py c60 2 vxc
cap vxcgemm vxc_i8_gemm 1 python tools/prof_sb.py c60 2 vxc
cap xc xc_unpol_kernel 1 python tools/prof_sb.py c60 2 both
cap aoeval ao_eval_sb_kernel 0 python tools/prof_sb.py c60 1 rho
cap becke becke_weights_kernel 0 python tools/prof_sb.py c60 1 rho
cap gridasm grid_assemble_kernel 0 python tools/prof_sb.py c60 1 rho
cap dfj1 dfj_pass1_kernel 1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline
cap dfj2 dfj_pass2_kernel 1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline
cap int3c int_dense_kernel 40 python bench.py --workload h2o-pbe-df --steps 1 --warmup 1 --no-cpu-baseline
cap gemv gemv_rows_kernel 2 python bench.py --workload benzene-lda-4c --steps 2 --warmup 1 --no-cpu-baseline
cap interi int_dense_kernel 60 python bench.py --workload benzene-lda-4c --steps 1 --warmup 1 --no-cpu-baseline
export B200QC_ERI_STORE_MAX_BYTES=0
cap jk jk_kernel 10 python bench.py --workload benzene-lda-4c --steps 1 --warmup 1 --no-cpu-baseline
unset B200QC_ERI_STORE_MAX_BYTES
cap dfk gemm_i8_kernel 2 python bench.py --workload c60-pbe0-df --steps 1 --warmup 1 --no-cpu-baseline
cap rhomgga rho_sb_gg_kernel 1 python bench.py --workload benzene-scan-4c --steps 1 --warmup 1 --no-cpu-baseline
du -sh $O; ls $O | head -80
