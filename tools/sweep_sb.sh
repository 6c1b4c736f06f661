for sbp in 256 512 1024 2048; do for eps in 1e-12 1e-10; do
echo "== SBP $sbp EPS $eps"; B200QC_SB_POINTS=$sbp B200QC_AO_SCREEN=$eps python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('ms %.2f kept %.3f rho %.2f vxcgemm %.2f vb %.2f gather %.2f aoGB %.1f'%(d['value'], d['config']['kept_ao_fraction'], k['rho_kernel']['ms_per_launch'], k['vxc_gemm_kernel']['ms_per_launch'], k['vxc_vb_kernel']['ms_per_launch'], k['sb_gather_dm_kernel']['ms_per_launch'], d['config']['ao_resident_gb']))"
done; done
