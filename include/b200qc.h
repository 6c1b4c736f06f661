/*
 * b200qc -- C ABI of the B200-native SCF Fock-build path.
 *
 * This is the drop-in boundary.  The reference (diffqc/dqc) reaches its native code through
 * ctypes handles to `dqclibs` (dqc/hamilton/intor/utils.py:15-19) and `pylibxc`
 * (dqc/xc/libxc.py:25-26); every entry point below names the reference call site / native symbol
 * it replaces.  Conventions (same ownership rule as dqclibs):
 *   - the caller allocates every output; the library only writes;
 *   - every data pointer is a DEVICE pointer (fp64 / int32) unless the name starts with `h_`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous
 *     with respect to the host unless stated otherwise;
 *   - return value 0 = success, non-zero = failure, message via b200qc_last_error();
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 * Basis data travel in libcint's (atm, bas, env) layout exactly as the reference packs them
 * (dqc/hamilton/intor/lcintwrap.py:36-117).
 */
#ifndef B200QC_H
#define B200QC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200qc_basis b200qc_basis; /* device-resident (atm, bas, env, ao_loc) */

const char *b200qc_last_error(void);
int b200qc_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t b200qc_launch_count(void);

/* per-kernel device timing for the benchmark: enable, run, read (ms_total / counts hold
 * b200qc_profile_nkernels() entries, names from b200qc_profile_name); read synchronises */
int b200qc_profile(int on);
int b200qc_profile_nkernels(void);
const char *b200qc_profile_name(int id);
int b200qc_profile_read(double *h_ms_total, int64_t *h_counts);
/* measured fp64 tensor-pipe (DMMA m8n8k4) peak of this device in TFLOP/s: the roofline
 * denominator of K2 / K4 (MEASURED_PEAKS.json has no fp64 entry).  scratch: >= 1 double (device) */
int b200qc_peak_fp64_dmma(int iters, double *scratch, double *h_tflops, void *stream);
/* measured plain fp64 pipe (DFMA) peak of this device in TFLOP/s: the roofline denominator of the J/K quartet kernels */
int b200qc_peak_fp64_fma(int iters, double *scratch, double *h_tflops, void *stream);
/* measured tcgen05.mma.kind::i8 issue-rate peak of this device in TOP/s (2 ops per multiply-add; one CTA per SM,
 * M = 128, N = 256, K = 32 MMAs on resident shared-memory operands): the denominator of the int8 (sliced fp64)
 * tensor rooflines -- bench.py measures it in-process instead of assuming 2 x the bf16 figure. */
int b200qc_peak_i8_mma(int iters, double *h_tops, void *stream);

/* ---- basis / tables ------------------------------------------------------------------ */
/* Replaces the (atm, bas, env) argument pack every dqclibs call receives
 * (molintor.py:629-638, gtoeval.py:219-233) and CINTcgto_spheric (lcintwrap.py:376-383):
 * the triplet is uploaded once and referred to by handle.  h_* are host pointers. */
int b200qc_basis_upload(const int *h_atm, int natm, const int *h_bas, int nbas, const double *h_env,
                        int nenv, const int *h_ao_loc, b200qc_basis **out);
int b200qc_basis_free(b200qc_basis *basis);
/* flag != 0: every integral entry point below returns RAW CARTESIAN blocks x^a y^b z^c sum_p c_p exp(-a_p r^2) for the
 * shells of this handle (libcint component order; the upload's h_ao_loc must then count (l + 1)(l + 2) / 2 functions
 * per shell).  Derivative integrals -- libcint's int1e_ip*, int2e_ip1 behind molintor.py:178-578 -- are assembled from
 * such blocks over shells of l + 1 and l - 1 (dqc_b200/hamilton/intor/deriv.py). */
int b200qc_basis_set_cartesian(b200qc_basis *basis, int flag);
/* the kernels' cartesian -> real-spherical matrix of angular momentum l: h_out[(2 l + 1)][(l + 1)(l + 2) / 2] (host) */
int b200qc_c2s_matrix(int l, double *h_out);
/* Rys-quadrature interpolation table (dqc_b200/data/rys_table.npz, tools/make_rys_table.py);
 * plays the role of libcint's built-in root tables / the `<intor>_optimizer` pair data
 * (molintor.py:695-708).  h_coef[n-1] points to (nint, 2n, deg+1) doubles, h_herm[n-1] to (2, n). */
int b200qc_rys_upload(int nmax, double h, int deg, double xmax, const double *const *h_coef,
                      const double *const *h_herm);
/* host arithmetic only: the table of one root count split into nsub sub-intervals per interval and truncated to ncoef
 * Chebyshev coefficients -- h_coef (nint, nf, deg + 1) -> h_out (nint * nsub, nf, ncoef).  b200qc_rys_upload keeps the
 * (2, 10) form beside the base table for the register-resident J/K engine (10 Clenshaw steps instead of 14). */
int b200qc_rys_refine(const double *h_coef, int nint, int nf, int deg, int nsub, int ncoef, double *h_out);

/* ---- K1: AO values on the grid -- replaces GTOval_sph / GTOval_ip_sph ------------------ */
/* gtoeval.py:196-260 (eval_gto / eval_gradgto / eval_laplgto with to_transpose=True).
 * coords: (ngrid, 3).  ao: [ncomp][ngrid_ld][ao_ld] with ncomp = 1 (deriv 0: values), 4 (deriv 1: value, d/dx,
 * d/dy, d/dz) or 5 (deriv 2: those and the Laplacian, the meta-GGA storage of hcgto.py:183-186); rows >= ngrid
 * and columns >= nao are left untouched. */
int b200qc_eval_gto(const b200qc_basis *basis, int sh0, int sh1, int deriv, const double *coords,
                    int64_t ngrid, double *ao, int64_t ngrid_ld, int64_t ao_ld, void *stream);

/* ---- molecular grid points -- replaces the host construction of lebedev_grid.py:33-102 and the per-atom translation
 * of multiatoms_grid.py:158-171.  Per atom TYPE: radial nodes [type_node_off[t], type_node_off[t+1]) with radius node_r,
 * radial weight node_dv, angular rule = rows [node_ang_off[n], ...) of ang (n, 5) = (sin theta, cos theta, sin phi,
 * cos phi, w) and first point node_pt_off[n] inside the atomic grid; atom a has type atom_type[a] and owns the points
 * [atom_pt_off[a], atom_pt_off[a+1]).  Outputs xyz (ngrid, 3), dvol (ngrid) = radial x angular weight (no partition
 * weight yet), owner (ngrid).  All pointers device. */
int b200qc_grid_assemble(int natom, const double *atompos, const int *atom_type, const int64_t *atom_pt_off,
                         const int *type_node_off, const double *node_r, const double *node_dv, const int *node_ang_off,
                         const int *node_pt_off, const double *ang, int64_t ngrid, double *xyz, double *dvol, int *owner,
                         void *stream);

/* ---- grid weights -- replaces the torch loop of multiatoms_grid.py:173-273 ------------- */
/* owner[g] = atom the point belongs to; aij = (natom, natom) hetero-nuclear shifts or NULL;
 * w[g] = P_owner / sum_k P_k. */
int b200qc_becke_weights(const double *xyz, const int *owner, int64_t ngrid, const double *atompos,
                         int natom, const double *aij, double *w, void *stream);

/* ---- K2: density on the grid -- replaces hcgto.py:371-443 (_dm2densinfo) --------------- */
/* dm: (ao_ld, ao_ld) symmetric AO-basis density, zero padded.  rho: (ngrid_ld);
 * grad: (3, ngrid_ld) or NULL (LDA).  ncomp of `ao` is 1 if grad == NULL else 4.
 * ngrid_ld must be a multiple of 128 and ao_ld a multiple of 64 (zero padded). */
int b200qc_rho(const double *ao, int64_t ngrid_ld, int64_t ao_ld, const double *dm, double *rho,
               double *grad, void *stream);

/* ---- K3: pointwise XC -- replaces pylibxc LibXCFunctional.compute ---------------------- */
/* libxc_wrapper.py:380-413 with the pre/post-processing of libxc.py:124-242 fused in:
 * input rho (n), grad (3, n) or NULL; outputs (each may be NULL):
 *   edens = sum_k coef_k zk_k rho            (energy per unit volume)
 *   vrho  = sum_k coef_k d e_k / d rho
 *   vgrad = sum_k coef_k 2 (d e_k/d sigma) grad   (3, n)
 * func ids: 1 lda_x, 2 lda_c_pw, 3 lda_c_pw_mod, 101 gga_x_pbe, 102 gga_c_pbe. */
int b200qc_xc_unpol(int nterm, const int *h_func_ids, const double *h_coefs, int64_t n, int64_t ld,
                    const double *rho, const double *grad, double *edens, double *vrho, double *vgrad,
                    void *stream);
/* meta-GGA (family 4) sums, unpolarised -- replaces the mgga branch of libxc.py:124-242 / libxc_wrapper.py:380-413:
 * inputs rho (n), grad (3, ld), lapl (n; accepted, no built-in functional reads it), tau (n); outputs edens, vrho,
 * vgrad = 2 (de/dsigma) grad, vlapl = de/d(lapl rho) (zero), vtau = de/dtau.  Extra func id: 201 mgga_x_scan (the
 * closed form of dqc/test/test_xc.py:427-455); LDA / GGA ids may be mixed in. */
int b200qc_xc_mgga_unpol(int nterm, const int *h_func_ids, const double *h_coefs, int64_t n, int64_t ld,
                         const double *rho, const double *grad, const double *lapl, const double *tau, double *edens,
                         double *vrho, double *vgrad, double *vlapl, double *vtau, void *stream);
/* spin-polarised variant: rho (2, n) = up, down; grad (2, 3, n); vrho (2, n); vgrad (2, 3, n) */
int b200qc_xc_pol(int nterm, const int *h_func_ids, const double *h_coefs, int64_t n, int64_t ld,
                  const double *rho, const double *grad, double *edens, double *vrho, double *vgrad,
                  void *stream);

/* ---- K4: Vxc integration -- replaces hcgto.py:445-495 (_get_vxc_from_potinfo) ---------- */
/* mat (ao_ld, ao_ld) = sum_g w_g phi_g^T (vrho_g phi_g + sum_d 2 vgrad_{d,g} d_d phi_g)
 * (no symmetrisation, no basis change: the caller does X^T M X and (M + M^T)/2 like the
 * reference).  work: scratch of b200qc_vxc_worksize() doubles. */
int64_t b200qc_vxc_worksize(int64_t ngrid_ld, int64_t ao_ld);
int b200qc_vxc_mat(const double *ao, int64_t ngrid_ld, int64_t ao_ld, const double *weights,
                   const double *vrho, const double *vgrad, double *mat, double *work, void *stream);

/* ---- block-sparse grid path ("superblocks") -- the production form of K1 / K2 / K4 -------- */
/* The grid is cut into superblocks (SB) of `sbp` consecutive points (multiple of 128); each SB keeps
 * only the shells whose envelope can exceed eps on it (the reference keeps everything: non0tab = 1,
 * gtoeval.py:211; eps = 0 reproduces that).  Per-SB descriptor (32 bytes, device array):
 *   { int64 ao_off; int64 d_off; int32 nsp; int32 idx_off; int32 shell_off; int32 nshell; }
 * ao_off: start of the SB's compact AO block [ncomp][sbp][nsp] (doubles); d_off: start of its gathered
 * density (nsp x nsp) in the scratch; nsp: kept AOs padded to a multiple of 64; idx: AO index of every
 * compact column (padding = nao); shell_ids / shell_col: kept shells and their first compact column. */
/* flags: (nsb, sh1 - sh0) bytes, 1 = shell kept on that SB; deriv = 1 also bounds the gradient, 2 the Laplacian */
int b200qc_ao_screen(const b200qc_basis *basis, int sh0, int sh1, const double *coords, int64_t ngrid, int sbp,
                     double eps, int deriv, unsigned char *flags, void *stream);
/* compact AO values (pre-zeroed buffer); same arithmetic as b200qc_eval_gto */
int b200qc_eval_gto_sb(const b200qc_basis *basis, int deriv, const double *coords, int64_t ngrid, int sbp, int nsb,
                       const void *sbdesc, const int *shell_ids, const int *shell_col, double *ao, void *stream);
/* rho (nsb * sbp), grad (3, nsb * sbp) or NULL from dm (nao, nao) symmetric AO-basis density
 * (hcgto.py:371-443); dsb: scratch of sum_sb nsp^2 doubles */
int b200qc_rho_sb(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                  const double *dm, int nao, double *dsb, double *rho, double *grad, void *stream);
/* mat (nao, nao) = sum_g w phi^T (vrho phi + 2 vgrad . grad phi) (hcgto.py:445-495), overwritten;
 * vb: scratch of sum_sb sbp * nsp doubles with per-SB offsets vb_off (device int64) */
int b200qc_vxc_sb(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                  const double *weights, const double *vrho, const double *vgrad, int nao, const int64_t *vb_off,
                  double *vb, double *mat, void *stream);

/* Meta-GGA forms of the two contractions on the 5-component storage [phi, dx, dy, dz, lapl] (b200qc_eval_gto_sb with
 * deriv = 2), fp64 DMMA engine -- hcgto.py:420-438 and 473-489:
 *   lapl = 2 (sum X lapl phi + gg), kin = gg / 2, gg = sum_d (d_d phi) D (d_d phi), X = phi D;
 *   mat = sum_g w [phi^T (vrho phi + 2 vgrad . grad phi + 2 vlapl lapl phi) + sum_d (d_d phi)^T (2 vlapl + vkin / 2) (d_d phi)]. */
int b200qc_rho_sb_mgga(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                       const double *dm, int nao, double *dsb, double *rho, double *grad, double *lapl, double *kin,
                       void *stream);
int b200qc_vxc_sb_mgga(const void *sbdesc, int nsb, int sbp, int max_nsp, const int *idx, const double *ao,
                       const double *weights, const double *vrho, const double *vgrad, const double *vlapl,
                       const double *vkin, int nao, const int64_t *vb_off, double *vb, double *mat, void *stream);

/* K4 on tcgen05: the same contraction as b200qc_vxc_sb with the GEMM done as an error-free sliced int8
 * product (Ozaki scheme, nslice = 5 or 6 slices of 7 bits, tcgen05.mma.kind::i8 with int32 accumulation in
 * TMEM, fp64 recombination; |error| ~ 1e-10 / 1e-12 of the fp64 result).  prepare slices the static AO
 * values once into the tiled operand order: aplanes = sum_sb nslice * sbp * ceil(nsp / 128) * 128 bytes
 * (zero-filled by the caller) at a_off[sb], ascale = sum_sb nsp doubles.  bplanes (sum_sb nslice * sbp * nsp
 * bytes at b_off[sb]) and bscale are per-call scratch; tile_off[sb] = exclusive prefix (device int32) of the
 * ceil(nsp / 128) * ceil(nsp / bn) output tiles per superblock, ntiles their total (the GEMM is one persistent
 * CTA per SM walking that list).  bn = N tile: 64, or 96 with nslice = 5 (tcgen05.mma re-reads both operands from
 * shared memory per instruction: the wider tile with one slice less moves 42 % fewer bytes per unit of work);
 * bplanes then holds nslice * sbp * ceil(nsp / bn) * bn bytes per superblock, zero-filled once by the caller.
 * colmax (sum_sb nsp * (sbp / 32) * ncomp floats, ncomp = 1 or 4 components of `ao`; optional) receives, per superblock
 * column and per group of 32 grid rows, an upper bound of max_g |ao_c[g][col]|; passing it to b200qc_vxc_sb_i8 selects
 * the FUSED operand preparation: vb = w (v phi + 2 g . grad phi) is cut into the int8 planes in the pass that forms it,
 * with block exponents from the bound max_groups [max|w v| colmax_0 + sum_d max|2 w g_d| colmax_d] instead of the exact
 * column maxima (vb never goes to HBM in fp64; vb / vb_off may then be NULL; every factor 2 the bound is loose costs one
 * of the 7 * nslice mantissa bits -- the slicer records the exact column maxima while it cuts, and 64-column blocks
 * whose bound came out more than 2^4 too large are flagged in fixflag (nsb * (max_nsp / 64) ints of scratch, needed with
 * colmax) and cut a second time with their exact exponents, so the planes never lose more than 4 bits against the
 * two-pass form).  colmax = NULL: two passes with exact maxima through the fp64 scratch vb. */
int b200qc_vxc_i8_prepare(const void *sbdesc, int nsb, int sbp, int max_nsp, int nslice, int ncomp, const double *ao,
                          const int64_t *a_off, signed char *aplanes, double *ascale, float *colmax, void *stream);
int b200qc_vxc_sb_i8(const void *sbdesc, int nsb, int sbp, int max_nsp, int nslice, const int *idx, const double *ao,
                     const double *weights, const double *vrho, const double *vgrad, int nao, const int64_t *vb_off,
                     double *vb, const float *colmax, int *fixflag, const signed char *aplanes, const int64_t *a_off,
                     const double *ascale, signed char *bplanes, const int64_t *b_off, double *bscale, int bn,
                     const int *tile_off, int ntiles, const int *ptile_off, int nptiles, double *mat, void *stream);
/* Scheduling switches of the tcgen05 kernels (bit mask; default 0): 1 = L2 evict_last hint on the re-used A planes
 * of K2; 2 = K4 in 2-CTA thread-block clusters, every A stage fetched half by each CTA and multicast to both (the
 * pair works on two N tiles of one M tile; needs ptile_off / nptiles = exclusive prefix and total of
 * ceil(nsp / 128) * ceil(nsp / 128) pair units per superblock, else NULL / 0); 4 = the same for K2 (the pair splits
 * the N tiles of one 128-row block; partial row sums are added atomically); 16 = K2 keeps the first K steps of a
 * unit's A tile in shared memory behind a 4-stage ring; bits 8..11 = depth of the K2 operand ring (2..8, 0 = 5).
 * All of these measured slower than or equal to the default on a B200. */
int b200qc_i8_mode(int flags);

/* timing experiments on the kernel above: 0 = normal, 1 = skip the epilogue, 2 = skip the MMAs */
int b200qc_i8_debug_variant(int v);

/* K2 on tcgen05: the same contraction as b200qc_rho_sb with X = phi D as an error-free sliced int8 GEMM
 * (both operands K-major, int32 accumulators in TMEM) and the row dots with phi / grad phi fused into the
 * epilogue.  prepare slices the static AO values row-wise once: aplanes = sum_sb nslice * sbp * nsp bytes at
 * a_off[sb], rscale = nsb * sbp doubles.  bplanes (sum_sb nslice * nsp * ceil(nsp / bn) * bn bytes at b_off[sb],
 * zero-filled once by the caller) and cscale (sum_sb nsp doubles) are per-call scratch for the gathered, sliced
 * density.  bn = row tile of the sliced density.  bn = 128 (default): point-stationary kernel -- M = 128 density rows,
 * N = 64 grid points whose phi planes (prepare with row_tile = 64) stay in shared memory for the whole unit, so every
 * plane is read from HBM once; the fp64 AO values of the epilogue are read coalesced and the row sums are formed by a
 * fixed-order warp reduce-scatter (bitwise reproducible).  bn = 64, or 96 with nslice = 5: the round-1 kernel (prepare
 * with row_tile = 128; the 128-row phi tile is streamed once per N tile). */
int b200qc_rho_i8_prepare(const void *sbdesc, int nsb, int sbp, int nslice, int row_tile, const double *ao,
                          const int64_t *a_off, signed char *aplanes, double *rscale, void *stream);
int b200qc_rho_sb_i8(const void *sbdesc, int nsb, int sbp, int max_nsp, int nslice, const int *idx, const double *ao,
                     const double *dm, int nao, const signed char *aplanes, const int64_t *a_off,
                     const double *rscale, signed char *bplanes, const int64_t *b_off, double *cscale, int bn,
                     double *rho, double *grad, void *stream);

/* ---- one- and two-electron integrals (Rys quadrature) ---------------------------------- */
/* kind: 0 int1e_ovlp, 1 int1e_kin, 2 int1e_nuc, 3 int1e_rinv (origin rinv_orig[3], host).
 * Replaces GTOint2c (molintor.py:624-644).  out: (nao_i, nao_j) row-major for shells
 * [ish0, ish1) x [jsh0, jsh1) -- i.e. already in the orientation the reference has after swapaxes. */
int b200qc_int1e(const b200qc_basis *basis, int kind, const int *h_shls_slice /*4*/,
                 const double *h_rinv_orig, double *out, void *stream);
/* (P|Q): replaces GTOint2c with int2c2e_sph (molintor.py:36-54). out (nP, nQ). */
int b200qc_int2c2e(const b200qc_basis *basis, const int *h_shls_slice /*4*/, double *out, void *stream);
/* (ij|P): replaces GTOnr3c_drv + GTOnr3c_fill_s1 with int3c2e_sph (molintor.py:646-665).
 * out (nao_i, nao_j, naux) row-major. */
int b200qc_int3c2e(const b200qc_basis *basis, const int *h_shls_slice /*6*/, double *out, void *stream);
/* (ij|kl) dense: replaces GTOnr2e_fill_drv + fills4 (molintor.py:667-688, symmetry.py:55-64).
 * out (ni, nj, nk, nl) row-major.  Meant for small systems / tests. */
int b200qc_int2e(const b200qc_basis *basis, const int *h_shls_slice /*8*/, double *out, void *stream);
/* Same integrals stored once per AO pair i >= j: out[(i(i+1)/2 + j) * ld + P] (ld >= naux; the
 * caller zero-fills the padding).  This is the resident DF tensor of K9: half the bytes of the
 * reference's (nao, nao, naux) j3c (dqc/df/dfmol.py:33-39).  i and j slices must be identical. */
int b200qc_int3c2e_packed(const b200qc_basis *basis, const int *h_shls_slice /*6*/, double *out, int64_t ld,
                          void *stream);

/* Stored-ERI regime for small molecules (2 nao^4 doubles fit HBM): eri_j[i][j][k][l] =
 * eri_k[i][k][j][l] = (ij|kl), filled from the quartets with i >= j, k >= l.  The reference stores
 * el_mat the same way (hcgto.py:129); J and K per iteration are then two HBM-bound GEMVs
 * (b200qc_gemv) instead of its einsums (hcgto.py:209,234). */
int b200qc_eri_store(const b200qc_basis *basis, int sh0, int sh1, double *eri_j, double *eri_k, void *stream);
/* y[r] = sum_c A[r * ld + c] x[c]; ld even, A and x 16-byte aligned */
int b200qc_gemv(const double *A, int64_t nrow, int64_t ncol, int64_t ld, const double *x, double *y, void *stream);

/* ---- K8: direct J/K -- replaces the dense-ERI einsums of hcgto.py:204-241 ---------------- */
/* without ever storing (ij|kl) (the reference holds nao^4 doubles, hcgto.py:129).
 * dm: (nset, nao, nao) SYMMETRIC, AO basis of shells [sh0, sh1); vj / vk: (nset, nao, nao) or NULL;
 *   vj[kl] = sum_ij dm[ij] (ij|kl),  vk[jk] = sum_il dm[il] (ij|kl)   (no -1/2 factor).
 * A plan = Schwarz-sorted shell-pair lists + work-item tables, built once per geometry
 * (plays the role of the `int2e_optimizer` pair cache, molintor.py:695-708).  Quartets with
 * Q_ij Q_kl < thresh are skipped (the reference skips nothing: use thresh <= 1e-12 for parity).
 * rank / world: this process digests work items rank, rank + world, ... and returns PARTIAL
 * matrices (multi-GPU: all-reduce them); rank = 0, world = 1 for the complete result. */
typedef struct b200qc_jkplan b200qc_jkplan;
int b200qc_jkplan_create(const b200qc_basis *basis, int sh0, int sh1, double thresh, b200qc_jkplan **out,
                         void *stream);
int64_t b200qc_jkplan_nquartets(const b200qc_jkplan *plan);
/* of those, the quartets of classes with l <= 1 on every shell: they run on the register-resident engine (one lane =
 * one contracted quartet, a warp = one bra pair x 128 kets; csrc/jk_reg.cuh), the others on the shared-memory engine */
int64_t b200qc_jkplan_nquartets_reg(const b200qc_jkplan *plan);
/* fp64 operations one build needs (Rys roots, 2-D recurrence tables and the sum over roots once per primitive quartet,
 * plus 2 flops per integral and tile contraction: 2 contractions for J, 4 for K): roofline numerator of the J/K kernels */
double b200qc_jkplan_flops(const b200qc_jkplan *plan, int with_j, int with_k);
int b200qc_jkplan_run(const b200qc_jkplan *plan, const double *dm, int nset, double *vj, double *vk, int rank,
                      int world, void *stream);
int b200qc_jkplan_free(b200qc_jkplan *plan);
/* one-shot convenience: plan (thresh 1e-13) + run + free */
int b200qc_jk_direct(const b200qc_basis *basis, int sh0, int sh1, const double *dm, int nset,
                     double *vj, double *vk, void *stream);

/* ---- K9: density-fitted J -- replaces dfmol.py:60-79 ----------------------------------- */
/* j3c: (npair, ld) packed AO pairs (b200qc_int3c2e_packed), ld even, padding zero;
 * dm: (nao, nao) AO basis (need not be symmetric); inv_j2c: (naux, naux);
 *   temp_P = sum_ij dm_ij (ij|P);  c = temp . inv_j2c;  vj_ij = sum_P (ij|P) c_P   (nao, nao).
 * work: b200qc_dfj_worksize(nao, ld) doubles.  pass1 / pass2 are the two halves for the
 * aux-sharded multi-GPU layout (each rank holds a column slice of j3c; temp is all-gathered,
 * the partial vj all-reduced); coef must be zero beyond naux up to ld. */
int64_t b200qc_dfj_worksize(int64_t nao, int64_t ld);
/* pair rows whose every |(ij|P)| (this rank's columns) is below thresh: mask[npair] = 1.  The masked variants of the
 * two passes do not read those rows (temp gets no contribution, J_ij = 0): -22 % of the HBM traffic at C60/def2-SVP,
 * -35 % on the 113-atom system with thresh = 1e-14; mask = NULL reads everything. */
int b200qc_dfj_rowmask(const double *j3c_packed, int64_t nao, int64_t naux, int64_t ld, double thresh,
                       unsigned char *mask, void *stream);
/* pass 1 over a list of pair rows (ascending indices into the npair rows, device; NULL = every row): with the rows the
 * mask keeps, the kernel is the plain streaming loop on fewer rows (the next index is fetched one step ahead) */
int b200qc_dfj_pass1_rows(const double *j3c_packed, int64_t nao, int64_t naux, int64_t ld, const double *dm,
                          double *temp, double *work, const int *rows, int64_t nrows, void *stream);
int b200qc_dfj_pass2_masked(const double *j3c_packed, int64_t nao, int64_t naux, int64_t ld, const double *coef,
                            double *vj, const unsigned char *mask, void *stream);
int b200qc_dfj(const double *j3c_packed, int64_t nao, int64_t naux, int64_t ld, const double *inv_j2c,
               const double *dm, double *vj, double *work, void *stream);
int b200qc_dfj_pass1(const double *j3c_packed, int64_t nao, int64_t naux, int64_t ld, const double *dm,
                     double *temp, double *work, void *stream);
int b200qc_dfj_pass2(const double *j3c_packed, int64_t nao, int64_t naux, int64_t ld, const double *coef,
                     double *vj, void *stream);
/* pack (nao, nao, naux) -> (npair, ld) */
int b200qc_pack_tril(const double *full, int64_t nao, int64_t naux, int64_t ld, double *packed, void *stream);

/* ---- K10 / K11: fp64-accurate GEMM on tcgen05 (density-fitted exact exchange) -------------------- */
/* The reference has no DF-K (hcgto.py:229-230 raises); SURVEY 8a defines the extension
 * K^DF_ij = sum_PQ (ik|P) (P|Q)^-1 (Q|jl) D_kl, built here from two batched GEMMs
 *     C[b][m][n] (=, +=) alpha * sum_k A[b][m][k] B[b][n][k]
 * done as error-free sliced int8 products (Ozaki scheme; nslice = 5 or 6 slices of 7 bits,
 * tcgen05.mma.kind::i8, int32 accumulators in TMEM, exact int64 recombination, power-of-two row scales).
 *
 * b200qc_i8_slice: fp64 operand -> int8 planes in the tiled UMMA order (bytes)
 *   [batch][row tile = r / tile_rows][k tile = k / 32][slice][(k % 32) / 16][(r % tile_rows) / 8][r % 8][k % 16]
 * tile_rows = 128 for an A operand, 64 for a B operand; planes = nbatch * Rpad * Kpad * nslice bytes,
 * scales = nbatch * Rpad doubles (one power of two per row and batch).  Element (b, r, k) is read from
 * src[b * sb + r * sr + k * sk]; rows >= R and columns >= K (K_last in the last batch) are zero.  pair_mode != 0
 * reads the packed (ij|P) layout instead: b = i, k = j, r = P at src[tri(i, j) * pair_ld + r].
 *
 * b200qc_gemm_i8: one persistent CTA per SM walks the (batch, M tile, N tile) list.  a_bstride / b_bstride are
 * the BYTES between the planes of consecutive batches (0 = operand shared by every batch), as_bstride /
 * bs_bstride the scale entries between batches; nk = K steps of 32 per batch (nk_last in the last batch; the
 * planes of every batch still span nk steps).  mode 0 stores C[b * c_bstride + m * ldc + n]; mode 1 adds
 * atomically (split-K: batches = K chunks, c_bstride = 0); mode 2 = mode 1 restricted to the tiles that touch
 * the lower triangle (symmetric results: the caller mirrors).  rowmax (optional, mode 0): see below. */
int b200qc_i8_slice(const double *src, int nbatch, int64_t sb, int64_t sr, int64_t sk, int pair_mode,
                    int64_t pair_ld, int R, int K, int K_last, int Rpad, int Kpad, int tile_rows, int nslice,
                    signed char *planes, double *scales, void *stream);
int b200qc_gemm_i8(const signed char *aplanes, const double *ascale, int64_t a_bstride, int64_t as_bstride,
                   const signed char *bplanes, const double *bscale, int64_t b_bstride, int64_t bs_bstride,
                   int nbatch, int mtiles, int ntiles, int nk, int nk_last, int nslice, int M, int N, double alpha,
                   double *C, int64_t c_bstride, int64_t ldc, int mode, double *rowmax, int64_t rm_bstride, int rm_div,
                   void *stream);
/* One-pass slicing of a K-contiguous operand into BOTH forms (128- and 64-row tiles) when the row maxima are
 * already known: element (b, r, k) at src[b * sb + r * sr + k], max_k |.| of (b, r) at rowmax[r * rm_ld + b]
 * (what b200qc_gemm_i8 leaves in its optional `rowmax` output: the bit patterns of non-negative doubles,
 * [batch * rm_bstride + row / rm_div], zero-initialised by the caller).  Padded sizes: rows to 128 (A form) and 64
 * (B form), K to Kpad. */
int b200qc_i8_slice_dual(const double *src, int nbatch, int64_t sb, int64_t sr, const double *rowmax, int64_t rm_ld,
                         int R, int K, int K_last, int Kpad, int nslice, signed char *planesA, double *scalesA,
                         signed char *planesB, double *scalesB, void *stream);

#ifdef __cplusplus
}
#endif
#endif
