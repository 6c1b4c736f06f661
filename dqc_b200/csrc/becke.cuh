// Becke partition weights -- replaces the per-atom torch loop of
// dqc/grid/multiatoms_grid.py:173-273 (O(natom^2 ngrid) on the host in the reference).
// One thread per grid point; the point's distances to all atoms sit in a shared-memory column
// (stride = blockDim, conflict-free); pair data (R_ij, a_ij) are warp-uniform global reads.
#pragma once
#include "common.cuh"

#define BECKE_THREADS 128

__global__ void becke_pairs_kernel(const double *__restrict__ pos, int nat, double *__restrict__ rinv) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nat * nat) return;
    int i = idx / nat, j = idx % nat;
    // the reference adds the identity to the displacement of the diagonal (:187-190) -> |(1,1,1)|
    double e = (i == j) ? 1.0 : 0.0;
    double dx = pos[3 * j] - pos[3 * i] + e, dy = pos[3 * j + 1] - pos[3 * i + 1] + e,
           dz = pos[3 * j + 2] - pos[3 * i + 2] + e;
    rinv[idx] = sqrt(dx * dx + dy * dy + dz * dz);  // R_ij itself: the weight kernel divides, like the reference
}

__global__ void __launch_bounds__(BECKE_THREADS)
becke_weights_kernel(const double *__restrict__ xyz, const int *__restrict__ owner, int64_t ngrid,
                     const double *__restrict__ pos, int nat, const double *__restrict__ rinv,
                     const double *__restrict__ aij, double *__restrict__ w) {
    extern __shared__ double rg[];  // [nat][blockDim.x]
    const int tid = threadIdx.x;
    const int BT = blockDim.x;       // 128, or fewer points per CTA when many atoms must fit in shared memory
    const int64_t g = (int64_t)blockIdx.x * BT + tid;
    if (g >= ngrid) return;  // no block-wide barrier below
    const double x = xyz[3 * g], y = xyz[3 * g + 1], z = xyz[3 * g + 2];
    for (int k = 0; k < nat; k++) {
        const double dx = x - pos[3 * k], dy = y - pos[3 * k + 1], dz = z - pos[3 * k + 2];
        rg[k * BT + tid] = sqrt(dx * dx + dy * dy + dz * dz);
    }
    const int own = owner[g];
    const double sdiag = 0.5 * (1.0 + 1e-12) + 0.5;  // the i == j factor, kept for fidelity (:259)
    double psum = 0.0, pown = 0.0;
    for (int j = 0; j < nat; j++) {
        const double rj = rg[j * BT + tid];
        double P = sdiag;
        bool keep = true;
        for (int i = 0; i < nat; i++) {
            if (i == j) continue;
            double mu = (rj - rg[i * BT + tid]) / rinv[i * nat + j];
            if (aij) mu += aij[i * nat + j] * (1.0 - mu * mu);
            if (!(mu < 0.74)) {
                keep = false;
                break;
            }
            double f = mu;
#pragma unroll
            for (int it = 0; it < 3; it++) f = 0.5 * f * (3.0 - f * f);
            P *= 0.5 * (1.0 + 1e-12 - f);
        }
        if (keep) {
            psum += P;
            if (j == own) pown = P;
        }
    }
    w[g] = pown / psum;
}

extern "C" int b200qc_becke_weights(const double *xyz, const int *owner, int64_t ngrid,
                                    const double *atompos, int natom, const double *aij, double *w,
                                    void *stream) {
    QC_REQUIRE(natom >= 1, "no atoms");
    if (ngrid == 0) return 0;
    cudaStream_t st = as_stream(stream);
    double *rinv = nullptr;
    QC_CHECK(cudaMallocAsync(&rinv, sizeof(double) * natom * natom, st));
    becke_pairs_kernel<<<(natom * natom + 255) / 256, 256, 0, st>>>(atompos, natom, rinv);
    QC_LAUNCHED(1);
    int bt = BECKE_THREADS;          // points per CTA: halve until the distance column fits (<= 800 atoms)
    while (bt > 32 && sizeof(double) * natom * bt > 200 * 1024) bt >>= 1;
    const size_t smem = sizeof(double) * natom * bt;
    QC_REQUIRE(smem <= 200 * 1024, "too many atoms for the shared-memory distance column (> 800)");
    QC_CHECK(cudaFuncSetAttribute(becke_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nblk = (int)((ngrid + bt - 1) / bt);
    prof_begin(PROF_BECKE, st);
    becke_weights_kernel<<<nblk, bt, smem, st>>>(xyz, owner, ngrid, atompos, natom, rinv, aij, w);
    prof_end(st);
    QC_LAUNCHED(1);
    QC_CHECK(cudaFreeAsync(rinv, st));
    return 0;
}

// ---- molecular grid assembly: radial x Lebedev products of every atom, pruned, translated to the nuclei ----
// Replaces the host-side torch construction of dqc/grid/lebedev_grid.py:33-60 (+ the per-slice products of
// TruncatedLebedevGrid, :63-102) and the per-atom translation / concatenation of multiatoms_grid.py:158-171.
// Per atom TYPE t the host passes only the 1-D radial rule and, per radial node, which Lebedev table it carries:
//   node n in [type_node_off[t], type_node_off[t + 1]):  r = node_r[n], radial weight node_dv[n] (4 pi r^2 dr/dx w),
//   angular rule = rows [node_ang_off[n], + node_nang[n]) of ang (sin theta, cos theta, sin phi, cos phi, w),
//   node_pt_off[n] = index of its first point inside the atomic grid of the type.
// One thread per molecular grid point: atom by binary search over atom_pt_off, node by binary search over the type's
// node_pt_off.  Point = r (sin theta cos phi, sin theta sin phi, cos theta) + R_atom, evaluated with the same
// operation order as the reference's torch code (no fused multiply-adds), so the grid is bit-identical to the host's.
__global__ void grid_assemble_kernel(int natom, const double *__restrict__ atompos, const int *__restrict__ atom_type,
                                     const int64_t *__restrict__ atom_pt_off, const int *__restrict__ type_node_off,
                                     const double *__restrict__ node_r, const double *__restrict__ node_dv,
                                     const int *__restrict__ node_ang_off, const int *__restrict__ node_pt_off,
                                     const double *__restrict__ ang, int64_t ngrid, double *__restrict__ xyz,
                                     double *__restrict__ dvol, int *__restrict__ owner) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngrid) return;
    int lo = 0, hi = natom;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (atom_pt_off[mid] <= g) lo = mid; else hi = mid;
    }
    const int a = lo, t = atom_type[a];
    const int p = (int)(g - atom_pt_off[a]);
    int n0 = type_node_off[t], n1 = type_node_off[t + 1];
    while (n1 - n0 > 1) {
        const int mid = (n0 + n1) >> 1;
        if (node_pt_off[mid] <= p) n0 = mid; else n1 = mid;
    }
    const int k = p - node_pt_off[n0];
    const double *row = ang + (int64_t)(node_ang_off[n0] + k) * 5;
    const double r = node_r[n0];
    const double rs = __dmul_rn(r, row[0]);
    xyz[3 * g] = __dadd_rn(__dmul_rn(rs, row[3]), atompos[3 * a]);
    xyz[3 * g + 1] = __dadd_rn(__dmul_rn(rs, row[2]), atompos[3 * a + 1]);
    xyz[3 * g + 2] = __dadd_rn(__dmul_rn(r, row[1]), atompos[3 * a + 2]);
    dvol[g] = __dmul_rn(node_dv[n0], row[4]);
    owner[g] = a;
}

extern "C" int b200qc_grid_assemble(int natom, const double *atompos, const int *atom_type, const int64_t *atom_pt_off,
                                    const int *type_node_off, const double *node_r, const double *node_dv,
                                    const int *node_ang_off, const int *node_pt_off, const double *ang, int64_t ngrid,
                                    double *xyz, double *dvol, int *owner, void *stream) {
    QC_REQUIRE(natom >= 1 && xyz && dvol && owner, "bad arguments");
    if (ngrid == 0) return 0;
    cudaStream_t st = as_stream(stream);
    grid_assemble_kernel<<<(unsigned)((ngrid + 255) / 256), 256, 0, st>>>(natom, atompos, atom_type, atom_pt_off, type_node_off,
                                                                         node_r, node_dv, node_ang_off, node_pt_off, ang, ngrid,
                                                                         xyz, dvol, owner);
    QC_LAUNCHED(1);
    return 0;
}
