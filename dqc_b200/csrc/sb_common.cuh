// Superblock descriptor and the vb = w (v phi + 2 g . grad phi) pass, shared by the fp64 (xc_sb.cuh) and the
// tcgen05 (vxc_i8.cuh, rho_i8.cuh) kernels -- the two live in different translation units.
#pragma once
#include "common.cuh"

// One digit of the sliced representation: q = round-to-nearest-even(y) as an int in [-64, 64], y <- (y - q) * 128.
// rint() / (int) compile to FRND.F64 + F2I.F64, two instructions of the 16-lane XU pipe per digit -- at C60 the slicers
// cut 2.8e9 digits per Fock build and were bound by exactly that.  The classic magic-number form stays on the fp64
// pipe: y + 1.5 * 2^52 has the rounded integer (two's complement) in its low mantissa word, valid for |y| < 2^31.
__device__ __forceinline__ int slice_digit(double &y) {
    const double magic = 6755399441055744.0;      // 1.5 * 2^52
    const double t = y + magic;
    const double q = t - magic;
    y = (y - q) * 128.0;
    return __double2loint(t);
}

struct SBDesc {
    int64_t ao_off;   // doubles: start of this SB's [ncomp][SBP][nsp] block
    int64_t d_off;    // doubles: start of this SB's gathered D (nsp x nsp) in the scratch
    int nsp;          // padded number of kept AOs (multiple of 64)
    int idx_off;      // start of this SB's AO index list (nsp entries, padding = nao)
    int shell_off;    // start of this SB's kept-shell list
    int nshell;       // number of kept shells
    int dsb_idx_off;  // idx_off of the superblock whose gathered, sliced D_sb this one uses (its own, or that of an earlier
                      // superblock with the same kept-AO list: consecutive radial shells of an atom mostly share it)
    int pad_;
};

template <int NCOMP>
__global__ void vxc_vb_sb_kernel(const SBDesc *__restrict__ sbd, const double *__restrict__ ao, int sbp,
                                 int64_t ngrid_ld, const double *__restrict__ w, const double *__restrict__ vrho,
                                 const double *__restrict__ vgrad, const int64_t *__restrict__ vb_off,
                                 double *__restrict__ vb, const double *__restrict__ vlapl = nullptr) {
    // NCOMP = 5 (meta-GGA): + 2 w vlapl lapl phi (hcgto.py:477)
    // one warp per grid row
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= ngrid_ld) return;
    const int sb = (int)(g / sbp), r = (int)(g - (int64_t)sb * sbp);
    const SBDesc d = sbd[sb];
    const int lane = threadIdx.x & 31;
    const double wg = w[g];
    const double c0 = wg * vrho[g];
    double c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (NCOMP == 5) c4 = 2.0 * wg * vlapl[g];
    if (NCOMP >= 4) {
        c1 = 2.0 * wg * vgrad[g];
        c2 = 2.0 * wg * vgrad[ngrid_ld + g];
        c3 = 2.0 * wg * vgrad[2 * ngrid_ld + g];
    }
    const int64_t ld = d.nsp, cs = (int64_t)sbp * ld;
    const double2 *p0 = reinterpret_cast<const double2 *>(ao + d.ao_off + (int64_t)r * ld);
    const double2 *p1 = reinterpret_cast<const double2 *>(ao + d.ao_off + cs + (int64_t)r * ld);
    const double2 *p2 = reinterpret_cast<const double2 *>(ao + d.ao_off + 2 * cs + (int64_t)r * ld);
    const double2 *p3 = reinterpret_cast<const double2 *>(ao + d.ao_off + 3 * cs + (int64_t)r * ld);
    const double2 *p4 = reinterpret_cast<const double2 *>(ao + d.ao_off + 4 * cs + (int64_t)r * ld);
    double2 *out = reinterpret_cast<double2 *>(vb + vb_off[sb] + (int64_t)r * ld);
    for (int c = lane; c < ld / 2; c += 32) {
        double2 v = p0[c];
        double2 o = make_double2(c0 * v.x, c0 * v.y);
        if (NCOMP >= 4) {
            v = p1[c]; o.x += c1 * v.x; o.y += c1 * v.y;
            v = p2[c]; o.x += c2 * v.x; o.y += c2 * v.y;
            v = p3[c]; o.x += c3 * v.x; o.y += c3 * v.y;
        }
        if (NCOMP == 5) {
            v = p4[c]; o.x += c4 * v.x; o.y += c4 * v.y;
        }
        out[c] = o;
    }
}

