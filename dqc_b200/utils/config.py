"""Run-time knobs.  Mirrors the reference's module-level ``config`` singleton
(dqc/utils/config.py:6-14: THRESHOLD_MEMORY, CHUNK_MEMORY, VERBOSE) and adds the GPU knobs."""
import os
from dataclasses import dataclass

__all__ = ["config"]


@dataclass
class _Config:
    # kept for drop-in compatibility; CHUNK_MEMORY only sizes the CPU restatement's chunks,
    # the CUDA path tiles the grid itself
    THRESHOLD_MEMORY: int = 10 * 1024 ** 3
    CHUNK_MEMORY: int = 16 * 1024 ** 2
    VERBOSE: int = 0
    # grid points per superblock of the block-sparse XC path (multiple of 128)
    SB_POINTS: int = int(os.environ.get("B200QC_SB_POINTS", "512"))
    # Vxc GEMM on tcgen05 as an error-free sliced int8 product: 0 = off (fp64 DMMA), 5 or 6 = number of slices
    VXC_I8_SLICES: int = int(os.environ.get("B200QC_VXC_I8", "5"))
    # N tile of that GEMM: 96 (only with 5 slices) or 64
    VXC_I8_BN: int = int(os.environ.get("B200QC_VXC_I8_BN", "96"))
    # density GEMM (K2) on tcgen05 the same way: 0 = off (fp64 DMMA), 5 or 6 slices
    RHO_I8_SLICES: int = int(os.environ.get("B200QC_RHO_I8", "5"))
    # form of that GEMM: 128 = point-stationary kernel (rows of the sliced density on the TMEM lanes, 64 grid points
    # per tile whose phi planes stay in shared memory); 64, or 96 (only with 5 slices) = the round-1 kernel with the
    # 128-row phi tile streamed once per N tile
    RHO_I8_BN: int = int(os.environ.get("B200QC_RHO_I8_BN", "128"))
    # K4 operand preparation: 1 = vb = w (v phi + 2 g . grad phi) is cut into the int8 planes in the pass that forms
    # it, block exponents from a column-maximum bound that is verified while cutting (loose blocks are cut again);
    # 0 = fp64 vb written to HBM, exact maxima, second slicing pass
    VXC_FUSED_VB: bool = os.environ.get("B200QC_VXC_FUSED_VB", "1") != "0"
    I8_VARIANT: int = int(os.environ.get("B200QC_I8_VARIANT", "0"))
    # experimental scheduling of the tcgen05 XC kernels (bit mask, default 0; all measured slower or equal, DESIGN.md
    # section 7): 1 = L2 evict_last hint on the K2 A planes, 2 = K4 in 2-CTA clusters with multicast A stages
    # (64-wide tiles only), 4 = the same for the round-1 K2, 16 = round-1 K2 with the first K steps of the A tile cached
    # in shared memory, bits 8..11 = depth of the round-1 K2 operand ring (2..8 stages, 0 = the default 5);
    # point-stationary K2: bits 12..16 = B cache slots, bits 17..18 = cluster size of the multicast D_sb stream
    # (0 = default, 1 = no clusters, 2, 3 = clusters of four)
    I8_MODE: int = int(os.environ.get("B200QC_I8_MODE", "0"))
    # density-fitted exact exchange (two batched GEMMs on tcgen05): 5 or 6 int8 slices
    DFK_I8_SLICES: int = int(os.environ.get("B200QC_DFK_I8", "6"))
    # the reference raises for exact exchange with density fitting (hcgto.py:229-230); set to 0 for that behaviour
    DF_EXCHANGE: bool = os.environ.get("B200QC_DF_EXCHANGE", "1") != "0"
    # without density fitting, keep both dense layouts of (ij|kl) in HBM when 2 * 8 * nao^4 bytes fit under
    # this (single GPU); beyond it J/K are built directly from Schwarz-screened quartets every iteration
    ERI_STORE_MAX_BYTES: int = int(float(os.environ.get("B200QC_ERI_STORE_MAX_BYTES", str(32 * 1024 ** 3))))
    # SCF fixed-point iteration: the convergence scalar of iteration k is read on the host while iteration k + lag is
    # being enqueued (0 = synchronous test every iteration)
    SCF_CHECK_LAG: int = int(os.environ.get("B200QC_SCF_CHECK_LAG", "1"))
    # sharded builds: run DF-J (whose fitting-coefficient exchange sits in the middle of it) on a second stream beside
    # the XC kernels
    DFJ_SIDE_STREAM: bool = os.environ.get("B200QC_DFJ_SIDE_STREAM", "1") != "0"
    # the same on a single GPU (no exchange to hide there: the HBM-bound DF-J passes then overlap the tensor-bound Vxc GEMM)
    DFJ_SIDE_STREAM_SINGLE: bool = os.environ.get("B200QC_DFJ_SIDE_STREAM_SINGLE", "0") != "0"
    # DF-J: pair rows of (ij|P) whose largest element is below this are skipped by both passes (0 reads every row)
    DFJ_ROW_SKIP: float = float(os.environ.get("B200QC_DFJ_ROW_SKIP", "1e-14"))
    # weight of the linear term of the superblock cost model nsp^2 + c nsp used to deal superblocks to ranks
    SB_COST_LINEAR: float = float(os.environ.get("B200QC_SB_COST_LINEAR", "900"))
    # a shell is dropped from a superblock when its envelope stays below this on every point of it
    # (0 keeps every shell everywhere, like the reference's non0tab = 1)
    AO_SCREEN: float = float(os.environ.get("B200QC_AO_SCREEN", "1e-12"))


config = _Config()
