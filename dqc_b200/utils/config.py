"""Run-time knobs.  Mirrors the reference's module-level ``config`` singleton
(dqc/utils/config.py:6-14: THRESHOLD_MEMORY, CHUNK_MEMORY, VERBOSE) and adds the GPU knobs."""
from dataclasses import dataclass

__all__ = ["config"]


@dataclass
class _Config:
    # kept for drop-in compatibility; CHUNK_MEMORY only sizes the CPU restatement's chunks,
    # the CUDA path tiles the grid itself
    THRESHOLD_MEMORY: int = 10 * 1024 ** 3
    CHUNK_MEMORY: int = 16 * 1024 ** 2
    VERBOSE: int = 0
    # grid points per block handled by one CTA pass in the XC kernels
    GRID_BLOCK: int = 128
    # integral screening threshold on the primitive-pair prefactor (0 => none, like the reference)
    INT_SCREEN: float = 0.0


config = _Config()
