"""Multi-GPU plumbing of the Fock build: one process per GPU (torch.distributed, NCCL over
NVLink), independent work units dealt to ranks and ONE all-reduce of the packed partial Fock per
SCF iteration.  The reference has no multi-device code at all; with world == 1 every
helper here is a no-op so that the single-GPU path is the same code."""
from typing import List, Tuple
import torch
import torch.distributed as dist

__all__ = ["ParallelContext", "get_context", "split_rows"]


class ParallelContext(object):
    def __init__(self, group=None):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1

    def allreduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks.  NCCL/gloo reduce in a fixed (rank-ordered ring/tree) order for a
        fixed world size, so the collective itself adds no run-to-run noise; the partial matrices it sums are
        reproducible to ~1e-13 relative, not bitwise (fp64 atomics in the Vxc scatter and the J/K digestion)."""
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_packed(self, tensors: List[torch.Tensor]) -> List[torch.Tensor]:
        """One collective for several partial results (J | K | Vxc | scalars): pack, reduce, unpack."""
        if self.world == 1:
            return tensors
        flat = torch.cat([t.reshape(-1) for t in tensors])
        self.allreduce_(flat)
        out, o = [], 0
        for t in tensors:
            out.append(flat[o:o + t.numel()].reshape(t.shape))
            o += t.numel()
        return out

    def allgather_cat(self, t: torch.Tensor, sizes: List[int]) -> torch.Tensor:
        """Concatenate the per-rank 1-D slices (sizes known to everyone, possibly uneven).  Done as
        a sum of zero-padded full-length vectors: one collective, and exact (x + 0 == x)."""
        if self.world == 1:
            return t
        off = sum(sizes[:self.rank])
        full = torch.zeros(sum(sizes), dtype=t.dtype, device=t.device)
        full[off:off + sizes[self.rank]] = t
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
        return full


_default = None


def get_context() -> ParallelContext:
    global _default
    if _default is None or (_default.world == 1 and dist.is_available() and dist.is_initialized()):
        _default = ParallelContext()
    return _default


def split_rows(n: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items for this rank; interior boundaries are multiples of `align`."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi
