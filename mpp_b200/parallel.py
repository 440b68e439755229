"""Multi-GPU plumbing: one process per GPU, columns sharded contiguously, one small collective per step.

Soil columns never interact in the hot path (vertical-only discretisation; the reference's Jacobian is block diagonal
per column, MeshType.F90:509-530), so the data path needs no collective at all.  What the reference keeps per MPI rank
and a batch driver wants globally are the mass-balance sums and the convergence flags of MPPVSFMALM_Driver.F90:556-601,
845-898.  Every StepDT leaves them in a 9-double device buffer owned by the library
(mppgpu_reduction_buffer_device: 4 sums, 4 maxima, worst SNES reason); `GlobalReductions.step()` gathers the 9 doubles
of every rank with ONE NCCL all-gather on the library's stream and folds them on the device.  torch.distributed is
plumbing only: rendezvous, the communicator and the stream.
"""
import torch
import torch.distributed as dist

NRED = 9          # sums[0:4] | maxs[4:8] | worst (minimum) SNES reason [8]


def shard_range(ncol_total, rank, world):
    """Contiguous column range [c0, c1) of `rank` (keeps the cell-ordered arrays contiguous per device; no halo)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return rank * ncol_total // world, (rank + 1) * ncol_total // world


def fold(gathered):
    """(world, 9) per-rank reduction buffers -> (9,) global: sums add, maxima take the max, the SNES reason the minimum."""
    out = torch.empty(NRED, dtype=gathered.dtype, device=gathered.device)
    out[0:4] = gathered[:, 0:4].sum(dim=0)
    out[4:8] = gathered[:, 4:8].max(dim=0).values
    out[8] = gathered[:, 8].min()
    return out


def device_view(ptr, n, device):
    """torch view of `n` doubles at raw device pointer `ptr` (no copy, no ownership)."""
    class _V:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_V(), device=device)


class GlobalReductions:
    """Per-step global mass-balance / convergence reductions of one solver handle."""

    def __init__(self, local, group=None):
        """`local`: (9,) float64 tensor holding this rank's reduction buffer (the library's device buffer on a GPU,
        any CPU tensor under gloo)."""
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.gathered = torch.empty((self.world, NRED), dtype=local.dtype, device=local.device)
        self.result = torch.empty(NRED, dtype=local.dtype, device=local.device)

    def step(self):
        """One collective: all ranks end up with the folded (9,) result (asynchronous on the current stream)."""
        if self.world == 1:
            self.result.copy_(self.local)
        else:
            dist.all_gather_into_tensor(self.gathered.view(-1), self.local, group=self.group)
            self.result.copy_(fold(self.gathered))
        return self.result

    def as_dict(self):
        r = self.result.tolist()
        return {"mass_begin": r[0], "mass_end": r[1], "source_dt": r[2], "boundary_exchanged": r[3],
                "max_abs_mass_error": r[4], "max_newton_its": int(r[5]), "any_diverged": bool(r[6]),
                "max_dt_cuts": int(r[7]), "worst_reason": int(r[8])}
