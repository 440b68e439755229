"""Multi-GPU plumbing: one process per GPU, columns sharded contiguously, one small collective per step.

Soil columns never interact in the hot path (vertical-only discretisation; the reference's Jacobian is block diagonal
per column, MeshType.F90:509-530), so the data path needs no collective at all.  What the reference keeps per MPI rank
and a batch driver wants globally are the mass-balance sums and the convergence flags of MPPVSFMALM_Driver.F90:556-601,
845-898.  The collective itself lives behind the C ABI (mppgpu_comm_init / mppgpu_global_reduce_async /
mppgpu_global_mass_balance: one ncclAllGather of 9 doubles per rank on the library's stream, folded on the device), so a
Fortran + MPI host model reaches it exactly as this module does; what is left here is the rendezvous -- getting rank 0's
NCCL unique id to every rank -- and, for the CPU test-suite, the same fold under gloo.
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
    """(world, 9) per-rank reduction buffers -> (9,) global: sums add, maxima take the max, the SNES reason the minimum
    (the host restatement of fold_reductions_kernel, used by the gloo tests)."""
    out = torch.empty(NRED, dtype=gathered.dtype, device=gathered.device)
    out[0:4] = gathered[:, 0:4].sum(dim=0)
    out[4:8] = gathered[:, 4:8].max(dim=0).values
    out[8] = gathered[:, 8].min()
    return out


def rendezvous_unique_id(make_id, group=None):
    """Rank 0 calls `make_id()` (mpp_b200.comm_unique_id); every rank returns the same 128 bytes.  Uses whatever
    torch.distributed backend is initialised (an MPI host model would MPI_Bcast instead)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return make_id()
    box = [make_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def init_comm(p, group=None):
    """Attach the library's NCCL communicator to solver handle `p` (one per handle); returns (rank, world)."""
    import mpp_b200
    if dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    uid = rendezvous_unique_id(mpp_b200.comm_unique_id, group) if world > 1 else None
    p.comm_init(world, rank, uid)
    return rank, world


class GlobalReductions:
    """Host-side mirror of the library's global reduction for tensors that do not live in a solver handle (the gloo tests of the
    CPU suite): all-gather of the (9,) local buffer + `fold`."""

    def __init__(self, local, group=None):
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.gathered = torch.empty((self.world, NRED), dtype=local.dtype, device=local.device)
        self.result = torch.empty(NRED, dtype=local.dtype, device=local.device)

    def step(self):
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
