"""Host logic of the multi-GPU layout (one process per GPU, torch.distributed).

* instances and the user rows (W_user, u_bias) are partitioned by ``user mod world`` --
  private to a rank, never communicated;
* the item side (W_item, i_bias, g_bias, and the item-indexed feedback rows) is replicated;
  every ``every`` steps each rank packs ``delta = current - snapshot`` into one contiguous
  device buffer (svdgpu_items_pack_delta), the deltas are summed with ONE all-reduce, and
  every rank applies ``snapshot + scale * sum`` (svdgpu_items_apply_delta).

Only plumbing lives here; the arithmetic is in the CUDA library.
"""
import numpy as np


def user_shard(uid, world):
    """Owner rank of a user id."""
    return np.asarray(uid) % world


def shard_rows(csr, rank, world):
    """Rows of a CSR batch owned by `rank`: exactly one user feature per row is required
    (a row with several user features would straddle shards)."""
    row_ptr, label, index, value = csr
    n = len(label)
    rp = row_ptr.astype(np.int64)
    nu = rp[2:3 * n + 1:3] - rp[1:3 * n:3]
    if not np.all(nu == 1):
        raise ValueError("user-hash sharding needs exactly one user feature per row")
    uid = index[rp[1:3 * n:3]]
    keep = np.nonzero(user_shard(uid, world) == rank)[0]
    seg_lo, seg_hi = rp[0:3 * n:3][keep], rp[3:3 * n + 1:3][keep]
    lens = seg_hi - seg_lo
    new_rp = np.zeros(3 * len(keep) + 1, np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    for j in range(3):
        new_rp[j:3 * len(keep):3] = off[:-1] + (rp[j:3 * n:3][keep] - seg_lo)
    new_rp[3 * len(keep)] = off[-1]
    take = np.concatenate([np.arange(a, b) for a, b in zip(seg_lo, seg_hi)]) if len(keep) else np.zeros(0, np.int64)
    return (new_rp.astype(np.int32), label[keep], index[take], value[take]), keep


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def device_tensor(ptr, n, device):
    """torch view of `n` floats of device memory owned by the CUDA library."""
    import torch

    return torch.as_tensor(_DevArray(ptr, n), device=device)


def exchange_deltas(delta, dist, scale=1.0):
    """Sum a packed delta buffer (torch tensor, any device) over all ranks in place and
    return the factor the caller applies (1 = keep the whole SGD step mass, 1/world = average)."""
    dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    return scale


class ItemExchange:
    """snapshot -> [train steps] -> sync() on an api.SvdGpu trainer, with the all-reduce done by
    torch.distributed (any backend).  The production path is the library's own exchange
    (api.SvdGpu.comm_init / allreduce_items: NCCL on the trainer's launch stream, svdgpu_comm.cu);
    this class remains for backends NCCL does not cover (the gloo tests).

    pack and apply run on the trainer's launch stream, the all-reduce on torch's current stream: the
    two are ordered here by synchronising each before the other continues."""

    def __init__(self, trainer, dist, device, scale=1.0):
        self.g, self.dist, self.device, self.scale = trainer, dist, device, scale
        self.g.items_snapshot()

    def sync(self):
        import torch

        ptr, n = self.g.items_pack_delta()
        self.g.sync()  # the packed buffer is complete before the collective reads it
        s = exchange_deltas(device_tensor(ptr, n, self.device), self.dist, self.scale)
        if torch.device(self.device).type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()  # ... and reduced before it is applied
        self.g.items_apply_delta(s)
