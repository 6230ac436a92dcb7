"""Multi-GPU plumbing: one process per GPU, torch.distributed for the rendezvous only.

The population is sharded by contiguous ranges of the (ordered) particle list.  Data-path exchanges
happen inside the CUDA library: the per-stage %M sums travel through NVLink peer mailboxes written by the step kernel
itself (odx_peer_export / odx_peer_attach), NCCL (odx_comm_init) carries the rarely used collectives (population sizes
after birth/death, error-norm MAX, %C allgather).  This module only distributes the opaque handles.
"""
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """contiguous, order-preserving split: first (n_total % world) ranks get one extra particle"""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_bytes(payload, src=0):
    """broadcast an opaque bytes object created on `src` (e.g. the 128-byte NCCL unique id)"""
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def all_gather_bytes(payload):
    """every rank contributes one bytes object; returns the list in rank order"""
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, payload)
    return out


def agreed(value, src=0):
    """the value rank `src` holds, on every rank.  Loop counts handed to the stepping calls (odx_run_fixed, ...) MUST go through
    this when they derive from anything rank-local (timings, sample counts): the step kernels exchange the mean-field sums with
    their peers at every stage, and a rank that runs one step more than the others waits for words that never arrive."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    box = [value if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def connect(pop, use_nccl=True, use_peers=True):
    """wire a Population into the process group: NCCL communicator + NVLink peer mailboxes"""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    if use_nccl:
        uid = broadcast_bytes(pop.unique_id() if rank == 0 else None, src=0)
        pop.comm_init(rank, world, uid)
    if use_peers:
        # every rank must take the same decision: attach only if the export / attach succeeded everywhere
        try:
            mine = pop.peer_export()
        except Exception:
            mine = None
        handles = all_gather_bytes(mine)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                pop.peer_attach(rank, world, handles)
            except Exception:
                ok = False
        if not all(all_gather_bytes(ok)):
            pop.peer_detach()       # no peer access on this box: per-stage sums go through ncclAllReduce instead
    dist.barrier()
