"""One process per GPU: partition of the work and the (tiny) exchange step.

Points / chains are identified by GLOBAL indices (Philox counters), so sharding is a
choice of index ranges.  Reductions use a FIXED set of N_SLOTS = 8 slots over the chunk
range of a call; each rank owns whole slots, slot partials are all-gathered (a few KB
over NVLink via NCCL; gloo in the CPU tests) and merged in slot order on every rank.
The reduction tree is therefore a function of the problem only and results are
bit-identical at 1, 2, 4 or 8 GPUs.
"""
import torch
import torch.distributed as dist

from ._lib import N_SLOTS


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def slot_span(rank=None, world_size=None):
    """Slots [s0, s1) owned by this rank."""
    if rank is None:
        rank, world_size = world()
    if N_SLOTS % world_size:
        raise RuntimeError("world size %d must divide the %d reduction slots" % (world_size, N_SLOTS))
    per = N_SLOTS // world_size
    return rank * per, (rank + 1) * per


def slot_begins(n_chunks):
    """Chunk index where each slot starts (N_SLOTS + 1 entries)."""
    return [(n_chunks * s) // N_SLOTS for s in range(N_SLOTS + 1)]


def local_span(n_total, chunk, rank=None, world_size=None):
    """This rank's part of a call over n_total items cut into chunks of `chunk`:
    returns (first_index, n_local, local_slot_begins) where local_slot_begins are the
    chunk offsets (relative to the rank's first chunk) of its slots."""
    n_chunks = (n_total + chunk - 1) // chunk
    begins = slot_begins(n_chunks)
    s0, s1 = slot_span(rank, world_size)
    c0, c1 = begins[s0], begins[s1]
    first = c0 * chunk
    n_local = max(0, min(n_total, c1 * chunk) - first)
    return first, n_local, [b - c0 for b in begins[s0:s1 + 1]]


def item_span(n_total, rank=None, world_size=None):
    """Contiguous split of n_total independent items (chains, RAMBO points): (first, count)."""
    if rank is None:
        rank, world_size = world()
    first = (n_total * rank) // world_size
    last = (n_total * (rank + 1)) // world_size
    return first, last - first


def gather_slots(local_rows, out=None):
    """(n_local_slots, P) on every rank -> (N_SLOTS, P), slot order, on every rank (into `out` if given)."""
    rank, ws = world()
    if ws == 1:
        return local_rows
    if out is None:
        out = torch.empty((N_SLOTS, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, local_rows.contiguous())
    return out


def sum_int(value_tensor):
    """All-reduce (sum) of integer counters; order-independent, hence deterministic."""
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(value_tensor, op=dist.ReduceOp.SUM)
    return value_tensor


# ---------------------------------------------------------------------------------------------
# Peer-memory slot exchange (multi-GPU VEGAS): the all-gather of the 8 packed slot rows done by the
# reduction kernel itself with plain stores into every peer's buffer over NVLink, and awaited by the
# grid-update kernel through per-slot flags (C ABI: hepmc_reduce_slots_p2p, hepmc_vegas_update_grid_p2p).
# PyTorch only provides the plumbing: a symmetric allocation mapped into all ranks of the node.
# HEPMC_B200_EXCHANGE = "p2p" (require it), "nccl" (never use it), "auto" (default: p2p if the
# symmetric allocation works, else NCCL's all-gather).
# ---------------------------------------------------------------------------------------------
class SlotExchange(object):
    _cache = {}

    def __init__(self, ncols):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        rank, ws = world()
        nbytes = int(_lib.load().hepmc_slot_exchange_bytes(int(ncols)))
        self.ncols = int(ncols)
        self.buffer = symm_mem.empty(nbytes // 8, dtype=torch.float64, device=_lib.device())
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, dist.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != ws or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric memory rendezvous returned %d peer pointers for %d ranks" % (len(ptrs), ws))
        self.peer_ptrs = _lib.i64_array(ptrs)
        self.timeout = torch.zeros(1, dtype=torch.int32, device=_lib.device())
        self.epoch = 0
        torch.cuda.synchronize()
        dist.barrier()                       # every buffer is zeroed before anyone stores into it

    def next_epoch(self):
        self.epoch += 1
        return self.epoch

    def check(self):
        """After the call's final synchronisation: did a grid update give up waiting for a peer?"""
        t = int(self.timeout.item())
        if t:
            self.timeout.zero_()
            raise RuntimeError("hepmc_b200: the peer-memory slot exchange timed out waiting for slot %d "
                               "(a rank of the job did not reach the same VEGAS iteration)" % (t - 1))

    @classmethod
    def get(cls, ncols):
        """The exchange for rows of `ncols` histogram columns, or None (-> NCCL all-gather)."""
        import os
        import sys
        mode = os.environ.get("HEPMC_B200_EXCHANGE", "auto").lower()
        rank, ws = world()
        if ws == 1 or mode == "nccl" or N_SLOTS % ws:
            return None
        if ncols not in cls._cache:
            try:
                cls._cache[ncols] = cls(ncols)
            except Exception as exc:        # no symmetric memory on this system / backend
                if mode == "p2p":
                    raise
                if rank == 0:
                    sys.stderr.write("hepmc_b200: peer-memory slot exchange unavailable (%s: %s); using the NCCL "
                                     "all-gather\n" % (type(exc).__name__, exc))
                cls._cache[ncols] = None
        return cls._cache[ncols]
