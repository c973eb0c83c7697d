"""world_size-2 test of the N>1 host logic on CPU (gloo): slot partition + all-gather +
ordered merge give the same (count, mean, M2 | hist) as a single process, bit for bit.
Chunk partial rows are synthesised from the oracle (the kernels are exercised on the GPU by
tests/test_gpu_parity.py::test_vegas_gpu_count_invariance_at_abi_level)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def chan_merge(a, b):
    if b[0] == 0:
        return a
    if a[0] == 0:
        return b
    n = a[0] + b[0]
    delta = b[1] - a[1]
    fb = b[0] / n
    return np.array([n, a[1] + delta * fb, a[2] + b[2] + delta * delta * a[0] * fb])


def chunk_rows(n_total, chunk, ncols):
    """Deterministic synthetic partial rows, one per chunk (function of the chunk index only)."""
    nc = (n_total + chunk - 1) // chunk
    rs = np.random.RandomState(5)
    vals = rs.lognormal(size=n_total)
    rows = np.zeros((nc, 3 + ncols))
    for c in range(nc):
        v = vals[c * chunk:(c + 1) * chunk]
        rows[c, 0], rows[c, 1], rows[c, 2] = v.size, v.mean(), ((v - v.mean()) ** 2).sum()
        rows[c, 3:] = rs.rand(ncols) * v.sum()
    return rows


def merge_range(rows, b, e):
    st = np.zeros(3)
    hist = np.zeros(rows.shape[1] - 3)
    for r in range(b, e):
        st = chan_merge(st, rows[r, :3])
        hist = hist + rows[r, 3:]
    return np.concatenate([st, hist])


def final_merge(slots):
    st = np.zeros(3)
    hist = np.zeros(slots.shape[1] - 3)
    for s in range(slots.shape[0]):
        st = chan_merge(st, slots[s, :3])
        hist = hist + slots[s, 3:]
    return np.concatenate([st, hist])


def _worker(rank, world, port, n_total, chunk, ncols, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hepmc_b200 import parallel
    rows = chunk_rows(n_total, chunk, ncols)
    first, n_local, begins = parallel.local_span(n_total, chunk)
    c0 = first // chunk
    local = np.stack([merge_range(rows, c0 + begins[i], c0 + begins[i + 1]) for i in range(len(begins) - 1)])
    slots = parallel.gather_slots(torch.from_numpy(local))
    total = final_merge(slots.numpy())
    acc = parallel.sum_int(torch.tensor([rank + 1], dtype=torch.int64))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([total, [float(acc.item())]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total,chunk", [(100000, 2048), (5000, 2048), (2048 * 8, 2048)])
def test_two_rank_merge_is_bitwise_identical_to_one_rank(tmp_path, n_total, chunk):
    ncols = 6
    world = 2
    port = 29500 + (os.getpid() + n_total) % 2000
    mp.spawn(_worker, args=(world, port, n_total, chunk, ncols, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from hepmc_b200 import parallel
    rows = chunk_rows(n_total, chunk, ncols)
    nc = rows.shape[0]
    begins = parallel.slot_begins(nc)
    single = final_merge(np.stack([merge_range(rows, begins[s], begins[s + 1]) for s in range(8)]))
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert np.array_equal(got[:-1], single)          # bit-identical on every rank
        assert got[-1] == 3.0                            # integer all-reduce
    # and the statistics are right
    vals = np.random.RandomState(5).lognormal(size=n_total)
    assert single[0] == n_total
    assert abs(single[1] - vals.mean()) < 1e-12 * abs(vals.mean())
    assert abs(single[2] / n_total - vals.var()) < 1e-10 * vals.var()
