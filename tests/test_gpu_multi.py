"""Two-rank NCCL check of the public API (needs >= 2 GPUs; skipped otherwise): every
integrator gives BIT-identical results at 1 and 2 ranks, ensembles shard by chain index."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
ROOTDIR = sys.path[0]
import numpy as np, torch, torch.distributed as dist
import hepmc_b200 as H
ws = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
H.set_outputs("torch")
out = {}
H.seed(11)
s = H.PlainMC(3)(H.densities.Camel(3), 300001, keep_samples=False)
out["plain"] = [s.integral, s.integral_err]
mc = H.VegasMC(4, divisions=9)
s = mc(H.densities.Camel(4), 200003, 4, keep_samples=False)
out["vegas"] = [s.integral, s.integral_err] + np.stack(mc.volumes.bounds).ravel().tolist()
s = H.ImportanceMC(H.densities.Rambo(4, 100.))(lambda *p: p[0] * p[4] / 1e3 + p[3] ** 2 / 1e3, 150001, keep_samples=False)
out["rambo_is"] = [s.integral, s.integral_err]
ch = H.MultiChannel([H.densities.Gaussian(2, mu=1 / 3, scale=0.1), H.densities.Gaussian(2, mu=2 / 3, scale=0.1), H.densities.Uniform(2)])
s = H.MultiChannelMC(ch)(H.densities.Camel(2), [30001] * 2, [50001] * 2, [], keep_samples=False)
out["multichannel"] = [s.integral, s.integral_err] + ch.channels_weight.tolist()
# ensembles: chains sharded by global index, no collective on the path
C = 1000
first, count = H.parallel.item_span(C)
met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
x0 = np.tile([0., 10.], (count, 1))
ms = met.sample(50, x0, store_chain=False, first_chain=first)
acc = ms.total_accepted.clone()
H.parallel.sum_int(acc)
out["rwm_total_accepted"] = int(acc.item())
out["rwm_last_sum"] = float(ms.data.sum().item()) if ws == 1 else None
# surrogate-gradient HMC: the cooperative / thread-per-chain layout follows the GLOBAL ensemble size, so the
# chains are bit-identical however they are sharded (600 chains: one call at 1 rank, 2 x 300 at 2 ranks)
e = dict(np.load(os.path.join(ROOTDIR, "tests", "golden", "elm.npz")))
tgt = H.densities.Camel(8)
H.surrogate.attach_surrogate_gradient(tgt, H.surrogate.GaussianBasis(8), (e["g100_centers"], e["g100_widths"], None),
                                      float(e["g100_bias"]), e["g100_weights"])
C = 600
first, count = H.parallel.item_span(C)
x0 = (1 / 3 + 0.05 * np.random.default_rng(3).standard_normal((C, 8)))[first:first + count]
H.seed(21)
hs = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8, cov=10.), 10, 0.1).sample(8, x0, store_chain=False, first_chain=first)
last = hs.data[:, 0].contiguous()
if ws > 1:
    parts = [torch.empty_like(last) for _ in range(ws)]
    dist.all_gather(parts, last)
    last = torch.cat(parts)
out["hmc_elm_last"] = last.cpu().numpy().ravel().tolist()
ex = H.parallel.SlotExchange._cache
out["exchange"] = "p2p" if any(v is not None for v in ex.values()) else "nccl"
if rank == 0:
    print("RESULT " + json.dumps(out))
if ws > 1:
    dist.destroy_process_group()
'''


def _run(nproc, exchange="auto"):
    code = WORKER % ROOT
    path = os.path.join(ROOT, "gpurun_out", "_multi_worker.py")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(code)
    if nproc == 1:
        cmd = [sys.executable, path]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
               "--master-addr", "127.0.0.1", "--master-port", "29533", path]
    env = dict(os.environ, HEPMC_B200_EXCHANGE=exchange)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_results_are_bit_identical_at_one_and_two_ranks():
    one, two, two_nccl = _run(1), _run(2), _run(2, "nccl")
    for key in ("plain", "vegas", "rambo_is", "multichannel", "hmc_elm_last"):
        assert one[key] == two[key], key
        assert one[key] == two_nccl[key], key
    assert one["rwm_total_accepted"] == two["rwm_total_accepted"]
    # the peer-memory exchange (NVLink stores + flags inside the kernels) and NCCL's all-gather are two transports
    # of the same packed slot rows
    assert two_nccl["exchange"] == "nccl"
    print("slot exchange at 2 ranks:", two["exchange"])
