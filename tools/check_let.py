"""torchrun check (N >= 2 GPUs): LetSimulation against the single-GPU engine on the same bodies."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402
from nbody_barnes_hut_cuda_b200.let import LetSimulation, _morton30_numpy, global_cube, split_by_keys  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n, steps = 400_000, 3
soa = bh.ic_refdisk(n, 42)
lo = np.array([soa[a].min() for a in range(3)], np.float32)
hi = np.array([soa[a].max() for a in range(3)], np.float32)
keys = _morton30_numpy(soa[0], soa[1], soa[2], global_cube(np.concatenate([lo, hi])[None, :]))
sel = np.nonzero(split_by_keys(keys, world)[rank])[0]
sim = LetSimulation(bh, [a[sel] for a in soa], sel.astype(np.int32), rank, world, local, dist,
                    capacity=n + 4096, cap_per_peer=n)
sim.step(steps)
torch.cuda.synchronize()
pos = np.zeros((n, 3), np.float32)
ids = sim.rank.ids.cpu().numpy()
pos[ids] = sim.rank.posm[:, :3].cpu().numpy()
allpos = torch.from_numpy(pos).to(f"cuda:{local}")
dist.all_reduce(allpos)                                   # every body is owned by exactly one rank
owned = torch.zeros(n, device=f"cuda:{local}")
owned[torch.from_numpy(ids.astype(np.int64)).to(f"cuda:{local}")] = 1
dist.all_reduce(owned)
if rank == 0:
    with bh.BHEngine(n, device=local) as ref:
        ref.load_soa(*soa)
        ref.simulation_step(steps)
        want = np.stack(ref.read_soa()[:3], 1).astype(np.float64)
    got = allpos.cpu().numpy().astype(np.float64)
    err = float(np.sqrt(((got - want) ** 2).sum() / (want ** 2).sum()))
    ok = err < 1e-5 and bool((owned == 1).all().item())
    print("LET_CHECK", "PASS" if ok else "FAIL", "world", world, "rel pos err", err, "stats", sim.stats)
sim.close()
dist.destroy_process_group()
