"""torchrun check (N >= 2 GPUs): the sliced run and its sharded host I/O reproduce the single-GPU run bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402
from nbody_barnes_hut_cuda_b200.sliced import SlicedSimulation  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n, steps = 300_001, 3
soa = bh.ic_refdisk(n, 42)
sim = SlicedSimulation(bh, soa, rank, world, local, dist)
sim.step(steps)            # overlapped all-gathers (default)
sim.finish()
torch.cuda.synchronize()
got = sim.eng.read_soa(want_acc=False)
# sharded host path: same physics starting from the same host state
host_in = [torch.from_numpy(x.copy()).pin_memory() for x in soa]
host_out = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(6)]
sim.step_host(host_in, host_out, steps)
chunk = (n + world - 1) // world
lo, hi = min(n, rank * chunk), min(n, (rank + 1) * chunk)
ok = True
with bh.BHEngine(n, device=local) as ref:
    ref.load_soa(*soa)
    ref.simulation_step(steps)
    want = ref.read_soa(want_acc=False)
for k in range(6):
    ok &= got[k].tobytes() == want[k].tobytes()
    ok &= host_out[k].numpy()[lo:hi].tobytes() == want[k][lo:hi].tobytes()
# and the plain (non-overlapped) loop gives the same bits
sim2 = SlicedSimulation(bh, soa, rank, world, local, dist)
sim2.step(steps, overlap=False)
torch.cuda.synchronize()
got2 = sim2.eng.read_soa(want_acc=False)
for k in range(6):
    ok &= got2[k].tobytes() == want[k].tobytes()
sim2.close()
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SLICED_CHECK", "PASS" if flag.item() == 1 else "FAIL", "world", world)
sim.close()
dist.destroy_process_group()
