import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import nbody_barnes_hut_cuda_b200 as bh
cub = C.CDLL(os.path.join(os.getcwd(), "oracle", "libcub_sort_baseline.so"))
dev = torch.device("cuda:0")
n = 16_000_000
rng = np.random.default_rng(1)
keys = torch.from_numpy(rng.integers(0, 1 << 30, n, dtype=np.uint32).view(np.int32)).to(dev)
vals = torch.arange(n, dtype=torch.int32, device=dev)
ko, vo = torch.empty_like(keys), torch.empty_like(vals)
tmp = torch.empty(bh.sort_pairs_u32(None, None, None, None, n, 0, 30), dtype=torch.uint8, device=dev)
for _ in range(3):
    bh.sort_pairs_u32(keys, vals, ko, vo, n, 0, 30, tmp)
torch.cuda.synchronize()
ms = C.c_float()
cub.cub_sort_pairs_ms(C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), C.c_void_p(ko.data_ptr()), C.c_void_p(vo.data_ptr()), C.c_longlong(n), 0, 30, 3, C.byref(ms))
torch.cuda.synchronize()
