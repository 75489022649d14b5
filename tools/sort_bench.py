"""Onesweep (csrc/bh_sort.cu) vs cub::DeviceRadixSort::SortPairs on the same keys."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402

cub = C.CDLL(os.path.join(ROOT, "oracle", "libcub_sort_baseline.so"))
dev = torch.device("cuda:0")
out = {}
tag = os.environ.get('BH_LIB', 'default').split('/')[-1]
for n in (1_000_000, 16_000_000):
    rng = np.random.default_rng(1)
    keys = torch.from_numpy(rng.integers(0, 1 << 30, n, dtype=np.uint32).view(np.int32)).to(dev)
    vals = torch.arange(n, dtype=torch.int32, device=dev)
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    tmp = torch.empty(bh.sort_pairs_u32(None, None, None, None, n, 0, 30), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for end_bit in (30,):
        for _ in range(3):
            bh.sort_pairs_u32(keys, vals, ko, vo, n, 0, end_bit, tmp)
        # correctness of whatever library variant is loaded: stable ascending order
        want_k, want_v = torch.sort(keys.to(torch.int64) & 0xFFFFFFFF, stable=True)
        assert bool((ko.to(torch.int64) & 0xFFFFFFFF == want_k).all()) and bool((vo.to(torch.int64) == want_v).all()), "sort result wrong"
        ts = []
        for _ in range(20):
            flush.fill_(0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            bh.sort_pairs_u32(keys, vals, ko, vo, n, 0, end_bit, tmp)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        mine = float(np.median(ts))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            bh.sort_pairs_u32(keys, vals, ko, vo, n, 0, end_bit, tmp)
        b.record()
        torch.cuda.synchronize()
        mine_b2b = a.elapsed_time(b) / 20
        ms = C.c_float()
        rc = cub.cub_sort_pairs_ms(C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), C.c_void_p(ko.data_ptr()),
                                   C.c_void_p(vo.data_ptr()), C.c_longlong(n), 0, end_bit, 20, C.byref(ms))
        assert rc == 0
        gb = (4 + 16 * 4) * n / 1e9
        out[f"n={n},bits={end_bit}"] = {"onesweep_ms_l2_flushed": mine, "onesweep_ms_back_to_back": mine_b2b, "cub_ms_back_to_back": ms.value,
                                        "onesweep_GBps_alg": gb / (mine * 1e-3), "cub_GBps_alg": gb / (ms.value * 1e-3)}
out['lib'] = tag
print(json.dumps(out, indent=1))
