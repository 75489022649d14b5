"""Standalone onesweep sort (replaces thrust::sort_by_key, bench:262-264) vs numpy's stable argsort."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_sort(bh, keys, begin_bit=0, end_bit=32):
    import torch

    n = len(keys)
    dev = torch.device("cuda:0")
    k_in = torch.from_numpy(keys.view(np.int32)).to(dev)
    v_in = torch.arange(n, dtype=torch.int32, device=dev)
    k_out, v_out = torch.empty_like(k_in), torch.empty_like(v_in)
    tmp = torch.empty(bh.sort_pairs_u32(None, None, None, None, n, begin_bit, end_bit), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    bh.sort_pairs_u32(k_in, v_in, k_out, v_out, n, begin_bit, end_bit, tmp, stream=0)
    torch.cuda.synchronize()
    assert (k_in.cpu().numpy().view(np.uint32) == keys).all()      # input untouched
    return k_out.cpu().numpy().view(np.uint32), v_out.cpu().numpy()


@pytest.mark.parametrize("n", [1, 2, 31, 4095, 4096, 4097, 100_000, 1_000_003])
@pytest.mark.parametrize("dist", ["random30", "random32", "equal", "sorted", "reversed", "few"])
def test_sort_is_the_stable_sort(bh, n, dist):
    rng = np.random.default_rng(n)
    if dist == "random30":
        keys, bits = rng.integers(0, 1 << 30, n, dtype=np.uint32), (0, 30)
    elif dist == "random32":
        keys, bits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32), (0, 32)
    elif dist == "equal":
        keys, bits = np.full(n, 0x2AAAAAAA, np.uint32), (0, 30)
    elif dist == "sorted":
        keys, bits = np.sort(rng.integers(0, 1 << 30, n, dtype=np.uint32)), (0, 30)
    elif dist == "reversed":
        keys, bits = np.sort(rng.integers(0, 1 << 30, n, dtype=np.uint32))[::-1].copy(), (0, 30)
    else:
        keys, bits = rng.integers(0, 7, n, dtype=np.uint32) << 11, (0, 30)
    k, v = run_sort(bh, keys, *bits)
    order = np.argsort(keys, kind="stable")
    assert (k == keys[order]).all()
    assert (v == order).all()


def test_partial_bit_ranges(bh):
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << 32, 50_000, dtype=np.uint64).astype(np.uint32)
    for lo, hi in [(0, 8), (8, 24), (3, 17), (16, 32)]:
        k, v = run_sort(bh, keys, lo, hi)
        digit = (keys >> lo) & ((1 << (hi - lo)) - 1)
        order = np.argsort(digit, kind="stable")
        assert (v == order).all() and (k == keys[order]).all()
