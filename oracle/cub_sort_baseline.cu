// cub_sort_baseline.cu — BASELINE ONLY: times cub::DeviceRadixSort::SortPairs, the
// library sort thrust::sort_by_key (nbody_v5_bench.cu:262-264) dispatches to, with its temporary
// storage pre-allocated, so the hand-written onesweep in csrc/bh_sort.cu has a bar to clear.
// Never linked into libbh.so.
#include <cub/device/device_radix_sort.cuh>
#include <cstdint>

extern "C" int cub_sort_pairs_ms(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                                 uint32_t* vals_out, long long n, int begin_bit, int end_bit, int iters,
                                 float* avg_ms) {
    void* tmp = nullptr;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit);
    if (cudaMalloc(&tmp, bytes) != cudaSuccess) return 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i)
        cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit);
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i)
        cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *avg_ms = ms / iters;
    cudaFree(tmp);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return (int)e;
}
