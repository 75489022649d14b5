// bh_diag.cu — diagnostics that sit beside the step path: double-precision direct sum for a
// sample of bodies (the accuracy reference north_star asks for), total energy, and two box
// probes (FP32 FMA issue rate, streaming copy bandwidth) used as roofline denominators.
// Force law: README.md:82-84 / nbody_v5_bench.cu:205-213,  a_i = G sum_j m_j d_ij / (d_ij^2 + soft)^(3/2).
#include "bh_common.cuh"

namespace {

constexpr int DT = 256;

__device__ __forceinline__ double block_sum(double v, double* s_buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (bh_lane() == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < DT / 32; ++w) t += s_buf[w];
    __syncthreads();
    return t;  // valid in thread 0
}

// one CTA per sampled body
__global__ void __launch_bounds__(DT) direct_kernel(const float4* __restrict__ posm, int64_t n,
                                                   const int32_t* __restrict__ slots, float soft, float G,
                                                   double* __restrict__ out) {
    __shared__ double s_buf[DT / 32];
    const float4 p = __ldg(posm + slots[blockIdx.x]);
    double fx = 0, fy = 0, fz = 0;
    for (int64_t j = threadIdx.x; j < n; j += DT) {
        const float4 q = __ldg(posm + j);
        const double dx = (double)q.x - p.x, dy = (double)q.y - p.y, dz = (double)q.z - p.z;
        const double r2 = dx * dx + dy * dy + dz * dz + (double)soft;
        const double inv = 1.0 / (r2 * sqrt(r2));
        const double f = (double)q.w * inv;
        fx += f * dx; fy += f * dy; fz += f * dz;
    }
    double sx = block_sum(fx, s_buf), sy = block_sum(fy, s_buf), sz = block_sum(fz, s_buf);
    if (threadIdx.x == 0) {
        out[3 * blockIdx.x] = (double)G * sx;
        out[3 * blockIdx.x + 1] = (double)G * sy;
        out[3 * blockIdx.x + 2] = (double)G * sz;
    }
}

// kinetic + softened pair potential; one CTA per body i, pairs j > i
__global__ void __launch_bounds__(DT) energy_kernel(const float4* __restrict__ posm, const float4* __restrict__ vel,
                                                   int64_t n, float soft, float G, double* __restrict__ ke_pe) {
    __shared__ double s_buf[DT / 32];
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const float4 p = __ldg(posm + i);
        double s = 0;
        for (int64_t j = i + 1 + threadIdx.x; j < n; j += DT) {
            const float4 q = __ldg(posm + j);
            const double dx = (double)q.x - p.x, dy = (double)q.y - p.y, dz = (double)q.z - p.z;
            s += (double)q.w / sqrt(dx * dx + dy * dy + dz * dz + (double)soft);
        }
        const double tot = block_sum(s, s_buf);
        if (threadIdx.x == 0) {
            const float4 v = __ldg(vel + i);
            atomicAdd(ke_pe + 0, 0.5 * p.w * ((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z));
            atomicAdd(ke_pe + 1, -(double)G * p.w * tot);
        }
    }
}

// ≙ updateVisualsKernel nbody_v5.cu:278-292, reading the Morton-ordered float4 state and scattering to
// the caller's original-order interleaved buffers.
__global__ void __launch_bounds__(DT) visuals_kernel(const float4* __restrict__ posm, const float4* __restrict__ vel,
                                                    const int32_t* __restrict__ ids, int64_t n, float* vbo_p, float* vbo_c) {
    for (int64_t i = (int64_t)blockIdx.x * DT + threadIdx.x; i < n; i += (int64_t)gridDim.x * DT) {
        const int32_t id = ids[i];
        if (id < 0 || (int64_t)id >= n) continue;   // ghosts / global ids have no slot in an n-vertex buffer
        const int64_t o = 3 * (int64_t)id;
        if (vbo_p) {
            const float4 p = __ldg(posm + i);
            vbo_p[o] = p.x; vbo_p[o + 1] = p.y; vbo_p[o + 2] = p.z;
        }
        if (vbo_c) {
            const float4 v = __ldg(vel + i);
            const float speed = __fsqrt_rn(__fmaf_rn(v.z, v.z, __fmaf_rn(v.x, v.x, __fmul_rn(v.y, v.y))));
            const float t = fminf(__fdiv_rn(speed, 150.0f), 1.0f);
            vbo_c[o] = __fmaf_rn(t, 0.6f, 0.4f); vbo_c[o + 1] = __fmaf_rn(t, 0.4f, 0.3f); vbo_c[o + 2] = __fmaf_rn(t, -0.7f, 1.0f);
        }
    }
}

__global__ void __launch_bounds__(DT) momentum_kernel(const float4* __restrict__ posm, const float4* __restrict__ vel,
                                                     int64_t n, double* __restrict__ out7) {
    __shared__ double s_buf[DT / 32];
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * DT + threadIdx.x; i < n; i += (int64_t)gridDim.x * DT) {
        const float4 p = __ldg(posm + i), v = __ldg(vel + i);
        const double m = p.w, px = m * v.x, py = m * v.y, pz = m * v.z;
        a[0] += m; a[1] += px; a[2] += py; a[3] += pz;
        a[4] += (double)p.y * pz - (double)p.z * py;
        a[5] += (double)p.z * px - (double)p.x * pz;
        a[6] += (double)p.x * py - (double)p.y * px;
    }
    for (int k = 0; k < 7; ++k) {
        const double t = block_sum(a[k], s_buf);
        if (threadIdx.x == 0) atomicAdd(out7 + k, t);
    }
}

// ---- probes -----------------------------------------------------------------------------
// 8 independent FMA chains per thread, register resident: measures the FP32 FMA issue rate.
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float a, float b) {
    float r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
            r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
        }
    }
    float s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
    if (s == 12345.678f) out[0] = s;  // keep the chains alive
}

// same with packed fma.rn.f32x2 (sm_100 FFMA2): two FMAs per issued instruction
__global__ void __launch_bounds__(256) fma2_probe_kernel(float* out, int iters, float a, float b) {
    unsigned long long pa, pb, r[8];
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float v = (float)(threadIdx.x + k);
        asm("mov.b64 %0, {%1, %1};" : "=l"(r[k]) : "f"(v));
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[k]) : "l"(pa), "l"(pb));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[k]));
        s += lo + hi;
    }
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_probe_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) dst[i] = src[i];
}

}  // namespace

int bh_direct_launch(const float4* posm, int64_t n, const int32_t* sample_slots, int k,
                     float softening, float G, double* acc_out, cudaStream_t st) {
    if (k <= 0) return 0;
    direct_kernel<<<k, DT, 0, st>>>(posm, n, sample_slots, softening, G, acc_out);
    return (int)cudaGetLastError();
}

int bh_energy_launch(const float4* posm, const float4* vel, int64_t n, float softening, float G,
                     double* ke_pe, cudaStream_t st) {
    BH_CUDA_TRY(cudaMemsetAsync(ke_pe, 0, 2 * sizeof(double), st));
    if (n <= 0) return 0;
    int grid = (int)(n < 148 * 16 ? n : 148 * 16);
    energy_kernel<<<grid, DT, 0, st>>>(posm, vel, n, softening, G, ke_pe);
    return (int)cudaGetLastError();
}

int bh_visuals_launch(const float4* posm, const float4* vel, const int32_t* ids, int64_t n, float* vbo_p, float* vbo_c,
                      cudaStream_t st) {
    if (n <= 0 || (!vbo_p && !vbo_c)) return 0;
    int64_t blocks = (n + DT - 1) / DT;
    if (blocks > 148 * 8) blocks = 148 * 8;
    visuals_kernel<<<(int)blocks, DT, 0, st>>>(posm, vel, ids, n, vbo_p, vbo_c);
    return (int)cudaGetLastError();
}

int bh_momentum_launch(const float4* posm, const float4* vel, int64_t n, double* out7, cudaStream_t st) {
    BH_CUDA_TRY(cudaMemsetAsync(out7, 0, 7 * sizeof(double), st));
    if (n <= 0) return 0;
    int64_t blocks = (n + DT - 1) / DT;
    if (blocks > 148 * 4) blocks = 148 * 4;
    momentum_kernel<<<(int)blocks, DT, 0, st>>>(posm, vel, n, out7);
    return (int)cudaGetLastError();
}

extern "C" int bh_probe_fp32_tflops(int device, float* tflops) {
    if (!tflops) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    BH_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    float* d_out = nullptr;
    BH_CUDA_TRY(cudaMalloc(&d_out, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_probe_kernel<<<blocks, 256>>>(d_out, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(d_out); return (int)e; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
        float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return 0;
}

extern "C" int bh_probe_fp32x2_tflops(int device, float* tflops) {
    if (!tflops) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    BH_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    float* d_out = nullptr;
    BH_CUDA_TRY(cudaMalloc(&d_out, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma2_probe_kernel<<<blocks, 256>>>(d_out, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(d_out); return (int)e; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 2 * 8 * 16 * (double)iters * 256.0 * blocks;
        float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return 0;
}

extern "C" int bh_probe_hbm_gbs(int device, float* gbs) {
    if (!gbs) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(device));
    const int64_t n4 = (int64_t)1 << 26;  // 1 GiB per buffer
    float4 *a = nullptr, *b = nullptr;
    BH_CUDA_TRY(cudaMalloc(&a, n4 * 16));
    if (cudaMalloc(&b, n4 * 16) != cudaSuccess) { cudaFree(a); return (int)cudaErrorMemoryAllocation; }
    cudaMemset(a, 0, n4 * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        copy_probe_kernel<<<148 * 16, 256>>>(a, b, n4);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(a); cudaFree(b); return (int)e; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        float g = (float)(2.0 * n4 * 16 / (ms * 1e-3) / 1e9);
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(a); cudaFree(b);
    *gbs = best;
    return 0;
}
