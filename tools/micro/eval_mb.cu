// eval_mb.cu — isolates the 32x32 interaction tile of bh_force.cu to find its ceiling on B200.
// Each warp owns 32 (or 64) bodies in registers and evaluates NT tiles of 32 sources read from shared
// memory (refilled rarely), no traversal.  Prints G interactions/s per variant.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct __align__(16) SrcPair { float4 xy, zm; };

// MODE 0: scalar, 1 body/lane.  MODE 1: packed, 1 body/lane, ILP pairs.  MODE 2: packed, 2 bodies/lane.
// MODE 3: scalar, 2 bodies/lane.
template <int MODE, int ILP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) mb_kernel(float* out, int ntiles, float soft) {
    __shared__ SrcPair s_src[WARPS][16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    {
        float* f = reinterpret_cast<float*>(&s_src[w][lane >> 1]) + (lane & 1);
        f[0] = lane * 1.1f; f[2] = lane * 0.7f; f[4] = lane * -0.3f; f[6] = 1.0f + lane;
    }
    __syncwarp();
    const float px = threadIdx.x * 0.01f, py = blockIdx.x * 0.02f, pz = 1.5f;
    const SrcPair* src = s_src[w];
    if (MODE == 0) {
        float ax = 0, ay = 0, az = 0;
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const float* f = reinterpret_cast<const float*>(src + (k >> 1)) + (k & 1);
                const float dx = f[0] - px, dy = f[2] - py, dz = f[4] - pz;
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, soft)));
                const float ri = rsq(r2);
                const float g = (f[6] * ri) * (ri * ri);
                ax = fmaf(g, dx, ax); ay = fmaf(g, dy, ay); az = fmaf(g, dz, az);
            }
        }
        if (ax + ay + az == 1.2345f) out[0] = ax;
    } else if (MODE == 1) {
        const f32x2 npx = pack2(-px, -px), npy = pack2(-py, -py), npz = pack2(-pz, -pz), soft2 = pack2(soft, soft);
        f32x2 ax = pack2(0, 0), ay = ax, az = ax;
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll 1
            for (int k = 0; k < 16; k += ILP) {
                f32x2 dx[ILP], dy[ILP], dz[ILP], r[ILP], m[ILP];
#pragma unroll
                for (int j = 0; j < ILP; ++j) {
                    const float4 xy = src[k + j].xy, zm = src[k + j].zm;
                    dx[j] = add2(pack2(xy.x, xy.y), npx); dy[j] = add2(pack2(xy.z, xy.w), npy);
                    dz[j] = add2(pack2(zm.x, zm.y), npz); m[j] = pack2(zm.z, zm.w);
                }
#pragma unroll
                for (int j = 0; j < ILP; ++j) r[j] = fma2(dx[j], dx[j], soft2);
#pragma unroll
                for (int j = 0; j < ILP; ++j) r[j] = fma2(dy[j], dy[j], r[j]);
#pragma unroll
                for (int j = 0; j < ILP; ++j) r[j] = fma2(dz[j], dz[j], r[j]);
#pragma unroll
                for (int j = 0; j < ILP; ++j) { float a, b; unpack2(r[j], a, b); r[j] = pack2(rsq(a), rsq(b)); }
#pragma unroll
                for (int j = 0; j < ILP; ++j) m[j] = mul2(m[j], r[j]);
#pragma unroll
                for (int j = 0; j < ILP; ++j) r[j] = mul2(r[j], r[j]);
#pragma unroll
                for (int j = 0; j < ILP; ++j) m[j] = mul2(m[j], r[j]);
#pragma unroll
                for (int j = 0; j < ILP; ++j) { ax = fma2(m[j], dx[j], ax); ay = fma2(m[j], dy[j], ay); az = fma2(m[j], dz[j], az); }
            }
        }
        float a, b; unpack2(ax, a, b); float c, d; unpack2(ay, c, d); float e, g; unpack2(az, e, g);
        if (a + b + c + d + e + g == 1.2345f) out[0] = a;
    } else if (MODE == 2) {
        // two bodies per lane: body A and body B share every source load
        const float qx = px + 3.f, qy = py - 2.f, qz = pz + 1.f;
        const f32x2 npx = pack2(-px, -px), npy = pack2(-py, -py), npz = pack2(-pz, -pz), soft2 = pack2(soft, soft);
        const f32x2 nqx = pack2(-qx, -qx), nqy = pack2(-qy, -qy), nqz = pack2(-qz, -qz);
        f32x2 ax = pack2(0, 0), ay = ax, az = ax, bx = ax, by = ax, bz = ax;
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll 1
            for (int k = 0; k < 16; k += ILP) {
#pragma unroll
                for (int j = 0; j < ILP; ++j) {
                    const float4 xy = src[k + j].xy, zm = src[k + j].zm;
                    const f32x2 sx = pack2(xy.x, xy.y), sy = pack2(xy.z, xy.w), sz = pack2(zm.x, zm.y), sm = pack2(zm.z, zm.w);
                    const f32x2 dxa = add2(sx, npx), dya = add2(sy, npy), dza = add2(sz, npz);
                    const f32x2 dxb = add2(sx, nqx), dyb = add2(sy, nqy), dzb = add2(sz, nqz);
                    f32x2 ra = fma2(dza, dza, fma2(dya, dya, fma2(dxa, dxa, soft2)));
                    f32x2 rb = fma2(dzb, dzb, fma2(dyb, dyb, fma2(dxb, dxb, soft2)));
                    float a0, a1, b0, b1; unpack2(ra, a0, a1); unpack2(rb, b0, b1);
                    ra = pack2(rsq(a0), rsq(a1)); rb = pack2(rsq(b0), rsq(b1));
                    const f32x2 fa = mul2(mul2(sm, ra), mul2(ra, ra)), fb = mul2(mul2(sm, rb), mul2(rb, rb));
                    ax = fma2(fa, dxa, ax); ay = fma2(fa, dya, ay); az = fma2(fa, dza, az);
                    bx = fma2(fb, dxb, bx); by = fma2(fb, dyb, by); bz = fma2(fb, dzb, bz);
                }
            }
        }
        float a, b, s = 0; unpack2(ax, a, b); s += a + b; unpack2(ay, a, b); s += a + b; unpack2(az, a, b); s += a + b;
        unpack2(bx, a, b); s += a + b; unpack2(by, a, b); s += a + b; unpack2(bz, a, b); s += a + b;
        if (s == 1.2345f) out[0] = s;
    }
}

template <int MODE, int ILP, int WARPS>
void run(const char* name, int ctas_per_sm) {
    float* d; cudaMalloc(&d, 4);
    const int ntiles = 2000, blocks = 148 * ctas_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    mb_kernel<MODE, ILP, WARPS><<<blocks, WARPS * 32>>>(d, 10, 50.f);
    cudaEventRecord(a);
    mb_kernel<MODE, ILP, WARPS><<<blocks, WARPS * 32>>>(d, ntiles, 50.f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bodies = (MODE == 2) ? 2.0 : 1.0;
    const double inter = (double)blocks * WARPS * 32 * bodies * 32.0 * ntiles;
    printf("%-34s warps/SM %2d  %8.3f ms  %7.1f G interactions/s  (%.1f TFLOP/s @20)\n", name, WARPS * ctas_per_sm, ms,
           inter / ms * 1e-6, inter * 20 / ms * 1e-9);
    cudaFree(d);
}

int main() {
    run<0, 1, 4>("scalar 1body", 4);  run<0, 1, 4>("scalar 1body", 8);
    run<1, 2, 4>("packed 1body ILP2", 4); run<1, 2, 4>("packed 1body ILP2", 8);
    run<1, 4, 4>("packed 1body ILP4", 4); run<1, 4, 4>("packed 1body ILP4", 6); run<1, 4, 4>("packed 1body ILP4", 8);
    run<1, 8, 4>("packed 1body ILP8", 4);
    run<2, 1, 4>("packed 2body ILP1", 4); run<2, 2, 4>("packed 2body ILP2", 4); run<2, 2, 4>("packed 2body ILP2", 8);
    run<2, 4, 4>("packed 2body ILP4", 4);
    return 0;
}
