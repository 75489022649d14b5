// bh_state.cu — streaming kernels over the body state: bounding cube, Morton keys,
// Morton reorder, kick-drift-clamp integrator, SoA import/export.
//
// All of them are HBM-streaming: 128-bit loads/stores on float4 SoA
// {x,y,z,m} / {vx,vy,vz,-}, grids sized to a multiple of the SM count with grid-stride loops.
//
//   bounds    replaces computeBoundingBoxKernel  nbody_v5_bench.cu:134-156 (one thread there)
//   keys      replaces computeMortonCodesKernel  nbody_v5_bench.cu:42-63
//   integrate replaces integrateKernel           nbody_v5_bench.cu:227-249
#include "bh_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t n, int per_thread, int num_sms_hint = BH_NUM_SMS_FALLBACK) {
    int64_t blocks = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    int64_t cap = (int64_t)num_sms_hint * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---- bounding cube ----------------------------------------------------------------------
__global__ void reset_step_scalars(BhDevScalars* sc) {
    if (threadIdx.x == 0) {
        // the reference starts its min/max at +-1e10f (bench:138); keep that so the
        // result is identical even for degenerate inputs
        sc->bbox_enc[0] = sc->bbox_enc[1] = sc->bbox_enc[2] = bh_f2ord(1e10f);
        sc->bbox_enc[3] = sc->bbox_enc[4] = sc->bbox_enc[5] = bh_f2ord(-1e10f);
    }
}

// CTA-wide min/max of per-thread partial boxes -> six ordered-integer atomics on sc->bbox_enc
__device__ __forceinline__ void block_minmax_to_enc(float lo[3], float hi[3], BhDevScalars* sc) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    __shared__ float s_lo[3][kThreads / 32], s_hi[3][kThreads / 32];
    const int w = threadIdx.x >> 5;
    if (bh_lane() == 0)
        for (int a = 0; a < 3; ++a) { s_lo[a][w] = lo[a]; s_hi[a][w] = hi[a]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float l = s_lo[threadIdx.x][0], h = s_hi[threadIdx.x][0];
        for (int k = 1; k < kThreads / 32; ++k) { l = fminf(l, s_lo[threadIdx.x][k]); h = fmaxf(h, s_hi[threadIdx.x][k]); }
        atomicMin(&sc->bbox_enc[threadIdx.x], bh_f2ord(l));
        atomicMax(&sc->bbox_enc[3 + threadIdx.x], bh_f2ord(h));
    }
}

__global__ void __launch_bounds__(kThreads) bounds_kernel(const float4* __restrict__ posm, int64_t n,
                                                         BhDevScalars* sc) {
    float lo[3] = {1e10f, 1e10f, 1e10f}, hi[3] = {-1e10f, -1e10f, -1e10f};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        float4 p = __ldg(posm + i);
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    block_minmax_to_enc(lo, hi, sc);
}

// bench:148-154: size = largest extent; cube anchored at the min corner.
__device__ __forceinline__ void cube_from_enc(const BhDevScalars* sc, float b[6]) {
    float minX = bh_ord2f(sc->bbox_enc[0]), minY = bh_ord2f(sc->bbox_enc[1]), minZ = bh_ord2f(sc->bbox_enc[2]);
    float maxX = bh_ord2f(sc->bbox_enc[3]), maxY = bh_ord2f(sc->bbox_enc[4]), maxZ = bh_ord2f(sc->bbox_enc[5]);
    float size = fmaxf(__fsub_rn(maxX, minX), fmaxf(__fsub_rn(maxY, minY), __fsub_rn(maxZ, minZ)));
    b[0] = minX; b[1] = minY; b[2] = minZ;
    b[3] = __fadd_rn(minX, size); b[4] = __fadd_rn(minY, size); b[5] = __fadd_rn(minZ, size);
}

// The min/max images are CONSUMED here: the cube is formed and the images go back to the reference's +-1e10f
// start values (bench:138), ready for the integrator of this step to accumulate the next step's box.
__global__ void finish_bounds_kernel(BhDevScalars* sc) {
    if (threadIdx.x == 0) {
        float b[6];
        cube_from_enc(sc, b);
        for (int k = 0; k < 6; ++k) sc->bounds[k] = b[k];
        sc->bbox_enc[0] = sc->bbox_enc[1] = sc->bbox_enc[2] = bh_f2ord(1e10f);
        sc->bbox_enc[3] = sc->bbox_enc[4] = sc->bbox_enc[5] = bh_f2ord(-1e10f);
    }
}

// ---- Morton keys --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // bench:42-49 bit spreading
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ uint32_t quantise(float p, float lo, float size) {
    // bench:58 — IEEE divide, multiply by 1023.0f (not 1024: SURVEY F5), truncate to u32
    // min(.,1023): a no-op for the reference's own cube (p <= min+size), it keeps the key on the grid when a
    // fixed cube (bh_set_fixed_bounds) is momentarily too small for a body
    return min(__float2uint_rz(__fmul_rn(__fdiv_rn(__fsub_rn(p, lo), size), 1023.0f)), 1023u);
}

__global__ void __launch_bounds__(kThreads) keys_kernel(const float4* __restrict__ posm, int64_t n,
                                                       const BhDevScalars* __restrict__ sc,
                                                       uint32_t* __restrict__ keys) {
    const float minX = sc->bounds[0], minY = sc->bounds[1], minZ = sc->bounds[2];
    const float size = fmaxf(__fsub_rn(sc->bounds[3], sc->bounds[0]), 1.0f);  // bench:57
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        float4 p = __ldg(posm + i);
        uint32_t x = quantise(p.x, minX, size), y = quantise(p.y, minY, size), z = quantise(p.z, minZ, size);
        keys[i] = (spread10(x) << 2) | (spread10(y) << 1) | spread10(z);  // bench:61, x most significant
    }
}

// key_bits = 60 (oracle: orc_morton_keys60): the fractional part of the SAME float t = (p-min)/size*1023 — exact,
// t and trunc(t) share their exponent range — gives ten more bits per axis; (hi << 30 | lo) refines the reference
// order without ever contradicting it.
__device__ __forceinline__ void quantise2(float p, float lo, float size, uint32_t& q, uint32_t& fr) {
    const float t = __fmul_rn(__fdiv_rn(__fsub_rn(p, lo), size), 1023.0f);
    q = min(__float2uint_rz(t), 1023u);
    const float g = __fmul_rn(__fsub_rn(t, __uint2float_rn(q)), 1024.0f);
    fr = !(g > 0.0f) ? 0u : (g >= 1023.0f ? 1023u : __float2uint_rz(g));
}

__global__ void __launch_bounds__(kThreads) keys60_kernel(const float4* __restrict__ posm, int64_t n,
                                                         const BhDevScalars* __restrict__ sc,
                                                         uint32_t* __restrict__ hi, uint32_t* __restrict__ lo) {
    const float minX = sc->bounds[0], minY = sc->bounds[1], minZ = sc->bounds[2];
    const float size = fmaxf(__fsub_rn(sc->bounds[3], sc->bounds[0]), 1.0f);  // bench:57
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const float4 p = __ldg(posm + i);
        uint32_t qx, qy, qz, fx, fy, fz;
        quantise2(p.x, minX, size, qx, fx); quantise2(p.y, minY, size, qy, fy); quantise2(p.z, minZ, size, qz, fz);
        hi[i] = (spread10(qx) << 2) | (spread10(qy) << 1) | spread10(qz);
        lo[i] = (spread10(fx) << 2) | (spread10(fy) << 1) | spread10(fz);
    }
}

__global__ void __launch_bounds__(kThreads) gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm,
                                                             uint32_t* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        dst[i] = __ldg(src + perm[i]);
}

__global__ void __launch_bounds__(kThreads) combine_keys_kernel(const uint32_t* __restrict__ hi_sorted,
                                                               const uint32_t* __restrict__ lo_unsorted,
                                                               const uint32_t* __restrict__ perm, uint64_t* __restrict__ keys64,
                                                               int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        keys64[i] = ((uint64_t)hi_sorted[i] << 30) | (uint64_t)__ldg(lo_unsorted + perm[i]);
}

// ---- Morton reorder -----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) reorder_kernel(const float4* __restrict__ posm_in,
                                                          const float4* __restrict__ vel_in,
                                                          const int32_t* __restrict__ ids_in,
                                                          const uint32_t* __restrict__ perm,
                                                          float4* __restrict__ posm_out,
                                                          float4* __restrict__ vel_out,
                                                          int32_t* __restrict__ ids_out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        uint32_t j = perm[i];
        float4 p = __ldg(posm_in + j), v = __ldg(vel_in + j);
        int32_t id = __ldg(ids_in + j);
        posm_out[i] = p; vel_out[i] = v; ids_out[i] = id;
    }
}

// the same in two parts (three-part step: positions are needed long before velocities and ids)
__global__ void __launch_bounds__(kThreads) reorder_posm_kernel(const float4* __restrict__ posm_in, const uint32_t* __restrict__ perm,
                                                               float4* __restrict__ posm_out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        posm_out[i] = __ldg(posm_in + perm[i]);
}

__global__ void __launch_bounds__(kThreads) reorder_rest_kernel(const float4* __restrict__ vel_in, const int32_t* __restrict__ ids_in,
                                                               const uint32_t* __restrict__ perm, float4* __restrict__ vel_out,
                                                               int32_t* __restrict__ ids_out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const uint32_t j = perm[i];
        vel_out[i] = __ldg(vel_in + j);
        ids_out[i] = __ldg(ids_in + j);
    }
}

// ---- kick-drift-clamp ----------------------------------------------------------------------
// bench:232-248 with the contraction nvcc applies to it:
//   v = fma(a,DT,v); s = fma(vz,vz,fma(vx,vx,vy*vy)); if s > MAX^2: v *= MAX/sqrt(s); p = fma(v,DT,p)
// Out of place: reads the Morton-sorted scratch copy the force phase used, writes the context's
// current state, so a captured CUDA graph sees the same pointers every step.
// The min/max of the NEW positions is reduced on the way out (bench:134-156 for the next step): the next
// step forms its cube from six atomics' worth of data instead of re-reading every position.
__global__ void __launch_bounds__(kThreads) integrate_kernel(const float4* __restrict__ posm_s,
                                                            const float4* __restrict__ vel_s,
                                                            const int32_t* __restrict__ ids_s,
                                                            const float4* __restrict__ acc, float4* __restrict__ posm,
                                                            float4* __restrict__ vel, int32_t* __restrict__ ids,
                                                            int64_t first, int64_t count, float dt, float max_speed,
                                                            BhDevScalars* sc) {
    const float vmax2 = __fmul_rn(max_speed, max_speed);
    float lo[3] = {1e10f, 1e10f, 1e10f}, hi[3] = {-1e10f, -1e10f, -1e10f};
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < count; t += (int64_t)gridDim.x * kThreads) {
        const int64_t i = first + t;
        float4 p = __ldg(posm_s + i), v = __ldg(vel_s + i);
        const float4 a = __ldg(acc + i);
        float x = __fmaf_rn(a.x, dt, v.x), y = __fmaf_rn(a.y, dt, v.y), z = __fmaf_rn(a.z, dt, v.z);
        float s = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        if (s > vmax2) {
            float scale = __fdiv_rn(max_speed, __fsqrt_rn(s));
            x = __fmul_rn(x, scale); y = __fmul_rn(y, scale); z = __fmul_rn(z, scale);
        }
        v.x = x; v.y = y; v.z = z;
        p.x = __fmaf_rn(x, dt, p.x); p.y = __fmaf_rn(y, dt, p.y); p.z = __fmaf_rn(z, dt, p.z);
        vel[i] = v; posm[i] = p; ids[i] = __ldg(ids_s + i);
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    block_minmax_to_enc(lo, hi, sc);
}

// ---- drop the ghosts (locally-essential-tree mode) ------------------------------------------
// Stable compaction of the bodies with id >= 0 (the rank's own bodies) out of a state that also holds
// imported point masses (id < 0), keeping the Morton order; vel.w of every kept body is set to the work
// its traversal chunk cost (acc.w).  Three launches: per-tile counts, one-block scan of the tile counts,
// scatter with warp ballots (one warp = 256 consecutive bodies, 32 at a time, so loads stay coalesced).
constexpr int kCompactTile = 2048;   // bodies per block: 8 warps x 256

__global__ void __launch_bounds__(kThreads) compact_count_kernel(const int32_t* __restrict__ ids, int64_t n,
                                                                int32_t* __restrict__ tile_count) {
    const int64_t base = (int64_t)blockIdx.x * kCompactTile;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kCompactTile / kThreads; ++k) {
        const int64_t i = base + k * kThreads + threadIdx.x;
        c += (i < n && __ldg(ids + i) >= 0) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int s[kThreads / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += s[w];
        tile_count[blockIdx.x] = t;
    }
}

// exclusive scan in place; the grand total lands in tile_count[tiles]
__global__ void __launch_bounds__(1024) compact_scan_kernel(int32_t* __restrict__ tile_count, int tiles) {
    __shared__ int s[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < tiles ? tile_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            s[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = carry + ((threadIdx.x >> 5) ? s[(threadIdx.x >> 5) - 1] : 0) + incl - v;
        if (i < tiles) tile_count[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_count[tiles] = carry;
}

__global__ void __launch_bounds__(kThreads) compact_scatter_kernel(const float4* __restrict__ posm, const float4* __restrict__ vel,
                                                                  const int32_t* __restrict__ ids, const float4* __restrict__ acc,
                                                                  int64_t n, const int32_t* __restrict__ tile_offset,
                                                                  float4* __restrict__ posm_out, float4* __restrict__ vel_out,
                                                                  int32_t* __restrict__ ids_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wbase = (int64_t)blockIdx.x * kCompactTile + warp * 256;
    int id[8];
    unsigned keep[8];
    int wcount = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        id[k] = i < n ? __ldg(ids + i) : -1;
        keep[k] = __ballot_sync(0xffffffffu, id[k] >= 0);
        wcount += __popc(keep[k]);
    }
    __shared__ int s[kThreads / 32];
    if (lane == 0) s[warp] = wcount;
    __syncthreads();
    int64_t o = tile_offset[blockIdx.x];
    for (int w = 0; w < warp; ++w) o += s[w];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        if (id[k] >= 0) {
            const int64_t d = o + __popc(keep[k] & ((1u << lane) - 1u));
            float4 v = __ldg(vel + i);
            v.w = __ldg(acc + i).w;
            posm_out[d] = __ldg(posm + i);
            vel_out[d] = v;
            ids_out[d] = id[k];
        }
        o += __popc(keep[k]);
    }
}

// ---- SoA import / export at the reference boundary (bench:32-35) --------------------------
__global__ void __launch_bounds__(kThreads) import_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                         const float* __restrict__ pz, const float* __restrict__ vx,
                                                         const float* __restrict__ vy, const float* __restrict__ vz,
                                                         const float* __restrict__ m, int64_t n,
                                                         float4* __restrict__ posm, float4* __restrict__ vel,
                                                         int32_t* __restrict__ ids) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        posm[i] = make_float4(px[i], py[i], pz[i], m[i]);
        vel[i] = make_float4(vx[i], vy[i], vz[i], 0.0f);
        ids[i] = (int32_t)i;
    }
}

// The same import in two parts, for the host-pointer step: positions first (all that bounds, keys and the radix
// sort read), masses + velocities + ids when their upload has landed.
__global__ void __launch_bounds__(kThreads) import_pos_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                                             const float* __restrict__ pz, int64_t n, float4* __restrict__ posm) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        posm[i] = make_float4(px[i], py[i], pz[i], 0.0f);
}

__global__ void __launch_bounds__(kThreads) import_mass_kernel(const float* __restrict__ m, int64_t n, float4* __restrict__ posm) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        reinterpret_cast<float*>(posm + i)[3] = m[i];
}

__global__ void __launch_bounds__(kThreads) import_vel_kernel(const float* __restrict__ vx, const float* __restrict__ vy,
                                                             const float* __restrict__ vz, int64_t n, float4* __restrict__ vel,
                                                             int32_t* __restrict__ ids) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        vel[i] = make_float4(vx[i], vy[i], vz[i], 0.0f);
        ids[i] = (int32_t)i;
    }
}

__global__ void __launch_bounds__(kThreads) export_kernel(const float4* __restrict__ posm, const float4* __restrict__ vel,
                                                         const float4* __restrict__ acc, const int32_t* __restrict__ ids,
                                                         int64_t n, float* px, float* py, float* pz, float* vx,
                                                         float* vy, float* vz, float* ax, float* ay, float* az) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int32_t o = ids[i];
        // slot i = body i only holds for ids that are a permutation of [0, n): contexts filled through
        // bh_import_state / a checkpoint may carry ghosts (-1) or global ids — those slots have no place in
        // the caller's n-element arrays (bh_direct_sample guards the same way)
        if (o < 0 || (int64_t)o >= n) continue;
        if (px || py || pz) {
            float4 p = __ldg(posm + i);
            if (px) px[o] = p.x;
            if (py) py[o] = p.y;
            if (pz) pz[o] = p.z;
        }
        if (vx || vy || vz) {
            float4 v = __ldg(vel + i);
            if (vx) vx[o] = v.x;
            if (vy) vy[o] = v.y;
            if (vz) vz[o] = v.z;
        }
        if (ax || ay || az) {
            float4 a = __ldg(acc + i);
            if (ax) ax[o] = a.x;
            if (ay) ay[o] = a.y;
            if (az) az[o] = a.z;
        }
    }
}

}  // namespace

// min/max images of all positions -> sc->bbox_enc (what the integrator leaves behind after a full step)
int bh_bounds_enc_launch(const float4* posm, int64_t n, BhDevScalars* sc, cudaStream_t st) {
    reset_step_scalars<<<1, 32, 0, st>>>(sc);
    bounds_kernel<<<grid_for(n, 4), kThreads, 0, st>>>(posm, n, sc);
    return (int)cudaGetLastError();
}

// bbox_enc -> sc->bounds (bench:148-154); consumes the images
int bh_bounds_finish_launch(BhDevScalars* sc, cudaStream_t st) {
    finish_bounds_kernel<<<1, 32, 0, st>>>(sc);
    return (int)cudaGetLastError();
}

int bh_keys_launch(const float4* posm, int64_t n, BhDevScalars* sc, uint32_t* keys, cudaStream_t st) {
    keys_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(posm, n, sc, keys);
    return (int)cudaGetLastError();
}

int bh_keys60_launch(const float4* posm, int64_t n, BhDevScalars* sc, uint32_t* hi, uint32_t* lo, cudaStream_t st) {
    keys60_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(posm, n, sc, hi, lo);
    return (int)cudaGetLastError();
}

int bh_gather_u32_launch(const uint32_t* src, const uint32_t* perm, uint32_t* dst, int64_t n, cudaStream_t st) {
    gather_u32_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(src, perm, dst, n);
    return (int)cudaGetLastError();
}

int bh_combine_keys_launch(const uint32_t* hi_sorted, const uint32_t* lo_unsorted, const uint32_t* perm, uint64_t* keys64,
                           int64_t n, cudaStream_t st) {
    combine_keys_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(hi_sorted, lo_unsorted, perm, keys64, n);
    return (int)cudaGetLastError();
}

int bh_reorder_launch(const float4* posm_in, const float4* vel_in, const int32_t* ids_in,
                      const uint32_t* perm, float4* posm_out, float4* vel_out, int32_t* ids_out,
                      int64_t n, cudaStream_t st) {
    reorder_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(posm_in, vel_in, ids_in, perm, posm_out, vel_out, ids_out, n);
    return (int)cudaGetLastError();
}

int bh_integrate_launch(const float4* posm_s, const float4* vel_s, const int32_t* ids_s, const float4* acc,
                        float4* posm, float4* vel, int32_t* ids, int64_t first_body, int64_t body_count,
                        float dt, float max_speed, BhDevScalars* sc, cudaStream_t st) {
    if (body_count <= 0) return 0;
    integrate_kernel<<<grid_for(body_count, 2), kThreads, 0, st>>>(posm_s, vel_s, ids_s, acc, posm, vel, ids, first_body,
                                                                  body_count, dt, max_speed, sc);
    return (int)cudaGetLastError();
}

int bh_import_launch(const float* px, const float* py, const float* pz, const float* vx,
                     const float* vy, const float* vz, const float* m, int64_t n, float4* posm,
                     float4* vel, int32_t* ids, cudaStream_t st) {
    import_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(px, py, pz, vx, vy, vz, m, n, posm, vel, ids);
    return (int)cudaGetLastError();
}

int bh_import_pos_launch(const float* px, const float* py, const float* pz, int64_t n, float4* posm, cudaStream_t st) {
    import_pos_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(px, py, pz, n, posm);
    return (int)cudaGetLastError();
}

int bh_import_mass_launch(const float* m, int64_t n, float4* posm, cudaStream_t st) {
    import_mass_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(m, n, posm);
    return (int)cudaGetLastError();
}

int bh_import_vel_launch(const float* vx, const float* vy, const float* vz, int64_t n, float4* vel, int32_t* ids, cudaStream_t st) {
    import_vel_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(vx, vy, vz, n, vel, ids);
    return (int)cudaGetLastError();
}

int bh_reorder_posm_launch(const float4* posm_in, const uint32_t* perm, float4* posm_out, int64_t n, cudaStream_t st) {
    reorder_posm_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(posm_in, perm, posm_out, n);
    return (int)cudaGetLastError();
}

int bh_reorder_rest_launch(const float4* vel_in, const int32_t* ids_in, const uint32_t* perm, float4* vel_out, int32_t* ids_out,
                           int64_t n, cudaStream_t st) {
    reorder_rest_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(vel_in, ids_in, perm, vel_out, ids_out, n);
    return (int)cudaGetLastError();
}

int bh_export_launch(const float4* posm, const float4* vel, const float4* acc, const int32_t* ids,
                     int64_t n, float* px, float* py, float* pz, float* vx, float* vy, float* vz,
                     float* ax, float* ay, float* az, cudaStream_t st) {
    export_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(posm, vel, acc, ids, n, px, py, pz, vx, vy, vz, ax, ay, az);
    return (int)cudaGetLastError();
}

// Body-order export as a GATHER (bh_step_host): where[id] = slot of body id in the state arrays, then each output array
// is written coalesced from one random 16-byte read per body.  The scatter of export_kernel costs six partial-sector
// writes per body; here it is one (the slot index).  ids must be a permutation of [0, n) (no ghosts).
__global__ void __launch_bounds__(kThreads) where_kernel(const int32_t* __restrict__ ids, int64_t n, int32_t* __restrict__ where) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int32_t o = ids[i];
        if (o >= 0 && (int64_t)o < n) where[o] = (int32_t)i;
    }
}
__global__ void __launch_bounds__(kThreads) gather3_kernel(const float4* __restrict__ src, const int32_t* __restrict__ where,
                                                          int64_t n, float* __restrict__ ox, float* __restrict__ oy,
                                                          float* __restrict__ oz) {
    for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < n; o += (int64_t)gridDim.x * kThreads) {
        const float4 v = __ldg(src + where[o]);
        ox[o] = v.x; oy[o] = v.y; oz[o] = v.z;
    }
}
int bh_where_launch(const int32_t* ids, int64_t n, int32_t* where, cudaStream_t st) {
    where_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(ids, n, where);
    return (int)cudaGetLastError();
}
int bh_gather3_launch(const float4* src, const int32_t* where, int64_t n, float* ox, float* oy, float* oz, cudaStream_t st) {
    gather3_kernel<<<grid_for(n, 2), kThreads, 0, st>>>(src, where, n, ox, oy, oz);
    return (int)cudaGetLastError();
}

// tile_scratch: n / 2048 + 2 ints.  The number of kept bodies is left in tile_scratch[tiles] (device).
int bh_compact_real_launch(const float4* posm, const float4* vel, const int32_t* ids, const float4* acc, int64_t n,
                           int32_t* tile_scratch, float4* posm_out, float4* vel_out, int32_t* ids_out, cudaStream_t st) {
    if (n <= 0) return 0;
    const int tiles = (int)((n + kCompactTile - 1) / kCompactTile);
    compact_count_kernel<<<tiles, kThreads, 0, st>>>(ids, n, tile_scratch);
    compact_scan_kernel<<<1, 1024, 0, st>>>(tile_scratch, tiles);
    compact_scatter_kernel<<<tiles, kThreads, 0, st>>>(posm, vel, ids, acc, n, tile_scratch, posm_out, vel_out, ids_out);
    return (int)cudaGetLastError();
}
