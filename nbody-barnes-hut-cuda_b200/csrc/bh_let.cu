// bh_let.cu — locally-essential-tree export (BASELINE.json north_star: "beyond ~100M bodies a
// locally-essential-tree exchange"; SURVEY §8e second bullet).  The reference has nothing comparable.
//
// A rank that owns only part of the bodies cannot traverse the others' trees.  Instead every rank walks
// ITS tree once per peer against the peer's whole domain box and emits the coarsest set of point masses
// that the peer may use in place of this rank's bodies: a cell that passes the acceptance test for the box
// (bench:207-208 squared, distance taken at the box — the same test bh_force.cu applies per group, so it is
// conservative for every group inside the box) is emitted as {centre of mass, mass}; a rejected cell is
// opened; loose bodies and the bodies of rejected identical-key buckets are emitted as they are.  The
// emitted masses always add up to the rank's total mass.  The peer merges what it receives with its own
// bodies and runs the ordinary step on the union (imports carry id -1 and are dropped afterwards).
//
// A rank's domain is a Morton-key range, which is not convex: one bounding box around it can be almost as
// large as the whole system, and boxes around equal RUNS of the sorted bodies still straddle coarse cell
// boundaries and swallow the neighbours' bodies (measured: 42 % of a rank inside its neighbour's 16 run
// boxes on the thin two-disc system).  The domain is therefore cut at octree-cell boundaries of the key
// space (domain_boxes_kernel: one tight body AABB per key interval; an interval that is a whole cell is
// convex and disjoint from every other rank's range) and the test uses the distance to the NEAREST box.
// Cells that pass the test already against the AABB of all of the peer's boxes skip the per-box loop.
//
// Level-synchronous: a queue of (cell to open, peer) pairs is expanded once per tree level by a grid-wide
// kernel (at most levels + 1 launches for all peers together); outputs are appended with atomics.
#include "bh_common.cuh"

namespace {

constexpr int LT = 256;

struct LetArgs {
    const int4* cell_meta;
    const int32_t* cell_child;
    const float4* cell_com;
    const float4* kid_src;
    const uint8_t* kid_lv;
    const float4* posm;        // sorted bodies of the tree (bucket ranges index it)
    const float* boxes;        // npeers x K x 6: centre xyz, half extent xyz; half.x < 0 marks an unused box
    const float* hull;         // npeers x 6: the same for the AABB of all of the peer's boxes
    int K;                     // boxes per peer
    float4* out;               // npeers x cap points
    unsigned int* out_count;   // npeers
    long long cap;
    float theta2, soft;
    int root_w2_bits;
    unsigned int* err;
};

// squared distance from a point to the nearest of the peer's K boxes
__device__ __forceinline__ float let_dist2(const LetArgs& a, const float* boxes, const float4 cm) {
    float best = 3.0e38f;
    for (int k = 0; k < a.K; ++k) {
        const float* box = boxes + 6 * k;
        if (box[3] < 0.0f) continue;
        const float dx = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.x, box[0])), box[3]));
        const float dy = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.y, box[1])), box[4]));
        const float dz = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.z, box[2])), box[5]));
        best = fminf(best, __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
    }
    return best;
}

__device__ __forceinline__ float let_box_dist2(const float* box, const float4 cm) {
    const float dx = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.x, box[0])), box[3]));
    const float dy = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.y, box[1])), box[4]));
    const float dz = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.z, box[2])), box[5]));
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ bool let_accepts(const LetArgs& a, int peer, const float4 cm, int level) {
    const float w2 = __int_as_float(a.root_w2_bits - (level << 24));
    // the hull is no farther than any box: passing there is passing everywhere
    if (w2 < __fmul_rn(a.theta2, __fadd_rn(let_box_dist2(a.hull + 6 * peer, cm), a.soft))) return true;
    const float d2 = let_dist2(a, a.boxes + (size_t)6 * a.K * peer, cm);
    return w2 < __fmul_rn(a.theta2, __fadd_rn(d2, a.soft));
}

__device__ __forceinline__ bool let_peer_active(const LetArgs& a, const float* boxes) {
    for (int k = 0; k < a.K; ++k)
        if (boxes[6 * k + 3] >= 0.0f) return true;
    return false;
}

__device__ __forceinline__ void let_emit(const LetArgs& a, int peer, const float4 p) {
    const unsigned int i = atomicAdd(a.out_count + peer, 1u);
    if ((long long)i < a.cap) a.out[(size_t)peer * a.cap + i] = p;
    else atomicOr(a.err, BH_DERR_LET_OVERFLOW);
}

__device__ void let_emit_bucket(const LetArgs& a, int peer, int first, int count) {
    for (int i = 0; i < count; ++i) let_emit(a, peer, __ldg(a.posm + first + i));
}

// the root is the only cell tested without a parent
__global__ void let_seed_kernel(LetArgs a, int npeers, const BhDevScalars* sc, int2* queue, unsigned int* qcount) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= npeers) return;
    const float* box = a.boxes + (size_t)6 * a.K * r;
    if (!let_peer_active(a, box)) return;
    const int root = sc->root;
    if (root < 0) return;
    const float4 cm = __ldg(a.cell_com + root);
    const int4 mt = __ldg(a.cell_meta + root);
    if (let_accepts(a, r, cm, mt.z & 0xFF)) let_emit(a, r, cm);
    else if ((mt.z >> 8) & 1) let_emit_bucket(a, r, mt.x, mt.y);
    else queue[atomicAdd(qcount, 1u)] = make_int2(root, r);   // <= npeers <= BH_LET_MAX_PEERS entries: always fits
}

// open every queued cell for its peer; children that are rejected in turn go to the next level's queue
__global__ void __launch_bounds__(LT) let_level_kernel(LetArgs a, const int2* __restrict__ qin,
                                                      const unsigned int* __restrict__ qin_count, int2* __restrict__ qout,
                                                      unsigned int* qout_count, long long qcap) {
    // the producer counts past the capacity without storing (and raises BH_DERR_LET_OVERFLOW): never read beyond it
    const unsigned int nin = (unsigned int)min((long long)*qin_count, qcap);
    for (unsigned int idx = blockIdx.x * LT + threadIdx.x; idx < nin; idx += gridDim.x * LT) {
        const int2 item = qin[idx];
        const int cell = item.x, peer = item.y;
        const int4* ch = reinterpret_cast<const int4*>(a.cell_child) + 2 * (size_t)cell;
        const int4 lo = __ldg(ch), hi = __ldg(ch + 1);
        const uint2 lv = __ldg(reinterpret_cast<const uint2*>(a.kid_lv) + cell);
        const int e[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        int r = 0;   // kid_src lines are dense: the r-th existing child
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (e[q] == BH_CHILD_EMPTY) continue;
            const float4 s = __ldg(a.kid_src + (size_t)cell * 8 + r);
            ++r;
            if (e[q] < 0) { let_emit(a, peer, s); continue; }
            const unsigned info = ((q < 4 ? lv.x : lv.y) >> (8 * (q & 3))) & 0xFFu;
            if (let_accepts(a, peer, s, (int)(info & 0x7Fu))) let_emit(a, peer, s);
            else if (info & 0x80u) {
                const int4 mt = __ldg(a.cell_meta + e[q]);
                let_emit_bucket(a, peer, mt.x, mt.y);
            } else {
                const unsigned int o = atomicAdd(qout_count, 1u);
                if ((long long)o < qcap) qout[o] = make_int2(e[q], peer);
                else atomicOr(a.err, BH_DERR_LET_OVERFLOW);
            }
        }
    }
}

// Tight AABB (lo xyz, hi xyz) of the bodies whose key lies in [cuts[k], cuts[k+1]) for every k < K; the keys are
// sorted, so each interval is a contiguous body range found by binary search.  An empty interval gets lo > hi.
__device__ __forceinline__ long long lower_bound_key(const uint32_t* __restrict__ keys, long long n, uint32_t v) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

constexpr int DOMAIN_SLICES = 48;   // blocks per interval: K x 48 blocks keep all SMs busy for K >= 4

__global__ void domain_boxes_init_kernel(unsigned int* __restrict__ enc, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 6 * K) enc[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;   // identity of min / max on the ordered images
}

__global__ void __launch_bounds__(LT) domain_boxes_kernel(const uint32_t* __restrict__ keys, const float4* __restrict__ posm,
                                                         long long n, const uint32_t* __restrict__ cuts, int K,
                                                         unsigned int* __restrict__ enc, int* __restrict__ counts) {
    const int k = blockIdx.x;
    const long long b0 = lower_bound_key(keys, n, cuts[k]);
    // the last cut may be 2^30 = "past every key"
    const long long b1 = cuts[k + 1] >= (1u << BH_KEY_BITS) ? n : lower_bound_key(keys, n, cuts[k + 1]);
    if (blockIdx.y == 0 && threadIdx.x == 0) counts[k] = (int)(b1 - b0);
    const long long per = (b1 - b0 + gridDim.y - 1) / gridDim.y;
    const long long s0 = b0 + per * blockIdx.y, s1 = min(b1, s0 + per);
    if (s0 >= s1) return;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (long long i = s0 + threadIdx.x; i < s1; i += LT) {
        const float4 p = __ldg(posm + i);
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    __shared__ float s[6][LT / 32];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 3; ++a) { s[a][threadIdx.x >> 5] = lo[a]; s[3 + a][threadIdx.x >> 5] = hi[a]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[threadIdx.x][0];
        for (int w = 1; w < LT / 32; ++w) v = threadIdx.x < 3 ? fminf(v, s[threadIdx.x][w]) : fmaxf(v, s[threadIdx.x][w]);
        if (threadIdx.x < 3) atomicMin(enc + 6 * k + threadIdx.x, bh_f2ord(v));
        else atomicMax(enc + 6 * k + threadIdx.x, bh_f2ord(v));
    }
}

__global__ void domain_boxes_finish_kernel(const unsigned int* __restrict__ enc, const int* __restrict__ counts, int K,
                                           float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * K) return;
    out[i] = counts[i / 6] > 0 ? bh_ord2f(enc[i]) : ((i % 6) < 3 ? 3.0e38f : -3.0e38f);
}

}  // namespace

// enc_dev: 6*K u32 scratch
int bh_domain_boxes_launch(const uint32_t* keys, const float4* posm, long long n, const uint32_t* cuts_dev, int K,
                           float* out_dev, int* counts_dev, unsigned int* enc_dev, cudaStream_t st) {
    domain_boxes_init_kernel<<<(6 * K + 255) / 256, 256, 0, st>>>(enc_dev, K);
    domain_boxes_kernel<<<dim3(K, DOMAIN_SLICES), LT, 0, st>>>(keys, posm, n, cuts_dev, K, enc_dev, counts_dev);
    domain_boxes_finish_kernel<<<(6 * K + 255) / 256, 256, 0, st>>>(enc_dev, counts_dev, K, out_dev);
    return (int)cudaGetLastError();
}

// boxes_dev: npeers x K x 6 (centre, half extent; half.x < 0 = skip).  queue: 2 x qcap int2 + 2 counters (scratch).
int bh_let_export_launch(const int4* cell_meta, const int32_t* cell_child, const float4* cell_com, const float4* kid_src,
                         const uint8_t* kid_lv, const float4* posm, BhDevScalars* sc, const float* boxes_dev,
                         const float* hull_dev, int npeers, int K, float4* out, unsigned int* out_count, long long cap, int2* queue, unsigned int* qcounts,
                         long long qcap, float theta, float softening, float root_w, int levels, cudaStream_t st) {
    LetArgs a;
    a.cell_meta = cell_meta; a.cell_child = cell_child; a.cell_com = cell_com; a.kid_src = kid_src; a.kid_lv = kid_lv;
    a.posm = posm; a.boxes = boxes_dev; a.hull = hull_dev; a.K = K; a.out = out; a.out_count = out_count; a.cap = cap;
    a.theta2 = theta * theta; a.soft = softening;
    const float w2 = root_w * root_w;
    memcpy(&a.root_w2_bits, &w2, 4);
    a.err = (unsigned int*)((char*)sc + offsetof(BhDevScalars, err));
    BH_CUDA_TRY(cudaMemsetAsync(out_count, 0, sizeof(unsigned int) * npeers, st));
    BH_CUDA_TRY(cudaMemsetAsync(qcounts, 0, 2 * sizeof(unsigned int), st));
    let_seed_kernel<<<(npeers + 63) / 64, 64, 0, st>>>(a, npeers, sc, queue, qcounts);
    for (int level = 0; level <= levels; ++level) {
        const int in = level & 1, outq = in ^ 1;
        BH_CUDA_TRY(cudaMemsetAsync(qcounts + outq, 0, sizeof(unsigned int), st));
        let_level_kernel<<<BH_NUM_SMS_FALLBACK * 8, LT, 0, st>>>(a, queue + (size_t)in * qcap, qcounts + in,
                                                              queue + (size_t)outq * qcap, qcounts + outq, qcap);
    }
    return (int)cudaGetLastError();
}
