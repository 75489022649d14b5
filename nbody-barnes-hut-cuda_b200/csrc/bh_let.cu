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
// Level-synchronous: a queue of (cell to open, peer) pairs is expanded once per tree level by a grid-wide
// kernel (at most BH_MAX_LEVEL + 1 launches for all peers together); outputs are appended with atomics.
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
    const float* boxes;        // npeers x 6: centre xyz, half extent xyz; half.x < 0 marks "skip this peer"
    float4* out;               // npeers x cap points
    unsigned int* out_count;   // npeers
    long long cap;
    float theta2, soft;
    int root_w2_bits;
    unsigned int* err;
};

__device__ __forceinline__ bool let_accepts(const LetArgs& a, const float* box, const float4 cm, int level) {
    const float dx = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.x, box[0])), box[3]));
    const float dy = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.y, box[1])), box[4]));
    const float dz = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.z, box[2])), box[5]));
    const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    const float w2 = __int_as_float(a.root_w2_bits - (level << 24));
    return w2 < __fmul_rn(a.theta2, __fadd_rn(d2, a.soft));
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
    const float* box = a.boxes + 6 * r;
    if (box[3] < 0.0f) return;
    const int root = sc->root;
    if (root < 0) return;
    const float4 cm = __ldg(a.cell_com + root);
    const int4 mt = __ldg(a.cell_meta + root);
    if (let_accepts(a, box, cm, mt.z & 0xFF)) let_emit(a, r, cm);
    else if ((mt.z >> 8) & 1) let_emit_bucket(a, r, mt.x, mt.y);
    else queue[atomicAdd(qcount, 1u)] = make_int2(root, r);
}

// open every queued cell for its peer; children that are rejected in turn go to the next level's queue
__global__ void __launch_bounds__(LT) let_level_kernel(LetArgs a, const int2* __restrict__ qin,
                                                      const unsigned int* __restrict__ qin_count, int2* __restrict__ qout,
                                                      unsigned int* qout_count, long long qcap) {
    const unsigned int nin = *qin_count;
    for (unsigned int idx = blockIdx.x * LT + threadIdx.x; idx < nin; idx += gridDim.x * LT) {
        const int2 item = qin[idx];
        const int cell = item.x, peer = item.y;
        const float* box = a.boxes + 6 * peer;
        const int4* ch = reinterpret_cast<const int4*>(a.cell_child) + 2 * (size_t)cell;
        const int4 lo = __ldg(ch), hi = __ldg(ch + 1);
        const uint2 lv = __ldg(reinterpret_cast<const uint2*>(a.kid_lv) + cell);
        const int e[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (e[q] == BH_CHILD_EMPTY) continue;
            const float4 s = __ldg(a.kid_src + (size_t)cell * 8 + q);
            if (e[q] < 0) { let_emit(a, peer, s); continue; }
            const unsigned info = ((q < 4 ? lv.x : lv.y) >> (8 * (q & 3))) & 0xFFu;
            if (let_accepts(a, box, s, (int)(info & 0x7Fu))) let_emit(a, peer, s);
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

}  // namespace

// boxes_dev: npeers x 6 (centre, half extent; half.x < 0 = skip).  queue: 2 x qcap int2 + 2 counters (scratch).
int bh_let_export_launch(const int4* cell_meta, const int32_t* cell_child, const float4* cell_com, const float4* kid_src,
                         const uint8_t* kid_lv, const float4* posm, BhDevScalars* sc, const float* boxes_dev, int npeers,
                         float4* out, unsigned int* out_count, long long cap, int2* queue, unsigned int* qcounts,
                         long long qcap, float theta, float softening, float root_w, cudaStream_t st) {
    LetArgs a;
    a.cell_meta = cell_meta; a.cell_child = cell_child; a.cell_com = cell_com; a.kid_src = kid_src; a.kid_lv = kid_lv;
    a.posm = posm; a.boxes = boxes_dev; a.out = out; a.out_count = out_count; a.cap = cap;
    a.theta2 = theta * theta; a.soft = softening;
    const float w2 = root_w * root_w;
    memcpy(&a.root_w2_bits, &w2, 4);
    a.err = (unsigned int*)((char*)sc + offsetof(BhDevScalars, err));
    BH_CUDA_TRY(cudaMemsetAsync(out_count, 0, sizeof(unsigned int) * npeers, st));
    BH_CUDA_TRY(cudaMemsetAsync(qcounts, 0, 2 * sizeof(unsigned int), st));
    let_seed_kernel<<<(npeers + 63) / 64, 64, 0, st>>>(a, npeers, sc, queue, qcounts);
    for (int level = 0; level <= BH_MAX_LEVEL; ++level) {
        const int in = level & 1, outq = in ^ 1;
        BH_CUDA_TRY(cudaMemsetAsync(qcounts + outq, 0, sizeof(unsigned int), st));
        let_level_kernel<<<BH_NUM_SMS_FALLBACK * 8, LT, 0, st>>>(a, queue + (size_t)in * qcap, qcounts + in,
                                                              queue + (size_t)outq * qcap, qcounts + outq, qcap);
    }
    return (int)cudaGetLastError();
}
