// bh_force.cu — warp-cooperative Barnes-Hut traversal.
//
// Replaces computeForceKernel  nbody_v5_bench.cu:191-225  (one thread per body, private
// int stack[64], 76-byte AoS nodes).
//
// One warp owns a GROUP of up to 32 Morton-consecutive bodies (one body per lane) and keeps ONE stack of cells
// TO OPEN in shared memory.  A stack word is `cell id << 3 | children - 1`; the children of a cell are a DENSE
// line in HBM (bh_tree.cu): kid_src[8c + r] = what child r contributes as a source — a loose body's {x,y,z,m}
// or a child cell's {centre of mass, mass} — and kid_info[8c + r] = {the child's own stack word | bucket flag,
// the child's squared width as float bits (-1 for a body: the acceptance test needs no case distinction)}.
// A round pops as many cells as have ROUND_ITEMS (= 32 x FORCE_ITEMS) children together and hands every lane
// FORCE_ITEMS CHILDREN (not cells: the octree holds ~3 children per cell, a lane that classifies 8 slots of one
// cell wastes most of them).  The lane -> (cell, child) map needs no shuffle chain: the popped words' child
// counts are prefix-summed with three independent ballots (one per bit of the count), a bit mask of the item
// numbers where a cell starts is OR-reduced over the warp, and a lane finds its cell as a population count of
// that mask.  Each child is then classified in registers:
//   - loose body          -> source
//   - child cell accepted -> source.  Test: w_L^2 < theta^2 * (d^2 + SOFTENING), d = distance from the
//     child's centre of mass to the group's bounding box (exact AABB of the positions).  This is
//     bench:207-208 (`width / sqrtf(d2 + SOFTENING) < THETA`) squared and taken at the group's closest
//     point, so it is uniform for the warp and every body of the group would have accepted too;
//   - child cell rejected -> its stack word is pushed (identical-key buckets: their body range is appended).
// One dependent global access per tree level (the child's data lives in the parent's line); ballots place
// sources and pushes; whenever 32 sources are pending the warp evaluates a 32x32 tile: each lane keeps its own
// body in registers and reads the sources as shared memory broadcasts (LDS.128), bench:205-213's arithmetic in
// packed FP32 with rsqrt.approx.ftz.
//
// Group splitting: a 32-slot chunk of the Morton order that straddles a coarse cell boundary would
// get a huge bounding box (measured on the 1M reference disk: median list 1,088 entries, worst
// chunk 61,493 — one warp then outlives the whole grid).  The warp therefore cuts its chunk at the
// coarsest key boundary whenever  ext(A) + ext(B) < alpha * ext(A u B)  (ext = sum of the box
// edges; redux.sync min/max on order-preserving integer images of the coordinates), recursively,
// and traverses each sub-group with all 32 lanes opening cells but only the sub-group's lanes
// keeping the result.  oracle/bh_oracle.cpp:orc_make_groups is the CPU restatement of the rule.
//
// Ghosts (locally-essential-tree mode): bodies with id < 0 are point masses imported from other ranks.
// They are sources like any other body (they sit in the tree) but nobody needs THEIR acceleration, and
// being coarse far-away cells they are spatially sparse — a sub-group box around 32 of them would span
// half the system and open most of the tree.  They are therefore left out of the sub-group boxes, a
// sub-group without real bodies is skipped, and a cut that separates ghosts from real bodies is always
// taken.  Without ghosts (ids == nullptr or all ids >= 0) nothing changes.
//
// Cross-tree pass (accumulate != 0; locally-essential-tree mode): the groups are still runs of `posm`, but the
// tree — its scalars tree_sc, its cells and the bodies src_posm its buckets index — belongs to ANOTHER body
// set (the points imported from the other ranks); the result is added to acc.  That pass hands the chunks out
// in plain Morton order and leaves the heavy-chunk bookkeeping of the main pass alone.
//
// acc[i].w carries the work of body i's chunk (its interaction-list entries, all sub-groups) so that a
// driver can balance key ranges by measured work (let.py).
#include "bh_common.cuh"

namespace {

constexpr int FORCE_WARPS = 4;
constexpr int FORCE_THREADS = FORCE_WARPS * 32;
#ifndef FORCE_MIN_CTAS
#define FORCE_MIN_CTAS 4
#endif
#ifndef FORCE_HEAVY_X10
#define FORCE_HEAVY_X10 15   // chunks above 1.5x the mean list length of the previous step are handed out first (measured: 1.1x 0.824, 1.3x 0.817, 1.5x 0.819, 2.0x 0.827, 2.5x 0.829 ms on the 1M disk; 1.5x is also the best on the Plummer sphere)
#endif
#ifndef FORCE_ITEMS
#define FORCE_ITEMS 2
#endif
constexpr int ITEMS = FORCE_ITEMS;             // children classified per lane and round
constexpr int ROUND_ITEMS = 32 * ITEMS;
constexpr int STACK_CAP = 640;
// A round that pops cells with T children pushes at most T words.  Rounds of ROUND_ITEMS children run only while
// that fits under the depth-first reserve; otherwise a round takes at most 8 children (one cell, or a few small
// ones): net growth <= 7 per tree level, the classic DFS bound — no overflow by construction.
template <int LEVELS> struct StackPlan {
    static constexpr int DFS_RESERVE = 7 * LEVELS + 8;
    static constexpr int WIDE_LIMIT = STACK_CAP - ROUND_ITEMS - DFS_RESERVE;
    static_assert(WIDE_LIMIT >= 64, "stack too small for wide rounds");
};
constexpr int SRC_CAP = 32 + ROUND_ITEMS + 32;   // pending sources: < 32 left over + one round + one bucket slab of 32
static_assert((SRC_CAP & (SRC_CAP - 1)) == 0, "the pending list is a ring of whole tiles");
constexpr unsigned LOOP_GUARD = 1u << 24;

constexpr int Q_CAP = 128;       // quadrupole option: pending CELL sources (3 float4 each), a ring of whole tiles like SRC_CAP
template <bool QUAD> struct __align__(16) WarpScratchT {
    float4 src[SRC_CAP];       // pending sources, stored as PAIRS (see SrcPair)
    unsigned stack[STACK_CAP]; // cells waiting to be opened: id << 3 | children - 1
    float4 qsrc[QUAD ? Q_CAP * 3 : 1];   // {com, mass}, {Qxx, Qxy, Qxz, Qyy}, {Qyz, Qzz, -, -}
};

// r2 >= SOFTENING > 0, never denormal: the flush-to-zero form is a bare MUFU.RSQ (the default
// rsqrtf adds a denormal rescue of three instructions per interaction).
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- packed FP32 (sm_100 FFMA2/FADD2/FMUL2): one issued instruction does two lanes' worth of work.
// The FMA pipe's flop rate is the same as with scalar FFMA (bh_probe_fp32x2_tflops: 73.8 vs 72.1
// TFLOP/s) but the traversal kernel is ISSUE bound, and the packed form halves the issue slots of the
// interaction arithmetic.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Source lists live in shared memory as PAIRS: 32 bytes {x0,x1,y0,y1 | z0,z1,m0,m1} per two sources,
// so two LDS.128 broadcasts feed one packed interaction step.
struct __align__(16) SrcPair { float4 xy, zm; };
__device__ __forceinline__ void store_source(SrcPair* list, int idx, const float4 v) {
    float* f = reinterpret_cast<float*>(list + (idx >> 1)) + (idx & 1);
    f[0] = v.x; f[2] = v.y; f[4] = v.z; f[6] = v.w;
}

// 32 sources (16 pairs) against this lane's body.  bench:205-213 with dist^-3 from one rsqrt:
// per pair 3 FADD2 + 3 FFMA2 (r^2 + soft) + 2 MUFU.RSQ + 3 FMUL2 (m r^-3) + 3 FFMA2 (accumulate).
// Four pairs (8 sources) are kept in flight per step, stage by stage, so the dependent chains
// (3 FFMA2 -> MUFU -> 2 FMUL2 -> FFMA2) of one pair are covered by the other three: the in-order issue
// of a single warp no longer waits on its own latency.
struct Accum { f32x2 x, y, z; };
#ifndef EVAL_ILP_N
#define EVAL_ILP_N 4
#endif
constexpr int EVAL_ILP = EVAL_ILP_N;
#ifndef EVAL_K_UNROLL
#define EVAL_K_UNROLL 4
#endif
constexpr int EVAL_KU = EVAL_K_UNROLL;   // copies of the 8-source step per loop trip (1 = a loop of four trips per tile)
__device__ __forceinline__ void eval_tile(const SrcPair* __restrict__ src, f32x2 npx, f32x2 npy, f32x2 npz,
                                          f32x2 soft2, Accum& a) {
#pragma unroll EVAL_KU
    for (int k = 0; k < 16; k += EVAL_ILP) {
        f32x2 dx[EVAL_ILP], dy[EVAL_ILP], dz[EVAL_ILP], r[EVAL_ILP], m[EVAL_ILP];
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) {
            const float4 xy = src[k + j].xy, zm = src[k + j].zm;   // uniform address: LDS.128 broadcast
            dx[j] = add2(pack2(xy.x, xy.y), npx);
            dy[j] = add2(pack2(xy.z, xy.w), npy);
            dz[j] = add2(pack2(zm.x, zm.y), npz);
            m[j] = pack2(zm.z, zm.w);
        }
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) r[j] = fma2(dx[j], dx[j], soft2);
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) r[j] = fma2(dy[j], dy[j], r[j]);
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) r[j] = fma2(dz[j], dz[j], r[j]);
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) {
            float r0, r1;
            unpack2(r[j], r0, r1);
            r[j] = pack2(rsqrt_fast(r0), rsqrt_fast(r1));
        }
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) m[j] = mul2(m[j], r[j]);
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) r[j] = mul2(r[j], r[j]);
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) m[j] = mul2(m[j], r[j]);
#pragma unroll
        for (int j = 0; j < EVAL_ILP; ++j) {
            a.x = fma2(m[j], dx[j], a.x);
            a.y = fma2(m[j], dy[j], a.y);
            a.z = fma2(m[j], dz[j], a.z);
        }
    }
}

// 32 accepted cells with quadrupoles against this lane's body (BH_FLAG_QUADRUPOLE):
//   a += M d / R^3 - Q d / R^5 + 5/2 (d.Q.d) d / R^7,   d = c - p,  R^2 = d^2 + soft     (G is applied at the end)
// The reference evaluates the first term only (bench:205-213).  Scalar FP32: this path is the accuracy knob, the
// packed monopole tile above stays the default.
__device__ __forceinline__ void eval_qtile(const float4* __restrict__ q, float px, float py, float pz, float soft, float& ax,
                                           float& ay, float& az) {
#pragma unroll 2
    for (int k = 0; k < 32; ++k) {
        const float4 c = q[3 * k], a = q[3 * k + 1], b = q[3 * k + 2];   // uniform addresses: LDS.128 broadcasts
        const float dx = c.x - px, dy = c.y - py, dz = c.z - pz;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, soft)));
        const float rinv = rsqrt_fast(r2);
        const float i2 = rinv * rinv, i3 = rinv * i2, i5 = i3 * i2, i7 = i5 * i2;
        const float qx = fmaf(a.z, dz, fmaf(a.y, dy, a.x * dx));      // Q d: xx xy xz / xy yy yz / xz yz zz
        const float qy = fmaf(b.x, dz, fmaf(a.w, dy, a.y * dx));
        const float qz = fmaf(b.y, dz, fmaf(b.x, dy, a.z * dx));
        const float dqd = fmaf(dz, qz, fmaf(dy, qy, dx * qx));
        const float s = fmaf(2.5f * dqd, i7, c.w * i3);
        ax += fmaf(s, dx, -i5 * qx);
        ay += fmaf(s, dy, -i5 * qy);
        az += fmaf(s, dz, -i5 * qz);
    }
}

template <int LEVELS, bool QUAD>
__global__ void __launch_bounds__(FORCE_THREADS, FORCE_MIN_CTAS) force_kernel(
    const float4* __restrict__ posm, const typename BhKey<LEVELS>::type* __restrict__ keys, const int32_t* __restrict__ ids,
    int64_t first_body, int64_t body_count, const int4* __restrict__ cell_meta, const float4* __restrict__ cell_com,
    const float4* __restrict__ kid_src, const uint2* __restrict__ kid_info, float4* __restrict__ acc, BhDevScalars* sc,
    uint32_t* __restrict__ heavy_list, uint32_t* __restrict__ heavy_flag, int64_t max_chunks, float theta, float soft,
    float G, float split_alpha, const float4* __restrict__ src_posm, const BhDevScalars* __restrict__ tree_sc,
    int accumulate, const float4* __restrict__ cell_quad, const float4* __restrict__ kid_quad, const BhFusedUpdate fu) {
    __shared__ WarpScratchT<QUAD> s_warp[FORCE_WARPS];
    __shared__ unsigned s_bbox[FORCE_WARPS][6];   // fused update: min/max images of this warp's new positions

    const int lane = bh_lane();
    WarpScratchT<QUAD>& W = s_warp[threadIdx.x >> 5];
    SrcPair* const slist = reinterpret_cast<SrcPair*>(W.src);
    const float theta2 = __fmul_rn(theta, theta);
    const int root = tree_sc->root;
    // bench:208 maxX - minX of the root; a level-L cell is root_w * 2^-L wide, so its squared width is
    // root_w^2 with 2L taken off the exponent (exact: power-of-two scaling commutes with rounding)
    const float root_w = fmaxf(__fsub_rn(tree_sc->bounds[3], tree_sc->bounds[0]), 1.0f);   // the key grid's size (bench:52)
    const int root_w2_bits = __float_as_int(__fmul_rn(root_w, root_w));
    const unsigned root_word = tree_sc->root_word;
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    const int64_t ngroups = (body_count + BH_GROUP - 1) / BH_GROUP;
    const int64_t end_body = first_body + body_count;

    unsigned long long tot_cell = 0, tot_body = 0, tot_entries = 0;
    unsigned max_sp = 0;
    unsigned* const bbox = s_bbox[threadIdx.x >> 5];
    if (lane < 6) bbox[lane] = bh_f2ord(lane < 3 ? 1e10f : -1e10f);   // the reference's start values (bench:138)
    __syncwarp();

    // heavy-first scheduling: tickets [0, heavy_n) replay last step's heavy chunks, the rest walk the
    // chunks in Morton order and skip the ones already handed out
    // heavy_flag holds launch-epoch tags instead of booleans, so nothing has to be cleared between launches:
    // a chunk is on the list being replayed iff its tag equals this launch's epoch
    const unsigned epoch = sc->epoch;
    const unsigned cur = epoch & 1u, nxt = cur ^ 1u;
    const unsigned heavy_n = accumulate ? 0u : min(sc->heavy_n[cur], (unsigned)min(ngroups, max_chunks));
    const unsigned heavy_thresh = sc->heavy_thresh;
    const uint32_t* list_cur = heavy_list + (size_t)cur * max_chunks;
    const uint32_t* flag_cur = heavy_flag + (size_t)cur * max_chunks;
    uint32_t* list_nxt = heavy_list + (size_t)nxt * max_chunks;
    uint32_t* flag_nxt = heavy_flag + (size_t)nxt * max_chunks;

    for (;;) {
        unsigned g = 0;
        if (lane == 0) {
            for (;;) {
                const unsigned t = atomicAdd(&sc->group_ticket, 1u);
                if (t < heavy_n) { g = list_cur[t]; break; }
                g = t - heavy_n;
                if ((int64_t)g >= ngroups || accumulate || flag_cur[g] != epoch) break;   // tagged chunks were served from the list
            }
        }
        g = __shfl_sync(0xffffffffu, g, 0);
        if ((int64_t)g >= ngroups) break;

        const int64_t my = first_body + (int64_t)g * BH_GROUP + lane;
        const bool valid = my < end_body;
        const int nb = (int)min((int64_t)BH_GROUP, end_body - (first_body + (int64_t)g * BH_GROUP));
        const float4 me = __ldg(posm + (valid ? my : end_body - 1));
        const bool sink = valid && (ids == nullptr || __ldg(ids + my) >= 0);   // someone wants this body's acceleration

        // ---- split the chunk into spatially compact sub-groups (see the file header) ----
        const typename BhKey<LEVELS>::type mykey = __ldg(keys + (valid ? my : end_body - 1));
        const typename BhKey<LEVELS>::type nextkey = __shfl_down_sync(0xffffffffu, mykey, 1);
        const int lvp = (lane + 1 < nb) ? bh_shared_digits_t<LEVELS>(mykey, nextkey) : 99;   // pair (lane, lane+1)
        const unsigned ox = bh_f2ord(me.x), oy = bh_f2ord(me.y), oz = bh_f2ord(me.z);
        const f32x2 npx = pack2(-me.x, -me.x), npy = pack2(-me.y, -me.y), npz = pack2(-me.z, -me.z);
        const f32x2 soft2 = pack2(soft, soft);
        float ax = 0.f, ay = 0.f, az = 0.f;      // this lane's body, final
        unsigned cuts = 0;                        // bit j: boundary between lanes j and j+1
        unsigned acc_cells_w = 0, dir_bodies_w = 0;   // weighted by sub-group size
        unsigned chunk_entries = 0;
        int ga = 0;
        while (ga < nb) {
        const unsigned pending = cuts >> ga;
        int gb = pending ? ga + __ffs(pending) : nb;
        unsigned sinks = __ballot_sync(0xffffffffu, lane >= ga && lane < gb && sink);
        if (!sinks) { ga = gb; continue; }        // ghosts only: nothing to compute
        float lox, loy, loz, hix, hiy, hiz;
        {
            const bool in = lane >= ga && lane < gb && sink;
            lox = bh_ord2f(__reduce_min_sync(0xffffffffu, in ? ox : 0xFFFFFFFFu));
            loy = bh_ord2f(__reduce_min_sync(0xffffffffu, in ? oy : 0xFFFFFFFFu));
            loz = bh_ord2f(__reduce_min_sync(0xffffffffu, in ? oz : 0xFFFFFFFFu));
            hix = bh_ord2f(__reduce_max_sync(0xffffffffu, in ? ox : 0u));
            hiy = bh_ord2f(__reduce_max_sync(0xffffffffu, in ? oy : 0u));
            hiz = bh_ord2f(__reduce_max_sync(0xffffffffu, in ? oz : 0u));
        }
        while (gb - ga >= 2 && split_alpha > 0.0f) {
            const int cand = (lane >= ga && lane + 1 < gb) ? ((lvp << 5) | lane) : 0x7FFFFFFF;
            const int best = __reduce_min_sync(0xffffffffu, cand);   // fewest shared digits, first such pair
            if ((best >> 5) >= LEVELS) break;
            const int gk = (best & 31) + 1;
            const bool inA = lane >= ga && lane < gk && sink, inB = lane >= gk && lane < gb && sink;
            const unsigned sA = __ballot_sync(0xffffffffu, inA), sB = __ballot_sync(0xffffffffu, inB);
            if (!sA) { cuts |= 1u << (gk - 1); ga = gk; continue; }   // ghosts in front: drop them, same box
            if (!sB) { cuts |= 1u << (gk - 1); gb = gk; continue; }   // ghosts behind: leave them to the next round
            const float alx = bh_ord2f(__reduce_min_sync(0xffffffffu, inA ? ox : 0xFFFFFFFFu));
            const float aly = bh_ord2f(__reduce_min_sync(0xffffffffu, inA ? oy : 0xFFFFFFFFu));
            const float alz = bh_ord2f(__reduce_min_sync(0xffffffffu, inA ? oz : 0xFFFFFFFFu));
            const float ahx = bh_ord2f(__reduce_max_sync(0xffffffffu, inA ? ox : 0u));
            const float ahy = bh_ord2f(__reduce_max_sync(0xffffffffu, inA ? oy : 0u));
            const float ahz = bh_ord2f(__reduce_max_sync(0xffffffffu, inA ? oz : 0u));
            const float blx = bh_ord2f(__reduce_min_sync(0xffffffffu, inB ? ox : 0xFFFFFFFFu));
            const float bly = bh_ord2f(__reduce_min_sync(0xffffffffu, inB ? oy : 0xFFFFFFFFu));
            const float blz = bh_ord2f(__reduce_min_sync(0xffffffffu, inB ? oz : 0xFFFFFFFFu));
            const float bhx = bh_ord2f(__reduce_max_sync(0xffffffffu, inB ? ox : 0u));
            const float bhy = bh_ord2f(__reduce_max_sync(0xffffffffu, inB ? oy : 0u));
            const float bhz = bh_ord2f(__reduce_max_sync(0xffffffffu, inB ? oz : 0u));
            const float eA = __fadd_rn(__fadd_rn(__fsub_rn(ahx, alx), __fsub_rn(ahy, aly)), __fsub_rn(ahz, alz));
            const float eB = __fadd_rn(__fadd_rn(__fsub_rn(bhx, blx), __fsub_rn(bhy, bly)), __fsub_rn(bhz, blz));
            const float eAB = __fadd_rn(__fadd_rn(__fsub_rn(hix, lox), __fsub_rn(hiy, loy)), __fsub_rn(hiz, loz));
            if (!(__fadd_rn(eA, eB) < __fmul_rn(split_alpha, eAB))) break;
            cuts |= 1u << (gk - 1);
            gb = gk;
            sinks = sA;
            lox = alx; loy = aly; loz = alz; hix = ahx; hiy = ahy; hiz = ahz;
        }
        const bool in_group = lane >= ga && lane < gb;
        const int gsize = __popc(sinks);
        const float cx = __fmul_rn(__fadd_rn(lox, hix), 0.5f), hx = __fmul_rn(__fsub_rn(hix, lox), 0.5f);
        const float cy = __fmul_rn(__fadd_rn(loy, hiy), 0.5f), hy = __fmul_rn(__fsub_rn(hiy, loy), 0.5f);
        const float cz = __fmul_rn(__fadd_rn(loz, hiz), 0.5f), hz = __fmul_rn(__fsub_rn(hiz, loz), 0.5f);

        // acceptance of a cell {com, level} for this sub-group (warp-uniform inputs except the cell)
        // w2 = squared width of the cell, root_w^2 * 4^-level (a loose body carries -1: always a source)
        auto accepts = [&](const float4 cm, float w2) -> bool {
            const float dx = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.x, cx)), hx));
            const float dy = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.y, cy)), hy));
            const float dz = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(cm.z, cz)), hz));
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            return w2 < __fmul_rn(theta2, __fadd_rn(d2, soft));
        };

        Accum t;                                  // this sub-group's pass (kept only by its lanes)
        t.x = t.y = t.z = pack2(0.f, 0.f);
        int sp = 0, ns = 0;                       // stack entries, pending sources
        unsigned my_cells = 0, my_bodies = 0;     // sources this lane emitted (summed over the warp at the end)
        unsigned bucket_bodies = 0;               // warp-uniform: bodies appended from rejected buckets

        // pending sources live in a ring of SRC_CAP / 32 tiles: [head, head + ns); head stays a multiple of 32
        int head = 0;
        int qhead = 0, nq = 0;                    // quadrupole option: the same for accepted cells
        float qax = 0.f, qay = 0.f, qaz = 0.f;
        // evaluate every full tile of 32 pending sources
        auto drain = [&]() {
            while (ns >= 32) {
                eval_tile(slist + (head >> 1), npx, npy, npz, soft2, t);
                head = (head + 32) & (SRC_CAP - 1);
                ns -= 32;
            }
            if (QUAD) {
                while (nq >= 32) {
                    eval_qtile(W.qsrc + 3 * qhead, me.x, me.y, me.z, soft, qax, qay, qaz);
                    qhead = (qhead + 32) & (Q_CAP - 1);
                    nq -= 32;
                }
            }
        };
        // append the bodies [bfirst, bfirst+bcount) of a rejected bucket
        auto append_bucket = [&](int bfirst, int bcount) {
            bucket_bodies += bcount;
            for (int b = 0; b < bcount; b += 32) {
                const int m = min(32, bcount - b);
                if (lane < m) store_source(slist, (head + ns + lane) & (SRC_CAP - 1), __ldg(src_posm + bfirst + b + lane));
                ns += m;
                __syncwarp();
                drain();
            }
        };

        // ---- the root is the only cell tested without a parent ----
        if (root >= 0) {
            const float4 rcm = __ldg(cell_com + root);
            const int4 rmt = __ldg(cell_meta + root);
            if (accepts(rcm, __int_as_float(root_w2_bits - ((rmt.z & 0xFF) << 24)))) {
                if (QUAD) {
                    if (lane == 0) {
                        W.qsrc[0] = rcm; W.qsrc[1] = __ldg(cell_quad + 2 * (size_t)root); W.qsrc[2] = __ldg(cell_quad + 2 * (size_t)root + 1);
                        my_cells = 1;
                    }
                    nq = 1;
                } else {
                    if (lane == 0) { store_source(slist, 0, rcm); my_cells = 1; }   // head = 0, ns = 0
                    ns = 1;
                }
            } else if ((rmt.z >> 8) & 1) {
                append_bucket(rmt.x, rmt.y);
            } else {
                if (lane == 0) W.stack[0] = root_word;
                sp = 1;
            }
        }
        __syncwarp();

        unsigned guard = 0;
        while (sp > 0) {
            if (++guard > LOOP_GUARD) { if (lane == 0) atomicOr(&sc->err, BH_DERR_LOOP); break; }
            // ---- pop: the cells on top of the stack whose children number <= ROUND_ITEMS together ----
            const int avail = min(sp, 32);
            const bool live = lane < avail;
            const unsigned ent = live ? W.stack[sp - 1 - lane] : 0u;
            const unsigned c1 = ent & 7u;                                    // children - 1
            const unsigned live_m = avail == 32 ? 0xffffffffu : (1u << avail) - 1u;
            // exclusive prefix of the child counts: three independent ballots instead of a shuffle chain
            const unsigned b0 = __ballot_sync(0xffffffffu, c1 & 1u), b1 = __ballot_sync(0xffffffffu, c1 & 2u),
                           b2 = __ballot_sync(0xffffffffu, c1 & 4u);
            const int excl = __popc(live_m & lt_mask) + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
            const int limit = (sp <= StackPlan<LEVELS>::WIDE_LIMIT) ? ROUND_ITEMS : 8;
            const unsigned tk = __ballot_sync(0xffffffffu, live && excl + (int)c1 + 1 <= limit);   // a prefix of the lanes, never empty
            const int E = __popc(tk);
            const int T = E + __popc(b0 & tk) + 2 * __popc(b1 & tk) + 4 * __popc(b2 & tk);        // children this round
            sp -= E;
            // item k belongs to the cell whose first item number is the last start <= k
            const bool taken = lane < E;
            const unsigned base = (ent & ~7u) - (unsigned)excl;             // 8 * cell - first item number
            unsigned starts[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
                starts[j] = __reduce_or_sync(0xffffffffu, (taken && (excl >> 5) == j) ? 1u << (excl & 31) : 0u);

            // ---- open: one child per (lane, j): its source and its info word ----
            float4 s[ITEMS];
            uint2 w[ITEMS];
            bool has[ITEMS];
            unsigned item[ITEMS];
            int before = 0;
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const int k = lane + 32 * j;
                has[j] = k < T;
                const int owner = before + __popc(starts[j] & le_mask) - 1;
                before += __popc(starts[j]);
                const unsigned idx = __shfl_sync(0xffffffffu, base, owner & 31) + (unsigned)k;
                item[j] = idx;
                if (has[j]) {
                    s[j] = __ldg(kid_src + idx);
                    w[j] = __ldg(kid_info + idx);
                }
            }
            // software pipelining: the tiles pending from the previous round are evaluated while the
            // loads above are in flight
            drain();
            unsigned m_src[ITEMS], m_push[ITEMS], m_bucket = 0;
            bool is_src[ITEMS], is_push[ITEMS], is_bucket[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const bool body = has[j] && (int)w[j].y < 0;
                const bool ok = has[j] && accepts(s[j], __uint_as_float(w[j].y));
                is_src[j] = QUAD ? body : ok;              // quadrupole option: accepted CELLS go to their own list
                is_bucket[j] = has[j] && !ok && (w[j].x & BH_KID_BUCKET);
                is_push[j] = has[j] && !ok && !(w[j].x & BH_KID_BUCKET);
                my_bodies += body;
                my_cells += ok && !body;
                m_src[j] = __ballot_sync(0xffffffffu, is_src[j]);
                m_push[j] = __ballot_sync(0xffffffffu, is_push[j]);
                m_bucket |= __ballot_sync(0xffffffffu, is_bucket[j]);
            }
            if (sp + T > STACK_CAP) { if (lane == 0) atomicOr(&sc->err, BH_DERR_STACK); break; }   // unreachable by construction
            if (QUAD) {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) {
                    const bool qcell = has[j] && (int)w[j].y >= 0 && !is_push[j] && !is_bucket[j];   // an accepted cell
                    const unsigned mq = __ballot_sync(0xffffffffu, qcell);
                    if (qcell) {
                        const int slot = (qhead + nq + __popc(mq & lt_mask)) & (Q_CAP - 1);
                        W.qsrc[3 * slot] = s[j];
                        W.qsrc[3 * slot + 1] = __ldg(kid_quad + 2 * (size_t)item[j]);
                        W.qsrc[3 * slot + 2] = __ldg(kid_quad + 2 * (size_t)item[j] + 1);
                    }
                    nq += __popc(mq);
                }
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                if (is_src[j]) store_source(slist, (head + ns + __popc(m_src[j] & lt_mask)) & (SRC_CAP - 1), s[j]);
                if (is_push[j]) W.stack[sp + __popc(m_push[j] & lt_mask)] = w[j].x;
                ns += __popc(m_src[j]);
                sp += __popc(m_push[j]);
            }
            max_sp = max(max_sp, (unsigned)sp);
            __syncwarp();

            // ---- rejected buckets (identical keys): their bodies are a contiguous range ----
            if (m_bucket) {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) {
                    unsigned bm = __ballot_sync(0xffffffffu, is_bucket[j]);
                    while (bm) {
                        const int srcl = __ffs(bm) - 1;
                        bm &= bm - 1;
                        const int id = (int)((__shfl_sync(0xffffffffu, w[j].x, srcl) & ~BH_KID_BUCKET) >> 3);
                        const int4 bmt = __ldg(cell_meta + id);
                        append_bucket(bmt.x, bmt.y);
                    }
                }
            }
        }
        const unsigned acc_cells = __reduce_add_sync(0xffffffffu, my_cells);
        const unsigned dir_bodies = __reduce_add_sync(0xffffffffu, my_bodies) + bucket_bodies;

        // ---- what is still pending: full tiles, then the last partial one (zero-mass padding adds 0) ----
        drain();
        if (ns > 0) {
            if (lane >= ns) store_source(slist, head + lane, make_float4(0.f, 0.f, 0.f, 0.f));
            __syncwarp();
            eval_tile(slist + (head >> 1), npx, npy, npz, soft2, t);
            __syncwarp();
        }

        if (QUAD && nq > 0) {   // the last partial tile of cells: zero mass and zero moments add exactly 0
            if (lane >= nq) {
                const int slot = (qhead + lane) & (Q_CAP - 1);
                W.qsrc[3 * slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                W.qsrc[3 * slot + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                W.qsrc[3 * slot + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
            eval_qtile(W.qsrc + 3 * qhead, me.x, me.y, me.z, soft, qax, qay, qaz);
            __syncwarp();
        }
        if (in_group) {
            float lo, hi;
            unpack2(t.x, lo, hi); ax = lo + hi + qax;
            unpack2(t.y, lo, hi); ay = lo + hi + qay;
            unpack2(t.z, lo, hi); az = lo + hi + qaz;
        }
        acc_cells_w += acc_cells * gsize;
        dir_bodies_w += dir_bodies * gsize;
        chunk_entries += acc_cells + dir_bodies;
        ga = gb;
        }   // sub-groups of the chunk
        tot_entries += chunk_entries;
        if (lane == 0 && !accumulate && chunk_entries > heavy_thresh && (int64_t)g < max_chunks) {
            const unsigned slot = atomicAdd(&sc->heavy_n[nxt], 1u);
            if ((int64_t)slot < max_chunks) { list_nxt[slot] = g; flag_nxt[g] = epoch + 1u; }
        }

        float4 a = make_float4(G * ax, G * ay, G * az, (float)chunk_entries);
        if (valid) {
            if (accumulate) { const float4 o = acc[my]; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
            acc[my] = a;
        }
        if (fu.posm_out != nullptr) {   // warp-uniform.  integrate_kernel's arithmetic (bh_state.cu), bit for bit
            float4 p = me, v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) v = __ldg(fu.vel_s + my);
            float x = __fmaf_rn(a.x, fu.dt, v.x), y = __fmaf_rn(a.y, fu.dt, v.y), z = __fmaf_rn(a.z, fu.dt, v.z);
            const float s2 = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
            if (s2 > __fmul_rn(fu.max_speed, fu.max_speed)) {
                const float scale = __fdiv_rn(fu.max_speed, __fsqrt_rn(s2));
                x = __fmul_rn(x, scale); y = __fmul_rn(y, scale); z = __fmul_rn(z, scale);
            }
            v.x = x; v.y = y; v.z = z;
            p.x = __fmaf_rn(x, fu.dt, p.x); p.y = __fmaf_rn(y, fu.dt, p.y); p.z = __fmaf_rn(z, fu.dt, p.z);
            if (valid) { fu.vel_out[my] = v; fu.posm_out[my] = p; fu.ids_out[my] = __ldg(fu.ids_s + my); }
            // min/max as integrate_kernel forms them (fminf / fmaxf against the start values), reduced in the ordered image
            const unsigned lx = __reduce_min_sync(0xffffffffu, valid ? bh_f2ord(fminf(1e10f, p.x)) : 0xFFFFFFFFu);
            const unsigned ly = __reduce_min_sync(0xffffffffu, valid ? bh_f2ord(fminf(1e10f, p.y)) : 0xFFFFFFFFu);
            const unsigned lz = __reduce_min_sync(0xffffffffu, valid ? bh_f2ord(fminf(1e10f, p.z)) : 0xFFFFFFFFu);
            const unsigned ux = __reduce_max_sync(0xffffffffu, valid ? bh_f2ord(fmaxf(-1e10f, p.x)) : 0u);
            const unsigned uy = __reduce_max_sync(0xffffffffu, valid ? bh_f2ord(fmaxf(-1e10f, p.y)) : 0u);
            const unsigned uz = __reduce_max_sync(0xffffffffu, valid ? bh_f2ord(fmaxf(-1e10f, p.z)) : 0u);
            if (lane == 0) {
                bbox[0] = min(bbox[0], lx); bbox[1] = min(bbox[1], ly); bbox[2] = min(bbox[2], lz);
                bbox[3] = max(bbox[3], ux); bbox[4] = max(bbox[4], uy); bbox[5] = max(bbox[5], uz);
            }
        }
        tot_cell += acc_cells_w;
        tot_body += dir_bodies_w;
    }

    if (lane == 0) {
        if (tot_cell) atomicAdd(&sc->inter_cell, tot_cell);
        if (tot_body) atomicAdd(&sc->inter_body, tot_body);
        if (tot_entries && !accumulate) atomicAdd(&sc->entries_total, tot_entries);
        atomicMax(&sc->max_stack, max_sp);
        if (fu.posm_out != nullptr) {
            for (int k = 0; k < 3; ++k) { atomicMin(&sc->bbox_enc[k], bbox[k]); atomicMax(&sc->bbox_enc[3 + k], bbox[3 + k]); }
        }
    }
}

__global__ void reset_force_scalars(BhDevScalars* sc, int64_t ngroups) {
    // advance the epoch, derive the heavy threshold from the previous launch, empty the list this launch fills
    if (threadIdx.x == 0) {
        sc->epoch += 1u;
        const unsigned nxt = (sc->epoch & 1u) ^ 1u;
        const unsigned long long prev = sc->entries_total;
        // FORCE_HEAVY_X10 / 10 times the mean list length of the previous launch; nothing is heavy on the first one
        sc->heavy_thresh = prev ? (unsigned)min((unsigned long long)0x7FFFFFFF, prev * (unsigned long long)FORCE_HEAVY_X10 / (10ull * (unsigned long long)ngroups) + 1ull)
                                : 0x7FFFFFFFu;
        sc->heavy_n[nxt] = 0;
        sc->entries_total = 0;
        sc->group_ticket = 0;
        sc->inter_cell = 0;
        sc->inter_body = 0;
        sc->max_stack = 0;
    }
}

__global__ void reset_ticket_only(BhDevScalars* sc) {
    if (threadIdx.x == 0) sc->group_ticket = 0;
}

__global__ void __launch_bounds__(256) zero_acc_kernel(float4* acc, int64_t first, int64_t count) {
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < count; t += (int64_t)gridDim.x * 256)
        acc[first + t] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

static int g_force_ctas_per_sm = 0, g_force_ctas_per_sm_quad = 0;
// occupancy query done once, outside any stream capture
int bh_force_prepare() {
    if (g_force_ctas_per_sm > 0) return 0;
    int a = 0, b = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, force_kernel<10, false>, FORCE_THREADS, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, force_kernel<20, false>, FORCE_THREADS, 0);
    if (e != cudaSuccess) return (int)e;
    const int max_ctas = a < b ? a : b;
    g_force_ctas_per_sm = max_ctas < 1 ? 1 : max_ctas;
    a = b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, force_kernel<10, true>, FORCE_THREADS, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, force_kernel<20, true>, FORCE_THREADS, 0);
    if (e != cudaSuccess) return (int)e;
    g_force_ctas_per_sm_quad = (a < b ? a : b) < 1 ? 1 : (a < b ? a : b);
    return 0;
}

int bh_force_launch(const float4* posm, const void* keys, int levels, const int32_t* ids, int64_t n, int64_t first_body,
                    int64_t body_count, const int4* cell_meta, const float4* cell_com,
                    const float4* kid_src, const uint2* kid_info,
                    float4* acc, BhDevScalars* sc, uint32_t* heavy_list, uint32_t* heavy_flag, int64_t max_chunks,
                    float theta, float softening, float G, float split_alpha, int num_sms, const float4* src_posm,
                    const BhDevScalars* tree_sc, int accumulate, const float4* cell_quad, const float4* kid_quad,
                    cudaStream_t st, const BhFusedUpdate* fused) {
    if (body_count <= 0) return 0;
    if (fused && (accumulate || n < 2 || ids != nullptr)) return BH_E_INVAL;   // the update needs final accelerations of every body
    const BhFusedUpdate fu = fused ? *fused : BhFusedUpdate{nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f};
    if (!src_posm) src_posm = posm;
    if (!tree_sc) tree_sc = sc;
    if (accumulate) {
        reset_ticket_only<<<1, 32, 0, st>>>(sc);
    } else {
        reset_force_scalars<<<1, 32, 0, st>>>(sc, (body_count + BH_GROUP - 1) / BH_GROUP);
    }
    if (n < 2 && !accumulate) {  // a single body feels nothing (its self term is exactly zero, bench:205-213)
        zero_acc_kernel<<<1, 256, 0, st>>>(acc, first_body, body_count);
        return (int)cudaGetLastError();
    }
    { int e = bh_force_prepare(); if (e) return e; }
    const bool quad = cell_quad != nullptr && kid_quad != nullptr;
    const int max_ctas = quad ? g_force_ctas_per_sm_quad : g_force_ctas_per_sm;
    const int64_t ngroups = (body_count + BH_GROUP - 1) / BH_GROUP;
    int64_t want = (ngroups + FORCE_WARPS - 1) / FORCE_WARPS;
    int64_t grid = (int64_t)(num_sms > 0 ? num_sms : BH_NUM_SMS_FALLBACK) * max_ctas;  // persistent: fill the chip once
    if (grid > want) grid = want;
#define BH_FORCE(L, Q, KT) force_kernel<L, Q><<<(int)grid, FORCE_THREADS, 0, st>>>(posm, (const KT*)keys, ids, first_body, body_count, cell_meta, \
        cell_com, kid_src, kid_info, acc, sc, heavy_list, heavy_flag, max_chunks, theta, softening, G, split_alpha, src_posm, tree_sc, accumulate, cell_quad, kid_quad, fu)
    if (levels == 20) { if (quad) BH_FORCE(20, true, uint64_t); else BH_FORCE(20, false, uint64_t); }
    else { if (quad) BH_FORCE(10, true, uint32_t); else BH_FORCE(10, false, uint32_t); }
#undef BH_FORCE
    return (int)cudaGetLastError();
}
