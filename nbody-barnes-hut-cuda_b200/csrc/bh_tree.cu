// bh_tree.cu — parallel emission of the compressed octree from the sorted 30-bit Morton keys,
// and the bottom-up centre-of-mass pass.
//
// Replaces, for the whole tree at once and without a zeroed 2N node pool:
//   cudaMemset + initRootKernel      nbody_v5_bench.cu:266-267, 65-81
//   insertParticlesKernel x N/1024   nbody_v5_bench.cu:83-132, 269-275
//   cudaMemcpy(&hCount ...)          nbody_v5_bench.cu:277-278   (no host round trip here)
//   computeCOMKernel/finalizeCOM     nbody_v5_bench.cu:158-189   (no float atomics here)
//
// Canonical tree (DESIGN.md §"Tree"): a CELL is a maximal run of >= 2 sorted bodies whose keys
// share exactly L leading 3-bit digits (L = level, 0..10).  L == 10 means identical keys: a
// BUCKET leaf holding its bodies as a contiguous range.  A cell with L < 10 has up to 8
// children, one per value of digit L+1: a single body, or another cell.  Bodies sit directly in
// their parent's child table (one body per leaf, as the reference's insertion tree).
//
// Karras-style construction, one thread per adjacent key pair j = (j, j+1):
//   lv(j)      = digits shared by K[j], K[j+1]   — pair j is a split point of the level-lv(j)
//                cell that contains it;
//   leader(j)  = the split point that ends the FIRST child of that cell (found with two
//                galloping searches on the sorted keys) — exactly one leader per cell;
//   cell id    = exclusive prefix sum of the leader flags (so ids ascend with leader position);
//   parent     = the cell owning the tighter of the two pairs that bound the range, lv(l-1) vs
//                lv(r) — its id is scan[leader(that pair)].
// Every cell and every loose body then writes itself into its parent's child table.
//
// Traversal view of the tree (written by the centre-of-mass pass, read by bh_force.cu / bh_let.cu): per cell one
// DENSE line of its children in digit order — kid_src[8c + r] = what child r contributes as a source (a loose
// body's {x,y,z,m}, a child cell's {centre of mass, mass}), kid_info[8c + r] = {stack word of the child cell
// (id << 3 | its own child count - 1 | bucket flag), its squared width as float bits (-1 for a body)}, r < number of children.  The octree of
// the reference disk holds 3.2 children per cell: dense lines let the traversal spend its lanes on children that
// exist.  cell_child stays the canonical digit-indexed table (parity tests, this pass).
#include "bh_common.cuh"

namespace {

constexpr int TB = 256;           // threads per CTA
constexpr int SCAN_ITEMS = 8;     // pairs per thread in the flag scan
constexpr int SCAN_TILE = TB * SCAN_ITEMS;

template <int LEVELS>
__device__ __forceinline__ int lv_pair(const typename BhKey<LEVELS>::type* __restrict__ K, int j, int n) {
    if (j < 0 || j >= n - 1) return -1;
    return bh_shared_digits_t<LEVELS>(__ldg(K + j), __ldg(K + j + 1));
}

// smallest index t <= j whose key shares L digits with K[j] (keys ascending)
template <int LEVELS>
__device__ int find_left(const typename BhKey<LEVELS>::type* __restrict__ K, int j, int L) {
    typedef typename BhKey<LEVELS>::type KeyT;
    if (L <= 0) return 0;
    const int s = 3 * LEVELS - 3 * L;
    const KeyT p = __ldg(K + j) >> s;
    int good = j, bad = -1, step = 1;
    for (;;) {
        int t = good - step;
        if (t < 0) { bad = -1; break; }
        if ((__ldg(K + t) >> s) == p) { good = t; step <<= 1; }
        else { bad = t; break; }
    }
    while (good - bad > 1) {
        int mid = (good + bad) >> 1;
        if ((__ldg(K + mid) >> s) == p) good = mid; else bad = mid;
    }
    return good;
}

// largest index t >= j whose key shares L digits with K[j]
template <int LEVELS>
__device__ int find_right(const typename BhKey<LEVELS>::type* __restrict__ K, int n, int j, int L) {
    typedef typename BhKey<LEVELS>::type KeyT;
    if (L <= 0) return n - 1;
    const int s = 3 * LEVELS - 3 * L;
    const KeyT p = __ldg(K + j) >> s;
    int good = j, bad = n, step = 1;
    for (;;) {
        int t = good + step;
        if (t >= n) { bad = n; break; }
        if ((__ldg(K + t) >> s) == p) { good = t; step <<= 1; }
        else { bad = t; break; }
    }
    while (bad - good > 1) {
        int mid = (good + bad) >> 1;
        if ((__ldg(K + mid) >> s) == p) good = mid; else bad = mid;
    }
    return good;
}

// ---- pass A: leader of every pair + per-tile leader counts ---------------------------------
template <int LEVELS>
__global__ void __launch_bounds__(TB) pair_kernel(const typename BhKey<LEVELS>::type* __restrict__ K, int n, int2* __restrict__ pair_info,
                                                 int32_t* __restrict__ tile_sums) {
    const int npairs = n - 1;
    const int tile0 = blockIdx.x * SCAN_TILE;
    int leaders = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int j = tile0 + k * TB + threadIdx.x;
        if (j < npairs) {
            const typename BhKey<LEVELS>::type a = __ldg(K + j), b = __ldg(K + j + 1);
            const int L = bh_shared_digits_t<LEVELS>(a, b);
            int lead, first;
            if (L == LEVELS) {  // inside a run of identical keys: the run's first pair leads
                first = j;
                lead = (j == 0 || __ldg(K + j - 1) != a) ? j : -1;
            } else {
                first = find_left<LEVELS>(K, j, L);
                // j leads iff it closes the first child, i.e. K[j] still shares L+1 digits with K[first]
                lead = (bh_shared_digits_t<LEVELS>(__ldg(K + first), a) >= L + 1) ? j : find_right<LEVELS>(K, n, first, L + 1);
            }
            pair_info[j] = make_int2(lead, first);
            leaders += (lead == j);
        }
    }
    // CTA reduction of the leader count
    __shared__ int s_w[TB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) leaders += __shfl_xor_sync(0xffffffffu, leaders, o);
    if (bh_lane() == 0) s_w[threadIdx.x >> 5] = leaders;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < TB / 32; ++w) t += s_w[w];
        tile_sums[blockIdx.x] = t;
    }
}

// ---- pass B: exclusive scan of the tile sums (single CTA), total -> num_cells -----------------
__global__ void __launch_bounds__(1024) scan_tiles_kernel(int32_t* tile_sums, int ntiles, BhDevScalars* sc) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? tile_sums[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)bh_lane() >= o) inc += t;
        }
        if (bh_lane() == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_w[threadIdx.x], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if ((int)bh_lane() >= o) winc += t;
            }
            s_w[threadIdx.x] = winc - w;  // exclusive warp bases
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + s_w[threadIdx.x >> 5] + inc - v;
        if (i < ntiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { sc->num_cells = s_carry; sc->root = -1; }
}

// ---- pass C: per-pair exclusive scan value = cell id of a leading pair -----------------------
__global__ void __launch_bounds__(TB) scan_pairs_kernel(const int2* __restrict__ pair_info, int n,
                                                       const int32_t* __restrict__ tile_sums,
                                                       int32_t* __restrict__ pair_scan) {
    const int npairs = n - 1;
    const int tile0 = blockIdx.x * SCAN_TILE;
    __shared__ int s_w[TB / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = tile_sums[blockIdx.x];
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int j = tile0 + k * TB + threadIdx.x;
        const int v = (j < npairs && pair_info[j].x == j) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        const int in_warp = __popc(bal & ((1u << bh_lane()) - 1u));
        if (bh_lane() == 0) s_w[threadIdx.x >> 5] = __popc(bal);
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TB / 32; ++w) {
            int c = s_w[w];
            if (w < (int)(threadIdx.x >> 5)) wbase += c;
            total += c;
        }
        const int carry = s_carry;
        if (j < npairs) pair_scan[j] = carry + wbase + in_warp;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

// ---- pass D: clear the child tables / arrival counters of the cells that exist --------------
__global__ void __launch_bounds__(TB) init_cells_kernel(int4* __restrict__ child4, int32_t* __restrict__ arrive,
                                                       const BhDevScalars* __restrict__ sc) {
    const int M = sc->num_cells;
    const int4 empty = make_int4(BH_CHILD_EMPTY, BH_CHILD_EMPTY, BH_CHILD_EMPTY, BH_CHILD_EMPTY);
    for (int c = blockIdx.x * TB + threadIdx.x; c < M; c += gridDim.x * TB) {
        child4[2 * c] = empty;
        child4[2 * c + 1] = empty;
        arrive[c] = 0;
    }
}

// ---- pass E: every cell and every loose body links itself under its parent -------------------
template <int LEVELS>
__global__ void __launch_bounds__(TB) link_kernel(const typename BhKey<LEVELS>::type* __restrict__ K, int n,
                                                 const int2* __restrict__ pair_info,
                                                 const int32_t* __restrict__ pair_scan, int4* __restrict__ cell_meta,
                                                 int32_t* __restrict__ cell_child,
                                                 uint8_t* __restrict__ kid_lv, BhDevScalars* sc) {
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB) {
        const typename BhKey<LEVELS>::type ki = __ldg(K + i);
        const int lv_left = lv_pair<LEVELS>(K, i - 1, n);   // pair (i-1, i)
        const int lv_here = lv_pair<LEVELS>(K, i, n);       // pair (i, i+1)

        // (1) body i as a loose child (skipped when it belongs to a bucket)
        const int Lb = max(lv_left, lv_here);
        if (Lb >= 0 && Lb < LEVELS) {
            const int sp = (lv_left >= lv_here) ? i - 1 : i;
            const int parent = pair_scan[pair_info[sp].x];
            const int slot = (int)(ki >> (3 * LEVELS - 3 * (Lb + 1))) & 7;
            cell_child[(size_t)parent * 8 + slot] = (int)(0x80000000u | (uint32_t)i);
        }

        // (2) pair i, if it leads a cell
        if (i < n - 1) {
            const int2 info = pair_info[i];
            if (info.x == i) {
                const int c = pair_scan[i];
                const int L = lv_here;
                const int l = info.y;
                const int r = find_right<LEVELS>(K, n, i, L);
                const int a = lv_pair<LEVELS>(K, l - 1, n), b = lv_pair<LEVELS>(K, r, n);
                const int Lp = max(a, b);
                int parent = -1, slot = 0;
                if (Lp >= 0) {
                    if (Lp >= L) atomicOr(&sc->err, BH_DERR_TREE);
                    const int sp = (a >= b) ? l - 1 : r;
                    parent = pair_scan[pair_info[sp].x];
                    slot = (int)(ki >> (3 * LEVELS - 3 * (Lp + 1))) & 7;
                    cell_child[(size_t)parent * 8 + slot] = c;
                    kid_lv[(size_t)parent * 8 + slot] = (uint8_t)(L | ((L == LEVELS) << 7));
                } else {
                    sc->root = c;
                }
                cell_meta[c] = make_int4(l, r - l + 1, L | ((L == LEVELS) << 8) | (slot << 12), parent);
            }
        }
    }
}

// ---- centre of mass: bottom-up with per-cell arrival counters --------------------------------
struct Moments { float m, x, y, z; };

__device__ __forceinline__ void add_body(Moments& s, const float4 p) {
    s.m = __fadd_rn(s.m, p.w);
    s.x = __fmaf_rn(p.w, p.x, s.x);
    s.y = __fmaf_rn(p.w, p.y, s.y);
    s.z = __fmaf_rn(p.w, p.z, s.z);
}

__device__ __forceinline__ float4 store_cell(float4* __restrict__ mom, float4* __restrict__ com, int c, const Moments& s) {
    __stcg(mom + c, make_float4(s.x, s.y, s.z, s.m));
    const float inv = (s.m > 1e-6f) ? __fdiv_rn(1.0f, s.m) : 0.0f;   // bench:181-183
    const float4 cm = make_float4(__fmul_rn(s.x, inv), __fmul_rn(s.y, inv), __fmul_rn(s.z, inv), s.m);
    com[c] = cm;
    return cm;
}

// Sums the children of a cell in slot (digit) order — run-to-run identical, equal to the oracle bit for bit —
// and writes the dense traversal entries of its loose bodies (child cells write their own entry when they
// finish, see below).  Returns the number of children.
__device__ __forceinline__ int sum_children(const int e[8], const float4* __restrict__ posm, const float4* __restrict__ mom,
                                            int c, float4* __restrict__ kid_src, uint2* __restrict__ kid_info, Moments& t) {
    int r = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (e[q] == BH_CHILD_EMPTY) continue;
        if (e[q] < 0) {
            const float4 p = __ldg(posm + (e[q] & 0x7FFFFFFF));
            add_body(t, p);
            kid_src[(size_t)c * 8 + r] = p;
            kid_info[(size_t)c * 8 + r] = make_uint2(0x7FFFFFFFu, BH_KID_BODY_W2);
        } else {
            const float4 cm = __ldcg(mom + e[q]);
            t.m = __fadd_rn(t.m, cm.w); t.x = __fadd_rn(t.x, cm.x);
            t.y = __fadd_rn(t.y, cm.y); t.z = __fadd_rn(t.z, cm.z);
        }
        ++r;
    }
    return r;
}

__global__ void __launch_bounds__(TB) com_kernel(const float4* __restrict__ posm, const int4* __restrict__ cell_meta,
                                                const int32_t* __restrict__ cell_child, int32_t* __restrict__ arrive,
                                                float4* __restrict__ mom, float4* __restrict__ com,
                                                float4* __restrict__ kid_src, uint2* __restrict__ kid_info, BhDevScalars* sc) {
    const int M = sc->num_cells;
    const int4* child4 = reinterpret_cast<const int4*>(cell_child);
    // squared width of a level-L cell: root_w^2 with 2L taken off the exponent (exact); root_w is the key
    // grid's size, clamped like bench:52
    const float root_w = fmaxf(__fsub_rn(sc->bounds[3], sc->bounds[0]), 1.0f);
    const int root_w2_bits = __float_as_int(__fmul_rn(root_w, root_w));
    for (int c0 = blockIdx.x * TB + threadIdx.x; c0 < M; c0 += gridDim.x * TB) {
        int c = c0;
        int4 mt = __ldg(cell_meta + c);
        Moments s = {0.f, 0.f, 0.f, 0.f};
        int nkids = 1;   // children of c (buckets are never opened as cells: the field is unused for them)
        if ((mt.z >> 8) & 1) {
            for (int i = mt.x; i < mt.x + mt.y; ++i) add_body(s, __ldg(posm + i));
        } else {
            const int4 lo = __ldg(child4 + 2 * c), hi = __ldg(child4 + 2 * c + 1);
            const int e[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            bool has_cell = false;
#pragma unroll
            for (int q = 0; q < 8; ++q) has_cell |= (e[q] >= 0 && e[q] != BH_CHILD_EMPTY);
            if (has_cell) continue;  // finished later by its last-arriving child cell
            nkids = sum_children(e, posm, mom, c, kid_src, kid_info, s);
        }
        float4 cm = store_cell(mom, com, c, s);
        // climb while this thread is the last child cell to arrive
        for (;;) {
            const int p = mt.w;
            if (p < 0) { sc->root_word = ((unsigned)c << 3) | (unsigned)(nkids - 1); break; }
            const int4 lo = __ldg(child4 + 2 * p), hi = __ldg(child4 + 2 * p + 1);
            const int e[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            const int myslot = (mt.z >> 12) & 7;
            int ncc = 0, rank = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ncc += (e[q] >= 0 && e[q] != BH_CHILD_EMPTY);
                rank += (q < myslot && e[q] != BH_CHILD_EMPTY);
            }
            // the parent's view of this cell: source + the word the traversal pushes to open it later
            kid_src[(size_t)p * 8 + rank] = cm;
            kid_info[(size_t)p * 8 + rank] =
                make_uint2(((unsigned)c << 3) | (unsigned)(nkids - 1) | (((mt.z >> 8) & 1) ? BH_KID_BUCKET : 0u),
                           (unsigned)(root_w2_bits - ((mt.z & 0xFF) << 24)));
            __threadfence();   // publish this cell's moments (st.cg) before announcing arrival
            const int old = atomicAdd(arrive + p, 1);
            if (old + 1 < ncc) break;
            // last arrival: the siblings' moments were fenced before their own atomics and are read with
            // ld.cg (L2, the coherence point) below, after the atomic's result is known — the pattern of the
            // CUDA threadFenceReduction sample; no second fence is needed
            Moments t = {0.f, 0.f, 0.f, 0.f};
            nkids = sum_children(e, posm, mom, p, kid_src, kid_info, t);
            c = p;
            mt = __ldg(cell_meta + c);
            cm = store_cell(mom, com, c, t);
        }
    }
}

inline int capped_grid(int64_t work_items, int per_block) {
    int64_t b = (work_items + per_block - 1) / per_block;
    if (b > BH_NUM_SMS_FALLBACK * 8) b = BH_NUM_SMS_FALLBACK * 8;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

// keys: ascending u32 keys of 10 digits (levels == 10, the reference's 30-bit key) or u64 keys of 20 digits
// (levels == 20, bh_params.key_bits = 60)
int bh_tree_launch(const void* keys, int levels, int64_t n64, int2* pair_info, int32_t* pair_scan,
                   int32_t* scan_block_sums, int4* cell_meta, int32_t* cell_child,
                   int32_t* cell_arrive, uint8_t* kid_lv, BhDevScalars* sc, cudaStream_t st) {
    const int n = (int)n64;
    if (n < 2) return 0;
    const int ntiles = (n - 1 + SCAN_TILE - 1) / SCAN_TILE;
    if (levels == 20) pair_kernel<20><<<ntiles, TB, 0, st>>>((const uint64_t*)keys, n, pair_info, scan_block_sums);
    else pair_kernel<10><<<ntiles, TB, 0, st>>>((const uint32_t*)keys, n, pair_info, scan_block_sums);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(scan_block_sums, ntiles, sc);
    scan_pairs_kernel<<<ntiles, TB, 0, st>>>(pair_info, n, scan_block_sums, pair_scan);
    init_cells_kernel<<<capped_grid(n, TB), TB, 0, st>>>(reinterpret_cast<int4*>(cell_child), cell_arrive, sc);
    if (levels == 20)
        link_kernel<20><<<capped_grid(n, TB), TB, 0, st>>>((const uint64_t*)keys, n, pair_info, pair_scan, cell_meta, cell_child, kid_lv, sc);
    else
        link_kernel<10><<<capped_grid(n, TB), TB, 0, st>>>((const uint32_t*)keys, n, pair_info, pair_scan, cell_meta, cell_child, kid_lv, sc);
    return (int)cudaGetLastError();
}

int bh_com_launch(const float4* posm, int64_t n, const int4* cell_meta, const int32_t* cell_child,
                  int32_t* cell_arrive, float4* cell_mom, float4* cell_com, float4* kid_src, uint2* kid_info,
                  BhDevScalars* sc, cudaStream_t st) {
    if (n < 2) return 0;
    // one thread per possible cell, no grid-stride: a thread that climbs towards the root must not delay
    // the leaf-level cells a strided loop would hand it next
    com_kernel<<<(int)((n + TB - 1) / TB), TB, 0, st>>>(posm, cell_meta, cell_child, cell_arrive, cell_mom, cell_com, kid_src, kid_info, sc);
    return (int)cudaGetLastError();
}
