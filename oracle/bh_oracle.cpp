// bh_oracle.cpp — CPU restatement of the Barnes-Hut step path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product path (nbody-barnes-hut-cuda_b200/csrc) never does.
//
// Parity status: the reference ships no tests and no golden vectors, so
// the pin is the reference ITSELF: oracle/ref_wrap.cu compiles the unmodified kernels and
// simulationStep() of /root/reference/nbody_v5_bench.cu for sm_100 into oracle/_ref/, and
// tests/test_gpu_reference_pin.py checks Oracle-L (below) against it on a B200 — bounds, keys
// and the sort permutation bit for bit, accelerations/positions to float-atomic noise.
// tests/golden/ holds outputs of that run for the CPU-only suite.
//
// Two oracles:
//   Oracle-L  literal transliteration of nbody_v5_bench.cu:42-249 including findings F1-F3
//             (node/body id aliasing => the root is accepted at once => 1 interaction/body).
//             Serial insertion in sorted order is one legal schedule of the racy kernel (F4).
//             `fixed`=1 applies the 6-line id fix of SURVEY H5 (node ids offset by n, body
//             branch in the force loop) = the algorithm README.md:72-80 documents.
//   Oracle-I  the algorithm the engine ships: canonical compressed octree over the sorted
//             30-bit keys, slot-ordered centre-of-mass sums, group (32 Morton-consecutive
//             bodies) acceptance test derived from bench:205-208, force law bench:210-213.
// Build: g++ -O2 -fopenmp -ffp-contract=off (explicit fmaf where the reference SASS has FFMA).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

double now_ms() {
#ifdef _OPENMP
    return omp_get_wtime() * 1e3;
#else
    return 0.0;
#endif
}

// ---- bench:42-49 -------------------------------------------------------------------------
inline uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

inline uint32_t quantise(float p, float lo, float size) {
    // bench:58 — IEEE divide, then multiply, then truncating float->u32 (F2I.U32.TRUNC,
    // which saturates negatives to 0; C leaves that undefined so clamp explicitly).
    float t = (p - lo) / size * 1023.0f;
    if (!(t > 0.0f)) return 0u;
    if (t >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

}  // namespace

extern "C" {

// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for the cores it reports
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// bench:134-156 — cube anchored at the min corner.
void orc_bounds(const float* px, const float* py, const float* pz, int64_t n, float b[6]) {
    float minX = 1e10f, minY = 1e10f, minZ = 1e10f, maxX = -1e10f, maxY = -1e10f, maxZ = -1e10f;
    for (int64_t i = 0; i < n; ++i) {
        minX = fminf(minX, px[i]); minY = fminf(minY, py[i]); minZ = fminf(minZ, pz[i]);
        maxX = fmaxf(maxX, px[i]); maxY = fmaxf(maxY, py[i]); maxZ = fmaxf(maxZ, pz[i]);
    }
    float size = fmaxf(maxX - minX, fmaxf(maxY - minY, maxZ - minZ));
    b[0] = minX; b[1] = minY; b[2] = minZ;
    b[3] = minX + size; b[4] = minY + size; b[5] = minZ + size;
}

// bench:51-63.
void orc_morton_keys(const float* px, const float* py, const float* pz, int64_t n,
                     const float b[6], uint32_t* keys, int32_t* idx) {
    const float size = fmaxf(b[3] - b[0], 1.0f);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t x = quantise(px[i], b[0], size);
        uint32_t y = quantise(py[i], b[1], size);
        uint32_t z = quantise(pz[i], b[2], size);
        keys[i] = (spread10(x) << 2) | (spread10(y) << 1) | spread10(z);
        if (idx) idx[i] = (int32_t)i;
    }
}

// bh_params.key_bits = 60: `hi` is the reference key above, `lo` interleaves ten more bits per axis taken
// from the fractional part of the same float t = (p-min)/size*1023.0f (t - trunc(t) is exact in float), so
// (hi << 30 | lo) refines the reference order without ever contradicting it.
void orc_morton_keys60(const float* px, const float* py, const float* pz, int64_t n,
                       const float b[6], uint32_t* hi, uint32_t* lo) {
    const float size = fmaxf(b[3] - b[0], 1.0f);
    auto split = [&](float p, float mn, uint32_t& q, uint32_t& fr) {
        float t = (p - mn) / size * 1023.0f;
        q = quantise(p, mn, size);
        float frac = t - (float)q;
        float g = frac * 1024.0f;
        fr = !(g > 0.0f) ? 0u : (g >= 1023.0f ? 1023u : (uint32_t)g);
    };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t qx, qy, qz, fx, fy, fz;
        split(px[i], b[0], qx, fx); split(py[i], b[1], qy, fy); split(pz[i], b[2], qz, fz);
        hi[i] = (spread10(qx) << 2) | (spread10(qy) << 1) | spread10(qz);
        lo[i] = (spread10(fx) << 2) | (spread10(fy) << 1) | spread10(fz);
    }
}

// stable ascending sort of 64-bit keys; idx follows
void orc_stable_sort64(uint64_t* keys, int32_t* idx, int64_t n) {
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> k2(n);
    std::vector<int32_t> i2(n);
    for (int64_t i = 0; i < n; ++i) { k2[i] = keys[order[i]]; i2[i] = idx[order[i]]; }
    std::memcpy(keys, k2.data(), n * sizeof(uint64_t));
    std::memcpy(idx, i2.data(), n * sizeof(int32_t));
}

// bench:262-264 — thrust::sort_by_key is a stable ascending radix sort.
void orc_stable_sort(uint32_t* keys, int32_t* idx, int64_t n) {
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int64_t a, int64_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> k2(n);
    std::vector<int32_t> i2(n);
    for (int64_t i = 0; i < n; ++i) { k2[i] = keys[order[i]]; i2[i] = idx[order[i]]; }
    std::memcpy(keys, k2.data(), n * sizeof(uint32_t));
    std::memcpy(idx, i2.data(), n * sizeof(int32_t));
}

// bench:227-249 with the FMA contraction nvcc applies:
// v = fma(a,DT,v); s = fma(vz,vz,fma(vx,vx,vy*vy)); clamp; p = fma(v,DT,p).
void orc_integrate(float* px, float* py, float* pz, float* vx, float* vy, float* vz,
                   const float* ax, const float* ay, const float* az, int64_t n,
                   float dt, float max_speed) {
    const float vmax2 = max_speed * max_speed;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float x = fmaf(ax[i], dt, vx[i]);
        float y = fmaf(ay[i], dt, vy[i]);
        float z = fmaf(az[i], dt, vz[i]);
        float s = fmaf(z, z, fmaf(x, x, y * y));
        if (s > vmax2) {
            float scale = max_speed / sqrtf(s);
            x *= scale; y *= scale; z *= scale;
        }
        vx[i] = x; vy[i] = y; vz[i] = z;
        px[i] = fmaf(x, dt, px[i]);
        py[i] = fmaf(y, dt, py[i]);
        pz[i] = fmaf(z, dt, pz[i]);
    }
}

}  // extern "C"

// =========================================================================================
// Oracle-L: the reference's insertion tree, literally (bench:20-28, 65-132, 158-225).
// =========================================================================================
namespace {

struct RefNode {  // bench:20-28
    float mass, comX, comY, comZ;
    float minX, minY, minZ, maxX, maxY, maxZ;
    int children[8];
    int parent;
};

struct RefTree {
    std::vector<RefNode> nodes;
    std::vector<int> leaf;  // d_leafNodeIdx — persists across steps like the device buffer
    int count = 0;
};

// bench:83-132.  `fixed` stores node ids offset by n so `child < n` really
// means "body".  One body at a time in sorted order.
void ref_insert(RefTree& t, const float* px, const float* py, const float* pz, int n,
                const int32_t* indices, const float b[6], bool fixed) {
    const int cap = 2 * n;                      // bench:321
    t.nodes.assign((size_t)cap, RefNode{});     // bench:266 memset 0
    if ((int)t.leaf.size() != n) t.leaf.assign(n, 0);
    RefNode& root = t.nodes[0];                 // bench:65-81
    root.minX = b[0]; root.minY = b[1]; root.minZ = b[2];
    root.maxX = b[3]; root.maxY = b[4]; root.maxZ = b[5];
    root.mass = 0; root.parent = -1;
    for (int k = 0; k < 8; ++k) root.children[k] = -1;
    int counter = 1;
    const int off = fixed ? n : 0;
    for (int s = 0; s < n; ++s) {
        int idx = indices[s];
        int nIdx = 0, depth = 0;
        while (depth < 25) {
            RefNode* node = &t.nodes[nIdx];
            float mx = (node->minX + node->maxX) * 0.5f;
            float my = (node->minY + node->maxY) * 0.5f;
            float mz = (node->minZ + node->maxZ) * 0.5f;
            int oct = (px[idx] >= mx) | ((py[idx] >= my) << 1) | ((pz[idx] >= mz) << 2);
            int child = node->children[oct];
            if (child == -1) {                                  // bench:101-105
                node->children[oct] = idx;
                t.leaf[idx] = nIdx;
                break;
            } else if (child < n) {                             // bench:106-125
                int oldB = child;
                int next = counter++;
                if (next >= cap) { counter = cap; break; }      // pool guard (reference has none)
                RefNode* nn = &t.nodes[next];
                nn->parent = nIdx; nn->mass = 0;
                for (int k = 0; k < 8; ++k) nn->children[k] = -1;
                nn->minX = (oct & 1) ? mx : node->minX; nn->maxX = (oct & 1) ? node->maxX : mx;
                nn->minY = (oct & 2) ? my : node->minY; nn->maxY = (oct & 2) ? node->maxY : my;
                nn->minZ = (oct & 4) ? mz : node->minZ; nn->maxZ = (oct & 4) ? node->maxZ : mz;
                int oldOct = (px[oldB] >= (nn->minX + nn->maxX) * 0.5f) |
                             ((py[oldB] >= (nn->minY + nn->maxY) * 0.5f) << 1) |
                             ((pz[oldB] >= (nn->minZ + nn->maxZ) * 0.5f) << 2);
                nn->children[oldOct] = oldB;
                t.leaf[oldB] = next;
                node->children[oct] = next + off;
                nIdx = next;
            } else {                                            // bench:126-129
                nIdx = child - off;
                if (nIdx < 0 || nIdx >= cap) break;
            }
            ++depth;
        }
    }
    t.count = counter;
}

// bench:158-189 — one body after another in id order (a legal order of the float atomics).
void ref_com(RefTree& t, const float* px, const float* py, const float* pz,
             const float* mass, int n) {
    for (int i = 0; i < n; ++i) {
        int curr = t.leaf[i];
        float m = mass[i], cx = px[i] * m, cy = py[i] * m, cz = pz[i] * m;
        int guard = 0;
        while (curr != -1 && guard++ < 64) {
            RefNode& nd = t.nodes[curr];
            nd.mass += m; nd.comX += cx; nd.comY += cy; nd.comZ += cz;
            curr = nd.parent;
        }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < t.count; ++i) {
        float m = t.nodes[i].mass;
        if (m > 1e-6f) {
            float inv = 1.0f / m;
            t.nodes[i].comX *= inv; t.nodes[i].comY *= inv; t.nodes[i].comZ *= inv;
        }
    }
}

// bench:191-225.  fixed=0: literal (idx < n accepts the root, F2).  fixed=1: ids >= n are
// nodes, ids < n are bodies read from pos/mass.  Returns interactions; tracks max stack.
int64_t ref_force(const RefTree& t, const float* px, const float* py, const float* pz,
                  const float* mass, int n, float* ax, float* ay, float* az,
                  float G, float theta, float soft, bool fixed, int* max_stack_out) {
    int64_t inter = 0;
    int max_stack = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : inter) reduction(max : max_stack)
    for (int i = 0; i < n; ++i) {
        float x = px[i], y = py[i], z = pz[i], fx = 0, fy = 0, fz = 0;
        std::vector<int> stack;
        stack.reserve(128);
        stack.push_back(fixed ? n : 0);                              // bench:198
        while (!stack.empty()) {
            int idx = stack.back();
            stack.pop_back();
            float m, cx, cy, cz, width = 0;
            bool is_body;
            const RefNode* node = nullptr;
            if (fixed) {
                is_body = idx < n;
                if (is_body) { m = mass[idx]; cx = px[idx]; cy = py[idx]; cz = pz[idx]; }
                else { node = &t.nodes[idx - n]; m = node->mass; cx = node->comX; cy = node->comY; cz = node->comZ; width = node->maxX - node->minX; }
            } else {
                is_body = idx < n;                                   // F2: always true for idx 0
                node = &t.nodes[idx];                                // F3: still reads the node record
                m = node->mass; cx = node->comX; cy = node->comY; cz = node->comZ;
                width = node->maxX - node->minX;
            }
            if (m <= 0) continue;                                    // bench:203
            float dx = cx - x, dy = cy - y, dz = cz - z;
            float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));          // SASS order, SURVEY R11
            float dist = sqrtf(d2 + soft);                           // bench:207
            if (is_body || width / dist < theta) {                   // bench:208
                float f = G * m / (dist * dist * dist);              // bench:210
                fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
                ++inter;
            } else {
                for (int k = 0; k < 8; ++k)                          // bench:217-219
                    if (node->children[k] != -1) stack.push_back(node->children[k]);
                if ((int)stack.size() > max_stack) max_stack = (int)stack.size();
            }
        }
        ax[i] = fx; ay[i] = fy; az[i] = fz;
    }
    if (max_stack_out) *max_stack_out = max_stack;
    return inter;
}

}  // namespace

extern "C" {

// One or more literal simulationStep() calls (bench:255-283) on SoA state in place.
// phase_ms[6] (accumulated): bounds+keys, sort, insert, com, force, integrate.
// info[0]=node count, info[1]=interactions of the last step, info[2]=max traversal stack.
int orc_reference_step(float* px, float* py, float* pz, float* vx, float* vy, float* vz,
                       float* ax, float* ay, float* az, const float* mass, int64_t n64,
                       int nsteps, int fixed, float G, float theta, float dt, float soft,
                       float max_speed, double* phase_ms, int64_t* info,
                       uint32_t* keys_out, int32_t* idx_out, float* bounds_out) {
    if (n64 <= 0 || n64 > (1 << 30)) return -1;
    const int n = (int)n64;
    std::vector<uint32_t> keys(n);
    std::vector<int32_t> idx(n);
    RefTree tree;
    float b[6];
    double ph[6] = {0, 0, 0, 0, 0, 0};
    int64_t inter = 0;
    int max_stack = 0;
    for (int s = 0; s < nsteps; ++s) {
        double t0 = now_ms();
        orc_bounds(px, py, pz, n, b);
        orc_morton_keys(px, py, pz, n, b, keys.data(), idx.data());
        double t1 = now_ms();
        orc_stable_sort(keys.data(), idx.data(), n);
        double t2 = now_ms();
        ref_insert(tree, px, py, pz, n, idx.data(), b, fixed != 0);
        double t3 = now_ms();
        ref_com(tree, px, py, pz, mass, n);
        double t4 = now_ms();
        inter = ref_force(tree, px, py, pz, mass, n, ax, ay, az, G, theta, soft, fixed != 0, &max_stack);
        double t5 = now_ms();
        if (s == nsteps - 1) {  // expose the pre-integration intermediates of the last step
            if (keys_out) std::memcpy(keys_out, keys.data(), n * sizeof(uint32_t));
            if (idx_out) std::memcpy(idx_out, idx.data(), n * sizeof(int32_t));
            if (bounds_out) std::memcpy(bounds_out, b, sizeof(b));
        }
        orc_integrate(px, py, pz, vx, vy, vz, ax, ay, az, n, dt, max_speed);
        double t6 = now_ms();
        ph[0] += t1 - t0; ph[1] += t2 - t1; ph[2] += t3 - t2;
        ph[3] += t4 - t3; ph[4] += t5 - t4; ph[5] += t6 - t5;
    }
    if (phase_ms) std::memcpy(phase_ms, ph, sizeof(ph));
    if (info) { info[0] = tree.count; info[1] = inter; info[2] = max_stack; }
    return 0;
}

}  // extern "C"

// =========================================================================================
// Oracle-I: the shipped algorithm.
// =========================================================================================
namespace {

// Keys are handled as 64-bit values with `levels` 3-bit digits: 10 digits = the reference's 30-bit key
// (bench:57-61), 20 digits = the same key extended by 10 fractional bits per axis (bh_params.key_bits = 60,
// the top 30 bits stay the reference key bit for bit).
constexpr int MAX_LEVELS_ANY = 20;
constexpr int CHILD_EMPTY = 0x7F7F7F7F;

// number of leading 3-bit digits two keys share (0..10)
inline int shared_digits(uint64_t a, uint64_t b, int levels) {
    uint64_t x = a ^ b;
    if (x == 0) return levels;
    int lead = __builtin_clzll(x) - (64 - 3 * levels);
    return lead / 3;
}
inline int digit_at(uint64_t key, int level /*1-based digit index*/, int levels) {
    return (int)((key >> (3 * (levels - level))) & 7);
}

struct Cell {
    int first, count, level, bucket, parent, slot;
    int leader;  // pair index that numbers the cell (see DESIGN.md "cell numbering")
    int child[8];
};

struct TreeBuilder {
    const uint64_t* k;
    int levels;
    std::vector<Cell> cells;

    // returns a child-table entry for the range [first, first+count)
    int build(int first, int count, int parent, int slot) {
        if (count == 1) return (int)(0x80000000u | (uint32_t)first);
        int L = shared_digits(k[first], k[first + count - 1], levels);
        Cell c;
        c.first = first; c.count = count; c.level = L; c.parent = parent; c.slot = slot;
        c.bucket = (L == levels);
        for (int q = 0; q < 8; ++q) c.child[q] = CHILD_EMPTY;
        int me = (int)cells.size();
        cells.push_back(c);
        if (L == levels) {
            cells[me].leader = first;  // first equal pair of the run
            return me;
        }
        int pos = first, end = first + count;
        bool first_child = true;
        while (pos < end) {
            int d = digit_at(k[pos], L + 1, levels);
            int e = pos + 1;
            while (e < end && digit_at(k[e], L + 1, levels) == d) ++e;
            if (first_child) { cells[me].leader = e - 1; first_child = false; }
            int entry = build(pos, e - pos, me, d);
            cells[me].child[d] = entry;
            pos = e;
        }
        return me;
    }
};

inline float w2_of_level(float root_w, int level) {
    float w = ldexpf(root_w, -level);  // exact scaling
    return w * w;
}

}  // namespace

extern "C" {

// Canonical tree over ascending keys.  meta: int4 per cell {first,count,level|bucket<<8|slot<<12,parent}
// (slot = which child of its parent the cell is),
// child: 8 ints per cell.  Cells are numbered by ascending leader pair (the numbering the
// parallel builder produces with a prefix sum).  Returns the cell count, or -1 if cap is short.
int orc_tree_build64(const uint64_t* sorted_keys, int64_t n64, int levels, int32_t* meta, int32_t* child,
                     int64_t cap, int32_t* root_out);

int orc_tree_build(const uint32_t* sorted_keys, int64_t n64, int32_t* meta, int32_t* child,
                   int64_t cap, int32_t* root_out) {
    std::vector<uint64_t> k64((size_t)std::max<int64_t>(n64, 1));
    for (int64_t i = 0; i < n64; ++i) k64[i] = sorted_keys[i];
    return orc_tree_build64(k64.data(), n64, 10, meta, child, cap, root_out);
}

// Same for keys of `levels` digits held in 64 bits (levels = 10 or 20).
int orc_tree_build64(const uint64_t* sorted_keys, int64_t n64, int levels, int32_t* meta, int32_t* child,
                     int64_t cap, int32_t* root_out) {
    const int n = (int)n64;
    if (root_out) *root_out = -1;
    if (n < 2) return 0;
    TreeBuilder tb;
    tb.k = sorted_keys;
    tb.levels = levels;
    tb.cells.reserve(n);
    tb.build(0, n, -1, 0);
    const int M = (int)tb.cells.size();
    if (M > cap) return -1;
    std::vector<int> order(M), rank(M);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(),
              [&](int a, int b) { return tb.cells[a].leader < tb.cells[b].leader; });
    for (int r = 0; r < M; ++r) rank[order[r]] = r;
    for (int old = 0; old < M; ++old) {
        const Cell& c = tb.cells[old];
        int id = rank[old];
        meta[4 * id + 0] = c.first;
        meta[4 * id + 1] = c.count;
        meta[4 * id + 2] = c.level | (c.bucket << 8) | ((c.parent < 0 ? 0 : c.slot) << 12);
        meta[4 * id + 3] = c.parent < 0 ? -1 : rank[c.parent];
        for (int q = 0; q < 8; ++q) {
            int e = c.child[q];
            child[8 * id + q] = (e == CHILD_EMPTY || e < 0) ? e : rank[e];
        }
        if (c.parent < 0 && root_out) *root_out = id;
    }
    return M;
}

// Centre of mass (engine: bh_tree.cu "centre of mass").  A cell is a contiguous range of the sorted bodies.  Its
// moments {m, m x, m y, m z} are taken in double (m x is exact in double): a cell of at most 16 bodies adds its
// bodies up directly, in order; a larger one takes a difference of prefix sums.  The reference adds the same terms
// up the parent chain with float atomics in schedule order (bench:158-176); any order is "the reference's", this
// one is fixed and more accurate than all of them.
// The order of every addition mirrors the kernels: blocks of 4,096 bodies; 16 consecutive bodies per thread in
// sequence (a "run"); Hillis-Steele across the 32 lanes of a warp; warps, and chunks of 1,024 block totals, in
// sequence; the prefix at a body = its run's prefix + the bodies of the run before it, in order.
// com = moment * (1/m) on the float-rounded sums as bench:181-186 (m > 1e-6f guard).
namespace {
struct D4 { double m, x, y, z; };
inline D4 d4_add(const D4& a, const D4& b) { return D4{a.m + b.m, a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D4 d4_sub(const D4& a, const D4& b) { return D4{a.m - b.m, a.x - b.x, a.y - b.y, a.z - b.z}; }
const D4 kZero{0.0, 0.0, 0.0, 0.0};

// vals: nw * 32 thread values -> exclusive prefix of every thread and the total (d4_block_exclusive in bh_tree.cu)
void block_exclusive(const D4* vals, int nw, D4* excl, D4& total) {
    std::vector<D4> incl(vals, vals + (size_t)nw * 32), tmp(32);
    for (int w = 0; w < nw; ++w) {
        D4* v = incl.data() + 32 * w;
        for (int o = 1; o < 32; o <<= 1) {
            std::copy(v, v + 32, tmp.begin());
            for (int i = o; i < 32; ++i) v[i] = d4_add(tmp[i], tmp[i - o]);
        }
    }
    D4 all = kZero;
    for (int w = 0; w < nw; ++w) {
        const D4 before = all;
        for (int l = 0; l < 32; ++l) excl[32 * w + l] = d4_add(before, l == 0 ? kZero : incl[32 * w + l - 1]);
        all = d4_add(all, incl[32 * w + 31]);
    }
    total = all;
}
}  // namespace

extern "C" void orc_tree_com(const float* posm, int64_t n64, const int32_t* meta, const int32_t* child,
                             int64_t M64, int32_t root, float* mom, float* com) {
    (void)child;
    const int M = (int)M64, n = (int)n64;
    if (M == 0 || root < 0) return;
    constexpr int CT = 256, CPT = 16, CB = CT * CPT;   // 4,096 bodies per block, 16 consecutive bodies per thread
    constexpr int DIRECT_MAX = 16;                      // cells up to this many bodies: direct sum
    const int nblocks = (n + CB - 1) / CB;
    std::vector<D4> runpre((size_t)n / CPT + 2, kZero), totals((size_t)nblocks, kZero), base((size_t)nblocks + 1, kZero);
    auto term = [&](int i) {
        const float* q = posm + 4 * (int64_t)i;
        const double m = (double)q[3];
        return D4{m, m * (double)q[0], m * (double)q[1], m * (double)q[2]};
    };
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nblocks; ++b) {
        D4 v[CT], pre[CT];
        for (int t = 0; t < CT; ++t) {
            const int i0 = b * CB + CPT * t;
            D4 sum = kZero;
            for (int k = 0; k < CPT; ++k)
                if (i0 + k < n) sum = d4_add(sum, term(i0 + k));
            v[t] = sum;
        }
        D4 total;
        block_exclusive(v, CT / 32, pre, total);
        for (int t = 0; t < CT; ++t) {
            const int i0 = b * CB + CPT * t;
            if (i0 <= n) runpre[i0 / CPT] = pre[t];
        }
        totals[b] = total;
    }
    if (nblocks * CB == n) runpre[n / CPT] = kZero;
    D4 carry = kZero;
    for (int c0 = 0; c0 < nblocks; c0 += 1024) {
        D4 v[1024], pre[1024], total;
        for (int t = 0; t < 1024; ++t) v[t] = c0 + t < nblocks ? totals[c0 + t] : kZero;
        block_exclusive(v, 32, pre, total);
        for (int t = 0; t < 1024 && c0 + t < nblocks; ++t) base[c0 + t] = d4_add(carry, pre[t]);
        carry = d4_add(carry, total);
    }
    base[nblocks] = carry;
    auto prefix_at = [&](int i) {   // d4_prefix_at in bh_tree.cu
        D4 v = runpre[i / CPT];
        for (int k = (i / CPT) * CPT; k < i; ++k) v = d4_add(v, term(k));
        return v;
    };
#pragma omp parallel for schedule(static)
    for (int c = 0; c < M; ++c) {
        const int32_t* mt = meta + 4 * (int64_t)c;
        const int first = mt[0], end = mt[0] + mt[1];
        D4 s = kZero;
        if (end - first <= DIRECT_MAX) {
            for (int k = first; k < end; ++k) s = d4_add(s, term(k));
        } else {
            s = d4_add(d4_sub(base[end / CB], base[first / CB]), d4_sub(prefix_at(end), prefix_at(first)));
        }
        const float m = (float)s.m, sx = (float)s.x, sy = (float)s.y, sz = (float)s.z;
        float* o = mom + 4 * (int64_t)c;
        o[0] = sx; o[1] = sy; o[2] = sz; o[3] = m;
        const float inv = (m > 1e-6f) ? 1.0f / m : 0.0f;
        float* cc = com + 4 * (int64_t)c;
        cc[0] = sx * inv; cc[1] = sy * inv; cc[2] = sz * inv; cc[3] = m;
    }
}

// Traceless quadrupole of every cell about its centre of mass, Q_ij = sum m (3 x_i x_j - |x|^2 delta_ij), x = s - c
// (bh_params.flags & BH_FLAG_QUADRUPOLE; the reference has monopoles only, bench:205-213 — this is the accuracy knob of
// SURVEY §8f N4).  Straight double sums over the cell's body range: deliberately NOT the kernels' prefix-sum route, so
// the two check each other.  quad: 6 floats per cell {xx, xy, xz, yy, yz, zz}.
void orc_tree_quad(const float* posm, const int32_t* meta, int64_t M64, float* quad) {
    const int M = (int)M64;
#pragma omp parallel for schedule(dynamic, 64)
    for (int c = 0; c < M; ++c) {
        const int first = meta[4 * (int64_t)c], count = meta[4 * (int64_t)c + 1];
        double m = 0, cx = 0, cy = 0, cz = 0;
        for (int i = first; i < first + count; ++i) {
            const float* q = posm + 4 * (int64_t)i;
            m += q[3]; cx += (double)q[3] * q[0]; cy += (double)q[3] * q[1]; cz += (double)q[3] * q[2];
        }
        if (m > 0) { cx /= m; cy /= m; cz /= m; }
        double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
        for (int i = first; i < first + count; ++i) {
            const float* q = posm + 4 * (int64_t)i;
            const double x = q[0] - cx, y = q[1] - cy, z = q[2] - cz, w = q[3];
            xx += w * x * x; xy += w * x * y; xz += w * x * z; yy += w * y * y; yz += w * y * z; zz += w * z * z;
        }
        const double tr = xx + yy + zz;
        float* o = quad + 6 * (int64_t)c;
        o[0] = (float)(3 * xx - tr); o[1] = (float)(3 * xy); o[2] = (float)(3 * xz);
        o[3] = (float)(3 * yy - tr); o[4] = (float)(3 * yz); o[5] = (float)(3 * zz - tr);
    }
}

// Force with the GROUP acceptance test (engine: bh_force.cu).
// Group = `group` Morton-consecutive bodies; box = exact AABB of their positions;
// d = distance from the cell's centre of mass to that box.  A cell is accepted for the whole
// group iff  w_L^2 < theta^2 * (d^2 + soft)  — bench:207-208's width/dist < THETA with
// dist^2 = d^2 + SOFTENING, squared, and d taken at the closest point of the group, so every
// body of the group would also accept it under the reference's per-body test.
// The interaction itself is bench:205-213 (f = G m / dist^3 with dist = sqrt(d2+soft)),
// evaluated in float, accumulated in double.  acc: float4 per body (sorted order).
// counts[0] = accepted (body,cell) pairs, counts[1] = direct (body,body) pairs.
void orc_force_groups(const float* posm, int64_t n64, const float* bounds,
                      const int32_t* meta, const int32_t* child, const float* com,
                      int64_t M64, int32_t root, const int32_t* gstart, int ngroups, float theta, float soft, float G,
                      float* acc, int64_t* counts, int32_t* group_entries /*nullable: list length per group*/);
// same with quadrupole corrections of the accepted cells (quad: 6 floats per cell, nullptr = monopoles only)
void orc_force_groups_quad(const float* posm, int64_t n64, const float* bounds,
                           const int32_t* meta, const int32_t* child, const float* com, const float* quad,
                           int64_t M64, int32_t root, const int32_t* gstart, int ngroups, float theta, float soft, float G,
                           float* acc, int64_t* counts, int32_t* group_entries);

void orc_force_group(const float* posm, int64_t n64, const float* bounds,
                     const int32_t* meta, const int32_t* child, const float* com,
                     int64_t M64, int32_t root, int group, float theta, float soft, float G,
                     float* acc, int64_t* counts, int32_t* group_entries) {
    const int n = (int)n64;
    const int ngroups = (n + group - 1) / group;
    std::vector<int32_t> gstart(ngroups + 1);
    for (int g = 0; g <= ngroups; ++g) gstart[g] = std::min(n, g * group);
    orc_force_groups(posm, n64, bounds, meta, child, com, M64, root, gstart.data(), ngroups, theta, soft, G, acc,
                     counts, group_entries);
}

// Same, for explicit groups: group g = sorted bodies [gstart[g], gstart[g+1]).
void orc_force_groups(const float* posm, int64_t n64, const float* bounds,
                      const int32_t* meta, const int32_t* child, const float* com,
                      int64_t M64, int32_t root, const int32_t* gstart, int ngroups, float theta, float soft, float G,
                      float* acc, int64_t* counts, int32_t* group_entries) {
    orc_force_groups_quad(posm, n64, bounds, meta, child, com, nullptr, M64, root, gstart, ngroups, theta, soft, G, acc, counts,
                          group_entries);
}

void orc_force_groups_quad(const float* posm, int64_t n64, const float* bounds,
                           const int32_t* meta, const int32_t* child, const float* com, const float* quad,
                           int64_t M64, int32_t root, const int32_t* gstart, int ngroups, float theta, float soft, float G,
                           float* acc, int64_t* counts, int32_t* group_entries) {
    const int n = (int)n64; (void)n;
    const float root_w = fmaxf(bounds[3] - bounds[0], 1.0f);   // the key grid's size, clamped like bench:52
    const float theta2 = theta * theta;
    float w2[MAX_LEVELS_ANY + 1];
    for (int L = 0; L <= MAX_LEVELS_ANY; ++L) w2[L] = w2_of_level(root_w, L);
    int64_t ncell = 0, nbody = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ncell, nbody)
    for (int g = 0; g < ngroups; ++g) {
        const int b0 = gstart[g], b1 = gstart[g + 1], nb = b1 - b0;
        if (nb <= 0) continue;
        float lo[3] = {posm[4 * (int64_t)b0], posm[4 * (int64_t)b0 + 1], posm[4 * (int64_t)b0 + 2]};
        float hi[3] = {lo[0], lo[1], lo[2]};
        for (int i = b0 + 1; i < b1; ++i)
            for (int a = 0; a < 3; ++a) {
                lo[a] = fminf(lo[a], posm[4 * (int64_t)i + a]);
                hi[a] = fmaxf(hi[a], posm[4 * (int64_t)i + a]);
            }
        float ctr[3], half[3];
        for (int a = 0; a < 3; ++a) { ctr[a] = (lo[a] + hi[a]) * 0.5f; half[a] = (hi[a] - lo[a]) * 0.5f; }
        std::vector<double> f(3 * (size_t)nb, 0.0);
        auto interact = [&](float sx, float sy, float sz, float sm) {
            for (int i = 0; i < nb; ++i) {
                const float* p = posm + 4 * (int64_t)(b0 + i);
                float dx = sx - p[0], dy = sy - p[1], dz = sz - p[2];
                float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                float dist = sqrtf(d2 + soft);
                float s = G * sm / (dist * dist * dist);
                f[3 * i] += (double)(s * dx); f[3 * i + 1] += (double)(s * dy); f[3 * i + 2] += (double)(s * dz);
            }
        };
        // an accepted cell with quadrupole Q: a = G [ M d / R^3 - Q d / R^5 + 5/2 (d.Q.d) d / R^7 ], d = c - p, R^2 = d^2 + soft
        auto interact_quad = [&](const float* cm, const float* Q) {
            for (int i = 0; i < nb; ++i) {
                const float* p = posm + 4 * (int64_t)(b0 + i);
                float dx = cm[0] - p[0], dy = cm[1] - p[1], dz = cm[2] - p[2];
                float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                float rinv = 1.0f / sqrtf(d2 + soft);
                float r2 = rinv * rinv, r3 = rinv * r2, r5 = r3 * r2, r7 = r5 * r2;
                float qx = Q[0] * dx + Q[1] * dy + Q[2] * dz, qy = Q[1] * dx + Q[3] * dy + Q[4] * dz, qz = Q[2] * dx + Q[4] * dy + Q[5] * dz;
                float dqd = dx * qx + dy * qy + dz * qz;
                float s = cm[3] * r3 + 2.5f * dqd * r7;
                f[3 * i] += (double)(G * (s * dx - r5 * qx)); f[3 * i + 1] += (double)(G * (s * dy - r5 * qy));
                f[3 * i + 2] += (double)(G * (s * dz - r5 * qz));
            }
        };
        std::vector<int> stack;
        int64_t entries = 0;
        if (root >= 0) stack.push_back(root);
        while (!stack.empty()) {
            int c = stack.back();
            stack.pop_back();
            const int32_t* mt = meta + 4 * (int64_t)c;
            const float* cm = com + 4 * (int64_t)c;
            int L = mt[2] & 0xFF;
            bool bucket = (mt[2] >> 8) & 1;
            float dx = fmaxf(0.0f, fabsf(cm[0] - ctr[0]) - half[0]);
            float dy = fmaxf(0.0f, fabsf(cm[1] - ctr[1]) - half[1]);
            float dz = fmaxf(0.0f, fabsf(cm[2] - ctr[2]) - half[2]);
            float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (w2[L] < theta2 * (d2 + soft)) {
                if (quad) interact_quad(cm, quad + 6 * (int64_t)c);
                else interact(cm[0], cm[1], cm[2], cm[3]);
                ncell += nb; ++entries;
            } else if (bucket) {
                for (int j = mt[0]; j < mt[0] + mt[1]; ++j) {
                    const float* q = posm + 4 * (int64_t)j;
                    interact(q[0], q[1], q[2], q[3]);
                }
                nbody += (int64_t)nb * mt[1]; entries += mt[1];
            } else {
                const int32_t* ch = child + 8 * (int64_t)c;
                for (int q = 0; q < 8; ++q) {
                    int e = ch[q];
                    if (e == CHILD_EMPTY) continue;
                    if (e < 0) {
                        const float* s = posm + 4 * (int64_t)(e & 0x7FFFFFFF);
                        interact(s[0], s[1], s[2], s[3]);
                        nbody += nb; ++entries;
                    } else stack.push_back(e);
                }
            }
        }
        for (int i = 0; i < nb; ++i) {
            float* o = acc + 4 * (int64_t)(b0 + i);
            o[0] = (float)f[3 * i]; o[1] = (float)f[3 * i + 1]; o[2] = (float)f[3 * i + 2]; o[3] = 0.f;
        }
        if (group_entries) group_entries[g] = (int32_t)std::min<int64_t>(entries, 0x7FFFFFFF);
    }
    if (counts) { counts[0] = ncell; counts[1] = nbody; }
}

// Traversal groups (engine: bh_force.cu "group splitting").  Bodies are taken in chunks of `chunk`
// Morton-consecutive slots; a chunk that straddles a high-level cell boundary would get a huge
// bounding box, so it is split recursively: cut the range at its coarsest key boundary (first
// pair with the fewest shared digits) and keep the cut iff
//     ext(A) + ext(B) < alpha * ext(A u B),   ext(box) = (hi-lo).x + (hi-lo).y + (hi-lo).z.
// Returns the number of groups; gstart gets ngroups+1 entries.
static float range_ext(const float* posm, int a, int b) {
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) lo[k] = hi[k] = posm[4 * (int64_t)a + k];
    for (int i = a + 1; i < b; ++i)
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], posm[4 * (int64_t)i + k]);
            hi[k] = fmaxf(hi[k], posm[4 * (int64_t)i + k]);
        }
    return ((hi[0] - lo[0]) + (hi[1] - lo[1])) + (hi[2] - lo[2]);
}

int orc_make_groups64(const float* posm, const uint64_t* keys, int64_t n64, int levels, int chunk, float alpha,
                      int32_t* gstart);

int orc_make_groups(const float* posm, const uint32_t* keys, int64_t n64, int chunk, float alpha,
                    int32_t* gstart) {
    std::vector<uint64_t> k64((size_t)std::max<int64_t>(n64, 1));
    for (int64_t i = 0; i < n64; ++i) k64[i] = keys[i];
    return orc_make_groups64(posm, k64.data(), n64, 10, chunk, alpha, gstart);
}

int orc_make_groups64(const float* posm, const uint64_t* keys, int64_t n64, int levels, int chunk, float alpha,
                      int32_t* gstart) {
    const int n = (int)n64;
    int ng = 0;
    for (int c0 = 0; c0 < n; c0 += chunk) {
        const int c1 = std::min(n, c0 + chunk);
        std::vector<std::pair<int, int>> todo;   // LIFO of ranges, processed left to right
        todo.push_back({c0, c1});
        while (!todo.empty()) {
            auto [a, b] = todo.back();
            todo.pop_back();
            bool split = false;
            if (b - a >= 2) {
                int best = a, bestlv = 99;
                for (int j = a; j < b - 1; ++j) {
                    int lv = shared_digits(keys[j], keys[j + 1], levels);
                    if (lv < bestlv) { bestlv = lv; best = j; }
                }
                const int k = best + 1;
                if (bestlv < levels) {
                    float eA = range_ext(posm, a, k), eB = range_ext(posm, k, b), eAB = range_ext(posm, a, b);
                    if (eA + eB < alpha * eAB) {
                        split = true;
                        todo.push_back({k, b});
                        todo.push_back({a, k});
                    }
                }
            }
            if (!split) gstart[ng++] = a;
        }
    }
    gstart[ng] = n;
    return ng;
}

// Same tree, the reference's PER-BODY test exactly as bench:205-208 (accuracy study only).
void orc_force_body(const float* posm, int64_t n64, const float* bounds,
                    const int32_t* meta, const int32_t* child, const float* com,
                    int64_t /*M*/, int32_t root, float theta, float soft, float G,
                    float* acc, int64_t* counts) {
    const int n = (int)n64;
    const float root_w = fmaxf(bounds[3] - bounds[0], 1.0f);   // the key grid's size, clamped like bench:52
    int64_t ncell = 0, nbody = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : ncell, nbody)
    for (int i = 0; i < n; ++i) {
        const float* p = posm + 4 * (int64_t)i;
        double fx = 0, fy = 0, fz = 0;
        auto interact = [&](const float* s) {
            float dx = s[0] - p[0], dy = s[1] - p[1], dz = s[2] - p[2];
            float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            float dist = sqrtf(d2 + soft);
            float f = G * s[3] / (dist * dist * dist);
            fx += (double)(f * dx); fy += (double)(f * dy); fz += (double)(f * dz);
        };
        std::vector<int> stack;
        if (root >= 0) stack.push_back(root);
        while (!stack.empty()) {
            int c = stack.back();
            stack.pop_back();
            const int32_t* mt = meta + 4 * (int64_t)c;
            const float* cm = com + 4 * (int64_t)c;
            float width = ldexpf(root_w, -(mt[2] & 0xFF));
            bool bucket = (mt[2] >> 8) & 1;
            float dx = cm[0] - p[0], dy = cm[1] - p[1], dz = cm[2] - p[2];
            float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            float dist = sqrtf(d2 + soft);
            if (width / dist < theta) { interact(cm); ++ncell; }
            else if (bucket) {
                for (int j = mt[0]; j < mt[0] + mt[1]; ++j) interact(posm + 4 * (int64_t)j);
                nbody += mt[1];
            } else {
                const int32_t* ch = child + 8 * (int64_t)c;
                for (int q = 0; q < 8; ++q) {
                    int e = ch[q];
                    if (e == CHILD_EMPTY) continue;
                    if (e < 0) { interact(posm + 4 * (int64_t)(e & 0x7FFFFFFF)); ++nbody; }
                    else stack.push_back(e);
                }
            }
        }
        float* o = acc + 4 * (int64_t)i;
        o[0] = (float)fx; o[1] = (float)fy; o[2] = (float)fz; o[3] = 0.f;
    }
    if (counts) { counts[0] = ncell; counts[1] = nbody; }
}

// O(N*k) direct sum in double for the sampled bodies (positions in SoA or float4 via stride).
// Force law README.md:82-84 / bench:205-213: a_i = G sum_j m_j d_ij / (|d_ij|^2 + soft)^(3/2).
void orc_direct_sum(const float* posm, int64_t n, const int32_t* sample, int k,
                    float soft, float G, double* acc) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int s = 0; s < k; ++s) {
        const float* p = posm + 4 * (int64_t)sample[s];
        double fx = 0, fy = 0, fz = 0;
        for (int64_t j = 0; j < n; ++j) {
            const float* q = posm + 4 * j;
            double dx = (double)q[0] - p[0], dy = (double)q[1] - p[1], dz = (double)q[2] - p[2];
            double r2 = dx * dx + dy * dy + dz * dz + (double)soft;
            double inv = 1.0 / (r2 * std::sqrt(r2));
            double f = (double)G * q[3] * inv;
            fx += f * dx; fy += f * dy; fz += f * dz;
        }
        acc[3 * s] = fx; acc[3 * s + 1] = fy; acc[3 * s + 2] = fz;
    }
}

// Kinetic energy and softened pair potential -G sum_{i<j} m_i m_j / sqrt(r^2+soft), double.
void orc_energy(const float* posm, const float* vel, int64_t n, float soft, float G,
                double* kinetic, double* potential) {
    double ke = 0, pe = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ke, pe)
    for (int64_t i = 0; i < n; ++i) {
        const float* p = posm + 4 * i;
        const float* v = vel + 4 * i;
        ke += 0.5 * p[3] * ((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
        double s = 0;
        for (int64_t j = i + 1; j < n; ++j) {
            const float* q = posm + 4 * j;
            double dx = (double)q[0] - p[0], dy = (double)q[1] - p[1], dz = (double)q[2] - p[2];
            s += q[3] / std::sqrt(dx * dx + dy * dy + dz * dz + (double)soft);
        }
        pe -= (double)G * p[3] * s;
    }
    *kinetic = ke; *potential = pe;
}

// Full shipped-algorithm step on the CPU, state in the engine's internal layout:
// posm/vel float4 in Morton order of the previous step, ids = original body id per slot.
// One call = bounds, keys, stable sort, reorder, tree, com, group force, kick-drift-clamp.
// Scratch outputs (any may be NULL) expose the intermediates of the LAST step.
int orc_engine_step(float* posm, float* vel, int32_t* ids, int64_t n64, int nsteps,
                    float G, float theta, float dt, float soft, float max_speed, int group, float split_alpha,
                    float* acc_out, uint32_t* keys_out, int32_t* perm_out, float* bounds_out,
                    int64_t* counts_out, double* phase_ms, int64_t slice_first, int64_t slice_count,
                    int key_bits, uint32_t* keys_lo_out) {
    // slice_count < 0: the whole range.  Otherwise only sorted slots [slice_first, slice_first+slice_count)
    // are traversed and integrated (multi-GPU Morton slices, engine: bh_set_slice); the other slots of
    // posm/vel/ids are left holding this step's sorted pre-drift state, to be overwritten by the
    // owners' all-gather.  slice_first must be a multiple of `group`.
    if (n64 <= 0 || n64 > (1 << 30)) return -1;
    const int n = (int)n64;
    std::vector<float> sx(n), sy(n), sz(n), p2(4 * (size_t)n), v2(4 * (size_t)n), acc(4 * (size_t)n);
    if (key_bits != 30 && key_bits != 60) return -2;
    const int levels = key_bits / 3;
    std::vector<uint32_t> keys(n), keys_lo(n, 0u);
    std::vector<uint64_t> k64(n);
    std::vector<int32_t> perm(n), id2(n), meta, child;
    std::vector<float> mom, com;
    double ph[6] = {0, 0, 0, 0, 0, 0};
    for (int s = 0; s < nsteps; ++s) {
        double t0 = now_ms();
        for (int i = 0; i < n; ++i) { sx[i] = posm[4 * (size_t)i]; sy[i] = posm[4 * (size_t)i + 1]; sz[i] = posm[4 * (size_t)i + 2]; }
        float b[6];
        orc_bounds(sx.data(), sy.data(), sz.data(), n, b);
        if (key_bits == 30) {
            orc_morton_keys(sx.data(), sy.data(), sz.data(), n, b, keys.data(), perm.data());
            for (int i = 0; i < n; ++i) k64[i] = keys[i];
        } else {
            orc_morton_keys60(sx.data(), sy.data(), sz.data(), n, b, keys.data(), keys_lo.data());
            for (int i = 0; i < n; ++i) { k64[i] = ((uint64_t)keys[i] << 30) | keys_lo[i]; perm[i] = i; }
        }
        double t1 = now_ms();
        orc_stable_sort64(k64.data(), perm.data(), n);
        for (int i = 0; i < n; ++i) {
            keys[i] = key_bits == 30 ? (uint32_t)k64[i] : (uint32_t)(k64[i] >> 30);
            keys_lo[i] = key_bits == 30 ? 0u : (uint32_t)(k64[i] & 0x3FFFFFFFu);
        }
        for (int i = 0; i < n; ++i) {
            std::memcpy(&p2[4 * (size_t)i], &posm[4 * (size_t)perm[i]], 16);
            std::memcpy(&v2[4 * (size_t)i], &vel[4 * (size_t)perm[i]], 16);
            id2[i] = ids[perm[i]];
        }
        std::memcpy(posm, p2.data(), 16 * (size_t)n);
        std::memcpy(vel, v2.data(), 16 * (size_t)n);
        std::memcpy(ids, id2.data(), 4 * (size_t)n);
        double t2 = now_ms();
        meta.assign(4 * (size_t)n, 0); child.assign(8 * (size_t)n, 0);
        int32_t root = -1;
        int M = orc_tree_build64(k64.data(), n, levels, meta.data(), child.data(), n, &root);
        double t3 = now_ms();
        mom.assign(4 * (size_t)std::max(M, 1), 0.f); com.assign(4 * (size_t)std::max(M, 1), 0.f);
        orc_tree_com(posm, n, meta.data(), child.data(), M, root, mom.data(), com.data());
        double t4 = now_ms();
        int64_t counts[2] = {0, 0};
        std::fill(acc.begin(), acc.end(), 0.f);
        std::vector<int32_t> gstart((size_t)n + 1);
        int ng = orc_make_groups64(posm, k64.data(), n, levels, group, split_alpha, gstart.data());
        const int s0 = slice_count < 0 ? 0 : (int)slice_first;
        const int s1 = slice_count < 0 ? n : (int)(slice_first + slice_count);
        int g0 = 0, g1 = ng;
        while (g0 < ng && gstart[g0] < s0) ++g0;
        while (g1 > g0 && gstart[g1 - 1] >= s1) --g1;
        orc_force_groups(posm, n, b, meta.data(), child.data(), com.data(), M, root, gstart.data() + g0, g1 - g0,
                         theta, soft, G, acc.data(), counts, nullptr);
        double t5 = now_ms();
        if (s == nsteps - 1) {
            if (acc_out) std::memcpy(acc_out, acc.data(), 16 * (size_t)n);
            if (keys_out) std::memcpy(keys_out, keys.data(), 4 * (size_t)n);
            if (keys_lo_out) std::memcpy(keys_lo_out, keys_lo.data(), 4 * (size_t)n);
            if (perm_out) std::memcpy(perm_out, perm.data(), 4 * (size_t)n);
            if (bounds_out) std::memcpy(bounds_out, b, sizeof(b));
            if (counts_out) { counts_out[0] = counts[0]; counts_out[1] = counts[1]; counts_out[2] = M; }
        }
        const float vmax2 = max_speed * max_speed;
#pragma omp parallel for schedule(static)
        for (int i = s0; i < s1; ++i) {  // bench:227-249, FMA pattern of SURVEY R12
            float* p = posm + 4 * (size_t)i; float* v = vel + 4 * (size_t)i; const float* a = &acc[4 * (size_t)i];
            float x = fmaf(a[0], dt, v[0]), y = fmaf(a[1], dt, v[1]), z = fmaf(a[2], dt, v[2]);
            float q = fmaf(z, z, fmaf(x, x, y * y));
            if (q > vmax2) { float sc = max_speed / sqrtf(q); x *= sc; y *= sc; z *= sc; }
            v[0] = x; v[1] = y; v[2] = z;
            p[0] = fmaf(x, dt, p[0]); p[1] = fmaf(y, dt, p[1]); p[2] = fmaf(z, dt, p[2]);
        }
        double t6 = now_ms();
        ph[0] += t1 - t0; ph[1] += t2 - t1; ph[2] += t3 - t2; ph[3] += t4 - t3; ph[4] += t5 - t4; ph[5] += t6 - t5;
    }
    if (phase_ms) std::memcpy(phase_ms, ph, sizeof(ph));
    return 0;
}

}  // extern "C"
