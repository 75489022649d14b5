// bh_engine.cu — context, step driver and the C ABI (include/bh.h).
//
// The step driver replaces simulationStep()  nbody_v5_bench.cu:255-283:
//   reference: bbox<<<1,1>>> -> keys -> thrust sort (alloc + sync) -> memset 2N nodes -> N/1024
//              insert launches -> blocking 4-byte D2H -> COM atomics -> force -> integrate
//   here:      bounds -> keys -> onesweep sort -> Morton reorder -> tree emission -> COM ->
//              group traversal -> kick-drift-clamp; ~20 launches, no host sync, no allocation,
//              replayed as one CUDA graph per step.
//
// State layout in HBM (DESIGN.md §"Data layout"):
//   current state   posm[n] float4 {x,y,z,m}, vel[n] float4 {vx,vy,vz,0}, ids[n] int32
//                   (Morton order of the last sort; ids = original body id of each slot)
//   sorted scratch  posm_s, vel_s, ids_s — this step's Morton order, read by force + integrate
//   acc[n] float4   accelerations, slot i pairs with posm_s[i]
#include "bh_common.cuh"

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

struct bh_ctx {
    int device = 0;
    int num_sms = BH_NUM_SMS_FALLBACK;
    bh_params prm{};
    int64_t n_max = 0, n_alloc = 0, n = 0;
    int64_t steps = 0;
    bool have_state = false, have_sorted = false;
    // sc->bbox_enc holds the min/max of the CURRENT positions (left there by the integrator of a full step, or by
    // bh_bounds_enc_launch); the keys phase consumes it.  False => the positions are reduced before the keys phase.
    bool bbox_fresh = false;
    bool external_writers = false;   // bh_state_ptrs handed the position array out: never trust bbox_fresh again
    // slice (multi-GPU Morton ranges); default = everything
    int rank = 0, world = 1;
    int64_t slice_first = 0, slice_count = 0;

    float4 *posm = nullptr, *vel = nullptr, *posm_s = nullptr, *vel_s = nullptr, *acc = nullptr;
    int32_t *ids = nullptr, *ids_s = nullptr;
    uint32_t *keys0 = nullptr, *keys1 = nullptr, *vals0 = nullptr, *vals1 = nullptr;
    uint32_t* perm = nullptr;            // where the step's sort leaves the permutation (depends on key width / pass parity)
    // 30-bit sort: pass 0 reads `keys_unsorted` and the passes ping-pong so that the sorted keys land in keys0 and the
    // permutation in vals0 — an odd pass count starts from keys1, an even one from keys0 (dead after pass 0)
    uint32_t* keys_unsorted = nullptr;
    bool keys_sorted = false;            // what BH_DBG_KEYS shows: the unsorted keys after the keys phase, the sorted ones later
    // key_bits = 60 only: unsorted low words, a sort ping-pong buffer, the sorted 60-bit keys
    uint32_t *klo = nullptr, *kaux = nullptr;
    uint64_t* keys64 = nullptr;
    int levels = BH_MAX_LEVEL;           // 3-bit digits per key: 10 (30-bit reference key) or 20
    void* sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    int2* pair_info = nullptr;
    int2* pair_aux = nullptr;    // leaders only: {last body of the cell, pair that names the parent}
    int32_t *pair_scan = nullptr, *tile_sums = nullptr;
    int4* cell_meta = nullptr;
    int32_t* cell_child = nullptr;
    float4* cell_com = nullptr;
    void* com_scratch = nullptr;   // prefix sums of the centre-of-mass pass (bh_com_scratch_bytes)
    // BH_FLAG_QUADRUPOLE only: second-moment prefix sums, per-cell quadrupoles, and their copy in the dense child lines
    bool quad = false;
    void* quad_scratch = nullptr;
    float4 *cell_quad = nullptr, *kid_quad = nullptr;
    float4* kid_src = nullptr;   // 8 per cell
    uint8_t* kid_lv = nullptr;   // 8 per cell (digit-indexed)
    uint2* kid_info = nullptr;   // 8 per cell (dense, pairs with kid_src)
    uint32_t* heavy_list = nullptr;   // 2 * max_chunks
    uint32_t* heavy_flag = nullptr;   // 2 * max_chunks epoch tags
    int64_t max_chunks = 0;
    BhDevScalars* sc = nullptr;
    float* stage = nullptr;  // 10 * n_alloc floats, lazily allocated for the host-pointer entry points
    double* d_scratch = nullptr;

    // locally-essential-tree mode
    bool fixed_bounds_set = false;
    float fixed_bounds[6] = {0, 0, 0, 0, 0, 0};
    float* let_boxes = nullptr;          // BH_LET_MAX_PEERS x (BH_LET_MAX_BOXES + 1) x 6: boxes, then one hull per peer
    unsigned int* let_counts = nullptr;  // BH_LET_MAX_PEERS + 2 queue counters
    int2* let_queue = nullptr;           // 2 x let_qcap
    long long let_qcap = 0;

    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_n = -1, graph_first = -1, graph_count = -1;
    // the step in three parts (bh_step_part): 0 = cube, keys, radix sort (positions only); 1 = reorder positions +
    // masses, tree, centre of mass, traversal; 2 = reorder velocities + ids, update
    cudaGraphExec_t half_exec[3] = {nullptr, nullptr, nullptr};
    bool no_fused_update = false;   // BH_NO_FUSED_UPDATE=1 in the environment at creation: A/B switch for measurements
    bool may_have_ghosts = false;   // ids < 0 possible (bh_import_state / checkpoint): the traversal must read the ids
    int64_t graph_half_n = -1, graph_half_first = -1, graph_half_count = -1;
    cudaStream_t own_stream = nullptr;
    cudaStream_t aux_stream = nullptr;       // second branch of the step: the centre-of-mass prefix sums run beside the tree build
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t copy_stream = nullptr;      // host uploads of bh_step_host (its events must not be the ones a capture re-records)
    cudaEvent_t ev_h2d_pos = nullptr, ev_h2d_mass = nullptr, ev_h2d_rest = nullptr;
    cudaEvent_t ev[BH_PHASE_COUNT + 1] = {};
    float phase_ms[BH_PHASE_COUNT] = {};
};

namespace {

template <class T>
cudaError_t dev_alloc(T** p, size_t count) {
    return cudaMalloc((void**)p, count * sizeof(T) + 256);
}

void free_all(bh_ctx* c) {
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    for (auto& g : c->half_exec) if (g) cudaGraphExecDestroy(g);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_h2d_pos) cudaEventDestroy(c->ev_h2d_pos);
    if (c->ev_h2d_rest) cudaEventDestroy(c->ev_h2d_rest);
    if (c->ev_h2d_mass) cudaEventDestroy(c->ev_h2d_mass);
    void* ptrs[] = {c->posm, c->vel, c->posm_s, c->vel_s, c->acc, c->ids, c->ids_s, c->keys0, c->keys1, c->vals0,
                    c->vals1, c->sort_tmp, c->pair_info, c->pair_aux, c->pair_scan, c->tile_sums, c->cell_meta, c->cell_child,
                    c->com_scratch, c->cell_com, c->sc, c->stage, c->d_scratch, c->heavy_list, c->heavy_flag, c->kid_src, c->kid_lv, c->kid_info, c->quad_scratch, c->cell_quad, c->kid_quad, c->let_boxes, c->let_counts, c->let_queue, c->klo, c->kaux, c->keys64};
    for (void* p : ptrs) if (p) cudaFree(p);
}

void default_slice(bh_ctx* c) {
    const int64_t groups = (c->n + BH_GROUP - 1) / BH_GROUP;
    const int64_t per = (groups + c->world - 1) / c->world;
    int64_t first = (int64_t)c->rank * per * BH_GROUP;
    int64_t last = (int64_t)(c->rank + 1) * per * BH_GROUP;
    if (first > c->n) first = c->n;
    if (last > c->n) last = c->n;
    c->slice_first = first;
    c->slice_count = last - first;
}

// ---- the phases ---------------------------------------------------------------------------
// Not part of a captured graph (whether it is needed changes from step to step): reduce the positions to the
// min/max images unless the integrator of the previous full step already left them behind.
int pre_keys(bh_ctx* c, cudaStream_t st) {
    if (c->fixed_bounds_set || c->bbox_fresh) return 0;
    int e = bh_bounds_enc_launch(c->posm, c->n, c->sc, st);
    if (e) return e;
    c->bbox_fresh = true;
    return 0;
}

// after the update phase: the images describe the new positions iff every body was integrated here
void post_update(bh_ctx* c) { c->bbox_fresh = !c->external_writers && c->slice_first == 0 && c->slice_count == c->n; }

int phase_keys(bh_ctx* c, cudaStream_t st) {
    int e = 0;
    if (c->fixed_bounds_set) {   // LET mode: the cube was agreed with the other ranks
        BH_CUDA_TRY(cudaMemcpyAsync((char*)c->sc + offsetof(BhDevScalars, bounds), c->fixed_bounds, sizeof(c->fixed_bounds),
                                    cudaMemcpyHostToDevice, st));
    } else {
        e = bh_bounds_finish_launch(c->sc, st);   // consumes the min/max images (pre_keys / the last integrator)
    }
    if (e) return e;
    if (c->levels == 20) return bh_keys60_launch(c->posm, c->n, c->sc, c->keys0, c->klo, st);
    return bh_keys_launch(c->posm, c->n, c->sc, c->keys_unsorted, st);
}

// key_bits = 60: LSD over the two 30-bit words with the same u32 sorter — low word first (values = iota), then the
// high word carried along in that order (stable), then the sorted 60-bit keys are assembled.  Buffer roles depend
// on the parity of the pass count (the result of a sort lands in its q pair iff the count is even).
int sort_keys60(bh_ctx* c, cudaStream_t st) {
    unsigned int* err = (unsigned int*)((char*)c->sc + offsetof(BhDevScalars, err));
    const bool even = (bh_sort_passes(BH_KEY_BITS) & 1) == 0;
    int in_q = 0, e = 0;
    // 1) order by low word -> vals0
    if (even) e = bh_sort_pairs_launch(c->klo, nullptr, c->keys1, c->vals1, c->kaux, c->vals0, c->n, 0, BH_KEY_BITS, c->sort_tmp, true, err, &in_q, st);
    else e = bh_sort_pairs_launch(c->klo, nullptr, c->kaux, c->vals0, c->keys1, c->vals1, c->n, 0, BH_KEY_BITS, c->sort_tmp, true, err, &in_q, st);
    if (e) return e;
    if ((in_q != 0) != even) return BH_E_UNSUPPORTED;
    e = bh_gather_u32_launch(c->keys0, c->vals0, c->keys1, c->n, st);               // high words in that order
    if (e) return e;
    // 2) stable sort by high word, carrying that order -> sorted high words in keys0, final order in c->perm
    if (even) e = bh_sort_pairs_launch(c->keys1, c->vals0, c->kaux, c->vals1, c->keys0, c->vals0, c->n, 0, BH_KEY_BITS, c->sort_tmp, false, err, &in_q, st);
    else e = bh_sort_pairs_launch(c->keys1, c->vals0, c->keys0, c->vals1, c->kaux, c->vals0, c->n, 0, BH_KEY_BITS, c->sort_tmp, false, err, &in_q, st);
    if (e) return e;
    if ((in_q != 0) != even) return BH_E_UNSUPPORTED;
    return bh_combine_keys_launch(c->keys0, c->klo, c->perm, c->keys64, c->n, st);
}

// the 30-bit sort of the unsorted keys; with_bodies: the last pass also writes posm_s / vel_s / ids_s
int sort_keys30(bh_ctx* c, bool with_bodies, cudaStream_t st) {
    int in_q = 0, e = 0;
    unsigned int* err = (unsigned int*)((char*)c->sc + offsetof(BhDevScalars, err));
    const float4* pin = with_bodies ? c->posm : nullptr;
    if (c->keys_unsorted == c->keys0)   // even pass count: keys0 -> keys1 -> keys0 ...
        e = bh_sort_pairs_move_launch(c->keys0, nullptr, c->keys1, c->vals1, c->keys0, c->vals0, c->n, 0, BH_KEY_BITS, c->sort_tmp,
                                      true, err, &in_q, pin, c->vel, c->ids, c->posm_s, c->vel_s, c->ids_s, st);
    else                                // odd: keys1 -> keys0 -> keys1 -> keys0
        e = bh_sort_pairs_move_launch(c->keys1, nullptr, c->keys0, c->vals0, c->keys1, c->vals1, c->n, 0, BH_KEY_BITS, c->sort_tmp,
                                      true, err, &in_q, pin, c->vel, c->ids, c->posm_s, c->vel_s, c->ids_s, st);
    if (e) return e;
    return ((in_q != 0) == (c->keys_unsorted == c->keys0)) ? 0 : BH_E_UNSUPPORTED;   // sorted keys in keys0, permutation in vals0
}

int sort_keys_only(bh_ctx* c, cudaStream_t st) {
    if (c->levels == 20) return sort_keys60(c, st);
    return sort_keys30(c, false, st);
}

int reorder_only(bh_ctx* c, cudaStream_t st) {
    return bh_reorder_launch(c->posm, c->vel, c->ids, c->perm, c->posm_s, c->vel_s, c->ids_s, c->n, st);
}

int phase_sort(bh_ctx* c, cudaStream_t st) {
    if (c->levels != 20) return sort_keys30(c, true, st);   // the last radix pass moves the bodies itself
    int e = sort_keys60(c, st);
    if (e) return e;
    return reorder_only(c, st);
}

int phase_build(bh_ctx* c, cudaStream_t st) {
    return bh_tree_launch(c->levels == 20 ? (const void*)c->keys64 : (const void*)c->keys0, c->levels, c->n, c->pair_info, c->pair_aux, c->pair_scan, c->tile_sums, c->cell_meta, c->cell_child,
                          c->kid_lv, c->sc, st);
}

int phase_com(bh_ctx* c, cudaStream_t st) {
    return bh_com_launch(c->posm_s, c->n, c->cell_meta, c->cell_child, c->com_scratch, c->cell_com, c->kid_src, c->kid_info, c->sc, c->quad_scratch,
                         c->cell_quad, c->kid_quad, st);
}

int phase_force(bh_ctx* c, cudaStream_t st) {
    return bh_force_launch(c->posm_s, c->levels == 20 ? (const void*)c->keys64 : (const void*)c->keys0, c->levels,
                           c->may_have_ghosts ? c->ids_s : nullptr, c->n, c->slice_first, c->slice_count, c->cell_meta,
                           c->cell_com, c->kid_src, c->kid_info, c->acc, c->sc, c->heavy_list, c->heavy_flag, c->max_chunks, c->prm.theta,
                           c->prm.softening, c->prm.G, c->prm.group_split, c->num_sms, nullptr, nullptr, 0, c->cell_quad, c->kid_quad, st);
}

// traversal + update in one launch (whole-step paths without ghosts; the phases stay separate everywhere else)
int phase_force_update(bh_ctx* c, cudaStream_t st) {
    if (c->may_have_ghosts || c->n < 2 || c->no_fused_update) {
        int e = phase_force(c, st);
        return e ? e : bh_integrate_launch(c->posm_s, c->vel_s, c->ids_s, c->acc, c->posm, c->vel, c->ids, c->slice_first,
                                           c->slice_count, c->prm.dt, c->prm.max_speed, c->sc, st);
    }
    const BhFusedUpdate fu{c->vel_s, c->ids_s, c->posm, c->vel, c->ids, c->prm.dt, c->prm.max_speed};
    return bh_force_launch(c->posm_s, c->levels == 20 ? (const void*)c->keys64 : (const void*)c->keys0, c->levels, nullptr, c->n,
                           c->slice_first, c->slice_count, c->cell_meta, c->cell_com, c->kid_src, c->kid_info, c->acc, c->sc,
                           c->heavy_list, c->heavy_flag, c->max_chunks, c->prm.theta, c->prm.softening, c->prm.G, c->prm.group_split,
                           c->num_sms, nullptr, nullptr, 0, c->cell_quad, c->kid_quad, st, &fu);
}

int phase_update(bh_ctx* c, cudaStream_t st) {
    return bh_integrate_launch(c->posm_s, c->vel_s, c->ids_s, c->acc, c->posm, c->vel, c->ids, c->slice_first,
                               c->slice_count, c->prm.dt, c->prm.max_speed, c->sc, st);
}

typedef int (*phase_fn)(bh_ctx*, cudaStream_t);
const phase_fn kPhases[BH_PHASE_TOTAL] = {phase_keys, phase_sort, phase_build, phase_com, phase_force, phase_update};

// build -> centre of mass -> force -> update, with the centre-of-mass prefix sums (which need the sorted bodies only)
// on a second stream beside the tree construction: fork after the sort, join before the per-cell kernel.  Captured
// into the step graph the two branches become parallel graph nodes.
int launch_tail_overlapped(bh_ctx* c, cudaStream_t st) {
    BH_CUDA_TRY(cudaEventRecord(c->ev_fork, st));
    BH_CUDA_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
    int e = bh_com_prefix_launch(c->posm_s, c->n, c->com_scratch, c->quad_scratch, c->sc, c->aux_stream);
    if (!e) e = phase_build(c, st);
    BH_CUDA_TRY(cudaEventRecord(c->ev_join, c->aux_stream));   // always join, also on error (an open fork breaks a capture)
    BH_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));
    if (e) return e;
    e = bh_com_cells_launch(c->posm_s, c->n, c->cell_meta, c->cell_child, c->com_scratch, c->cell_com, c->kid_src, c->kid_info, c->sc,
                            c->quad_scratch, c->cell_quad, c->kid_quad, st);
    if (!e) e = phase_force_update(c, st);
    return e;
}

int launch_all_phases(bh_ctx* c, cudaStream_t st) {
    int e = phase_keys(c, st);
    if (!e) e = phase_sort(c, st);
    if (!e) e = launch_tail_overlapped(c, st);
    return e;
}

// The step in THREE parts, so that transfers can hide behind compute (NVLink all-gathers in bh_mg_step, PCIe uploads in
// bh_step_host): part 0 — cube, keys, radix sort — reads positions only; part 1 — reorder of positions + masses, tree,
// centre of mass, traversal — still needs no velocity (and no id unless ghosts are possible); only part 2 — reorder of
// velocities + ids, kick-drift-clamp — needs the rest of the state.
int launch_part(bh_ctx* c, int part, cudaStream_t st) {
    if (part == 0) {
        int e = phase_keys(c, st);
        if (e) return e;
        return sort_keys_only(c, st);
    }
    if (part == 1) {
        int e = c->may_have_ghosts ? reorder_only(c, st) : bh_reorder_posm_launch(c->posm, c->perm, c->posm_s, c->n, st);
        if (e) return e;
        BH_CUDA_TRY(cudaEventRecord(c->ev_fork, st));
        BH_CUDA_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        e = bh_com_prefix_launch(c->posm_s, c->n, c->com_scratch, c->quad_scratch, c->sc, c->aux_stream);
        if (!e) e = phase_build(c, st);
        BH_CUDA_TRY(cudaEventRecord(c->ev_join, c->aux_stream));
        BH_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));
        if (e) return e;
        e = bh_com_cells_launch(c->posm_s, c->n, c->cell_meta, c->cell_child, c->com_scratch, c->cell_com, c->kid_src, c->kid_info, c->sc,
                            c->quad_scratch, c->cell_quad, c->kid_quad, st);
        if (!e) e = phase_force(c, st);
        return e;
    }
    int e = c->may_have_ghosts ? 0 : bh_reorder_rest_launch(c->vel, c->ids, c->perm, c->vel_s, c->ids_s, c->n, st);
    if (!e) e = phase_update(c, st);
    return e;
}

int ensure_half_graphs(bh_ctx* c) {
    if (c->half_exec[0] && c->half_exec[1] && c->half_exec[2] && c->graph_half_n == c->n && c->graph_half_first == c->slice_first &&
        c->graph_half_count == c->slice_count)
        return 0;
    for (int h = 0; h < 3; ++h) {
        if (c->half_exec[h]) { cudaGraphExecDestroy(c->half_exec[h]); c->half_exec[h] = nullptr; }
        cudaGraph_t graph = nullptr;
        BH_CUDA_TRY(cudaStreamBeginCapture(c->own_stream, cudaStreamCaptureModeThreadLocal));
        int e = launch_part(c, h, c->own_stream);
        cudaError_t ce = cudaStreamEndCapture(c->own_stream, &graph);
        if (e) { if (graph) cudaGraphDestroy(graph); return e; }
        if (ce != cudaSuccess) return (int)ce;
        ce = cudaGraphInstantiate(&c->half_exec[h], graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return (int)ce;
    }
    c->graph_half_n = c->n; c->graph_half_first = c->slice_first; c->graph_half_count = c->slice_count;
    return 0;
}

int ensure_graph(bh_ctx* c) {
    if (c->graph_exec && c->graph_n == c->n && c->graph_first == c->slice_first && c->graph_count == c->slice_count)
        return 0;
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    BH_CUDA_TRY(cudaStreamBeginCapture(c->own_stream, cudaStreamCaptureModeThreadLocal));
    int e = launch_all_phases(c, c->own_stream);
    cudaError_t ce = cudaStreamEndCapture(c->own_stream, &graph);
    if (e) { if (graph) cudaGraphDestroy(graph); return e; }
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return (int)ce;
    c->graph_n = c->n; c->graph_first = c->slice_first; c->graph_count = c->slice_count;
    return 0;
}

// the captured graphs bake in whether the traversal reads the ids: drop them when that changes
int set_ghosts_possible(bh_ctx* c, bool yes) {
    if (c->may_have_ghosts == yes) return 0;
    c->may_have_ghosts = yes;
    BH_CUDA_TRY(cudaDeviceSynchronize());
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    for (auto& g : c->half_exec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    return 0;
}

int ensure_stage(bh_ctx* c) {
    if (c->stage) return 0;
    return (int)dev_alloc(&c->stage, (size_t)10 * c->n_alloc);
}

}  // namespace

extern "C" {

void bh_default_params(bh_params* p) {
    if (!p) return;
    p->theta = 0.5f; p->G = 0.5f; p->dt = 0.02f; p->softening = 50.0f; p->max_speed = 500.0f;
    p->key_bits = BH_KEY_BITS; p->leaf_cap = 1; p->flags = 0; p->group_split = 0.5f;
}

int bh_abi_version(void) { return BH_ABI_VERSION; }
int bh_group_size(void) { return BH_GROUP; }

const char* bh_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case BH_E_INVAL: return "invalid argument";
        case BH_E_NOMEM: return "host allocation failed";
        case BH_E_STATE: return "call order violated";
        case BH_E_UNSUPPORTED: return "unsupported parameter combination";
        case BH_E_DEVICE: return "device-side error flag raised";
        case BH_E_IO: return "file i/o failed or not a checkpoint";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int bh_create(bh_ctx** out, int64_t n_max, const bh_params* params, int device) {
    if (!out || n_max <= 0 || n_max >= ((int64_t)1 << 28)) return BH_E_INVAL;   // traversal stack words are cell id << 3 | flag
    *out = nullptr;
    bh_params prm;
    if (params) prm = *params; else bh_default_params(&prm);
    if ((prm.key_bits != BH_KEY_BITS && prm.key_bits != 2 * BH_KEY_BITS) || prm.leaf_cap != 1) return BH_E_UNSUPPORTED;
    if (!(prm.softening > 0.0f) || !(prm.theta >= 0.0f) || !(prm.max_speed > 0.0f)) return BH_E_INVAL;
    if (!(prm.group_split >= 0.0f) || prm.group_split > 1.0f) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(device));
    { int e0 = bh_force_prepare(); if (e0) return e0; }
    bh_ctx* c = new (std::nothrow) bh_ctx();
    if (!c) return BH_E_NOMEM;
    c->device = device; c->prm = prm; c->n_max = n_max;
    { const char* v = getenv("BH_NO_FUSED_UPDATE"); c->no_fused_update = v && v[0] == '1'; }
    c->levels = prm.key_bits / 3;
    c->n_alloc = ((n_max + 4095) / 4096 + 1) * 4096;  // room for slice padding in all-gathers
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->num_sms = sms;
    const size_t na = (size_t)c->n_alloc;
    BhSortPlan plan = bh_sort_plan(c->n_alloc);
    c->sort_tmp_bytes = plan.total_bytes;
    cudaError_t e = cudaSuccess;
#define TRYA(x) if (e == cudaSuccess) e = (x)
    TRYA(dev_alloc(&c->posm, na)); TRYA(dev_alloc(&c->vel, na)); TRYA(dev_alloc(&c->posm_s, na));
    TRYA(dev_alloc(&c->vel_s, na)); TRYA(dev_alloc(&c->acc, na)); TRYA(dev_alloc(&c->ids, na)); TRYA(dev_alloc(&c->ids_s, na));
    TRYA(dev_alloc(&c->keys0, na)); TRYA(dev_alloc(&c->keys1, na)); TRYA(dev_alloc(&c->vals0, na)); TRYA(dev_alloc(&c->vals1, na));
    TRYA(cudaMalloc(&c->sort_tmp, plan.total_bytes));
    {
        const bool even = (bh_sort_passes(BH_KEY_BITS) & 1) == 0;
        c->keys_unsorted = even ? c->keys0 : c->keys1;
        c->perm = (c->levels == 20 && !even) ? c->vals1 : c->vals0;
    }
    if (c->levels == 20) { TRYA(dev_alloc(&c->klo, na)); TRYA(dev_alloc(&c->kaux, na)); TRYA(dev_alloc(&c->keys64, na)); }
    TRYA(dev_alloc(&c->pair_info, na)); TRYA(dev_alloc(&c->pair_aux, na)); TRYA(dev_alloc(&c->pair_scan, na)); TRYA(dev_alloc(&c->tile_sums, na / 2048 + 16));
    TRYA(dev_alloc(&c->cell_meta, na)); TRYA(dev_alloc(&c->cell_child, na * 8));
    TRYA(dev_alloc(&c->cell_com, na)); TRYA(cudaMalloc(&c->com_scratch, bh_com_scratch_bytes(c->n_alloc)));
    c->quad = (prm.flags & BH_FLAG_QUADRUPOLE) != 0;
    if (c->quad) {
        TRYA(cudaMalloc(&c->quad_scratch, bh_quad_scratch_bytes(c->n_alloc)));
        TRYA(dev_alloc(&c->cell_quad, na * 2)); TRYA(dev_alloc(&c->kid_quad, na * 16));
    }
    TRYA(dev_alloc(&c->kid_src, na * 8)); TRYA(dev_alloc(&c->kid_lv, na * 8)); TRYA(dev_alloc(&c->kid_info, na * 8));
    TRYA(dev_alloc(&c->sc, 1)); TRYA(dev_alloc(&c->d_scratch, 8));
    c->max_chunks = (int64_t)(na / BH_GROUP + 1);
    TRYA(dev_alloc(&c->heavy_list, 2 * (size_t)c->max_chunks)); TRYA(dev_alloc(&c->heavy_flag, 2 * (size_t)c->max_chunks));
    TRYA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    TRYA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    TRYA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    TRYA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    TRYA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    TRYA(cudaEventCreateWithFlags(&c->ev_h2d_pos, cudaEventDisableTiming));
    TRYA(cudaEventCreateWithFlags(&c->ev_h2d_rest, cudaEventDisableTiming));
    TRYA(cudaEventCreateWithFlags(&c->ev_h2d_mass, cudaEventDisableTiming));
    for (auto& ev : c->ev) TRYA(cudaEventCreate(&ev));
#undef TRYA
    if (e == cudaSuccess) e = cudaMemset(c->sc, 0, sizeof(BhDevScalars));
    if (e == cudaSuccess) e = cudaMemset(c->acc, 0, na * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemset(c->heavy_flag, 0, 8 * (size_t)c->max_chunks);
    if (e != cudaSuccess) { free_all(c); delete c; return (int)e; }
    *out = c;
    return 0;
}

void bh_destroy(bh_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_all(c);
    delete c;
}

int bh_import_soa(bh_ctx* c, const float* px, const float* py, const float* pz, const float* vx, const float* vy,
                  const float* vz, const float* mass, int64_t n, void* stream) {
    if (!c || !px || !py || !pz || !vx || !vy || !vz || !mass || n <= 0 || n > c->n_max) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    c->n = n; c->steps = 0; c->have_sorted = false; c->bbox_fresh = false;
    { int eg = set_ghosts_possible(c, false); if (eg) return eg; }
    default_slice(c);
    // scheduling history of the previous body set is meaningless now
    BH_CUDA_TRY(cudaMemsetAsync((char*)c->sc + offsetof(BhDevScalars, epoch), 0,
                                sizeof(BhDevScalars) - offsetof(BhDevScalars, epoch), (cudaStream_t)stream));
    BH_CUDA_TRY(cudaMemsetAsync(c->heavy_flag, 0, 8 * (size_t)c->max_chunks, (cudaStream_t)stream));
    int e = bh_import_launch(px, py, pz, vx, vy, vz, mass, n, c->posm, c->vel, c->ids, (cudaStream_t)stream);
    if (e) return e;
    c->have_state = true;
    return 0;
}

static int import_host_impl(bh_ctx* c, const float* px, const float* py, const float* pz, const float* vx, const float* vy,
                            const float* vz, const float* mass, int64_t n, bool sync) {
    if (!c || !px || !py || !pz || !vx || !vy || !vz || !mass || n <= 0 || n > c->n_max) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    int e = ensure_stage(c);
    if (e) return e;
    const float* src[7] = {px, py, pz, vx, vy, vz, mass};
    for (int k = 0; k < 7; ++k)
        BH_CUDA_TRY(cudaMemcpyAsync(c->stage + (size_t)k * c->n_alloc, src[k], (size_t)n * 4, cudaMemcpyHostToDevice, c->own_stream));
    float* s = c->stage;
    const size_t na = (size_t)c->n_alloc;
    e = bh_import_soa(c, s, s + na, s + 2 * na, s + 3 * na, s + 4 * na, s + 5 * na, s + 6 * na, n, c->own_stream);
    if (e) return e;
    // own_stream does not order against the caller's streams: finish before returning
    if (sync) BH_CUDA_TRY(cudaStreamSynchronize(c->own_stream));
    return 0;
}

int bh_import_soa_host(bh_ctx* c, const float* px, const float* py, const float* pz, const float* vx, const float* vy,
                       const float* vz, const float* mass, int64_t n) {
    return import_host_impl(c, px, py, pz, vx, vy, vz, mass, n, true);
}

int bh_set_fixed_bounds(bh_ctx* c, const float b[6]) {
    if (!c) return BH_E_INVAL;
    const bool was = c->fixed_bounds_set;
    c->fixed_bounds_set = b != nullptr;
    if (b) memcpy(c->fixed_bounds, b, sizeof(c->fixed_bounds));
    if (was != c->fixed_bounds_set) {   // the captured graphs contain the other variant of the keys phase
        BH_CUDA_TRY(cudaSetDevice(c->device));
        BH_CUDA_TRY(cudaDeviceSynchronize());
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        for (auto& g : c->half_exec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    }
    return 0;
}

int bh_local_bounds(bh_ctx* c, float lohi[6]) {
    if (!c || !lohi) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    int e = bh_bounds_enc_launch(c->posm, c->n, c->sc, 0);
    if (e) return e;
    c->bbox_fresh = true;   // the images now describe the current positions: the next keys phase may use them
    BhDevScalars h;
    BH_CUDA_TRY(cudaMemcpy(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 6; ++k) {   // undo the order-preserving integer image (bh_f2ord)
        const unsigned int u = h.bbox_enc[k];
        const unsigned int bits = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
        memcpy(&lohi[k], &bits, 4);
    }
    return 0;
}

int bh_import_state(bh_ctx* c, const void* posm, const void* vel, const int32_t* ids, int64_t n, void* stream) {
    if (!c || !posm || !vel || !ids || n <= 0 || n > c->n_max) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    c->n = n; c->steps = 0; c->have_sorted = false; c->bbox_fresh = false;
    { int eg = set_ghosts_possible(c, true); if (eg) return eg; }   // the caller's ids may mark ghosts (id < 0)
    default_slice(c);
    BH_CUDA_TRY(cudaMemsetAsync((char*)c->sc + offsetof(BhDevScalars, epoch), 0,
                                sizeof(BhDevScalars) - offsetof(BhDevScalars, epoch), st));
    BH_CUDA_TRY(cudaMemsetAsync(c->heavy_flag, 0, 8 * (size_t)c->max_chunks, st));
    BH_CUDA_TRY(cudaMemcpyAsync(c->posm, posm, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
    BH_CUDA_TRY(cudaMemcpyAsync(c->vel, vel, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
    BH_CUDA_TRY(cudaMemcpyAsync(c->ids, ids, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    c->have_state = true;
    return 0;
}

static int let_scratch(bh_ctx* c) {
    if (c->let_boxes) return 0;
    c->let_qcap = c->n_alloc / 4 > (1 << 20) ? c->n_alloc / 4 : (1 << 20);
    BH_CUDA_TRY(dev_alloc(&c->let_boxes, (size_t)BH_LET_MAX_PEERS * (BH_LET_MAX_BOXES + 1) * 6));
    BH_CUDA_TRY(dev_alloc(&c->let_counts, BH_LET_MAX_PEERS + 2));
    BH_CUDA_TRY(dev_alloc(&c->let_queue, 2 * (size_t)c->let_qcap));
    return 0;
}

int bh_let_domain_boxes(bh_ctx* c, const uint32_t* cuts, int K, float* lohi, int32_t* body_counts) {
    if (!c || !cuts || !lohi || K < 1 || K > BH_LET_MAX_BOXES) return BH_E_INVAL;
    for (int k = 0; k < K; ++k)
        if (cuts[k] > cuts[k + 1]) return BH_E_INVAL;
    if (!c->have_state || !c->have_sorted) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    int e = let_scratch(c);
    if (e) return e;
    BH_CUDA_TRY(cudaDeviceSynchronize());
    // scratch: boxes at the front of let_boxes, cuts and counts behind them (the export call rewrites all of it)
    uint32_t* d_cuts = (uint32_t*)(c->let_boxes + (size_t)6 * BH_LET_MAX_BOXES);
    int* d_counts = (int*)(d_cuts + BH_LET_MAX_BOXES + 1);
    unsigned int* d_enc = (unsigned int*)(d_counts + BH_LET_MAX_BOXES);
    BH_CUDA_TRY(cudaMemcpy(d_cuts, cuts, sizeof(uint32_t) * (K + 1), cudaMemcpyHostToDevice));
    e = bh_domain_boxes_launch(c->keys0, c->posm_s, c->n, d_cuts, K, c->let_boxes, d_counts, d_enc, 0);
    if (e) return e;
    BH_CUDA_TRY(cudaMemcpy(lohi, c->let_boxes, sizeof(float) * 6 * K, cudaMemcpyDeviceToHost));
    if (body_counts) BH_CUDA_TRY(cudaMemcpy(body_counts, d_counts, sizeof(int) * K, cudaMemcpyDeviceToHost));
    return 0;
}

int bh_sorted_ptrs(bh_ctx* c, void** keys, void** posm, void** vel, void** ids, void** acc, int64_t* n) {
    if (!c) return BH_E_INVAL;
    if (!c->have_state || !c->have_sorted) return BH_E_STATE;
    if (keys) *keys = c->keys0;
    if (posm) *posm = c->posm_s;
    if (vel) *vel = c->vel_s;
    if (ids) *ids = c->ids_s;
    if (acc) *acc = c->acc;
    if (n) *n = c->n;
    return 0;
}

int bh_sort_coarse(bh_ctx* c, void* stream) {
    if (!c) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    int e = 0;
    if (c->fixed_bounds_set) {
        BH_CUDA_TRY(cudaMemcpyAsync((char*)c->sc + offsetof(BhDevScalars, bounds), c->fixed_bounds, sizeof(c->fixed_bounds),
                                    cudaMemcpyHostToDevice, st));
    } else {
        e = pre_keys(c, st);
        if (!e) e = bh_bounds_finish_launch(c->sc, st);
        if (e) return e;
        c->bbox_fresh = false;
    }
    e = bh_keys_launch(c->posm, c->n, c->sc, c->keys_unsorted, st);
    if (e) return e;
    e = sort_keys30(c, true, st);   // sorted 30-bit keys in keys0, bodies in posm_s / vel_s / ids_s
    if (e) return e;
    c->keys_sorted = true;
    c->have_sorted = true;
    return 0;
}

int bh_force_from(bh_ctx* c, bh_ctx* src, void* stream) {
    if (!c || !src || c == src) return BH_E_INVAL;
    if (!c->have_state || !c->have_sorted || !src->have_state || !src->have_sorted) return BH_E_STATE;
    if (c->device != src->device || c->levels != src->levels) return BH_E_INVAL;
    if (c->quad || src->quad) return BH_E_UNSUPPORTED;   // exported points carry monopoles only
    if (!c->fixed_bounds_set || !src->fixed_bounds_set || memcmp(c->fixed_bounds, src->fixed_bounds, sizeof(c->fixed_bounds)) != 0)
        return BH_E_STATE;   // cell widths are derived from the cube: both trees must live on the same grid
    BH_CUDA_TRY(cudaSetDevice(c->device));
    if (src->n < 2) return BH_E_UNSUPPORTED;   // a single body has no tree
    return bh_force_launch(c->posm_s, c->levels == 20 ? (const void*)c->keys64 : (const void*)c->keys0, c->levels, c->ids_s, c->n,
                           c->slice_first, c->slice_count, src->cell_meta, src->cell_com, src->kid_src,
                           src->kid_info, c->acc, c->sc, c->heavy_list, c->heavy_flag, c->max_chunks, c->prm.theta,
                           c->prm.softening, c->prm.G, c->prm.group_split, c->num_sms, src->posm_s, src->sc, 1, nullptr, nullptr,
                           (cudaStream_t)stream);
}

int bh_export_real(bh_ctx* c, void* posm_out, void* vel_out, int32_t* ids_out, int64_t* n_real, void* stream) {
    if (!c || !posm_out || !vel_out || !ids_out || !n_real) return BH_E_INVAL;
    if (!c->have_state || !c->have_sorted) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    // tile_sums holds n_alloc/2048 + 16 ints: enough for the per-tile counts and the total
    int e = bh_compact_real_launch(c->posm, c->vel, c->ids, c->acc, c->n, c->tile_sums, (float4*)posm_out, (float4*)vel_out,
                                   ids_out, st);
    if (e) return e;
    const int tiles = (int)((c->n + 2047) / 2048);
    int32_t total = 0;
    BH_CUDA_TRY(cudaMemcpyAsync(&total, c->tile_sums + tiles, 4, cudaMemcpyDeviceToHost, st));
    BH_CUDA_TRY(cudaStreamSynchronize(st));
    *n_real = total;
    return 0;
}

int bh_let_export(bh_ctx* c, const float* boxes_lohi, int npeers, int K, void* out, int64_t cap_per_peer, int32_t* counts,
                  void* stream) {
    if (!c || !boxes_lohi || !out || !counts || npeers < 1 || npeers > BH_LET_MAX_PEERS || K < 1 || K > BH_LET_MAX_BOXES ||
        cap_per_peer < 1)
        return BH_E_INVAL;
    if (!c->have_state || !c->have_sorted) return BH_E_STATE;
    if (c->quad) return BH_E_UNSUPPORTED;   // the export emits {centre of mass, mass} points
    BH_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    { int e0 = let_scratch(c); if (e0) return e0; }
    // layout: npeers*K boxes, then npeers hulls (the AABB of each peer's used boxes)
    std::vector<float> boxes(((size_t)npeers * K + npeers) * 6);
    std::vector<char> active((size_t)npeers, 0);
    std::vector<float> hull((size_t)npeers * 6);
    for (int p = 0; p < npeers; ++p)
        for (int a = 0; a < 3; ++a) { hull[6 * p + a] = 3.0e38f; hull[6 * p + 3 + a] = -3.0e38f; }
    auto centre_half = [](const float* b, float* o) {   // centre / half extent exactly as the force kernel forms them
        for (int a = 0; a < 3; ++a) {
            o[a] = (b[a] + b[3 + a]) * 0.5f;
            o[3 + a] = (b[3 + a] - b[a]) * 0.5f;
        }
    };
    for (int r = 0; r < npeers * K; ++r) {
        const float* b = boxes_lohi + 6 * r;
        centre_half(b, &boxes[6 * r]);
        if (b[0] > b[3]) { boxes[6 * r + 3] = -1.0f; continue; }
        active[r / K] = 1;
        float* h = &hull[6 * (r / K)];
        for (int a = 0; a < 3; ++a) {
            h[a] = b[a] < h[a] ? b[a] : h[a];
            h[3 + a] = b[3 + a] > h[3 + a] ? b[3 + a] : h[3 + a];
        }
    }
    for (int p = 0; p < npeers; ++p) {
        float* o = &boxes[((size_t)npeers * K + p) * 6];
        if (active[p]) centre_half(&hull[6 * p], o);
        else { for (int a = 0; a < 6; ++a) o[a] = 0.0f; }
    }
    BH_CUDA_TRY(cudaStreamSynchronize(st));
    BhDevScalars h;
    BH_CUDA_TRY(cudaMemcpy(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost));
    // overflow is reported per call: a list that did not fit last time must not fail this (larger) attempt
    if (h.err & BH_DERR_LET_OVERFLOW) {
        const unsigned int cleared = h.err & ~(unsigned int)BH_DERR_LET_OVERFLOW;
        BH_CUDA_TRY(cudaMemcpy((char*)c->sc + offsetof(BhDevScalars, err), &cleared, 4, cudaMemcpyHostToDevice));
    }
    BH_CUDA_TRY(cudaMemcpy(c->let_boxes, boxes.data(), sizeof(float) * boxes.size(), cudaMemcpyHostToDevice));
    if (c->n >= 2) {
        int e = bh_let_export_launch(c->cell_meta, c->cell_child, c->cell_com, c->kid_src, c->kid_lv, c->posm_s, c->sc,
                                     c->let_boxes, c->let_boxes + (size_t)npeers * K * 6, npeers, K, (float4*)out, c->let_counts, cap_per_peer, c->let_queue,
                                     c->let_counts + BH_LET_MAX_PEERS, c->let_qcap, c->prm.theta, c->prm.softening,
                                     fmaxf(h.bounds[3] - h.bounds[0], 1.0f), c->levels, st);
        if (e) return e;
        unsigned int hc[BH_LET_MAX_PEERS];
        BH_CUDA_TRY(cudaMemcpyAsync(hc, c->let_counts, sizeof(unsigned int) * npeers, cudaMemcpyDeviceToHost, st));
        BH_CUDA_TRY(cudaStreamSynchronize(st));
        for (int r = 0; r < npeers; ++r) counts[r] = (int32_t)(hc[r] < (unsigned long long)cap_per_peer ? hc[r] : cap_per_peer);
    } else {   // a single body has no tree: it is its own essential set
        for (int r = 0; r < npeers; ++r) {
            counts[r] = 0;
            if (!active[r]) continue;
            BH_CUDA_TRY(cudaMemcpyAsync((float4*)out + (size_t)r * cap_per_peer, c->posm_s, 16, cudaMemcpyDeviceToDevice, st));
            counts[r] = 1;
        }
        BH_CUDA_TRY(cudaStreamSynchronize(st));
    }
    BH_CUDA_TRY(cudaMemcpy(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost));
    return (h.err & BH_DERR_LET_OVERFLOW) ? BH_E_DEVICE : 0;
}

int bh_set_slice(bh_ctx* c, int rank, int world) {
    if (!c || world < 1 || rank < 0 || rank >= world) return BH_E_INVAL;
    c->rank = rank; c->world = world;
    c->bbox_fresh = false;
    if (c->have_state) {
        default_slice(c);
        BH_CUDA_TRY(cudaSetDevice(c->device));
        BH_CUDA_TRY(cudaDeviceSynchronize());
        // chunk indices are relative to the slice: drop the heavy-chunk history
        BH_CUDA_TRY(cudaMemset((char*)c->sc + offsetof(BhDevScalars, epoch), 0, sizeof(BhDevScalars) - offsetof(BhDevScalars, epoch)));
        BH_CUDA_TRY(cudaMemset(c->heavy_flag, 0, 8 * (size_t)c->max_chunks));
    }
    return 0;
}

int bh_state_ptrs(bh_ctx* c, void** posm, void** vel, void** ids, int64_t* n, int64_t* slice_first, int64_t* slice_count) {
    if (!c) return BH_E_INVAL;
    if (posm) { *posm = c->posm; c->external_writers = true; c->bbox_fresh = false; }
    if (vel) *vel = c->vel;
    if (ids) *ids = c->ids;
    if (n) *n = c->n;
    if (slice_first) *slice_first = c->slice_first;
    if (slice_count) *slice_count = c->slice_count;
    return 0;
}

int bh_step(bh_ctx* c, int nsteps, void* stream) {
    if (!c || nsteps < 0) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool timer = (c->prm.flags & BH_FLAG_PHASE_TIMER) != 0;
    const bool direct = timer || (c->prm.flags & BH_FLAG_NO_GRAPH) != 0;
    if (timer) memset(c->phase_ms, 0, sizeof(c->phase_ms));
    if (!direct) {
        int e = ensure_graph(c);
        if (e) return e;
    }
    for (int s = 0; s < nsteps; ++s) {
        if (timer) BH_CUDA_TRY(cudaEventRecord(c->ev[0], st));
        { int e = pre_keys(c, st); if (e) return e; }   // only when the last integrator did not cover every body
        if (!direct) {
            BH_CUDA_TRY(cudaGraphLaunch(c->graph_exec, st));
        } else if (!timer) {
            int e = launch_all_phases(c, st);
            if (e) return e;
        } else {
            for (int p = 0; p < BH_PHASE_TOTAL; ++p) {
                if (p > 0) BH_CUDA_TRY(cudaEventRecord(c->ev[p], st));
                int e = kPhases[p](c, st);
                if (e) return e;
            }
            BH_CUDA_TRY(cudaEventRecord(c->ev[BH_PHASE_TOTAL], st));
            BH_CUDA_TRY(cudaEventSynchronize(c->ev[BH_PHASE_TOTAL]));
            for (int p = 0; p < BH_PHASE_TOTAL; ++p) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, c->ev[p], c->ev[p + 1]);
                c->phase_ms[p] += ms;
                c->phase_ms[BH_PHASE_TOTAL] += ms;
            }
        }
        post_update(c);
    }
    if (nsteps > 0) { c->steps += nsteps; c->have_sorted = true; c->keys_sorted = true; }
    return 0;
}

int bh_step_part(bh_ctx* c, int part, void* stream) {
    if (!c || part < 0 || part > 2) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    if (part == 0) { int e = pre_keys(c, (cudaStream_t)stream); if (e) return e; }
    if (c->prm.flags & (BH_FLAG_NO_GRAPH | BH_FLAG_PHASE_TIMER)) {
        int e = launch_part(c, part, (cudaStream_t)stream);
        if (e) return e;
    } else {
        int e = ensure_half_graphs(c);
        if (e) return e;
        BH_CUDA_TRY(cudaGraphLaunch(c->half_exec[part], (cudaStream_t)stream));
    }
    if (part == 0) { c->bbox_fresh = false; c->keys_sorted = true; }   // images consumed by the keys phase; keys sorted
    if (part == 2) { c->steps += 1; c->have_sorted = true; post_update(c); }
    return 0;
}

int bh_step_half(bh_ctx* c, int half, void* stream) {
    if (!c || half < 0 || half > 1) return BH_E_INVAL;
    if (half == 0) return bh_step_part(c, 0, stream);
    int e = bh_step_part(c, 1, stream);
    return e ? e : bh_step_part(c, 2, stream);
}

int bh_run_phase(bh_ctx* c, int phase, void* stream) {
    if (!c || phase < 0 || phase >= BH_PHASE_TOTAL) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    if (phase == BH_PHASE_KEYS) { int e0 = pre_keys(c, (cudaStream_t)stream); if (e0) return e0; }
    int e = kPhases[phase](c, (cudaStream_t)stream);
    if (e) return e;
    if (phase == BH_PHASE_KEYS) { c->bbox_fresh = false; c->keys_sorted = false; }   // images consumed; keys unsorted
    if (phase == BH_PHASE_SORT) c->keys_sorted = true;
    if (phase == BH_PHASE_UPDATE) post_update(c);
    if (phase >= BH_PHASE_SORT) c->have_sorted = true;
    return 0;
}

int bh_set_flags(bh_ctx* c, int flags) {
    if (!c) return BH_E_INVAL;
    c->prm.flags = (flags & ~BH_FLAG_QUADRUPOLE) | (c->quad ? BH_FLAG_QUADRUPOLE : 0);   // the moment arrays are allocated at creation
    return 0;
}

int bh_phase_ms(bh_ctx* c, float out[BH_PHASE_COUNT]) {
    if (!c || !out) return BH_E_INVAL;
    memcpy(out, c->phase_ms, sizeof(c->phase_ms));
    return 0;
}

int bh_export_soa(bh_ctx* c, float* px, float* py, float* pz, float* vx, float* vy, float* vz, float* ax, float* ay,
                  float* az, void* stream) {
    if (!c) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    return bh_export_launch(c->posm, c->vel, c->acc, c->ids, c->n, px, py, pz, vx, vy, vz, ax, ay, az, (cudaStream_t)stream);
}

int bh_export_soa_host(bh_ctx* c, float* px, float* py, float* pz, float* vx, float* vy, float* vz, float* ax,
                       float* ay, float* az) {
    if (!c) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    int e = ensure_stage(c);
    if (e) return e;
    float* dst[9] = {px, py, pz, vx, vy, vz, ax, ay, az};
    float* dev[9];
    const size_t na = (size_t)c->n_alloc;
    for (int k = 0; k < 9; ++k) dev[k] = dst[k] ? c->stage + (size_t)k * na : nullptr;
    // order the export after whatever the caller queued on the legacy stream / our stream
    BH_CUDA_TRY(cudaDeviceSynchronize());
    e = bh_export_launch(c->posm, c->vel, c->acc, c->ids, c->n, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6],
                         dev[7], dev[8], c->own_stream);
    if (e) return e;
    for (int k = 0; k < 9; ++k)
        if (dst[k]) BH_CUDA_TRY(cudaMemcpyAsync(dst[k], dev[k], (size_t)c->n * 4, cudaMemcpyDeviceToHost, c->own_stream));
    BH_CUDA_TRY(cudaStreamSynchronize(c->own_stream));
    return 0;
}

// Host state in -> nsteps -> host state out, every copy inside the call.  The uploads are ordered so that compute
// starts early: positions first; bounds, keys and the radix sort (they read positions only) run while masses and
// velocities are still crossing PCIe on the copy stream — the trick bh_mg_step plays for NVLink.
#define BH_STEP_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail((int)_e); } while (0)
int bh_step_host(bh_ctx* c, float* px, float* py, float* pz, float* vx, float* vy, float* vz, const float* mass,
                 int64_t n, int nsteps) {
    if (!c || !px || !py || !pz || !vx || !vy || !vz || !mass || n <= 0 || n > c->n_max || nsteps < 0) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    int e = ensure_stage(c);
    if (e) return e;
    const size_t na = (size_t)c->n_alloc, bytes = (size_t)n * 4;
    float* s = c->stage;
    cudaStream_t cs = c->own_stream, xs = c->copy_stream;   // compute / copies
    // on any error the caller's host buffers must not be read (or written) behind its back any more
    auto fail = [&](int code) { cudaStreamSynchronize(xs); cudaStreamSynchronize(cs); return code; };
#ifdef BH_TRACE_STEP_HOST   // build-time diagnostic: device timeline of one call on stderr (tools/build_variants.sh)
    cudaEvent_t tr[12]; int ntr = 0; const char* trn[12];
    for (auto& t : tr) cudaEventCreate(&t);
    auto mark = [&](cudaStream_t q, const char* name) { trn[ntr] = name; cudaEventRecord(tr[ntr++], q); };
    cudaStreamSynchronize(cs); cudaStreamSynchronize(xs);
    mark(xs, "start"); cudaStreamWaitEvent(cs, tr[0], 0);
#define BH_MARK(q, name) mark(q, name)
#else
#define BH_MARK(q, name) do {} while (0)
#endif
    if (nsteps == 0 || (c->prm.flags & BH_FLAG_PHASE_TIMER)) {   // nothing to overlap with / phases are timed one by one
        e = import_host_impl(c, px, py, pz, vx, vy, vz, mass, n, false);
        if (!e) e = bh_step(c, nsteps, cs);
        if (e) return fail(e);
    } else {
        // upload order = order of need: positions (cube, keys, sort), masses (tree, centre of mass, traversal),
        // velocities (update only) — the whole velocity upload hides behind the tree build and the traversal
        const float* src[7] = {px, py, pz, mass, vx, vy, vz};
        const int slot[7] = {0, 1, 2, 6, 3, 4, 5};
        for (int k = 0; k < 7; ++k) {
            BH_STEP_TRY(cudaMemcpyAsync(s + (size_t)slot[k] * na, src[k], bytes, cudaMemcpyHostToDevice, xs));
            if (k == 2) { BH_STEP_TRY(cudaEventRecord(c->ev_h2d_pos, xs)); BH_MARK(xs, "h2d pos"); }
            if (k == 3) { BH_STEP_TRY(cudaEventRecord(c->ev_h2d_mass, xs)); BH_MARK(xs, "h2d mass"); }
        }
        BH_STEP_TRY(cudaEventRecord(c->ev_h2d_rest, xs));
        BH_MARK(xs, "h2d vel");
        // what bh_import_soa does, in three parts.  A caller that steps the same system call after call hands back
        // the bodies the last call returned: after the sort the chunks hold the same bodies again, so the record of
        // which chunks were expensive (a scheduling hint, bh_force.cu) stays valid — it is only dropped when n changes.
        const bool same_system = c->have_state && c->n == n && c->world == 1;   // a stale hint costs time, never correctness
        c->n = n; c->steps = 0; c->have_sorted = false; c->bbox_fresh = false;
        e = set_ghosts_possible(c, false);
        if (e) return fail(e);
        default_slice(c);
        if (!same_system) {
            BH_STEP_TRY(cudaMemsetAsync((char*)c->sc + offsetof(BhDevScalars, epoch), 0, sizeof(BhDevScalars) - offsetof(BhDevScalars, epoch), cs));
            BH_STEP_TRY(cudaMemsetAsync(c->heavy_flag, 0, 8 * (size_t)c->max_chunks, cs));
        }
        BH_STEP_TRY(cudaStreamWaitEvent(cs, c->ev_h2d_pos, 0));
        e = bh_import_pos_launch(s, s + na, s + 2 * na, n, c->posm, cs);
        if (e) return fail(e);
        c->have_state = true;
        e = bh_step_part(c, 0, cs);                       // cube, keys, radix sort: positions only
        if (e) return fail(e);
        BH_MARK(cs, "part0 done");
        BH_STEP_TRY(cudaStreamWaitEvent(cs, c->ev_h2d_mass, 0));
        e = bh_import_mass_launch(s + 6 * na, n, c->posm, cs);
        if (!e) e = bh_step_part(c, 1, cs);               // reorder positions + masses, tree, centre of mass, traversal
        if (e) return fail(e);
        BH_MARK(cs, "part1 done");
        BH_STEP_TRY(cudaStreamWaitEvent(cs, c->ev_h2d_rest, 0));
        e = bh_import_vel_launch(s + 3 * na, s + 4 * na, s + 5 * na, n, c->vel, c->ids, cs);
        if (!e) e = bh_step_part(c, 2, cs);               // reorder velocities + ids, update
        if (!e && nsteps > 1) e = bh_step(c, nsteps - 1, cs);
        if (e) return fail(e);
        BH_MARK(cs, "part2 done");
    }
    // export on the compute stream as a gather (slot of every body id, then positions, then velocities), downloads on
    // the copy stream behind each: the positions are already crossing PCIe while the velocities are still being
    // gathered.  The second value buffer is sort scratch, free until the next step.
    float* dst[6] = {px, py, pz, vx, vy, vz};
    int32_t* where = (int32_t*)(c->perm == c->vals1 ? c->vals0 : c->vals1);   // the value buffer that is not the permutation
    e = bh_where_launch(c->ids, c->n, where, cs);
    if (!e) e = bh_gather3_launch(c->posm, where, c->n, s, s + na, s + 2 * na, cs);
    if (e) return fail(e);
    BH_STEP_TRY(cudaEventRecord(c->ev_h2d_pos, cs));
    e = bh_gather3_launch(c->vel, where, c->n, s + 3 * na, s + 4 * na, s + 5 * na, cs);
    if (e) return fail(e);
    BH_STEP_TRY(cudaEventRecord(c->ev_h2d_rest, cs));
    BH_MARK(cs, "export done");
    BH_STEP_TRY(cudaStreamWaitEvent(xs, c->ev_h2d_pos, 0));
    for (int k = 0; k < 3; ++k)
        BH_STEP_TRY(cudaMemcpyAsync(dst[k], s + (size_t)k * na, bytes, cudaMemcpyDeviceToHost, xs));
    BH_STEP_TRY(cudaStreamWaitEvent(xs, c->ev_h2d_rest, 0));
    for (int k = 3; k < 6; ++k)
        BH_STEP_TRY(cudaMemcpyAsync(dst[k], s + (size_t)k * na, bytes, cudaMemcpyDeviceToHost, xs));
    BH_MARK(xs, "d2h done");
    BH_STEP_TRY(cudaStreamSynchronize(xs));
    BH_STEP_TRY(cudaStreamSynchronize(cs));
#ifdef BH_TRACE_STEP_HOST
    for (int i = 1; i < ntr; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, tr[0], tr[i]); fprintf(stderr, "[step_host] %-12s %.3f ms\n", trn[i], ms); }
    for (auto& t : tr) cudaEventDestroy(t);
#endif
#undef BH_MARK
    return 0;
}
#undef BH_STEP_TRY

static int dbg_locate(bh_ctx* c, int what, void** ptr, size_t* bytes) {
    BhDevScalars h;
    int M = 0;
    if (what == BH_DBG_CELL_META || what == BH_DBG_CELL_COM || what == BH_DBG_CELL_CHILD || what == BH_DBG_CELL_QUAD) {
        BH_CUDA_TRY(cudaMemcpy(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost));
        M = h.num_cells;
    }
    const size_t n = (size_t)c->n;
    switch (what) {
        case BH_DBG_BOUNDS: *ptr = (char*)c->sc + offsetof(BhDevScalars, bounds); *bytes = 24; break;
        case BH_DBG_KEYS: *ptr = (c->levels == 20 || c->keys_sorted) ? c->keys0 : c->keys_unsorted; *bytes = n * 4; break;
        case BH_DBG_KEYS64:
            if (c->levels != 20) return BH_E_UNSUPPORTED;
            *ptr = c->keys64; *bytes = n * 8; break;
        case BH_DBG_PERM: *ptr = c->perm; *bytes = n * 4; break;
        case BH_DBG_IDS: *ptr = c->ids; *bytes = n * 4; break;
        case BH_DBG_POSM: *ptr = c->posm; *bytes = n * 16; break;
        case BH_DBG_VEL: *ptr = c->vel; *bytes = n * 16; break;
        case BH_DBG_ACC: *ptr = c->acc; *bytes = n * 16; break;
        case BH_DBG_CELL_META: *ptr = c->cell_meta; *bytes = (size_t)M * 16; break;
        case BH_DBG_CELL_COM: *ptr = c->cell_com; *bytes = (size_t)M * 16; break;
        case BH_DBG_CELL_CHILD: *ptr = c->cell_child; *bytes = (size_t)M * 32; break;
        case BH_DBG_CELL_QUAD:
            if (!c->quad) return BH_E_UNSUPPORTED;
            *ptr = c->cell_quad; *bytes = (size_t)M * 32; break;
        case BH_DBG_POSM_SORTED: *ptr = c->posm_s; *bytes = n * 16; break;
        case BH_DBG_VEL_SORTED: *ptr = c->vel_s; *bytes = n * 16; break;
        case BH_DBG_IDS_SORTED: *ptr = c->ids_s; *bytes = n * 4; break;
        default: return BH_E_INVAL;
    }
    return 0;
}

int bh_debug_get(bh_ctx* c, int what, void* dst, size_t bytes) {
    if (!c || !dst) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    void* p = nullptr; size_t want = 0;
    int e = dbg_locate(c, what, &p, &want);
    if (e) return e;
    if (bytes != want) return BH_E_INVAL;
    if (bytes == 0) return 0;
    return (int)cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost);
}

int bh_debug_set(bh_ctx* c, int what, const void* src, size_t bytes) {
    if (!c || !src) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    void* p = nullptr; size_t want = 0;
    int e = dbg_locate(c, what, &p, &want);
    if (e) return e;
    if (bytes != want) return BH_E_INVAL;
    if (bytes == 0) return 0;
    c->bbox_fresh = false;   // the caller may have moved bodies
    if (what == BH_DBG_IDS || what == BH_DBG_IDS_SORTED) { int eg = set_ghosts_possible(c, true); if (eg) return eg; }
    return (int)cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
}

int64_t bh_stat(bh_ctx* c, int which) {
    if (!c) return BH_E_INVAL;
    if (which == BH_STAT_N) return c->n;
    if (which == BH_STAT_STEPS) return c->steps;
    if (cudaSetDevice(c->device) != cudaSuccess) return BH_E_INVAL;
    if (cudaDeviceSynchronize() != cudaSuccess) return BH_E_DEVICE;
    BhDevScalars h;
    if (cudaMemcpy(&h, c->sc, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return BH_E_DEVICE;
    switch (which) {
        case BH_STAT_CELLS: return h.num_cells;
        case BH_STAT_ROOT: return h.root;
        case BH_STAT_INTERACTIONS_CELL: return (int64_t)h.inter_cell;
        case BH_STAT_INTERACTIONS_BODY: return (int64_t)h.inter_body;
        case BH_STAT_DEVICE_ERROR: return h.err;
        case BH_STAT_MAX_STACK: return h.max_stack;
        default: return BH_E_INVAL;
    }
}

int bh_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                      int64_t n, int begin_bit, int end_bit, void* tmp, size_t* tmp_bytes, void* stream) {
    if (!tmp_bytes || n < 0) return BH_E_INVAL;
    BhSortPlan plan = bh_sort_plan(n);
    const size_t arr = ((size_t)n * 4 + 255) / 256 * 256;
    const size_t plan_bytes = (plan.total_bytes + 255) / 256 * 256;
    const size_t need = plan_bytes + 2 * arr + 256;
    if (!tmp) { *tmp_bytes = need; return 0; }
    if (*tmp_bytes < need || !keys_in || !vals_in || !keys_out || !vals_out) return BH_E_INVAL;
    char* base = (char*)tmp;
    uint32_t* tk = (uint32_t*)(base + plan_bytes);
    uint32_t* tv = (uint32_t*)(base + plan_bytes + arr);
    unsigned int* err = (unsigned int*)(base + plan_bytes + 2 * arr);
    BH_CUDA_TRY(cudaMemsetAsync(err, 0, 4, (cudaStream_t)stream));
    const int passes = bh_sort_passes(end_bit - begin_bit);
    int in_q = 0;
    // make the final pass land in (keys_out, vals_out)
    if (passes & 1)
        return bh_sort_pairs_launch(keys_in, vals_in, keys_out, vals_out, tk, tv, n, begin_bit, end_bit, tmp, false, err, &in_q, (cudaStream_t)stream);
    return bh_sort_pairs_launch(keys_in, vals_in, tk, tv, keys_out, vals_out, n, begin_bit, end_bit, tmp, false, err, &in_q, (cudaStream_t)stream);
}

int bh_direct_sample(bh_ctx* c, const int32_t* sample, int k, double* acc_out) {
    if (!c || !sample || !acc_out || k <= 0) return BH_E_INVAL;
    if (!c->have_state || !c->have_sorted) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    std::vector<int32_t> ids((size_t)c->n), inv((size_t)c->n, -1), slots((size_t)k);
    BH_CUDA_TRY(cudaMemcpy(ids.data(), c->ids_s, (size_t)c->n * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < c->n; ++i)
        if (ids[i] >= 0 && ids[i] < c->n) inv[ids[i]] = (int32_t)i;
    for (int s = 0; s < k; ++s) {
        if (sample[s] < 0 || sample[s] >= c->n || inv[sample[s]] < 0) return BH_E_INVAL;
        slots[s] = inv[sample[s]];
    }
    int32_t* d_slots = nullptr; double* d_out = nullptr;
    BH_CUDA_TRY(cudaMalloc(&d_slots, (size_t)k * 4));
    if (cudaMalloc(&d_out, (size_t)k * 24) != cudaSuccess) { cudaFree(d_slots); return (int)cudaErrorMemoryAllocation; }
    cudaMemcpy(d_slots, slots.data(), (size_t)k * 4, cudaMemcpyHostToDevice);
    int e = bh_direct_launch(c->posm_s, c->n, d_slots, k, c->prm.softening, c->prm.G, d_out, 0);
    cudaError_t ce = cudaMemcpy(acc_out, d_out, (size_t)k * 24, cudaMemcpyDeviceToHost);
    cudaFree(d_slots); cudaFree(d_out);
    return e ? e : (int)ce;
}

int bh_export_visuals(bh_ctx* c, float* vbo_p, float* vbo_c, void* stream) {
    if (!c) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    return bh_visuals_launch(c->posm, c->vel, c->ids, c->n, vbo_p, vbo_c, (cudaStream_t)stream);
}

int bh_momentum(bh_ctx* c, double out[7]) {
    if (!c || !out) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    int e = bh_momentum_launch(c->posm, c->vel, c->n, c->d_scratch, 0);
    if (e) return e;
    return (int)cudaMemcpy(out, c->d_scratch, 7 * sizeof(double), cudaMemcpyDeviceToHost);
}

// ---- state I/O ----------------------------------------------------------------
namespace {
struct CkptHeader {       // explicit fixed-width fields: the file format does not follow the ABI struct
    char magic[8];        // "BHB200\0\0"
    int32_t version;      // 2
    int32_t header_bytes; // sizeof(CkptHeader)
    int64_t n;
    int64_t steps;
    float theta, G, dt, softening, max_speed;
    int32_t key_bits, leaf_cap;
    int32_t reserved;
};
const int32_t kCkptVersion = 2;
const char kMagic[8] = {'B', 'H', 'B', '2', '0', '0', 0, 0};
}  // namespace

int bh_dump_text(bh_ctx* c, const char* path) {
    if (!c || !path) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    const size_t n = (size_t)c->n;
    std::vector<float> a[6];
    for (auto& v : a) v.resize(n);
    int e = bh_export_soa_host(c, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), nullptr,
                               nullptr, nullptr);
    if (e) return e;
    FILE* f = fopen(path, "w");
    if (!f) return BH_E_IO;
    // header as output_bh.txt:1-4
    fprintf(f, "# Barnes-Hut N-Body Simulation Results\n");
    fprintf(f, "# Final positions and velocities after %lld steps\n", (long long)c->steps);
    fprintf(f, "# Bodies: %lld, Theta: %.2f, dt: %.3f\n", (long long)c->n, c->prm.theta, c->prm.dt);
    fprintf(f, "# Format: x y z vx vy vz\n");
    for (size_t i = 0; i < n; ++i)
        fprintf(f, "%.6f %.6f %.6f %.6f %.6f %.6f\n", a[0][i], a[1][i], a[2][i], a[3][i], a[4][i], a[5][i]);
    return fclose(f) == 0 ? 0 : BH_E_IO;
}

int bh_save_checkpoint(bh_ctx* c, const char* path) {
    if (!c || !path) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    const size_t n = (size_t)c->n;
    std::vector<float4> posm(n), vel(n);
    std::vector<int32_t> ids(n);
    BH_CUDA_TRY(cudaMemcpy(posm.data(), c->posm, n * 16, cudaMemcpyDeviceToHost));
    BH_CUDA_TRY(cudaMemcpy(vel.data(), c->vel, n * 16, cudaMemcpyDeviceToHost));
    BH_CUDA_TRY(cudaMemcpy(ids.data(), c->ids, n * 4, cudaMemcpyDeviceToHost));
    CkptHeader h{};
    memcpy(h.magic, kMagic, 8);
    h.version = kCkptVersion; h.header_bytes = (int32_t)sizeof(CkptHeader); h.n = c->n; h.steps = c->steps;
    h.theta = c->prm.theta; h.G = c->prm.G; h.dt = c->prm.dt; h.softening = c->prm.softening; h.max_speed = c->prm.max_speed;
    h.key_bits = c->prm.key_bits; h.leaf_cap = c->prm.leaf_cap;
    FILE* f = fopen(path, "wb");
    if (!f) return BH_E_IO;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(posm.data(), 16, n, f) == n && fwrite(vel.data(), 16, n, f) == n &&
              fwrite(ids.data(), 4, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : BH_E_IO;
}

int bh_load_checkpoint(bh_ctx* c, const char* path) {
    if (!c || !path) return BH_E_INVAL;
    FILE* f = fopen(path, "rb");
    if (!f) return BH_E_IO;
    CkptHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0 || h.version != kCkptVersion ||
        h.header_bytes != (int32_t)sizeof(CkptHeader) || h.n <= 0 || h.n > c->n_max) {
        fclose(f);
        return BH_E_IO;
    }
    // the same checks bh_create applies; a different key width or leaf capacity builds a different tree, which
    // would break the bit-for-bit resume
    if (!(h.softening > 0.0f) || !(h.theta >= 0.0f) || !(h.max_speed > 0.0f)) { fclose(f); return BH_E_IO; }
    if (h.key_bits != c->prm.key_bits || h.leaf_cap != c->prm.leaf_cap) { fclose(f); return BH_E_UNSUPPORTED; }
    const size_t n = (size_t)h.n;
    std::vector<float4> posm(n), vel(n);
    std::vector<int32_t> ids(n);
    bool ok = fread(posm.data(), 16, n, f) == n && fread(vel.data(), 16, n, f) == n && fread(ids.data(), 4, n, f) == n;
    fclose(f);
    if (!ok) return BH_E_IO;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    BH_CUDA_TRY(cudaMemcpy(c->posm, posm.data(), n * 16, cudaMemcpyHostToDevice));
    BH_CUDA_TRY(cudaMemcpy(c->vel, vel.data(), n * 16, cudaMemcpyHostToDevice));
    BH_CUDA_TRY(cudaMemcpy(c->ids, ids.data(), n * 4, cudaMemcpyHostToDevice));
    // simulation parameters travel with the state; flags and tuning knobs stay the context's own
    c->prm.theta = h.theta; c->prm.G = h.G; c->prm.dt = h.dt;
    c->prm.softening = h.softening; c->prm.max_speed = h.max_speed;
    c->n = h.n; c->steps = h.steps; c->have_state = true; c->have_sorted = false; c->bbox_fresh = false;
    { int eg = set_ghosts_possible(c, true); if (eg) return eg; }   // ids come from the file
    default_slice(c);
    BH_CUDA_TRY(cudaMemset((char*)c->sc + offsetof(BhDevScalars, epoch), 0, sizeof(BhDevScalars) - offsetof(BhDevScalars, epoch)));
    BH_CUDA_TRY(cudaMemset(c->heavy_flag, 0, 8 * (size_t)c->max_chunks));
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }   // theta/dt may have changed
    for (auto& g : c->half_exec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    return 0;
}

int bh_energy(bh_ctx* c, double* kinetic, double* potential) {
    if (!c || !kinetic || !potential) return BH_E_INVAL;
    if (!c->have_state) return BH_E_STATE;
    BH_CUDA_TRY(cudaSetDevice(c->device));
    BH_CUDA_TRY(cudaDeviceSynchronize());
    int e = bh_energy_launch(c->posm, c->vel, c->n, c->prm.softening, c->prm.G, c->d_scratch, 0);
    if (e) return e;
    double h[2];
    BH_CUDA_TRY(cudaMemcpy(h, c->d_scratch, 16, cudaMemcpyDeviceToHost));
    *kinetic = h[0]; *potential = h[1];
    return 0;
}

}  // extern "C"
