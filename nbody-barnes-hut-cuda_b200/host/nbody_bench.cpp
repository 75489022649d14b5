// nbody_bench — headless front-end on libbh.so that stands in for the reference's bench binary.
//
// Mirrors main() of /root/reference/nbody_v5_bench.cu:285-390: same banner, same
// "Frame | Trajanje (ms) | FPS" table from one cudaEvent pair per frame (bench:346-367), same disk
// initial condition (bench:294-308) — plus what README.md:23-29,56-60 promises and the reference never
// prints: per-phase times, transfer timings, body-steps/s and interactions/s.  Options replace the
// reference's edit-and-recompile knobs (N is a global at bench:31, 1000 frames at bench:353).
//
//   nbody_bench [--n N] [--frames F] [--ic refdisk|uniform|plummer|twodisk] [--theta T] [--key-bits 30|60]
//               [--quiet] [--phases] [--dump out.txt] [--save ckpt.bin] [--resume ckpt.bin] [--gpus G]
//
// --gpus G (Morton-slice mode, include/bh.h bh_mg_*): one PROCESS per GPU.  The process started by the user is
// rank 0; it re-executes itself G-1 times (--mg-rank r --mg-id FILE), writes the 128-byte NCCL id to FILE and
// every rank then runs the same frame loop on bh_mg_step; rank 0 prints the table.  No MPI, no Python.
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bh.h"

static void die(const char* what, int code) {
    fprintf(stderr, "%s failed: %d (%s)\n", what, code, bh_error_string(code));
    exit(1);
}
#define CHECK(call) do { int _e = (call); if (_e) die(#call, _e); } while (0)

int main(int argc, char** argv) {
    long long n = 1000000;   // README.md:23 (the code's global says 500000, bench:31)
    int frames = 100, quiet = 0, phases = 0, key_bits = 30, gpus = 1, mg_rank = 0;
    float theta = 0.5f;
    std::string ic = "refdisk", dump, save, resume, mg_id_file;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--n") n = atoll(next());
        else if (a == "--frames") frames = atoi(next());
        else if (a == "--ic") ic = next();
        else if (a == "--theta") theta = (float)atof(next());
        else if (a == "--key-bits") key_bits = atoi(next());
        else if (a == "--quiet") quiet = 1;
        else if (a == "--phases") phases = 1;
        else if (a == "--dump") dump = next();
        else if (a == "--save") save = next();
        else if (a == "--resume") resume = next();
        else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--mg-rank") mg_rank = atoi(next());      // internal: a worker started by rank 0
        else if (a == "--mg-id") mg_id_file = next();           // internal
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (gpus < 1 || gpus > 64 || mg_rank < 0 || mg_rank >= gpus) { fprintf(stderr, "bad --gpus / --mg-rank\n"); return 2; }
    // ---- multi-GPU: rank 0 starts the other ranks as processes of their own BEFORE touching CUDA
    std::vector<pid_t> workers;
    unsigned char mg_id[BH_MG_ID_BYTES];
    if (gpus > 1 && mg_rank == 0) {
        char tmpl[] = "/tmp/nbody_bench_mgid_XXXXXX";
        const int fd = mkstemp(tmpl);
        if (fd < 0) { perror("mkstemp"); return 1; }
        close(fd);
        mg_id_file = tmpl;
        for (int r = 1; r < gpus; ++r) {
            const pid_t pid = fork();
            if (pid < 0) { perror("fork"); return 1; }
            if (pid == 0) {
                std::vector<std::string> args(argv, argv + argc);
                args.push_back("--mg-rank"); args.push_back(std::to_string(r));
                args.push_back("--mg-id"); args.push_back(mg_id_file);
                args.push_back("--quiet");
                std::vector<char*> cargs;
                for (auto& x : args) cargs.push_back(const_cast<char*>(x.c_str()));
                cargs.push_back(nullptr);
                execv("/proc/self/exe", cargs.data());
                perror("execv");
                _exit(127);
            }
            workers.push_back(pid);
        }
        CHECK(bh_mg_unique_id(mg_id));
        FILE* f = fopen((mg_id_file + ".tmp").c_str(), "wb");
        if (!f || fwrite(mg_id, 1, BH_MG_ID_BYTES, f) != BH_MG_ID_BYTES || fclose(f) != 0) { fprintf(stderr, "cannot write the NCCL id\n"); return 1; }
        rename((mg_id_file + ".tmp").c_str(), mg_id_file.c_str());   // atomic: a reader sees all 128 bytes or nothing
    } else if (gpus > 1) {
        for (int tries = 0;; ++tries) {   // wait for rank 0's id
            FILE* f = fopen(mg_id_file.c_str(), "rb");
            size_t got = f ? fread(mg_id, 1, BH_MG_ID_BYTES, f) : 0;
            if (f) fclose(f);
            if (got == BH_MG_ID_BYTES) break;
            if (tries > 6000) { fprintf(stderr, "rank %d: no NCCL id from rank 0\n", mg_rank); return 1; }
            usleep(10000);
        }
    }
    const bool lead = mg_rank == 0;
    if (lead) printf("Pokretanje Benchmarka za N = %lld...\n", n);   // bench:287

    bh_params p;
    bh_default_params(&p);
    p.theta = theta;
    p.key_bits = key_bits;   // 60: the reference key + 10 more bits per axis (deep trees for > 10^7 bodies)
    if (phases) p.flags |= BH_FLAG_PHASE_TIMER;
    bh_ctx* ctx = nullptr;
    const int device = gpus > 1 ? mg_rank : 0;
    CHECK(bh_create(&ctx, n, &p, device));
    cudaSetDevice(device);

    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    float h2d_ms = 0.f, d2h_ms = 0.f;
    if (!resume.empty()) {
        CHECK(bh_load_checkpoint(ctx, resume.c_str()));
    } else {
        std::vector<float> a[7];
        for (auto& v : a) v.resize((size_t)n);
        if (ic == "refdisk") CHECK(bh_ic_refdisk(n, 42, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data()));
        else if (ic == "uniform") CHECK(bh_ic_uniform_cube(n, 42, 1000.0f, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data()));
        else if (ic == "plummer") CHECK(bh_ic_plummer(n, 42, 200.0f, 10.0f, 4.5f, 0.5f, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data()));
        else if (ic == "twodisk") CHECK(bh_ic_two_disks(n, 42, 4000.0f, 20.0f, 8.0f, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data()));
        else { fprintf(stderr, "unknown --ic %s\n", ic.c_str()); return 2; }
        cudaEventRecord(start);
        CHECK(bh_import_soa_host(ctx, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data(), n));
        cudaEventRecord(stop);
        cudaEventSynchronize(stop);
        cudaEventElapsedTime(&h2d_ms, start, stop);
    }

    bh_mg* mg = nullptr;
    if (gpus > 1) CHECK(bh_mg_create(&mg, ctx, mg_id, mg_rank, gpus, device));     // NCCL communicator + this rank's Morton slice
    if (lead) {
        printf("------------------------------------------\n");                   // bench:350
        printf("\n%-10s | %-15s | %-10s\n", "Frame", "Trajanje (ms)", "FPS");     // bench:351
    }
    double total_ms = 0.0, phase_sum[BH_PHASE_COUNT] = {0};
    for (int frame = 0; frame < frames; ++frame) {                                // bench:353-367
        cudaEventRecord(start);
        if (mg) { CHECK(bh_mg_step(mg, 1, nullptr)); CHECK(bh_mg_finish(mg, nullptr)); }   // sliced simulationStep() + gathers
        else CHECK(bh_step(ctx, 1, nullptr));                                     // simulationStep()
        cudaEventRecord(stop);
        cudaEventSynchronize(stop);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, start, stop);
        total_ms += ms;
        if (!quiet) printf("%-10d | %-15.3f | %-10.1f\n", frame, ms, 1000.0f / ms);
        if (phases) {
            float ph[BH_PHASE_COUNT];
            bh_phase_ms(ctx, ph);
            for (int k = 0; k < BH_PHASE_COUNT; ++k) phase_sum[k] += ph[k];
        }
    }
    long long err = bh_stat(ctx, BH_STAT_DEVICE_ERROR);
    if (err) { fprintf(stderr, "device error flag %lld\n", err); return 1; }
    if (!lead) {   // workers: done (the state is replicated; rank 0 reports and writes the files)
        bh_mg_destroy(mg);
        bh_destroy(ctx);
        return 0;
    }

    // README.md:56-60 — the per-section metrics
    const double inter = ((double)bh_stat(ctx, BH_STAT_INTERACTIONS_CELL) + (double)bh_stat(ctx, BH_STAT_INTERACTIONS_BODY)) *
                         (mg ? (double)gpus : 1.0);   // a rank counts its own slice; slices hold equal body counts
    printf("------------------------------------------\n");
    if (mg) printf("%d GPUs, one process each: replicated sort + tree, Morton-slice traversal, in-place NCCL all-gather\n", gpus);
    printf("frames %d  mean %.3f ms  %.1f FPS  %.1f M body-steps/s  %.1f G interactions/s (%.1f per body)\n", frames,
           total_ms / frames, 1000.0 * frames / total_ms, n * 1e-3 * frames / total_ms, inter * 1e-6 * frames / total_ms, inter / n);
    if (phases) {
        const char* names[] = {"morton keys", "sort+reorder", "octree build", "centre of mass", "force", "update"};
        for (int k = 0; k < BH_PHASE_TOTAL; ++k) printf("  %-16s %8.3f ms/frame\n", names[k], phase_sum[k] / frames);
    }
    std::vector<float> out((size_t)n * 6);
    cudaEventRecord(start);
    CHECK(bh_export_soa_host(ctx, &out[0], &out[n], &out[2 * n], &out[3 * n], &out[4 * n], &out[5 * n], nullptr, nullptr, nullptr));
    cudaEventRecord(stop);
    cudaEventSynchronize(stop);
    cudaEventElapsedTime(&d2h_ms, start, stop);
    printf("  transfers: host->device %.3f ms (28 B/body), device->host %.3f ms (24 B/body)\n", h2d_ms, d2h_ms);
    if (!dump.empty()) CHECK(bh_dump_text(ctx, dump.c_str()));
    if (!save.empty()) CHECK(bh_save_checkpoint(ctx, save.c_str()));
    bh_mg_destroy(mg);
    bh_destroy(ctx);
    int rc = 0;
    for (pid_t pid : workers) {
        int status = 0;
        if (waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = 1;
    }
    if (!mg_id_file.empty() && lead) unlink(mg_id_file.c_str());
    return rc;
}
