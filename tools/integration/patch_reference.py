"""Compiled proof of INTEGRATION.md: applies the patch a maintainer would make to the reference's
nbody_v5_bench.cu — by line number, so that no reference text is stored in this repo — and builds the result
against libbh.so.

    python tools/integration/patch_reference.py        # here (needs /root/reference): writes the patched
                                                       # source to /tmp and the binary to oracle/_ref/
    oracle/_ref/nbody_v5_bench_libbh                   # on a GPU box: the reference's own main(), banner and
                                                       # frame table, its simulationStep() now one bh_step

Edits (line numbers of /root/reference/nbody_v5_bench.cu, checked by the anchors below):
  1          + #include "bh.h" and the context pointer
  255-283    simulationStep(): body replaced by bh_step(g_bh, 1, nullptr)          (INTEGRATION.md "Patch")
  336        after the H2D copies (bench:329-335): bh_create + bh_import_soa
  353        1000 frames -> BH_FRAMES frames (the only knob; the reference hard-codes it)
  369        before the cudaFrees: bh_export_soa into the reference's own arrays, checksum print, bh_destroy
Everything else — IC generation, the 16 cudaMallocs, the copies, the event pair, the table — is the reference's.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REF_SRC", "/root/reference/nbody_v5_bench.cu")
OUT_SRC = "/tmp/nbody_v5_bench_libbh.cu"
OUT_BIN = os.path.join(ROOT, "oracle", "_ref", "nbody_v5_bench_libbh")


def main():
    if not os.path.exists(REF):
        print("reference source not present; keeping prebuilt oracle/_ref/nbody_v5_bench_libbh if any")
        return 0
    src = open(REF).read().split("\n")
    L = lambda n: src[n - 1]   # 1-based like the citations
    assert L(255).startswith("void simulationStep()") and L(256).strip() == "{" and L(283).strip() == "}", "simulationStep moved"
    assert "cudaMemcpy(d_mass" in L(335) and "frame < 1000" in L(353) and "// Cleanup" in L(369), "main() moved"
    out = ['#include "bh.h"', "static bh_ctx* g_bh = nullptr;", "#ifndef BH_FRAMES", "#define BH_FRAMES 100", "#endif"]
    for n in range(1, len(src) + 1):
        if 257 <= n <= 282:
            if n == 257:
                out.append("    bh_step(g_bh, 1, /*stream=*/nullptr);   // legacy default stream, like the reference; asynchronous")
            continue
        line = L(n)
        if n == 353:
            line = line.replace("1000", "BH_FRAMES")
        out.append(line)
        if n == 335:
            out += ["    bh_params bh_p; bh_default_params(&bh_p);   // THETA/G_CONST/DT/SOFTENING/MAX_SPEED of bench:13-18",
                    "    if (bh_create(&g_bh, N, &bh_p, 0) != 0) return 1;",
                    "    if (bh_import_soa(g_bh, d_posX, d_posY, d_posZ, d_velX, d_velY, d_velZ, d_mass, N, nullptr) != 0) return 1;"]
        if n == 368:
            out += ["    // what display() reads after the step (nbody_v5.cu:335): the SoA arrays, body i in slot i",
                    "    if (bh_export_soa(g_bh, d_posX, d_posY, d_posZ, d_velX, d_velY, d_velZ, d_accX, d_accY, d_accZ, nullptr) != 0) return 1;",
                    "    { float* h = new float[N]; double s = 0; cudaMemcpy(h, d_posX, N * 4, cudaMemcpyDeviceToHost);",
                    "      for (int i = 0; i < N; i++) s += h[i]; delete[] h;",
                    '      printf("libbh: sum(posX) after %d frames = %.6f, interactions/body = %.1f, device error flag = %lld\\n", BH_FRAMES, s,',
                    "             (double)(bh_stat(g_bh, BH_STAT_INTERACTIONS_CELL) + bh_stat(g_bh, BH_STAT_INTERACTIONS_BODY)) / N,",
                    "             (long long)bh_stat(g_bh, BH_STAT_DEVICE_ERROR)); }",
                    "    bh_destroy(g_bh);"]
    open(OUT_SRC, "w").write("\n".join(out))
    os.makedirs(os.path.dirname(OUT_BIN), exist_ok=True)
    pkg = os.path.join(ROOT, "nbody-barnes-hut-cuda_b200")
    cmd = ["nvcc", "-o", OUT_BIN, OUT_SRC, "-arch=sm_100", "-O3", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
           "-L" + pkg, "-lbh", "-Xlinker", "-rpath=$ORIGIN/../../nbody-barnes-hut-cuda_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-3000:])
        return 1
    print("built", OUT_BIN)
    return 0


if __name__ == "__main__":
    sys.exit(main())
