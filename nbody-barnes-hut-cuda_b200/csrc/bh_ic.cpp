// bh_ic.cpp — synthetic initial conditions (host side, no CUDA).
//
// refdisk  : the generator in main(), /root/reference/nbody_v5_bench.cu:294-308
//            (srand(42) + five rand() draws per body in the order r, angle, z,
//            mass, vz).  Uses the C library's rand(), like the reference, so on
//            glibc it yields the reference's bodies bit for bit.
// uniform  : BASELINE.json configs[0] — cube, masses as bench:302, v = 0.
// plummer  : BASELINE.json configs[2..3] — Aarseth/Henon/Wielen sampling.
// The last two use a counter-based SplitMix64 stream so they are identical on
// every libc.
#include "../../include/bh.h"

#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <thread>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // uniform in [0,1) with 24 bits — exactly representable as float
    float uniform() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
    double uniform_d() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

inline float unit_rand() { return (float)rand() / RAND_MAX; }  // bench:297 idiom

}  // namespace

extern "C" int bh_ic_refdisk(int64_t n, unsigned seed,
                             float* px, float* py, float* pz,
                             float* vx, float* vy, float* vz, float* mass) {
    if (n < 0 || !px || !py || !pz || !vx || !vy || !vz || !mass) return BH_E_INVAL;
    const float G = 0.5f;  // G_CONST, bench:14
    srand(seed);           // bench:294 uses 42
    for (int64_t i = 0; i < n; ++i) {
        float r = 200.0f + unit_rand() * 1500.0f;                     // bench:297
        // bench:298 — float*float, then promoted to double by M_PI, stored as float
        float a = (float)((double)(unit_rand() * 2.0f) * M_PI);
        float ca = cosf(a), sa = sinf(a);   // cos/sin(float) resolve to the float overloads
        px[i] = r * ca;                                               // bench:299
        py[i] = r * sa;                                               // bench:300
        pz[i] = (unit_rand() - 0.5f) * (r * 0.05f);                   // bench:301
        mass[i] = 2.0f + unit_rand() * 5.0f;                          // bench:302
        float inside = 50000.0f + r * 100.0f;                         // bench:303
        float vmag = sqrtf(G * inside / r);                           // bench:304
        vx[i] = -sa * vmag;                                           // bench:305
        vy[i] = ca * vmag;                                            // bench:306
        vz[i] = (unit_rand() - 0.5f) * 2.0f;                          // bench:307
    }
    return 0;
}

extern "C" int bh_ic_uniform_cube(int64_t n, uint64_t seed, float half_edge,
                                  float* px, float* py, float* pz,
                                  float* vx, float* vy, float* vz, float* mass) {
    if (n < 0 || !px || !py || !pz || !vx || !vy || !vz || !mass) return BH_E_INVAL;
    SplitMix64 rng(seed);
    for (int64_t i = 0; i < n; ++i) {
        px[i] = (rng.uniform() * 2.0f - 1.0f) * half_edge;
        py[i] = (rng.uniform() * 2.0f - 1.0f) * half_edge;
        pz[i] = (rng.uniform() * 2.0f - 1.0f) * half_edge;
        mass[i] = 2.0f + rng.uniform() * 5.0f;  // same mass law as bench:302
        vx[i] = vy[i] = vz[i] = 0.0f;
    }
    return 0;
}

extern "C" int bh_ic_plummer(int64_t n, uint64_t seed, float scale_a, float rcut_in_a,
                             float body_mass, float G,
                             float* px, float* py, float* pz,
                             float* vx, float* vy, float* vz, float* mass) {
    if (n < 0 || !px || !py || !pz || !vx || !vy || !vz || !mass) return BH_E_INVAL;
    if (!(scale_a > 0) || !(rcut_in_a > 0) || !(body_mass > 0)) return BH_E_INVAL;
    SplitMix64 rng(seed);
    const double a = scale_a, Mtot = (double)body_mass * (double)n;
    for (int64_t i = 0; i < n; ++i) {
        double r;
        do {  // radius from the cumulative mass profile M(r)/M = r^3 (r^2+a^2)^-3/2
            double u = rng.uniform_d();
            if (u < 1e-12) u = 1e-12;
            r = a / std::sqrt(std::pow(u, -2.0 / 3.0) - 1.0);
        } while (r > rcut_in_a * a);
        double cz = 2.0 * rng.uniform_d() - 1.0, ph = 2.0 * M_PI * rng.uniform_d();
        double sz = std::sqrt(1.0 - cz * cz);
        px[i] = (float)(r * sz * std::cos(ph));
        py[i] = (float)(r * sz * std::sin(ph));
        pz[i] = (float)(r * cz);
        // speed: q = v/v_esc sampled from g(q) = q^2 (1-q^2)^(7/2) by rejection
        double q, y;
        do {
            q = rng.uniform_d();
            y = 0.1 * rng.uniform_d();
        } while (y > q * q * std::pow(1.0 - q * q, 3.5));
        double vesc = std::sqrt(2.0 * G * Mtot / a) * std::pow(1.0 + (r * r) / (a * a), -0.25);
        double v = q * vesc;
        cz = 2.0 * rng.uniform_d() - 1.0;
        ph = 2.0 * M_PI * rng.uniform_d();
        sz = std::sqrt(1.0 - cz * cz);
        vx[i] = (float)(v * sz * std::cos(ph));
        vy[i] = (float)(v * sz * std::sin(ph));
        vz[i] = (float)(v * cz);
        mass[i] = body_mass;
    }
    return 0;
}

// BASELINE.json configs[4] / SURVEY §8d(5): two discs built like main()'s (bench:297-307), centred at
// -/+ (sep/2, 0, 0) and approaching with -/+ (vx, vy, 0).  Body i draws from its own SplitMix64 stream
// (seed, i), so the result does not depend on the thread count.
// Bodies [first, first + n) of that system into arrays of n floats (slot i - first): a rank of a multi-GPU run
// generates only its own share.
extern "C" int bh_ic_two_disks_range(int64_t first, int64_t n, uint64_t seed, float sep, float vx0, float vy0,
                                     float* px_, float* py_, float* pz_, float* vx_, float* vy_, float* vz_, float* mass_) {
    if (first < 0 || n < 0 || !px_ || !py_ || !pz_ || !vx_ || !vy_ || !vz_ || !mass_) return BH_E_INVAL;
    const float G = 0.5f;
    float *px = px_ - first, *py = py_ - first, *pz = pz_ - first, *vx = vx_ - first, *vy = vy_ - first, *vz = vz_ - first,
          *mass = mass_ - first;
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = first + lo; i < first + hi; ++i) {
            SplitMix64 rng(seed * 0x9E3779B97F4A7C15ull + (uint64_t)i * 0xD1B54A32D192ED03ull);
            const float side = (i & 1) ? 1.0f : -1.0f;
            float r = 200.0f + rng.uniform() * 1500.0f;
            float a = (float)((double)(rng.uniform() * 2.0f) * M_PI);
            float ca = cosf(a), sa = sinf(a);
            px[i] = r * ca + side * 0.5f * sep;
            py[i] = r * sa;
            pz[i] = (rng.uniform() - 0.5f) * (r * 0.05f);
            mass[i] = 2.0f + rng.uniform() * 5.0f;
            float vmag = sqrtf(G * (50000.0f + r * 100.0f) / r);
            vx[i] = -sa * vmag - side * vx0;
            vy[i] = ca * vmag - side * vy0;
            vz[i] = (rng.uniform() - 0.5f) * 2.0f;
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 16) nt = 16;
    if (n < 100000) { work(0, n); return 0; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& x : th) x.join();
    return 0;
}

extern "C" int bh_ic_two_disks(int64_t n, uint64_t seed, float sep, float vx0, float vy0,
                               float* px, float* py, float* pz, float* vx, float* vy, float* vz, float* mass) {
    return bh_ic_two_disks_range(0, n, seed, sep, vx0, vy0, px, py, pz, vx, vy, vz, mass);
}
