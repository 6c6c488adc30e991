// Shared device helpers for the fdfdpy_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

typedef double2 cplx;   // (re, im), layout-compatible with numpy complex128

extern thread_local char g_fdfd_err[512];
extern unsigned long long g_fdfd_launches;   // kernels launched by this library (bench bookkeeping)

#define FDFD_CHECK(call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            snprintf(g_fdfd_err, sizeof(g_fdfd_err), "%s:%d: %s -> %s", __FILE__, __LINE__, \
                     #call, cudaGetErrorString(e__));                                     \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

#define FDFD_FAIL(...)                                                 \
    do {                                                               \
        snprintf(g_fdfd_err, sizeof(g_fdfd_err), __VA_ARGS__);         \
        return -1;                                                     \
    } while (0)

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx cneg(cplx a) { return make_double2(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
// a += b*c
__host__ __device__ __forceinline__ void cfma(cplx& a, cplx b, cplx c) {
    a.x = fma(b.x, c.x, a.x);
    a.x = fma(-b.y, c.y, a.x);
    a.y = fma(b.x, c.y, a.y);
    a.y = fma(b.y, c.x, a.y);
}
__host__ __device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    // Smith's algorithm (no spurious overflow for the huge |S| values deep in the PML)
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return make_double2((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return make_double2((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
__host__ __device__ __forceinline__ cplx crecip(cplx b) { return cdiv(make_double2(1.0, 0.0), b); }

__device__ __forceinline__ cplx ldg_c(const cplx* p) { return __ldg(p); }

// Field vectors are stored as complex128 (cplx) or complex64 (cplx32); arithmetic is always fp64.
typedef float2 cplx32;
__device__ __forceinline__ cplx vload(const cplx* p) { return __ldg(p); }
__device__ __forceinline__ cplx vload(const cplx32* p) { float2 v = __ldg(p); return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ void vstore(cplx* p, cplx v) { *p = v; }
__device__ __forceinline__ void vstore(cplx32* p, cplx v) { *p = make_float2((float)v.x, (float)v.y); }

// ---- optional live phase timing (CUDA events on the launching stream), read by bench.py ----
enum FdfdPhase { PH_ASSEMBLE = 0, PH_PIVOT, PH_PANEL, PH_ROWGEMM, PH_COPY, PH_UPDATE, PH_EXTRACT, PH_SOLVE_FWD,
                 PH_SOLVE_BWD, PH_STENCIL, PH_GGEMM, PH_SCHUR, PH_SMALL, PH_COUNT };
struct PhaseTiming {
    bool on = false;
    int level = -1;                 // elimination-tree level the next scopes belong to (-1: none)
    std::vector<cudaEvent_t> ev;
    std::vector<int> cat;
    std::vector<int> lvl;
};
extern PhaseTiming g_phase_timing;
struct PhaseScope {
    cudaStream_t st;
    bool active;
    PhaseScope(int cat, cudaStream_t s) : st(s), active(g_phase_timing.on) {
        if (!active) return;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        g_phase_timing.ev.push_back(e0);
        g_phase_timing.ev.push_back(e1);
        g_phase_timing.cat.push_back(cat);
        g_phase_timing.lvl.push_back(g_phase_timing.level);
    }
    ~PhaseScope() {
        if (active) cudaEventRecord(g_phase_timing.ev.back(), st);
    }
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
