// extern "C" surface of libfdfd_b200.so; declarations and reference citations in include/fdfd_b200.h
#include <mutex>
#include <vector>
#include "krylov.cuh"
#include "mode.cuh"
#include "zgemm.cuh"
#include "../../include/fdfd_b200.h"

namespace {
struct DevBuf {
    cplx* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t count) {
        FDFD_CHECK(cudaMalloc(&p, sizeof(cplx) * count));
        return 0;
    }
};
}  // namespace

ZgemmTiming g_zgemm_timing;
int g_zgemm_variant = 0;
thread_local int g_zgemm_max_ctas = 148;          // per host thread: in-process ranks factorise concurrently
thread_local ZgemmHelper* g_zgemm_helper = nullptr;
unsigned* zgemm_tile_counter(cudaStream_t stream) {
    // ring of counters per device: launches in flight at the same time (main / side streams, several solver handles,
    // in-process ranks on several threads) never share one
    constexpr int RING = 4096, MAXDEV = 16;
    static std::mutex mu;
    static unsigned* pool[MAXDEV] = {nullptr};
    static int next[MAXDEV] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;   // falls back to static hand-out
    std::lock_guard<std::mutex> lock(mu);
    if (!pool[dev] && cudaMalloc(&pool[dev], sizeof(unsigned) * RING) != cudaSuccess) { pool[dev] = nullptr; return nullptr; }
    unsigned* c = pool[dev] + (next[dev]++ % RING);
    cudaMemsetAsync(c, 0, sizeof(unsigned), stream);
    return c;
}

// register-resident DMMA loop: the practical FP64 tensor-pipe ceiling at the clocks the board runs at
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
/* the same probe with the SM clock it ran at measured INSIDE the kernel (clock64 cycles / globaltimer ns of one
 * resident warp), so the result can be held against the pipe rate at the clock the board actually sustained:
 * out[0] = TFLOP/s, out[1] = SM MHz during the probe, out[2] = ms, out[3] = real flops */
template <int NACC>
__global__ void __launch_bounds__(1024) dmma_probe_clocked_kernel(double* out, unsigned long long* clk, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
    unsigned long long t0 = 0, c0 = 0;
    if (threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (threadIdx.x == 0) {
        unsigned long long c1 = clock64(), t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        clk[2 * blockIdx.x] = c1 - c0 + (s == 123.456 ? 1 : 0);
        clk[2 * blockIdx.x + 1] = t1 - t0;
    }
    if (s == 123.456) out[0] = s;
}
/* DMMA issue-ORDER probe: the 3M inner loop of the GEMM (2 x 4 accumulator tiles x 3 terms = 24 accumulators, operands
 * in registers, two k-steps) with the same 48 DMMAs issued in different orders.
 *   0: the k-step order of the GEMM (k-step outer; T1/T2 of every tile, then T3 of every tile)
 *   1: accumulator-major (every accumulator takes its two k-steps back to back)
 *   2: term-major inside a k-step (all T1, all T2, all T3)
 *   3: tile-major inside a k-step (T1, T2, T3 of a tile back to back) */
template <int PATTERN>
__global__ void __launch_bounds__(256) dmma_pattern_kernel(double* out, unsigned long long* clk, int iters) {
    double t[3][2][4][2];
    double a[2][2][3], b[2][4][3];       // [k-step][tile][re, im, re + im]
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { t[q][i][j][0] = threadIdx.x * 1e-9 * (q + 1); t[q][i][j][1] = 0.0; }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int q = 0; q < 3; ++q) a[k][i][q] = 1.0 + 1e-12 * (threadIdx.x + 7 * k + 3 * i + q);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 3; ++q) b[k][j][q] = 1.0 - 1e-12 * (threadIdx.x + 5 * k + 11 * j + q);
    }
    unsigned long long t0 = 0, c0 = 0;
    if (threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
    }
    for (int it = 0; it < iters; ++it) {
        if (PATTERN == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        dmma884(t[0][i][j][0], t[0][i][j][1], a[k][i][0], b[k][j][0]);
                        dmma884(t[1][i][j][0], t[1][i][j][1], a[k][i][1], b[k][j][1]);
                    }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(t[2][i][j][0], t[2][i][j][1], a[k][i][2], b[k][j][2]);
            }
        } else if (PATTERN == 1) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int k = 0; k < 2; ++k) dmma884(t[q][i][j][0], t[q][i][j][1], a[k][i][q], b[k][j][q]);
        } else if (PATTERN == 2) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma884(t[q][i][j][0], t[q][i][j][1], a[k][i][q], b[k][j][q]);
        } else {
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int q = 0; q < 3; ++q) dmma884(t[q][i][j][0], t[q][i][j][1], a[k][i][q], b[k][j][q]);
        }
        // keep the operands live and changing (one cheap op per iteration, as the fragment loads of the GEMM do)
        a[0][0][0] += 1e-13;
    }
    double sacc = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sacc += t[q][i][j][0] + t[q][i][j][1];
    if (threadIdx.x == 0) {
        unsigned long long c1 = clock64(), t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        clk[2 * blockIdx.x] = c1 - c0 + (sacc == 123.456 ? 1 : 0);
        clk[2 * blockIdx.x + 1] = t1 - t0;
    }
    if (sacc == 123.456) out[0] = sacc;
}

/* the GEMM's k-step verbatim (fragments by LDS.128 from a shared-memory tile, re + im sums by DADD, 24 DMMAs) in an
 * endless k loop over ONE resident tile: no global loads, no cp.async, MODE 0: no barrier at all; MODE 1: a CTA barrier
 * every 8 k-steps (the k-tile cadence of the kernel); MODE 2: barrier + 16 cp.async of 16 bytes per thread per k-tile
 * into a second buffer (the ring's traffic without its latency). */
template <int MODE>
__global__ void __launch_bounds__(256, 1) dmma_smem_kernel(double* out, unsigned long long* clk, const double2* src, int iters) {
    constexpr int BK = 32, LDA = BK + 4;
    extern __shared__ __align__(16) unsigned char ps_smem[];
    double2* As = reinterpret_cast<double2*>(ps_smem);           // [64][LDA]
    double2* Bs = As + 64 * LDA;                                   // [64][LDA]
    double2* Dump = Bs + 64 * LDA;                                 // cp.async landing zone, 2 x 64 x LDA
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp / 2, wn = warp % 2, gq = lane >> 2, tq = lane & 3;
    for (int i = tid; i < 2 * 64 * LDA; i += 256) As[i] = make_double2(1.0 + 1e-9 * i, 1.0 - 1e-9 * i);
    __syncthreads();
    double t1[2][4][2], t2[2][4][2], t3[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { t1[i][j][0] = t1[i][j][1] = t2[i][j][0] = t2[i][j][1] = t3[i][j][0] = t3[i][j][1] = 0.0; }
    unsigned long long t0 = 0, c0 = 0;
    if (tid == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
    }
    struct Frag { double2 a[2], b[4]; double as_[2], bs_[4]; };
    auto load_frag = [&](Frag& f, int kk) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { f.a[mt] = As[(wm * 16 + mt * 8 + gq) * LDA + kk + tq]; f.as_[mt] = f.a[mt].x + f.a[mt].y; }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { f.b[nt] = Bs[(wn * 32 + nt * 8 + gq) * LDA + kk + tq]; f.bs_[nt] = f.b[nt].x + f.b[nt].y; }
    };
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 1) __syncthreads();
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) cp_async16(Dump + (i * 256 + tid) % (2 * 64 * LDA), src + (i * 256 + tid), true);
            cp_async_commit();
            cp_async_wait<1>();
        }
        Frag f[2];
        load_frag(f[0], 0);
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            if (ks + 1 < BK / 4) load_frag(f[(ks + 1) & 1], 4 * (ks + 1));
            const Frag& c = f[ks & 1];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    dmma884(t1[mt][nt][0], t1[mt][nt][1], c.a[mt].x, c.b[nt].x);
                    dmma884(t2[mt][nt][0], t2[mt][nt][1], c.a[mt].y, c.b[nt].y);
                }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(t3[mt][nt][0], t3[mt][nt][1], c.as_[mt], c.bs_[nt]);
        }
    }
    cp_async_wait<0>();
    double sacc = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sacc += t1[i][j][0] + t1[i][j][1] + t2[i][j][0] + t2[i][j][1] + t3[i][j][0] + t3[i][j][1];
    if (tid == 0) {
        unsigned long long c1 = clock64(), tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        clk[2 * blockIdx.x] = c1 - c0 + (sacc == 123.456 ? 1 : 0);
        clk[2 * blockIdx.x + 1] = tt - t0;
    }
    if (sacc == 123.456) out[0] = sacc;
}

extern "C" {

int fdfd_dmma_smem_probe(int mode, double* out4) {
    double* d = nullptr;
    double2* src = nullptr;
    unsigned long long* clk = nullptr;
    FDFD_CHECK(cudaMalloc(&d, sizeof(double)));
    FDFD_CHECK(cudaMalloc(&src, sizeof(double2) * 4096));
    FDFD_CHECK(cudaMemset(src, 0, sizeof(double2) * 4096));
    FDFD_CHECK(cudaMalloc(&clk, sizeof(unsigned long long) * 2 * 148));
    cudaEvent_t e0, e1;
    FDFD_CHECK(cudaEventCreate(&e0));
    FDFD_CHECK(cudaEventCreate(&e1));
    const int iters = 4000;
    const size_t sm = sizeof(double2) * 4 * 64 * 36;
    auto launch = [&](int it) {
        if (mode == 0) { cudaFuncSetAttribute(dmma_smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); dmma_smem_kernel<0><<<148, 256, sm>>>(d, clk, src, it); }
        else if (mode == 1) { cudaFuncSetAttribute(dmma_smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); dmma_smem_kernel<1><<<148, 256, sm>>>(d, clk, src, it); }
        else { cudaFuncSetAttribute(dmma_smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); dmma_smem_kernel<2><<<148, 256, sm>>>(d, clk, src, it); }
    };
    launch(50);
    FDFD_CHECK(cudaDeviceSynchronize());
    FDFD_CHECK(cudaEventRecord(e0));
    launch(iters);
    FDFD_CHECK(cudaEventRecord(e1));
    FDFD_CHECK(cudaEventSynchronize(e1));
    FDFD_CHECK(cudaGetLastError());
    float ms = 0;
    FDFD_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long h[2 * 148];
    FDFD_CHECK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double mhz = 0;
    for (int i = 0; i < 148; ++i) mhz += h[2 * i + 1] ? (double)h[2 * i] / (double)h[2 * i + 1] * 1e3 : 0.0;
    const double fl = 148.0 * 8 * (double)iters * 8.0 * 24.0 * 512.0;
    out4[0] = fl / (ms * 1e-3) / 1e12;
    out4[1] = mhz / 148.0;
    out4[2] = ms;
    out4[3] = out4[0] / (148 * 128 * out4[1] * 1e-6);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d); cudaFree(clk); cudaFree(src);
    return 0;
}

int fdfd_dmma_pattern_probe(int pattern, int warps, double* out4) {
    if (warps < 1 || warps > 8) FDFD_FAIL("warps per SM: 1..8");
    double* d = nullptr;
    unsigned long long* clk = nullptr;
    FDFD_CHECK(cudaMalloc(&d, sizeof(double)));
    FDFD_CHECK(cudaMalloc(&clk, sizeof(unsigned long long) * 2 * 148));
    cudaEvent_t e0, e1;
    FDFD_CHECK(cudaEventCreate(&e0));
    FDFD_CHECK(cudaEventCreate(&e1));
    const int iters = 8000;
    auto launch = [&](int it) {
        if (pattern == 0) dmma_pattern_kernel<0><<<148, warps * 32>>>(d, clk, it);
        else if (pattern == 1) dmma_pattern_kernel<1><<<148, warps * 32>>>(d, clk, it);
        else if (pattern == 2) dmma_pattern_kernel<2><<<148, warps * 32>>>(d, clk, it);
        else dmma_pattern_kernel<3><<<148, warps * 32>>>(d, clk, it);
    };
    launch(100);
    FDFD_CHECK(cudaDeviceSynchronize());
    FDFD_CHECK(cudaEventRecord(e0));
    launch(iters);
    FDFD_CHECK(cudaEventRecord(e1));
    FDFD_CHECK(cudaEventSynchronize(e1));
    FDFD_CHECK(cudaGetLastError());
    float ms = 0;
    FDFD_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long h[2 * 148];
    FDFD_CHECK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double mhz = 0;
    for (int i = 0; i < 148; ++i) mhz += h[2 * i + 1] ? (double)h[2 * i] / (double)h[2 * i + 1] * 1e3 : 0.0;
    const double fl = 148.0 * warps * (double)iters * 48.0 * 512.0;
    out4[0] = fl / (ms * 1e-3) / 1e12;
    out4[1] = mhz / 148.0;
    out4[2] = ms;
    out4[3] = out4[0] / (148 * 128 * out4[1] * 1e-6);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d); cudaFree(clk);
    return 0;
}

int fdfd_dmma_peak(double* tflops) {
    double* d = nullptr;
    FDFD_CHECK(cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    FDFD_CHECK(cudaEventCreate(&e0));
    FDFD_CHECK(cudaEventCreate(&e1));
    const int iters = 4000, blocks = 148 * 4;
    dmma_peak_kernel<<<blocks, 256>>>(d, 100);
    FDFD_CHECK(cudaDeviceSynchronize());
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        FDFD_CHECK(cudaEventRecord(e0));
        dmma_peak_kernel<<<blocks, 256>>>(d, iters);
        FDFD_CHECK(cudaEventRecord(e1));
        FDFD_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        FDFD_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = (double)blocks * 8 /*warps*/ * iters * 16.0 * 512.0;   // 8x8x4 MACs x 2 per DMMA
        best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return 0;
}

/* DMMA issue-rate probe: `warps` warps per SM (one CTA per SM), `nacc` independent accumulators */
int fdfd_dmma_probe(int warps, int nacc, double* tflops) {
    double* d = nullptr;
    FDFD_CHECK(cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    FDFD_CHECK(cudaEventCreate(&e0));
    FDFD_CHECK(cudaEventCreate(&e1));
    const int iters = 20000 / nacc;
    auto launch = [&](int it) {
        if (nacc == 1) dmma_probe_kernel<1><<<148, warps * 32>>>(d, it);
        else if (nacc == 2) dmma_probe_kernel<2><<<148, warps * 32>>>(d, it);
        else if (nacc == 4) dmma_probe_kernel<4><<<148, warps * 32>>>(d, it);
        else if (nacc == 8) dmma_probe_kernel<8><<<148, warps * 32>>>(d, it);
        else dmma_probe_kernel<16><<<148, warps * 32>>>(d, it);
    };
    launch(10);
    FDFD_CHECK(cudaDeviceSynchronize());
    FDFD_CHECK(cudaEventRecord(e0));
    launch(iters);
    FDFD_CHECK(cudaEventRecord(e1));
    FDFD_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    FDFD_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    int na = nacc >= 16 ? 16 : nacc;
    *tflops = 148.0 * warps * iters * na * 512.0 / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return 0;
}

int fdfd_dmma_probe_clocked(int warps, int nacc, double* out4) {
    if (warps < 1 || warps > 32) FDFD_FAIL("warps per SM: 1..32");
    double* d = nullptr;
    unsigned long long* clk = nullptr;
    FDFD_CHECK(cudaMalloc(&d, sizeof(double)));
    FDFD_CHECK(cudaMalloc(&clk, sizeof(unsigned long long) * 2 * 148));
    cudaEvent_t e0, e1;
    FDFD_CHECK(cudaEventCreate(&e0));
    FDFD_CHECK(cudaEventCreate(&e1));
    const int na = nacc >= 16 ? 16 : nacc >= 8 ? 8 : nacc >= 4 ? 4 : nacc >= 2 ? 2 : 1;
    const int iters = 400000 / na;                      // ~50 ms: long enough for the clock to settle
    auto launch = [&](int it) {
        if (na == 1) dmma_probe_clocked_kernel<1><<<148, warps * 32>>>(d, clk, it);
        else if (na == 2) dmma_probe_clocked_kernel<2><<<148, warps * 32>>>(d, clk, it);
        else if (na == 4) dmma_probe_clocked_kernel<4><<<148, warps * 32>>>(d, clk, it);
        else if (na == 8) dmma_probe_clocked_kernel<8><<<148, warps * 32>>>(d, clk, it);
        else dmma_probe_clocked_kernel<16><<<148, warps * 32>>>(d, clk, it);
    };
    launch(1000);
    FDFD_CHECK(cudaDeviceSynchronize());
    FDFD_CHECK(cudaEventRecord(e0));
    launch(iters);
    FDFD_CHECK(cudaEventRecord(e1));
    FDFD_CHECK(cudaEventSynchronize(e1));
    FDFD_CHECK(cudaGetLastError());
    float ms = 0;
    FDFD_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long h[2 * 148];
    FDFD_CHECK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double mhz = 0;
    for (int i = 0; i < 148; ++i) mhz += h[2 * i + 1] ? (double)h[2 * i] / (double)h[2 * i + 1] * 1e3 : 0.0;
    const double fl = 148.0 * warps * (double)iters * na * 512.0;
    out4[0] = fl / (ms * 1e-3) / 1e12;
    out4[1] = mhz / 148.0;
    out4[2] = ms;
    out4[3] = fl;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d); cudaFree(clk);
    return 0;
}

int fdfd_gemm_timing(int enable) {
    for (cudaEvent_t e : g_zgemm_timing.ev) cudaEventDestroy(e);
    g_zgemm_timing.ev.clear(); g_zgemm_timing.flops.clear(); g_zgemm_timing.big.clear(); g_zgemm_timing.tflops_exec.clear();
    g_zgemm_timing.on = enable != 0;
    return 0;
}
/* totals since fdfd_gemm_timing(1): out[0..2] = ms, real flops, launches of the 64x64-tile kernel;
 * out[3..5] the same for the 32x32-tile kernel.  Synchronises the device. */
int fdfd_gemm_timing_read(double* out) {
    FDFD_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < 6; ++i) out[i] = 0;
    for (size_t i = 0; i < g_zgemm_timing.flops.size(); ++i) {
        float ms = 0;
        FDFD_CHECK(cudaEventElapsedTime(&ms, g_zgemm_timing.ev[2 * i], g_zgemm_timing.ev[2 * i + 1]));
        int o = g_zgemm_timing.big[i] ? 0 : 3;
        out[o] += ms; out[o + 1] += g_zgemm_timing.flops[i]; out[o + 2] += 1;
    }
    return 0;
}
/* real flops the tensor pipe executed in the launches timed so far (6 per complex multiply-add with the 3M form) */
int fdfd_gemm_timing_exec_flops(double* out) {
    double t = 0;
    for (double v : g_zgemm_timing.tflops_exec) t += v;
    *out = t;
    return 0;
}
int fdfd_phase_timing(int enable) {
    for (cudaEvent_t e : g_phase_timing.ev) cudaEventDestroy(e);
    g_phase_timing.ev.clear(); g_phase_timing.cat.clear(); g_phase_timing.lvl.clear();
    g_phase_timing.on = enable != 0;
    g_phase_timing.level = -1;
    return 0;
}
/* per-phase totals in ms since fdfd_phase_timing(1): assemble, pivot, panel, rowgemm, copy, update,
 * expand, solve_fwd, solve_bwd, stencil, ggemm, schur, small (13 doubles). Synchronises the device. */
int fdfd_phase_timing_read(double* out13) {
    FDFD_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < PH_COUNT; ++i) out13[i] = 0;
    for (size_t i = 0; i < g_phase_timing.cat.size(); ++i) {
        float ms = 0;
        FDFD_CHECK(cudaEventElapsedTime(&ms, g_phase_timing.ev[2 * i], g_phase_timing.ev[2 * i + 1]));
        out13[g_phase_timing.cat[i]] += ms;
    }
    return 0;
}
/* the same totals split by elimination-tree level: out[level * 13 + phase], levels 0..max_levels-1 */
int fdfd_phase_timing_read_levels(double* out, int max_levels) {
    FDFD_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < max_levels * PH_COUNT; ++i) out[i] = 0;
    for (size_t i = 0; i < g_phase_timing.cat.size(); ++i) {
        int l = g_phase_timing.lvl[i];
        if (l < 0 || l >= max_levels) continue;
        float ms = 0;
        FDFD_CHECK(cudaEventElapsedTime(&ms, g_phase_timing.ev[2 * i], g_phase_timing.ev[2 * i + 1]));
        out[l * PH_COUNT + g_phase_timing.cat[i]] += ms;
    }
    return 0;
}
int fdfd_timer_start(fdfd_op* op) {
    if (!op->ev0) { FDFD_CHECK(cudaEventCreate(&op->ev0)); FDFD_CHECK(cudaEventCreate(&op->ev1)); }
    FDFD_CHECK(cudaEventRecord(op->ev0, op->stream));
    return 0;
}
int fdfd_timer_stop(fdfd_op* op, double* ms) {
    FDFD_CHECK(cudaEventRecord(op->ev1, op->stream));
    FDFD_CHECK(cudaEventSynchronize(op->ev1));
    float t = 0;
    FDFD_CHECK(cudaEventElapsedTime(&t, op->ev0, op->ev1));
    *ms = t;
    return 0;
}

int fdfd_version(void) { return 100; }
const char* fdfd_last_error(void) { return g_fdfd_err; }

int fdfd_device_count(int* count) { FDFD_CHECK(cudaGetDeviceCount(count)); return 0; }
int fdfd_set_device(int device) { FDFD_CHECK(cudaSetDevice(device)); return 0; }
int fdfd_mem_info(double* free_bytes, double* total_bytes) {
    size_t f, t;
    FDFD_CHECK(cudaMemGetInfo(&f, &t));
    *free_bytes = (double)f; *total_bytes = (double)t;
    return 0;
}
int fdfd_malloc(void** dev_ptr, double bytes) { FDFD_CHECK(cudaMalloc(dev_ptr, (size_t)bytes)); return 0; }
int fdfd_free(void* dev_ptr) { FDFD_CHECK(cudaFree(dev_ptr)); return 0; }
int fdfd_memcpy_h2d(void* dev, const void* host, double bytes) {
    FDFD_CHECK(cudaMemcpy(dev, host, (size_t)bytes, cudaMemcpyHostToDevice));
    return 0;
}
int fdfd_memcpy_d2h(void* host, const void* dev, double bytes) {
    FDFD_CHECK(cudaMemcpy(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost));
    return 0;
}
double fdfd_launch_count(int reset) {
    double v = (double)g_fdfd_launches;
    if (reset) g_fdfd_launches = 0;
    return v;
}
int fdfd_host_register(void* host, double bytes) {
    FDFD_CHECK(cudaHostRegister(host, (size_t)bytes, cudaHostRegisterDefault));
    return 0;
}
int fdfd_host_alloc(void** host_ptr, double bytes) {
    FDFD_CHECK(cudaHostAlloc(host_ptr, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}
int fdfd_host_free(void* host_ptr) { FDFD_CHECK(cudaFreeHost(host_ptr)); return 0; }
int fdfd_host_unregister(void* host) { FDFD_CHECK(cudaHostUnregister(host)); return 0; }
int fdfd_op_sync(fdfd_op* op) { FDFD_CHECK(cudaStreamSynchronize(op->stream)); return 0; }

int fdfd_op_create(fdfd_op** out, int nx, int ny, double omega, double dl, int npml_x, int npml_y, int pol,
                   double L0) {
    if (!(omega > 0)) FDFD_FAIL("omega must be positive");
    if (!(dl > 0)) FDFD_FAIL("dl must be positive");
    if (!(L0 > 0)) FDFD_FAIL("L0 must be positive");
    if (npml_x < 0 || npml_y < 0) FDFD_FAIL("NPML entries must be >= 0");
    return op_create(out, nx, ny, omega, dl, npml_x, npml_y, pol, L0);
}
void fdfd_op_destroy(fdfd_op* op) { op_destroy(op); }

int fdfd_op_assemble_dev(fdfd_op* op, const void* d_eps_r, const void* d_eps_nl, int averaging) {
    return op_assemble_dev(op, (const cplx*)d_eps_r, (const cplx*)d_eps_nl, averaging);
}
int fdfd_op_assemble_host(fdfd_op* op, const double* eps_r, const double* eps_nl, int averaging) {
    size_t n = op->n();
    FDFD_CHECK(cudaMemcpyAsync(op->eps_r, eps_r, sizeof(cplx) * n, cudaMemcpyHostToDevice, op->stream));
    if (eps_nl) FDFD_CHECK(cudaMemcpyAsync(op->eps_nl, eps_nl, sizeof(cplx) * n, cudaMemcpyHostToDevice, op->stream));
    if (op_assemble_dev(op, op->eps_r, eps_nl ? op->eps_nl : nullptr, averaging)) return -1;
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    return 0;
}
int fdfd_op_assemble_host_f64(fdfd_op* op, const double* eps_r_f64, int averaging) {
    // real permittivity: half the PCIe bytes, widened to complex on the device
    size_t n = op->n();
    cplx* io = nullptr;
    if (op_io_buffer(op, &io)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(io, eps_r_f64, sizeof(double) * n, cudaMemcpyHostToDevice, op->stream));
    if (op_scale_expand(op, io, 1, make_double2(1.0, 0.0), op->eps_r, n)) return -1;
    if (op_assemble_dev(op, op->eps_r, nullptr, averaging)) return -1;
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    return 0;
}
int fdfd_op_get_sfactors_host(fdfd_op* op, double* isxf, double* isxb, double* isyf, double* isyb) {
    FDFD_CHECK(cudaMemcpy(isxf, op->isxf, sizeof(cplx) * op->nx, cudaMemcpyDeviceToHost));
    FDFD_CHECK(cudaMemcpy(isxb, op->isxb, sizeof(cplx) * op->nx, cudaMemcpyDeviceToHost));
    FDFD_CHECK(cudaMemcpy(isyf, op->isyf, sizeof(cplx) * op->ny, cudaMemcpyDeviceToHost));
    FDFD_CHECK(cudaMemcpy(isyb, op->isyb, sizeof(cplx) * op->ny, cudaMemcpyDeviceToHost));
    return 0;
}
int fdfd_op_get_planes_host(fdfd_op* op, double* planes) {
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    FDFD_CHECK(cudaMemcpy(planes, op->planes, sizeof(cplx) * op->n() * 5, cudaMemcpyDeviceToHost));
    return 0;
}
int fdfd_op_apply_dev(fdfd_op* op, const void* d_x, void* d_y, int nvec, int fused) {
    return fused ? op_apply_fused(op, (const cplx*)d_x, (cplx*)d_y, nvec)
                 : op_apply_planes(op, (const cplx*)d_x, (cplx*)d_y, nvec);
}
int fdfd_op_apply_host(fdfd_op* op, const double* x, double* y, int nvec, int fused) {
    size_t cnt = op->n() * nvec;
    DevBuf dx, dy;
    if (dx.alloc(cnt) || dy.alloc(cnt)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(dx.p, x, sizeof(cplx) * cnt, cudaMemcpyHostToDevice, op->stream));
    if (fdfd_op_apply_dev(op, dx.p, dy.p, nvec, fused)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(y, dy.p, sizeof(cplx) * cnt, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    return 0;
}
int fdfd_op_derive_fields_dev(fdfd_op* op, const void* d_x, void* d_f1, void* d_f2, int averaging) {
    return op_derive_fields(op, (const cplx*)d_x, (cplx*)d_f1, (cplx*)d_f2, averaging);
}
int fdfd_op_derive_fields_host(fdfd_op* op, const double* x, double* f1, double* f2, int averaging) {
    size_t n = op->n();
    DevBuf buf;
    if (buf.alloc(3 * n)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(buf.p, x, sizeof(cplx) * n, cudaMemcpyHostToDevice, op->stream));
    if (op_derive_fields(op, buf.p, buf.p + n, buf.p + 2 * n, averaging)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(f1, buf.p + n, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaMemcpyAsync(f2, buf.p + 2 * n, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    return 0;
}

int fdfd_direct_create(fdfd_direct** out, int nx, int ny, int tile) { return nd_create(out, nx, ny, tile); }
int fdfd_direct_add_level(fdfd_direct* s, const fdfd_level_desc* d) {
    NdLevelDesc x;
    x.kind = d->kind; x.nb = d->nb; x.kmax = d->kmax; x.mmax = d->mmax; x.ncls = d->ncls;
    x.child_mmax = d->child_mmax; x.cls = d->cls; x.k_cls = d->k_cls; x.ch1 = d->ch1; x.ch2 = d->ch2;
    x.c1map = d->c1map; x.c2map = d->c2map; x.x0 = d->x0; x.y0 = d->y0; x.slot_lx = d->slot_lx;
    x.slot_ly = d->slot_ly; x.slot_right = d->slot_right; x.slot_up = d->slot_up;
    x.send_to = d->send_to; x.recv_from = d->recv_from;
    return nd_add_level(s, &x);
}
void fdfd_direct_destroy(fdfd_direct* s) { nd_destroy(s); }
int fdfd_direct_factor(fdfd_direct* s, fdfd_op* op) {
    if (op->halo) FDFD_FAIL("the direct solver takes the whole-grid operator (sharded: fdfd_direct_set_comm), not a slab");
    return nd_factor(s, op);
}
int fdfd_direct_stats(fdfd_direct* s, double* factor_bytes, double* factor_flops) {
    *factor_bytes = (double)s->factor_bytes;
    *factor_flops = s->factor_flops;
    return 0;
}
int fdfd_direct_solve_dev(fdfd_direct* s, fdfd_op* op, const void* d_b, void* d_x, int nrhs, int max_refine,
                          double tol, double* relres, int* refine_steps) {
    double rr = -1.0;
    int steps = 0;
    if (max_refine < 0) {          // plain substitution, no residual evaluation
        if (nd_solve(s, op, (const cplx*)d_b, (cplx*)d_x, nrhs)) return -1;
    } else if (refine_solve(s, op, (const cplx*)d_b, (cplx*)d_x, nrhs, max_refine, tol, &rr, &steps)) {
        return -1;
    }
    if (relres) *relres = rr;
    if (refine_steps) *refine_steps = steps;
    return 0;
}
int fdfd_direct_solve_host(fdfd_direct* s, fdfd_op* op, const double* b, double* x, int nrhs, int max_refine,
                           double tol, double* relres, int* refine_steps) {
    size_t cnt = op->n() * nrhs;
    DevBuf db, dx;
    if (db.alloc(cnt) || dx.alloc(cnt)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(db.p, b, sizeof(cplx) * cnt, cudaMemcpyHostToDevice, op->stream));
    if (fdfd_direct_solve_dev(s, op, db.p, dx.p, nrhs, max_refine, tol, relres, refine_steps)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(x, dx.p, sizeof(cplx) * cnt, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    return 0;
}

int fdfd_solve_fields_host(fdfd_direct* s, fdfd_op* op, const double* src, int src_is_real, double scale_re,
                           double scale_im, double* x, double* f1, double* f2, int averaging, int max_refine,
                           double tol, double* relres, int* refine_steps) {
    const size_t n = op->n();
    cplx* io = nullptr;
    if (op_io_buffer(op, &io)) return -1;
    cplx *b = io, *xx = io + n, *g1 = io + 2 * n, *g2 = io + 3 * n;
    FDFD_CHECK(cudaMemcpyAsync(g1, src, (src_is_real ? sizeof(double) : sizeof(cplx)) * n, cudaMemcpyHostToDevice,
                               op->stream));
    if (op_scale_expand(op, g1, src_is_real, make_double2(scale_re, scale_im), b, n)) return -1;
    if (fdfd_direct_solve_dev(s, op, b, xx, 1, max_refine, tol, relres, refine_steps)) return -1;
    if (op_derive_fields(op, xx, g1, g2, averaging)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(x, xx, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaMemcpyAsync(f1, g1, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaMemcpyAsync(f2, g2, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    return 0;
}

int fdfd_factor_solve_fields_host(fdfd_direct* s, fdfd_op* op, const double* src, int src_is_real, double scale_re,
                                  double scale_im, double* x, double* f1, double* f2, int averaging, int max_refine,
                                  double tol, double* relres, int* refine_steps, double* factor_ms) {
    // solve_fields with the factorisation inside the call: the factorisation is QUEUED first, the source then crosses
    // PCIe on a second stream while it runs (a pageable 134 MB array at 4096^2 keeps the host busy for ~12 ms), and the
    // singular-pivot flag is read at the call's own final synchronisation.  An all-zero source still returns zero
    // fields (linalg.py:129-130), the host does not have to scan for it.
    if (op->halo) FDFD_FAIL("the direct solver takes the whole-grid operator, not a slab");
    const size_t n = op->n();
    cplx* io = nullptr;
    if (op_io_buffer(op, &io)) return -1;
    cplx *b = io, *xx = io + n, *g1 = io + 2 * n, *g2 = io + 3 * n;
    if (!op->up_stream) {
        FDFD_CHECK(cudaStreamCreateWithFlags(&op->up_stream, cudaStreamNonBlocking));
        FDFD_CHECK(cudaEventCreateWithFlags(&op->ev_up, cudaEventDisableTiming));
    }
    // factors of another operator, or of this operator before its last assembly, are not this system's factors
    const bool factor = !s->factored || s->fact_op != op || s->fact_version != op->version;
    struct EventPair {                                   // destroyed on every exit path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } ev;
    cudaEvent_t &e0 = ev.a, &e1 = ev.b;
    if (factor) {
        FDFD_CHECK(cudaEventCreate(&e0));
        FDFD_CHECK(cudaEventCreate(&e1));
        FDFD_CHECK(cudaEventRecord(e0, op->stream));
        if (nd_factor(s, op, true)) return -1;
        FDFD_CHECK(cudaEventRecord(e1, op->stream));
    }
    FDFD_CHECK(cudaMemcpyAsync(g1, src, (src_is_real ? sizeof(double) : sizeof(cplx)) * n, cudaMemcpyHostToDevice,
                               op->up_stream));
    FDFD_CHECK(cudaEventRecord(op->ev_up, op->up_stream));
    FDFD_CHECK(cudaStreamWaitEvent(op->stream, op->ev_up, 0));
    if (op_scale_expand(op, g1, src_is_real, make_double2(scale_re, scale_im), b, n)) return -1;
    int rc = fdfd_direct_solve_dev(s, op, b, xx, 1, max_refine, tol, relres, refine_steps);
    if (factor) {
        // (the solve synchronised the stream more than once; a singular pivot shows up here first, whatever the solve said)
        if (nd_factor_check(s, op)) rc = -1;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && factor_ms) *factor_ms = ms;
    } else if (factor_ms) *factor_ms = 0.0;
    if (rc) return -1;
    if (op_derive_fields(op, xx, g1, g2, averaging)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(x, xx, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaMemcpyAsync(f1, g1, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaMemcpyAsync(f2, g2, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    return 0;
}
int fdfd_op_eps_flags(fdfd_op* op, int* flags) { return op_eps_flags(op, flags); }

int fdfd_comm_load(const char* libnccl_path) { return comm_load(libnccl_path); }
int fdfd_comm_unique_id(void* id128) { return comm_unique_id(id128); }
int fdfd_comm_create(fdfd_comm** out, const void* id128, int rank, int world) {
    return comm_create(out, id128, rank, world);
}
void fdfd_comm_destroy(fdfd_comm* c) { comm_destroy(c); }
int fdfd_comm_allreduce_sum_dev(fdfd_comm* c, fdfd_op* op, void* d_buf, double count) {
    return comm_allreduce_sum(c, d_buf, (size_t)count, op->stream);
}
int fdfd_slab_op_create(fdfd_op** out, fdfd_comm* comm, int gnx, int ny, int x0, int nxl, double omega, double dl,
                        int npml_x, int npml_y, int pol, double L0) {
    if (!(omega > 0) || !(dl > 0) || !(L0 > 0)) FDFD_FAIL("omega, dl and L0 must be positive");
    if (npml_x < 0 || npml_y < 0) FDFD_FAIL("NPML entries must be >= 0");
    return op_create_slab(out, comm, gnx, ny, x0, nxl, omega, dl, npml_x, npml_y, pol, L0);
}
int fdfd_schwarz_sub_create(fdfd_op** out, fdfd_op* slab, int overlap, int npml_sub) {
    return op_create_schwarz_sub(out, slab, overlap, npml_sub);
}
int fdfd_slab_set_schwarz(fdfd_op* slab, fdfd_op* sub, fdfd_direct* sub_factors, int overlap, int npml_sub) {
    if (sub && (!sub_factors || !sub_factors->factored)) FDFD_FAIL("factorise the Schwarz subdomain before attaching it");
    return schwarz_attach(slab, sub, sub_factors, overlap, npml_sub);
}
int fdfd_comm_create_local(fdfd_comm** out, int world) { return comm_create_local(out, world); }
void fdfd_comm_abort(fdfd_comm* c) { comm_abort(c); }
int fdfd_direct_add_dist_front(fdfd_direct* s, const fdfd_dist_front_desc* d) {
    NdDistFrontDesc x;
    x.level0 = d->level0; x.nsteps = d->nsteps; x.gbase = d->gbase; x.gsize = d->gsize; x.n = d->n; x.nblk = d->nblk;
    x.bstart = d->bstart; x.bowner = d->bowner; x.mc1 = d->mc1; x.mc2 = d->mc2; x.inv1 = d->inv1; x.inv2 = d->inv2;
    return nd_add_dist_front(s, &x);
}
int fdfd_direct_set_comm(fdfd_direct* s, fdfd_comm* c) {
    s->comm = c;
    s->factored = false;
    return 0;
}

int fdfd_krylov_solve_dev(fdfd_op* op, fdfd_direct* precond, const void* d_b, void* d_x, int method, double tol,
                          int maxiter, int fused, int check_every, const void* d_c12, int real_inner, int* iters,
                          double* relres, int* converged) {
    KrylovResult r;
    int rc;
    if (method == 0)
        rc = krylov_bicgstab(op, precond, (const cplx*)d_b, (cplx*)d_x, tol, maxiter, fused, check_every,
                             (const cplx*)d_c12, real_inner, &r);
    else if (method == 1) {
        if (precond || d_c12) FDFD_FAIL("COCG takes neither a preconditioner nor an anti-linear term");
        rc = krylov_cocg(op, (const cplx*)d_b, (cplx*)d_x, tol, maxiter, fused, check_every, &r);
    } else if (method == 2) {
        if (d_c12) FDFD_FAIL("GMRES takes no anti-linear term");
        rc = krylov_gmres(op, precond, (const cplx*)d_b, (cplx*)d_x, tol, maxiter, check_every, fused, &r);
    } else FDFD_FAIL("unknown Krylov method %d", method);
    if (rc) return -1;
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    if (iters) *iters = r.iters;
    if (relres) *relres = r.relres;
    if (converged) *converged = r.converged;
    return 0;
}
int fdfd_krylov_solve_host(fdfd_op* op, fdfd_direct* precond, const double* b, double* x, int method, double tol,
                           int maxiter, int fused, int check_every, const double* c12, int real_inner, int* iters,
                           double* relres, int* converged) {
    size_t n = op->n();
    DevBuf db, dx, dc;
    if (db.alloc(n) || dx.alloc(n)) return -1;
    if (c12) {
        if (dc.alloc(n)) return -1;
        FDFD_CHECK(cudaMemcpyAsync(dc.p, c12, sizeof(cplx) * n, cudaMemcpyHostToDevice, op->stream));
    }
    FDFD_CHECK(cudaMemcpyAsync(db.p, b, sizeof(cplx) * n, cudaMemcpyHostToDevice, op->stream));
    FDFD_CHECK(cudaMemcpyAsync(dx.p, x, sizeof(cplx) * n, cudaMemcpyHostToDevice, op->stream));   // initial guess
    if (fdfd_krylov_solve_dev(op, precond, db.p, dx.p, method, tol, maxiter, fused, check_every, dc.p, real_inner,
                              iters, relres, converged))
        return -1;
    FDFD_CHECK(cudaMemcpy(x, dx.p, sizeof(cplx) * n, cudaMemcpyDeviceToHost));
    return 0;
}

int fdfd_nl_solve_host(fdfd_op* op_nl, fdfd_direct* lin, fdfd_direct* work, const double* K, const double* b, double* E,
                       int method, int strategy, double conv_threshold, int max_iter, double* conv, int* iters,
                       int* inner_iters) {
    if (method != 0 && method != 1) FDFD_FAIL("nonlinear method: 0 = born, 1 = newton");
    if (max_iter < 1) FDFD_FAIL("max_iter must be >= 1");
    const size_t n = op_nl->n();
    DevBuf dk, db, de;
    if (dk.alloc(n) || db.alloc(n) || de.alloc(n)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(dk.p, K, sizeof(cplx) * n, cudaMemcpyHostToDevice, op_nl->stream));
    FDFD_CHECK(cudaMemcpyAsync(db.p, b, sizeof(cplx) * n, cudaMemcpyHostToDevice, op_nl->stream));
    FDFD_CHECK(cudaMemcpyAsync(de.p, E, sizeof(cplx) * n, cudaMemcpyHostToDevice, op_nl->stream));
    if (nl_solve(op_nl, lin, work, dk.p, db.p, de.p, method, strategy, conv_threshold, max_iter, conv, iters, inner_iters))
        return -1;
    FDFD_CHECK(cudaMemcpyAsync(E, de.p, sizeof(cplx) * n, cudaMemcpyDeviceToHost, op_nl->stream));
    FDFD_CHECK(cudaStreamSynchronize(op_nl->stream));
    return 0;
}

/* ---- complex64 storage (fp64 arithmetic): stencil and Krylov loop ---- */
int fdfd_op_apply_dev_c64(fdfd_op* op, const void* d_x, void* d_y, int fused) {
    return fused ? op_apply_fused_t<cplx32>(op, (const cplx32*)d_x, (cplx32*)d_y, 1)
                 : op_apply_planes_t<cplx32>(op, (const cplx32*)d_x, (cplx32*)d_y, 1);
}
int fdfd_op_apply_host_c64(fdfd_op* op, const float* x, float* y, int fused) {
    size_t cnt = op->n();
    cplx32 *dx = nullptr, *dy = nullptr;
    FDFD_CHECK(cudaMalloc(&dx, sizeof(cplx32) * cnt));
    FDFD_CHECK(cudaMalloc(&dy, sizeof(cplx32) * cnt));
    int rc = 0;
    if (cudaMemcpyAsync(dx, x, sizeof(cplx32) * cnt, cudaMemcpyHostToDevice, op->stream) != cudaSuccess) rc = -1;
    if (!rc && cudaMemsetAsync(dy, 0, sizeof(cplx32) * cnt, op->stream) != cudaSuccess) rc = -1;
    if (!rc) rc = fdfd_op_apply_dev_c64(op, dx, dy, fused);
    if (!rc && cudaMemcpyAsync(y, dy, sizeof(cplx32) * cnt, cudaMemcpyDeviceToHost, op->stream) != cudaSuccess) rc = -1;
    if (cudaStreamSynchronize(op->stream) != cudaSuccess) rc = -1;
    cudaFree(dx); cudaFree(dy);
    if (rc && !g_fdfd_err[0]) snprintf(g_fdfd_err, sizeof(g_fdfd_err), "complex64 apply failed");
    return rc;
}
int fdfd_krylov_solve_dev_c64(fdfd_op* op, const void* d_b, void* d_x, int method, double tol, int maxiter, int fused,
                              int check_every, int* iters, double* relres, int* converged) {
    KrylovResult r;
    int rc;
    if (method == 0) rc = krylov_bicgstab_c64(op, (const cplx32*)d_b, (cplx32*)d_x, tol, maxiter, fused, check_every, &r);
    else if (method == 1) rc = krylov_cocg_c64(op, (const cplx32*)d_b, (cplx32*)d_x, tol, maxiter, fused, check_every, &r);
    else FDFD_FAIL("unknown Krylov method %d", method);
    if (rc) return -1;
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    if (iters) *iters = r.iters;
    if (relres) *relres = r.relres;
    if (converged) *converged = r.converged;
    return 0;
}
int fdfd_krylov_solve_host_c64(fdfd_op* op, const float* b, float* x, int method, double tol, int maxiter, int fused,
                               int check_every, int* iters, double* relres, int* converged) {
    size_t n = op->n();
    cplx32 *db = nullptr, *dx = nullptr;
    FDFD_CHECK(cudaMalloc(&db, sizeof(cplx32) * n));
    FDFD_CHECK(cudaMalloc(&dx, sizeof(cplx32) * n));
    int rc = 0;
    if (cudaMemcpyAsync(db, b, sizeof(cplx32) * n, cudaMemcpyHostToDevice, op->stream) != cudaSuccess) rc = -1;
    if (!rc && cudaMemcpyAsync(dx, x, sizeof(cplx32) * n, cudaMemcpyHostToDevice, op->stream) != cudaSuccess) rc = -1;
    if (!rc) rc = fdfd_krylov_solve_dev_c64(op, db, dx, method, tol, maxiter, fused, check_every, iters, relres, converged);
    if (!rc && cudaMemcpy(x, dx, sizeof(cplx32) * n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
    cudaFree(db); cudaFree(dx);
    if (rc && !g_fdfd_err[0]) snprintf(g_fdfd_err, sizeof(g_fdfd_err), "complex64 Krylov solve failed");
    return rc;
}

int fdfd_zgemm_batched_host(const double* A, const double* B, double* Cm, int M, int N, int K, int batch, int mode,
                            int transb, int lower) {
    // test hook: C[b] = A[b] op(B[b]) (mode 0) or C[b] -= A[b] op(B[b]) (mode 1), dense row-major, packed batches
    DevBuf a, b, c;
    size_t sa = (size_t)M * K, sb = (size_t)K * N, sc = (size_t)M * N;
    if (a.alloc(sa * batch) || b.alloc(sb * batch) || c.alloc(sc * batch)) return -1;
    FDFD_CHECK(cudaMemcpy(a.p, A, sizeof(cplx) * sa * batch, cudaMemcpyHostToDevice));
    FDFD_CHECK(cudaMemcpy(b.p, B, sizeof(cplx) * sb * batch, cudaMemcpyHostToDevice));
    FDFD_CHECK(cudaMemcpy(c.p, Cm, sizeof(cplx) * sc * batch, cudaMemcpyHostToDevice));
    GemmBatch g;
    g.A = a.p; g.sA = sa; g.lda = K; g.B = b.p; g.sB = sb; g.ldb = transb ? K : N; g.C = c.p; g.sC = sc; g.ldc = N;
    g.M = M; g.N = N; g.K = K; g.batch = batch; g.mode = mode; g.transb = transb; g.lower = lower;
    if (zgemm_batched(g, 0)) return -1;
    FDFD_CHECK(cudaDeviceSynchronize());
    FDFD_CHECK(cudaMemcpy(Cm, c.p, sizeof(cplx) * sc * batch, cudaMemcpyDeviceToHost));
    return 0;
}

extern int g_fused_rows, g_fused_rows32, g_hz_halo_lanes, g_hz_chunk;
int fdfd_stencil_set_variant(int rows_per_thread, int complex64) {
    if (rows_per_thread != 2 && rows_per_thread != 4 && rows_per_thread != 8) FDFD_FAIL("rows per thread: 2, 4 or 8");
    if (complex64) g_fused_rows32 = rows_per_thread;
    else g_fused_rows = rows_per_thread;
    return 0;
}
int fdfd_stencil_set_hz_variant(int chunk_rows, int halo_lanes) {
    if (chunk_rows != 0 && chunk_rows != -4 && chunk_rows != -8 && (chunk_rows < 4 || chunk_rows > 1024 || (chunk_rows & 3)))
        FDFD_FAIL("Hz stencil: 0 (auto), a multiple of 4 in 4 ... 1024 (rows per CTA), or -4 / -8 (one-shot kernel)");
    g_hz_chunk = chunk_rows;
    g_hz_halo_lanes = halo_lanes != 0;
    return 0;
}
int fdfd_zgemm_set_variant(int v) { g_zgemm_variant = v; return 0; }
extern int g_small_front_enabled;
int fdfd_direct_set_small_fronts(int enable) { g_small_front_enabled = enable != 0; return 0; }

int fdfd_zgemm_bench(int M, int N, int K, int batch, int mode, int transb, int lower, int iters,
                     double* ms_per_launch) {
    // device-only timing of one GEMM shape (uninitialised-but-finite operands), CUDA events
    DevBuf a, b, c;
    size_t sa = (size_t)M * K, sb = (size_t)K * N, sc = (size_t)M * N;
    if (a.alloc(sa * batch) || b.alloc(sb * batch) || c.alloc(sc * batch)) return -1;
    FDFD_CHECK(cudaMemset(a.p, 0, sizeof(cplx) * sa * batch));
    FDFD_CHECK(cudaMemset(b.p, 0, sizeof(cplx) * sb * batch));
    FDFD_CHECK(cudaMemset(c.p, 0, sizeof(cplx) * sc * batch));
    GemmBatch g;
    g.A = a.p; g.sA = sa; g.lda = K; g.B = b.p; g.sB = sb; g.ldb = transb ? K : N; g.C = c.p; g.sC = sc; g.ldc = N;
    g.M = M; g.N = N; g.K = K; g.batch = batch; g.mode = mode; g.transb = transb; g.lower = lower;
    cudaEvent_t e0, e1;
    FDFD_CHECK(cudaEventCreate(&e0));
    FDFD_CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) if (zgemm_batched(g, 0)) return -1;
    FDFD_CHECK(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) if (zgemm_batched(g, 0)) return -1;
    FDFD_CHECK(cudaEventRecord(e1, 0));
    FDFD_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    FDFD_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_launch = ms / iters;
    return 0;
}

int fdfd_mode_solve_host(const double* eps_line, int n, double omega, double dl, int pol, double L0, double neff,
                         int order, int averaged, double* vals, double* vecs) {
    return mode_solve(eps_line, n, omega, dl, pol, L0, neff, order, averaged, vals, vecs);
}

}  // extern "C"
