// Modal-source eigensolve on the device.  Replaces the ARPACK shift-invert call the reference
// makes from source/mode.py:92 (solver_eigs, linalg.py:104-115) for the 1-D waveguide operator
//     A = w^2 mu0' eps + Dxf Dxb                      (Ez, mode.py:85)
//     A = w^2 mu0' eps + eps Dxf ex^-1 Dxb            (Hz, mode.py:88)
// on the periodic source line.  A = E K with E = diag(eps) > 0 and K symmetric, so the similar
// symmetric matrix B = E^1/2 K E^1/2 is used; eigenvectors map back as u = E^1/2 w.
//
// The line is short (tens to a few thousand cells), so one CTA does everything:
//   dense LU with partial pivoting of (B - sigma I), then shift-invert subspace iteration with a
//   Rayleigh-Ritz projection (generalised p x p problem, Cholesky + cyclic Jacobi) per sweep.
#include <vector>
#include "mode.cuh"
#include "operator.cuh"

#define MODE_THREADS 256
#define MODE_PMAX 16

__device__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < MODE_THREADS / 32; ++i) t += red[i];
    return t;
}

// y = B x for the cyclic symmetric tridiagonal B (d on the diagonal, off[i] couples i and i+1 mod n)
__device__ __forceinline__ double tri_apply(const double* d, const double* off, const double* x, int i, int n) {
    int ip = i + 1 == n ? 0 : i + 1, im = i == 0 ? n - 1 : i - 1;
    double v = d[i] * x[i];
    if (n > 1) v += off[i] * x[ip] + off[im] * x[im];
    return v;
}

__global__ void __launch_bounds__(MODE_THREADS)
mode_kernel(const double* __restrict__ eps_line, int n, double omega, double dl, int pol, double L0, double neff,
            int order, int averaged, int p, double* __restrict__ M, int* __restrict__ piv, double* __restrict__ d,
            double* __restrict__ off, double* __restrict__ sq, double* __restrict__ X, double* __restrict__ Y,
            double* __restrict__ BY, double* __restrict__ vals, double* __restrict__ vecs, int* __restrict__ info) {
    __shared__ double red[MODE_THREADS / 32];
    __shared__ double G[MODE_PMAX][MODE_PMAX], H[MODE_PMAX][MODE_PMAX], V[MODE_PMAX][MODE_PMAX];
    __shared__ double theta[MODE_PMAX];
    __shared__ int s_piv, s_done;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = MODE_THREADS / 32;
    const double e0 = FDFD_EPS0 * L0, m0 = FDFD_MU0 * L0;
    const double sigma = (omega * sqrt(m0 * e0) * neff) * (omega * sqrt(m0 * e0) * neff);
    const double idl2 = 1.0 / (dl * dl);

    // ---- tridiagonal coefficients
    for (int i = tid; i < n; i += MODE_THREADS) {
        int ip = i + 1 == n ? 0 : i + 1, im = i == 0 ? n - 1 : i - 1;
        double e = e0 * eps_line[i], ep = e0 * eps_line[ip], em = e0 * eps_line[im];
        if (pol == 0) {
            d[i] = omega * omega * m0 * e - 2.0 * idl2;
            off[i] = idl2;
            sq[i] = 1.0;
        } else {
            double ex_i = averaged ? (em + e) / 2 : e;          // edge average on the lower face of i
            double ex_p = averaged ? (e + ep) / 2 : ep;         // ... of i+1
            d[i] = omega * omega * m0 * e - e * idl2 / ex_p - e * idl2 / ex_i;
            off[i] = sqrt(e * ep) * idl2 / ex_p;
            sq[i] = sqrt(e);
        }
    }
    __syncthreads();
    // ---- dense M = B - sigma I
    for (long long i = tid; i < (long long)n * n; i += MODE_THREADS) M[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += MODE_THREADS) M[(size_t)i * n + i] = d[i] - sigma;
    __syncthreads();
    if (n > 1) {
        // serial accumulation of the (few) off-diagonal entries keeps n == 2 (double wrap) exact
        for (int i = tid; i < n; i += MODE_THREADS) {
            int ip = i + 1 == n ? 0 : i + 1;
            atomicAdd(&M[(size_t)i * n + ip], off[i]);
            atomicAdd(&M[(size_t)ip * n + i], off[i]);
        }
    }
    __syncthreads();
    // ---- LU with partial pivoting (row swaps recorded in piv)
    for (int k = 0; k < n; ++k) {
        if (warp == 0) {
            double best = -1.0;
            int bi = k;
            for (int r = k + lane; r < n; r += 32) {
                double v = fabs(M[(size_t)r * n + k]);
                if (v > best) { best = v; bi = r; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_down_sync(0xffffffffu, best, o);
                int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                s_piv = bi;
                piv[k] = bi;
                if (!(best > 0.0)) *info = 1;
            }
        }
        __syncthreads();
        int pr = s_piv;
        if (pr != k)
            for (int c = tid; c < n; c += MODE_THREADS) {
                double a = M[(size_t)k * n + c];
                M[(size_t)k * n + c] = M[(size_t)pr * n + c];
                M[(size_t)pr * n + c] = a;
            }
        __syncthreads();
        double ip = 1.0 / M[(size_t)k * n + k];
        for (int r = k + 1 + tid; r < n; r += MODE_THREADS) M[(size_t)r * n + k] *= ip;
        __syncthreads();
        int rem = n - k - 1;
        for (long long e = tid; e < (long long)rem * rem; e += MODE_THREADS) {
            int r = k + 1 + (int)(e / rem), c = k + 1 + (int)(e % rem);
            M[(size_t)r * n + c] -= M[(size_t)r * n + k] * M[(size_t)k * n + c];
        }
        __syncthreads();
    }
    // ---- start vectors: smooth, linearly independent, deterministic
    for (int e = tid; e < n * p; e += MODE_THREADS) {
        int i = e / p, c = e % p;
        double t = (i + 0.5) / n;
        X[e] = cos(3.141592653589793 * c * t) + 0.01 * sin(12.9898 * (i + 1) * (c + 1));
    }
    if (tid == 0) s_done = 0;
    __syncthreads();

    for (int iter = 0; iter < 400; ++iter) {
        // ---- Y = (B - sigma)^-1 X : one warp per column, row-oriented substitution
        for (int c = warp; c < p; c += nwarp) {
            for (int i = lane; i < n; i += 32) Y[(size_t)i * p + c] = X[(size_t)i * p + c];
            __syncwarp();
            if (lane == 0)
                for (int k = 0; k < n; ++k) {
                    int pr = piv[k];
                    if (pr != k) {
                        double a = Y[(size_t)k * p + c];
                        Y[(size_t)k * p + c] = Y[(size_t)pr * p + c];
                        Y[(size_t)pr * p + c] = a;
                    }
                }
            __syncwarp();
            for (int i = 1; i < n; ++i) {          // L y = b (unit lower)
                double s = 0;
                for (int j = lane; j < i; j += 32) s += M[(size_t)i * n + j] * Y[(size_t)j * p + c];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
                if (lane == 0) Y[(size_t)i * p + c] -= s;
                __syncwarp();
            }
            for (int i = n - 1; i >= 0; --i) {     // U x = y
                double s = 0;
                for (int j = i + 1 + lane; j < n; j += 32) s += M[(size_t)i * n + j] * Y[(size_t)j * p + c];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
                if (lane == 0) Y[(size_t)i * p + c] = (Y[(size_t)i * p + c] - s) / M[(size_t)i * n + i];
                __syncwarp();
            }
            // scale the column to unit norm (keeps the Gram matrix well conditioned)
            double s = 0;
            for (int i = lane; i < n; i += 32) s += Y[(size_t)i * p + c] * Y[(size_t)i * p + c];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            s = __shfl_sync(0xffffffffu, s, 0);
            double inv = rsqrt(s);
            for (int i = lane; i < n; i += 32) Y[(size_t)i * p + c] *= inv;
        }
        __syncthreads();
        // ---- BY, then G = Y^T Y and H = Y^T B Y (one warp per entry of the upper triangle)
        for (int e = tid; e < n * p; e += MODE_THREADS) {
            int i = e / p, c = e % p;
            int ip = i + 1 == n ? 0 : i + 1, im = i == 0 ? n - 1 : i - 1;
            double v = d[i] * Y[(size_t)i * p + c];
            if (n > 2) v += off[i] * Y[(size_t)ip * p + c] + off[im] * Y[(size_t)im * p + c];
            else if (n == 2) v += (off[0] + off[1]) * Y[(size_t)ip * p + c];
            BY[e] = v;
        }
        __syncthreads();
        for (int e = warp; e < p * p; e += nwarp) {
            int a = e / p, b = e % p;
            if (b < a) continue;
            double g = 0, h = 0;
            for (int i = lane; i < n; i += 32) {
                double ya = Y[(size_t)i * p + a];
                g += ya * Y[(size_t)i * p + b];
                h += ya * BY[(size_t)i * p + b];
            }
            for (int o = 16; o > 0; o >>= 1) {
                g += __shfl_down_sync(0xffffffffu, g, o);
                h += __shfl_down_sync(0xffffffffu, h, o);
            }
            if (lane == 0) { G[a][b] = G[b][a] = g; H[a][b] = H[b][a] = h; }
        }
        __syncthreads();
        // ---- small generalised symmetric eigenproblem H v = theta G v  (thread 0)
        if (tid == 0) {
            // Cholesky G = L L^T (L stored in the lower triangle of G)
            for (int j = 0; j < p; ++j) {
                double s = G[j][j];
                for (int k = 0; k < j; ++k) s -= G[j][k] * G[j][k];
                if (!(s > 1e-28)) { s = 1e-28; }
                G[j][j] = sqrt(s);
                for (int i = j + 1; i < p; ++i) {
                    double t = G[i][j];
                    for (int k = 0; k < j; ++k) t -= G[i][k] * G[j][k];
                    G[i][j] = t / G[j][j];
                }
            }
            // C = L^-1 H L^-T  (in H)
            for (int j = 0; j < p; ++j)           // H <- L^-1 H
                for (int i = 0; i < p; ++i) {
                    double t = H[i][j];
                    for (int k = 0; k < i; ++k) t -= G[i][k] * H[k][j];
                    H[i][j] = t / G[i][i];
                }
            for (int i = 0; i < p; ++i)           // H <- H L^-T
                for (int j = 0; j < p; ++j) {
                    double t = H[i][j];
                    for (int k = 0; k < j; ++k) t -= H[i][k] * G[j][k];
                    H[i][j] = t / G[j][j];
                }
            for (int i = 0; i < p; ++i)
                for (int j = 0; j < p; ++j) V[i][j] = i == j ? 1.0 : 0.0;
            // cyclic Jacobi on the symmetric C
            for (int sweep = 0; sweep < 30; ++sweep) {
                double offn = 0;
                for (int i = 0; i < p; ++i)
                    for (int j = i + 1; j < p; ++j) offn += H[i][j] * H[i][j];
                double diagn = 0;
                for (int i = 0; i < p; ++i) diagn += H[i][i] * H[i][i];
                if (offn <= 1e-32 * diagn) break;
                for (int a = 0; a < p; ++a)
                    for (int b = a + 1; b < p; ++b) {
                        double apq = 0.5 * (H[a][b] + H[b][a]);
                        if (apq == 0.0) continue;
                        double tau = (H[b][b] - H[a][a]) / (2.0 * apq);
                        double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
                        for (int k = 0; k < p; ++k) {
                            double hka = H[k][a], hkb = H[k][b];
                            H[k][a] = cs * hka - sn * hkb;
                            H[k][b] = sn * hka + cs * hkb;
                        }
                        for (int k = 0; k < p; ++k) {
                            double hak = H[a][k], hbk = H[b][k];
                            H[a][k] = cs * hak - sn * hbk;
                            H[b][k] = sn * hak + cs * hbk;
                        }
                        for (int k = 0; k < p; ++k) {
                            double vka = V[k][a], vkb = V[k][b];
                            V[k][a] = cs * vka - sn * vkb;
                            V[k][b] = sn * vka + cs * vkb;
                        }
                    }
            }
            for (int j = 0; j < p; ++j) theta[j] = H[j][j];
            // V <- L^-T V  (back substitution), then sort columns by |theta - sigma|
            for (int j = 0; j < p; ++j)
                for (int i = p - 1; i >= 0; --i) {
                    double t = V[i][j];
                    for (int k = i + 1; k < p; ++k) t -= G[k][i] * V[k][j];
                    V[i][j] = t / G[i][i];
                }
            for (int a = 0; a < p; ++a) {
                int best = a;
                for (int b = a + 1; b < p; ++b)
                    if (fabs(theta[b] - sigma) < fabs(theta[best] - sigma)) best = b;
                if (best != a) {
                    double t = theta[a]; theta[a] = theta[best]; theta[best] = t;
                    for (int k = 0; k < p; ++k) { double v = V[k][a]; V[k][a] = V[k][best]; V[k][best] = v; }
                }
            }
        }
        __syncthreads();
        // ---- X = Y V ; residuals of the wanted pairs from BY V - theta X
        double rmax_local = 0.0;
        for (int e = tid; e < n * p; e += MODE_THREADS) {
            int i = e / p, c = e % p;
            double x = 0, bx = 0;
            for (int k = 0; k < p; ++k) {
                x += Y[(size_t)i * p + k] * V[k][c];
                bx += BY[(size_t)i * p + k] * V[k][c];
            }
            X[e] = x;
            if (c < order) {
                double r = bx - theta[c] * x;
                rmax_local += r * r;
            }
        }
        double r2 = block_sum(rmax_local, red);
        double scale = fabs(sigma) + 4.0 * idl2;
        if (tid == 0 && sqrt(r2) <= 2e-14 * scale && iter >= 2) s_done = 1;
        __syncthreads();
        if (s_done) break;
    }
    // ---- outputs: eigenvalues and unit-norm eigenvectors of A (u = E^1/2 w)
    for (int c = 0; c < order; ++c) {
        double s = 0;
        for (int i = tid; i < n; i += MODE_THREADS) {
            double u = sq[i] * X[(size_t)i * p + c];
            s += u * u;
        }
        double nrm = sqrt(block_sum(s, red));
        for (int i = tid; i < n; i += MODE_THREADS) vecs[(size_t)c * n + i] = sq[i] * X[(size_t)i * p + c] / nrm;
        if (tid == 0) vals[c] = theta[c];
        __syncthreads();
    }
    if (tid == 0 && !s_done) *info = *info | 2;
}

int mode_solve(const double* eps_line, int n, double omega, double dl, int pol, double L0, double neff, int order,
               int averaged, double* vals, double* vecs) {
    if (n < 2) FDFD_FAIL("mode plane must span at least 2 cells, got %d", n);
    if (order < 1 || order > n || order > MODE_PMAX - 2) FDFD_FAIL("mode order %d out of range", order);
    int p = order + 6;
    if (p > MODE_PMAX) p = MODE_PMAX;
    if (p > n) p = n;
    double* buf = nullptr;
    size_t nn = (size_t)n * n, np = (size_t)n * p;
    size_t total = nn + 3 * (size_t)n + 3 * np + n /*eps*/ + order + (size_t)order * n;
    FDFD_CHECK(cudaMalloc(&buf, sizeof(double) * total + sizeof(int) * (n + 1)));
    double *M = buf, *d = M + nn, *off = d + n, *sq = off + n, *X = sq + n, *Y = X + np, *BY = Y + np;
    double *eps = BY + np, *dvals = eps + n, *dvecs = dvals + order;
    int* piv = reinterpret_cast<int*>(dvecs + (size_t)order * n);
    int* info = piv + n;
    cudaError_t e = cudaMemcpy(eps, eps_line, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(info, 0, sizeof(int));
    if (e == cudaSuccess) {
        { mode_kernel<<<1, MODE_THREADS>>>(eps, n, omega, dl, pol, L0, neff, order, averaged, p, M, piv, d, off, sq, X, Y,
                                         BY, dvals, dvecs, info); ++g_fdfd_launches; }
        e = cudaGetLastError();
    }
    int h_info = 0;
    if (e == cudaSuccess) e = cudaMemcpy(vals, dvals, sizeof(double) * order, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(vecs, dvecs, sizeof(double) * order * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (e != cudaSuccess) FDFD_FAIL("mode_solve: %s", cudaGetErrorString(e));
    if (h_info & 1) FDFD_FAIL("mode_solve: shifted operator is singular (sigma is an exact eigenvalue)");
    if (h_info & 2) FDFD_FAIL("mode_solve: subspace iteration did not converge");
    return 0;
}
