// lu_warp.cu — operator-level batched MNA factor + solve, one circuit per warp (sm_100a).
//
// The analysis kernels (codegen.cpp) keep one circuit per THREAD: for the bundled n <= 10 the whole matrix fits in
// a thread's registers and the LU needs no communication.  That stops scaling at n ~ 12 (n^2 doubles per thread).
// This kernel is the other mapping the north star names — "one circuit per warp, the matrix resident in registers
// and warp shuffles for the pivot-row broadcast" — and the drop-in for the reference's matrix OPERATOR
// (pkg/matrix/circuit.go:126-150 Solve = sparse Factor + Solve; matrix/device.go:3-8) for hosts that stamp
// themselves: many systems A x = b of one order n <= 32 sharing the pivot order of the reference's symbolic pass.
//
// Mapping.  W = 8, 16 or 32 lanes own one system (4, 2 or 1 systems per warp).  The system is staged through
// shared memory with coalesced loads and read back PERMUTED: lane k holds internal row k = external row prow[k],
// register a[j] holds internal column j = external column pcol[j].  After that the frozen-order LU is plain
// Doolittle on internal indices with every register index a compile-time literal:
//   step k:  rp = 1/a_kk (lane k), row k scaled by rp (the U row, as Sparse 1.3 stores it),
//            for j > k: u = shfl(a[j], k); lanes i > k: a[j] -= u * a[k]        (a[k] of lane i = L_ik)
//   forward: c_k *= rp_k (skipped when c_k == 0, as spSolve does); t = shfl(c, k); lanes i > k: c -= t * a[k]
//   back:    fast build  : for j = n-1..1: t = shfl(c, j); lanes i < j: c -= a[j] * t      (n steps, all rows at once)
//            strict build: row by row, columns ascending — the summation order of spSolve (bit parity)
// Same operation order per element as Sparse 1.3 (right-looking here, left-looking in its re-factorisation: each
// element receives the same updates in the same ascending-k order), so the strict build (no FMA contraction, IEEE
// division) reproduces the CPU bits.
// The fast build is the textbook variant of the same elimination, shaped for the instruction count (3 per (k, j)
// pair: two 32-bit shuffles + one DFMA): the multiplier m_i = a_ik / a_kk is formed once per lane and step instead
// of scaling the pivot row in one lane (a_ij -= a_kj * m_i is the same product re-associated), lanes that do not
// take part use m = 0 instead of a select per element, the U rows stay unscaled and the reciprocal pivots are
// applied during back-substitution; reciprocals from the hardware seed + one cubic step.
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsb {

namespace {

__device__ __forceinline__ double shfl_d(double v, int src, int width) {
    return __shfl_sync(0xffffffffu, v, src, width);
}
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}
template <bool STRICT> __device__ __forceinline__ double nmuladd(double a, double u, double l) {   // a - u*l
    return STRICT ? __dsub_rn(a, __dmul_rn(u, l)) : fma(-u, l, a);
}

constexpr int LU_THREADS = 128;
// 1: pivot-row broadcast through shared memory (one lane stores the row with STS.128, all lanes read it back with
// broadcast LDS.128: 2 instead of 3 instructions per (k, j) pair).  Measured SLOWER than the shuffles on B200
// (n = 32: 1.18e8 vs 1.30e8 systems/s; profiles/r01_lu_operator.txt), so the shuffles stay.
#ifndef TSB_LU_SMEM_BCAST
#define TSB_LU_SMEM_BCAST 0
#endif

// A: [n_inst][n][n] row-major, b / x: [n_inst][n]  (instance-major: what a per-instance stamper produces)
template <int W, bool STRICT>
__global__ void __launch_bounds__(LU_THREADS, 4) tsb_k_lu_warp(const double* __restrict__ A, const double* __restrict__ b,
                                                            double* __restrict__ x, int* __restrict__ status, long long n_inst,
                                                            int n, const int* __restrict__ prow, const int* __restrict__ pcol, int fence) {
    constexpr int GROUPS = LU_THREADS / W;          // systems per block and pass
    constexpr int LD = W + 1;                       // padded row stride of the staging tile (bank-conflict-free columns)
    extern __shared__ double smem[];
    double* tile_all = smem;                        // GROUPS * W * LD doubles
    int* perm = reinterpret_cast<int*>(smem + GROUPS * W * LD);   // 4 * W ints, see below
    const int lane = threadIdx.x % W;
    const int group = threadIdx.x / W;
    // perm[0..W): prow, perm[W..2W): pcol (identity beyond n); perm[2W..3W): inverse of prow, perm[3W..4W): inverse of pcol
    if (threadIdx.x < 2 * W) {
        const int k = threadIdx.x % W;
        const int* src = threadIdx.x < W ? prow : pcol;
        const int v = k < n ? src[k] : k;
        perm[threadIdx.x] = v;
        perm[(threadIdx.x < W ? 2 * W : 3 * W) + v] = k;
    }
    double* tile = tile_all + group * W * LD;
    // The tile holds the system in INTERNAL order (row k = pivot row k, column j = pivot column j) inside an identity
    // matrix of order W: the padding is written once, the n x n part is overwritten for every system, and a lane
    // reads its row with compile-time offsets — no index arithmetic or selects per element.
    for (int e = lane; e < W * W; e += W) tile[(e / W) * LD + lane] = (e / W == lane) ? 1.0 : 0.0;
    __syncthreads();
    const int my_row = perm[lane];                  // external row owned by this lane (output / right-hand side)
    const int my_col_int = perm[3 * W + lane];      // internal column of external column `lane` (staging)

    for (long long base = (long long)blockIdx.x * GROUPS; base < n_inst; base += (long long)gridDim.x * GROUPS) {
        const long long inst = base + group;
        const bool live = inst < n_inst;
        // ---- stage the system row by row: lanes = external columns (coalesced), scattered into internal order ----
        if (live && lane < n) {
            const double* Ai = A + inst * (long long)n * n;
            for (int r = 0; r < n; ++r) tile[perm[2 * W + r] * LD + my_col_int] = __ldcs(Ai + r * n + lane);   // (unrolling by 8 measured slower)
        }
        __syncwarp();
        double a[W];
        double c = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) a[j] = tile[lane * LD + j];
        if (live && lane < n) c = __ldcs(b + inst * (long long)n + my_row);
        __syncwarp();
        bool ok = true;
        if (STRICT) {
            // ---- factor (Sparse 1.3 operation order) ----------------------------------------------------
#pragma unroll
            for (int k = 0; k < W; ++k) {
                if (k < n) {                                          // uniform: padded steps are skipped
                    const double piv = shfl_d(a[k], k, W);
                    ok = ok && (piv != 0.0);
                    const double rp = 1.0 / piv;
                    if (lane == k) {
                        a[k] = rp;
#pragma unroll
                        for (int j = k + 1; j < W; ++j) a[j] = __dmul_rn(a[j], rp);
                    }
#pragma unroll
                    for (int j = k + 1; j < W; ++j) {
                        const double u = shfl_d(a[j], k, W);
                        if (lane > k) a[j] = nmuladd<true>(a[j], u, a[k]);
                    }
                }
            }
            // ---- forward substitution (spSolve: zero entries skipped) -------------------------------------
#pragma unroll
            for (int k = 0; k < W; ++k) {
                if (k < n) {
                    if (lane == k && c != 0.0) c = __dmul_rn(c, a[k]);
                    const double t = shfl_d(c, k, W);
                    if (t != 0.0 && lane > k) c = nmuladd<true>(c, t, a[k]);
                }
            }
            // ---- back substitution: row by row, columns ascending (spSolve's summation order) --------------
#pragma unroll
            for (int i = W - 2; i >= 0; --i) {
                if (i < n - 1) {
#pragma unroll
                    for (int j = i + 1; j < W; ++j) {
                        if (j < n) {
                            const double t = shfl_d(c, j, W);
                            if (lane == i) c = __dsub_rn(c, __dmul_rn(a[j], t));
                        }
                    }
                }
            }
        } else {
            // ---- factor: multipliers in the L part, U unscaled, reciprocal pivots on the diagonal -----------
            // (no `k < n` tests: the identity padding makes the steps beyond n exact no-ops, and they are the short ones)
#pragma unroll
            for (int k = 0; k < W; ++k) {
                // `fence` is the lane width passed at run time, i.e. always true: the uniform branch keeps ptxas from
                // hoisting the shuffles of later steps across this one (255 registers and spills without it)
                if (k < fence) {
#if TSB_LU_SMEM_BCAST
                    // pivot-row broadcast through shared memory (the staging tile is free by now): one lane stores the
                    // row, every lane reads it back with broadcast loads — 1 load per element instead of 2 shuffles
                    if (lane == k) {
#pragma unroll
                        for (int j = k; j < W; ++j) tile[j] = a[j];
                    }
                    __syncwarp();
                    const double piv = tile[k];
#else
                    const double piv = shfl_d(a[k], k, W);
#endif
                    ok = ok && (piv != 0.0);
                    const double rp = rcp_fast(piv);
                    const double m = lane > k ? a[k] * rp : 0.0;      // 0 for rows that are finished: their update is a no-op
                    a[k] = lane == k ? rp : (lane > k ? m : a[k]);    // rows above k keep U_ik for the back-substitution
#if TSB_LU_SMEM_BCAST
#pragma unroll
                    for (int j = k + 1; j < W; ++j) a[j] = fma(-tile[j], m, a[j]);
                    __syncwarp();
#else
#pragma unroll
                    for (int j = k + 1; j < W; ++j) a[j] = fma(-shfl_d(a[j], k, W), m, a[j]);
#endif
                }
            }
            // ---- forward: y_i = b_i - sum_k m_ik y_k  (a[k] of the rows <= k is rp or 0-multiplier: masked) -----
#pragma unroll
            for (int k = 0; k < W - 1; ++k) {
                const double t = shfl_d(c, k, W);
                c = fma(-t, lane > k ? a[k] : 0.0, c);
            }
            // ---- back: x_j = y_j / a_jj, then eliminated from the rows above ---------------------------------
#pragma unroll
            for (int j = W - 1; j >= 0; --j) {
                if (lane == j) c *= a[j];
                const double t = shfl_d(c, j, W);
                c = fma(-t, lane < j ? a[j] : 0.0, c);
            }
        }
        if (live && lane < n) __stcs(x + inst * (long long)n + perm[W + lane], c);
        if (live && lane == 0) status[inst] = ok ? 0 : 1;
    }
}

template <int W, bool STRICT>
cudaError_t launch_lu(const double* A, const double* b, double* x, int* status, long long n_inst, int n, const int* prow,
                      const int* pcol, int sms, cudaStream_t s) {
    constexpr int GROUPS = LU_THREADS / W;
    const size_t smem = (size_t)GROUPS * W * (W + 1) * sizeof(double) + 4 * W * sizeof(int);
    auto kern = tsb_k_lu_warp<W, STRICT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long want = (n_inst + GROUPS - 1) / GROUPS;
    long long resident = (long long)sms * 16;                // grid-stride over systems: SM count x max resident blocks
    int blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    kern<<<blocks, LU_THREADS, smem, s>>>(A, b, x, status, n_inst, n, prow, pcol, W);
    return cudaGetLastError();
}

}  // namespace

// prow / pcol: device arrays of n 0-based external indices (internal step k -> external row / column).
cudaError_t launch_lu_warp(const double* A, const double* b, double* x, int* status, long long n_inst, int n, const int* prow,
                           const int* pcol, int strict, int sms, cudaStream_t s) {
    if (n <= 8) return strict ? launch_lu<8, true>(A, b, x, status, n_inst, n, prow, pcol, sms, s)
                              : launch_lu<8, false>(A, b, x, status, n_inst, n, prow, pcol, sms, s);
    if (n <= 16) return strict ? launch_lu<16, true>(A, b, x, status, n_inst, n, prow, pcol, sms, s)
                               : launch_lu<16, false>(A, b, x, status, n_inst, n, prow, pcol, sms, s);
    return strict ? launch_lu<32, true>(A, b, x, status, n_inst, n, prow, pcol, sms, s)
                  : launch_lu<32, false>(A, b, x, status, n_inst, n, prow, pcol, sms, s);
}

}  // namespace tsb
