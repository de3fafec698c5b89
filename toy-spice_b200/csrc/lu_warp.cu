// lu_warp.cu — operator-level batched MNA factor + solve, one circuit per warp (sm_100a).
//
// The analysis kernels (codegen.cpp) keep one circuit per THREAD: for the bundled n <= 10 the whole matrix fits in
// a thread's registers and the LU needs no communication.  That stops scaling at n ~ 12 (n^2 doubles per thread).
// This kernel is the other mapping the north star names — "one circuit per warp [here: per 4 - 16 lanes of one], the
// matrix resident in registers and warp shuffles for the pivot-row broadcast" — and the drop-in for the reference's matrix OPERATOR
// (pkg/matrix/circuit.go:126-150 Solve = sparse Factor + Solve; matrix/device.go:3-8) for hosts that stamp
// themselves: many systems A x = b of one order n <= 32 sharing the pivot order of the reference's symbolic pass.
//
// Mapping.  W = 4, 8 or 16 lanes own one system (8, 4 or 2 systems per warp) and every lane holds R = 2 or 3 of its rows
// (rows i, i + W, i + 2W of the INTERNAL order: cyclic, so the slots above the pivot's drop out of the unrolled code as the
// elimination proceeds).  What bounds this kernel is the delivery of pivot-row elements to the lanes — a 64-bit value to 32
// lanes costs two shuffles = 2 clocks of the SM's 128 B/clock shuffle / shared-memory path whichever way it travels — so
// every delivered element should feed as many DFMA as the register file allows: R per lane.  The system is staged through
// shared memory — the NEXT one with cp.async while the current one is eliminated — and read back PERMUTED: slot s of lane
// l holds internal row l + sW = external row prow[l + sW], register a[s][j] holds internal column j = external column
// pcol[j].  After that the frozen-order LU is plain Doolittle on internal indices with every register index a
// compile-time literal (static_for over the steps):
//   step k (slot sk = k / W, lane lk = k % W):
//            rp = 1/a_kk (lane lk), row k scaled by rp (the U row, as Sparse 1.3 stores it),
//            for j > k: u = shfl(a[sk][j], lk); rows i > k: a[s][j] -= u * a[s][k]      (a[s][k] of row i = L_ik)
//   forward: c_k *= rp_k (skipped when c_k == 0, as spSolve does); t = shfl(c[sk], lk); rows i > k: c -= t * a[s][k]
//   back:    fast build  : for j = n-1..1: t = shfl(c[sj], lj); rows i < j: c -= a[s][j] * t  (n steps, all rows at once)
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
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

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

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Compile-time loop over the elimination steps: the step index is a constant expression inside the body, so every register
// index of the matrix is a literal from the start (a `#pragma unroll`ed loop leaves that to the optimiser's pass order, which
// put the whole matrix into local memory for 3 and 4 rows per lane).
template <int B, int E, class F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

constexpr int LU_THREADS = 128;
// 1: also build the shared-memory pivot-row broadcast variants (BCAST below).  Measured SLOWER than the shuffles at every order
// and rows-per-lane on B200 (n = 32, 2 rows: 4.7 vs 5.6 TFLOP/s; profiles/r02_lu_operator.txt): a 128-bit shared-memory load
// delivers its 512 bytes to the warp at the same 128 bytes per clock as four shuffles do, and the one-lane stores come on top.
#ifndef TSB_LU_WITH_BCAST
#define TSB_LU_WITH_BCAST 0
#endif
// The pivot order travels as a kernel parameter (256 bytes of the constant bank): no device allocation, no copy and no
// stream-ordering question per call (a cudaMallocAsync + copy + cudaFreeAsync per call cost 0.2 - 3 ms between
// synchronisations: the pool gives its memory back at every one).
struct LuPerm { int prow[32], pcol[32]; };

// Variant chosen per order n at launch (launch_lu_warp): W lanes x R rows per lane, BCAST = where the pivot row of the
// fast build's elimination comes from.  $TSB_LU_VARIANT="R,B" overrides the choice for one process (A/B measurements:
// tests/gpu_lu_perf.py); results of the strict build do not depend on it (same operations per element in the same
// order), those of the fast build agree to rounding.
//   R = 1: one row per lane, 32 / n systems per warp — per (k, j) pair and system 2 SHFL.32 + 1 DFMA;
//   R = 2 / 4: rows i, i + W, .. of a system in one lane (cyclic: the slots below the pivot's drop out of the unrolled
//          code as the elimination proceeds), twice / four times as many systems per warp — one shuffled pivot-row
//          element feeds R DFMA: per pair and system 2/R SHFL.32 + 1 DFMA;
//   BCAST: the pivot row goes through shared memory (the lane that owns it stores it with STS.128 into a double-buffered
//          row of the — by then free — staging tile, every lane of the group reads it back with broadcast LDS.128): per
//          pair and system 1/R shared-memory instructions, one __syncwarp per step.
//   ASYNC: the NEXT system of a group is copied into the staging tile with cp.async (8 bytes per lane, scattered into
//          internal order as before, no registers) right after the current one has been read out of it, so the copy
//          runs under the elimination; otherwise LDG + STS at the top of every pass.
// A: [n_inst][n][n] row-major, b / x: [n_inst][n]  (instance-major: what a per-instance stamper produces)
template <int W, int R, bool STRICT, bool BCAST, bool ASYNC>
__global__ void __launch_bounds__(LU_THREADS, (R * W * R >= 64 ? 3 : 4))
tsb_k_lu_warp(const double* __restrict__ A, const double* __restrict__ b, double* __restrict__ x, int* __restrict__ status,
              long long n_inst, int n, const __grid_constant__ LuPerm pv, int fence) {
    constexpr int N = W * R;                        // order of the padded system
    constexpr int GROUPS = LU_THREADS / W;          // systems per block and pass
    constexpr int LD = N + 2;                       // padded row stride of the staging tile: even (a lane reads its rows with
                                                    // LDS.128), 17 sixteen-byte units for N = 32 (conflict-free rows and columns)
    constexpr int TILE = N * LD;                    // doubles per system
    static_assert(!BCAST || !STRICT, "shared-memory pivot-row broadcast: fast build only");
    extern __shared__ __align__(16) double smem[];
    double* tile_all = smem;                        // GROUPS * TILE doubles
    double* rowbuf_all = smem + GROUPS * TILE;      // GROUPS * 2 * N doubles (BCAST: double-buffered pivot row)
    int* perm = reinterpret_cast<int*>(rowbuf_all + GROUPS * 2 * N);   // 4 * N ints, see below
    const int lane = threadIdx.x % W;
    const int group = threadIdx.x / W;
    // perm[0..N): prow, perm[N..2N): pcol (identity beyond n); perm[2N..3N): inverse of prow, perm[3N..4N): inverse of pcol
    if (threadIdx.x < 2 * N) {
        const int k = threadIdx.x % N;
        const int v = k < n ? (threadIdx.x < N ? pv.prow[k] : pv.pcol[k]) : k;
        perm[threadIdx.x] = v;
        perm[(threadIdx.x < N ? 2 * N : 3 * N) + v] = k;
    }
    double* tile = tile_all + group * TILE;
    double* rowbufs = rowbuf_all + group * 2 * N;
    // The tile holds the system in INTERNAL order (row k = pivot row k, column j = pivot column j) inside an identity
    // matrix of order N: the padding is written once, the n x n part is overwritten for every system, and a lane
    // reads its rows with compile-time offsets — no index arithmetic or selects per element.
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int cc = 0; cc < R; ++cc) tile[r * LD + lane + cc * W] = (r == lane + cc * W) ? 1.0 : 0.0;
    __syncthreads();
    int my_row[R], my_col_int[R], my_col_out[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        my_row[s] = perm[lane + s * W];                 // external row of internal row lane + s W (right-hand side)
        my_col_int[s] = perm[3 * N + lane + s * W];     // internal column of external column lane + s W (staging)
        my_col_out[s] = perm[N + lane + s * W];         // external column of internal column lane + s W (solution)
    }
    // ---- staging of one system, row by row: lanes = external columns (coalesced), scattered into internal order ----
    auto stage = [&](long long inst_s) {
        if (inst_s < n_inst) {
            const double* Ai = A + inst_s * (long long)n * n;
            for (int r = 0; r < n; ++r) {
                const int ir = perm[2 * N + r] * LD;
#pragma unroll
                for (int cc = 0; cc < R; ++cc)
                    if (lane + cc * W < n) {
                        if (ASYNC) cp_async8(tile + ir + my_col_int[cc], Ai + r * n + lane + cc * W);
                        else tile[ir + my_col_int[cc]] = __ldcs(Ai + r * n + lane + cc * W);
                    }
            }
        }
    };
    const long long stride = (long long)gridDim.x * GROUPS;
    if (ASYNC) stage((long long)blockIdx.x * GROUPS + group);

    for (long long base = (long long)blockIdx.x * GROUPS; base < n_inst; base += stride) {
        const long long inst = base + group;
        const bool live = inst < n_inst;
        if (ASYNC) cp_async_wait_all();
        else stage(inst);
        __syncwarp();
        double a[R][N];
        double c[R];
#pragma unroll
        for (int s = 0; s < R; ++s) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                const double2 v = *reinterpret_cast<const double2*>(tile + (lane + s * W) * LD + j);
                a[s][j] = v.x;
                a[s][j + 1] = v.y;
            }
            c[s] = 0.0;
            if (live && lane + s * W < n) c[s] = __ldcs(b + inst * (long long)n + my_row[s]);
        }
        __syncwarp();
        if (ASYNC) stage(inst + stride);                 // lands in the tile while this system is eliminated
        bool ok = true;
        if (STRICT) {
            // The padded steps are skipped (uniform tests against n).  Each phase tests its own laundered copy of n: with one
            // variable the compiler specialises the later phases for every exit point of the earlier ones (4x the code).
            int n_fa = n, n_fw = n, n_bk = n;
            asm volatile("mov.u32 %0, %0;" : "+r"(n_fa));
            asm volatile("mov.u32 %0, %0;" : "+r"(n_fw));
            asm volatile("mov.u32 %0, %0;" : "+r"(n_bk));
            // ---- factor (Sparse 1.3 operation order) ----------------------------------------------------
            static_for<0, N>([&](auto kc_) {
                constexpr int k = decltype(kc_)::value;
                if (k < n_fa) {                                       // uniform: padded steps are skipped
                    constexpr int sk = k / W, lk = k % W;                 // slot and lane of the pivot row
                    const double piv = shfl_d(a[sk][k], lk, W);
                    ok = ok && (piv != 0.0);
                    const double rp = 1.0 / piv;
                    if (lane == lk) {
                        a[sk][k] = rp;
#pragma unroll
                        for (int j = k + 1; j < N; ++j) a[sk][j] = __dmul_rn(a[sk][j], rp);
                    }
#pragma unroll
                    for (int j = k + 1; j < N; ++j) {
                        const double u = shfl_d(a[sk][j], lk, W);
#pragma unroll
                        for (int s = sk; s < R; ++s)
                            if (s > sk || lane > lk) a[s][j] = nmuladd<true>(a[s][j], u, a[s][k]);
                    }
                }
            });
            // ---- forward substitution (spSolve: zero entries skipped) -------------------------------------
            static_for<0, N>([&](auto kc_) {
                constexpr int k = decltype(kc_)::value;
                if (k < n_fw) {
                    constexpr int sk = k / W, lk = k % W;
                    if (lane == lk && c[sk] != 0.0) c[sk] = __dmul_rn(c[sk], a[sk][k]);
                    const double t = shfl_d(c[sk], lk, W);
                    if (t != 0.0) {
#pragma unroll
                        for (int s = sk; s < R; ++s)
                            if (s > sk || lane > lk) c[s] = nmuladd<true>(c[s], t, a[s][k]);
                    }
                }
            });
            // ---- back substitution: row by row, columns ascending (spSolve's summation order) --------------
            static_for<0, N - 1>([&](auto kc_) {
                constexpr int i = N - 2 - decltype(kc_)::value;
                if (i < n_bk - 1) {
                    constexpr int si = i / W, li = i % W;
#pragma unroll
                    for (int j = i + 1; j < N; ++j) {
                        if (j < n_bk) {
                            const double t = shfl_d(c[j / W], j % W, W);
                            if (lane == li) c[si] = __dsub_rn(c[si], __dmul_rn(a[si][j], t));
                        }
                    }
                }
            });
        } else {
            // ---- factor: multipliers in the L part, U unscaled, reciprocal pivots on the diagonal -----------
            // (no `k < n` tests: the identity padding makes the steps beyond n exact no-ops, and they are the short ones)
            static_for<0, N>([&](auto kc_) {
                constexpr int k = decltype(kc_)::value;
                // `fence` is the lane width passed at run time, i.e. always true: the uniform branch keeps ptxas from
                // hoisting the shuffles of later steps across this one (255 registers and spills without it)
                if (k < fence) {
                    constexpr int sk = k / W, lk = k % W;
                    constexpr int j0 = k & ~1;                            // first element of the stored part of the pivot row
                    double* rowbuf = rowbufs + (k & 1) * N;           // BCAST: double-buffered, 16-byte aligned
                    double piv;
                    if (BCAST) {
                        if (lane == lk) {
#pragma unroll
                            for (int j = j0; j < N; j += 2)
                                *reinterpret_cast<double2*>(rowbuf + j) = make_double2(a[sk][j], a[sk][j + 1]);
                        }
                        __syncwarp();
                        piv = rowbuf[k];
                    } else {
                        piv = shfl_d(a[sk][k], lk, W);
                    }
                    ok = ok && (piv != 0.0);
                    const double rp = rcp_fast(piv);
                    double m[R];
#pragma unroll
                    for (int s = 0; s < R; ++s) {
                        if (s < sk) m[s] = 0.0;                       // rows above the pivot's slot are finished
                        else if (s == sk) {
                            m[s] = lane > lk ? a[s][k] * rp : 0.0;    // 0 for rows that are finished: their update is a no-op
                            a[s][k] = lane == lk ? rp : (lane > lk ? m[s] : a[s][k]);   // rows above k keep U_ik for the back-substitution
                        } else {
                            m[s] = a[s][k] * rp;
                            a[s][k] = m[s];
                        }
                    }
                    if (BCAST) {
                        constexpr int je = (k + 2) & ~1;                  // first even column beyond k: aligned pairs from there
                        if (k + 1 < je && k + 1 < N) {
                            const double u = rowbuf[k + 1];
#pragma unroll
                            for (int s = sk; s < R; ++s) a[s][k + 1] = fma(-u, m[s], a[s][k + 1]);
                        }
#pragma unroll
                        for (int j = je; j < N; j += 2) {             // one LDS.128 for columns j and j + 1
                            const double2 uu = *reinterpret_cast<const double2*>(rowbuf + j);
#pragma unroll
                            for (int s = sk; s < R; ++s) {
                                a[s][j] = fma(-uu.x, m[s], a[s][j]);
                                a[s][j + 1] = fma(-uu.y, m[s], a[s][j + 1]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = k + 1; j < N; ++j) {
                            const double u = shfl_d(a[sk][j], lk, W);
#pragma unroll
                            for (int s = sk; s < R; ++s) a[s][j] = fma(-u, m[s], a[s][j]);
                        }
                    }
                }
            });
            // ---- forward: y_i = b_i - sum_k m_ik y_k  (a[k] of the rows <= k is rp or 0-multiplier: masked) -----
            static_for<0, N - 1>([&](auto kc_) {
                constexpr int k = decltype(kc_)::value;
                constexpr int sk = k / W, lk = k % W;
                const double t = shfl_d(c[sk], lk, W);
#pragma unroll
                for (int s = sk; s < R; ++s) c[s] = fma(-t, (s > sk || lane > lk) ? a[s][k] : 0.0, c[s]);
            });
            // ---- back: x_j = y_j / a_jj, then eliminated from the rows above ---------------------------------
            static_for<0, N>([&](auto kc_) {
                constexpr int j = N - 1 - decltype(kc_)::value;
                constexpr int sj = j / W, lj = j % W;
                if (lane == lj) c[sj] *= a[sj][j];
                const double t = shfl_d(c[sj], lj, W);
#pragma unroll
                for (int s = 0; s <= sj; ++s) c[s] = fma(-t, (s < sj || lane < lj) ? a[s][j] : 0.0, c[s]);
            });
        }
#pragma unroll
        for (int s = 0; s < R; ++s)
            if (live && lane + s * W < n) __stcs(x + inst * (long long)n + my_col_out[s], c[s]);
        if (live && lane == 0) status[inst] = ok ? 0 : 1;
    }
}

template <int W, int R, bool STRICT, bool BCAST, bool ASYNC>
cudaError_t launch_lu(const double* A, const double* b, double* x, int* status, long long n_inst, int n, const LuPerm& pv, int sms,
                      cudaStream_t s) {
    constexpr int N = W * R;
    constexpr int GROUPS = LU_THREADS / W;
    constexpr int TILE = N * (N + 2);
    const size_t smem = (size_t)GROUPS * (TILE + 2 * N) * sizeof(double) + 4 * N * sizeof(int);
    auto kern = tsb_k_lu_warp<W, R, STRICT, BCAST, ASYNC>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    long long want = (n_inst + GROUPS - 1) / GROUPS;
    long long resident = (long long)sms * 16;                // grid-stride over systems: SM count x max resident blocks
    int blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    kern<<<blocks, LU_THREADS, smem, s>>>(A, b, x, status, n_inst, n, pv, N);
    return cudaGetLastError();
}

template <int W, int R>
cudaError_t launch_lu_wr(const double* A, const double* b, double* x, int* status, long long n_inst, int n, const LuPerm& pv,
                         int strict, int bcast, int async, int sms, cudaStream_t s) {
#define TSB_LU_ARGS A, b, x, status, n_inst, n, pv, sms, s
    if (strict) return async ? launch_lu<W, R, true, false, true>(TSB_LU_ARGS) : launch_lu<W, R, true, false, false>(TSB_LU_ARGS);
#if TSB_LU_WITH_BCAST
    if (bcast) return async ? launch_lu<W, R, false, true, true>(TSB_LU_ARGS) : launch_lu<W, R, false, true, false>(TSB_LU_ARGS);
#endif
    return async ? launch_lu<W, R, false, false, true>(TSB_LU_ARGS) : launch_lu<W, R, false, false, false>(TSB_LU_ARGS);
#undef TSB_LU_ARGS
}

}  // namespace

// prow / pcol: HOST arrays of n 0-based external indices (internal step k -> external row / column).
cudaError_t launch_lu_warp(const double* A, const double* b, double* x, int* status, long long n_inst, int n, const int* prow,
                           const int* pcol, int strict, int sms, cudaStream_t s) {
    // default variant per order (measured: profiles/r02_lu_operator.txt); $TSB_LU_VARIANT = "R,B,A" forces rows per lane,
    // the shared-memory pivot-row broadcast and the asynchronous staging for one process
    const int forced = [] {                         // 16 + 4 * R + 2 * bcast + async (read per call: one process can A/B)
        const char* e = getenv("TSB_LU_VARIANT");
        int r = 0, bc = 0, as = 0;
        if (e && sscanf(e, "%d,%d,%d", &r, &bc, &as) >= 1 && r >= 1 && r <= 4) return 16 + 4 * r + (bc ? 2 : 0) + (as ? 1 : 0);
        return 0;
    }();
    LuPerm pv;
    for (int k = 0; k < 32; ++k) { pv.prow[k] = k < n ? prow[k] : k; pv.pcol[k] = k < n ? pcol[k] : k; }
    // defaults, both builds (measured on B200, profiles/r02_lu_operator.txt): asynchronous staging, shuffles, and as many rows per
    // lane as give the smallest padded order that fits the register file — n <= 8: 4 lanes x 2 rows, <= 12: 4 x 3,
    // <= 16: 8 x 2, <= 24: 8 x 3, <= 32: 16 x 2  (fast build n = 32: 3.8 -> 5.7 TFLOP/s, n = 16: 2.7 -> 4.8, n = 10: 0.8 -> 2.7;
    // strict build n = 32: 3.8 -> 3.2 ms, n = 10: 4.2 -> 0.9 ms)
    int R = (n <= 8 || (n > 12 && n <= 16) || n > 24) ? 2 : 3, bc = 0, as = 1;
    if (forced) { R = (forced - 16) / 4; bc = (forced >> 1) & 1; as = forced & 1; }
    if (R == 4 && (n > 16 || (n > 8 && strict))) R = 2;           // 4 rows: where they fit the register file
    if (R == 3 && n > 24) R = 2;                                   // 3 rows: orders up to 12 (4 lanes) and up to 24 (8 lanes)
#define TSB_LU_GO(W_, R_) return launch_lu_wr<W_, R_>(A, b, x, status, n_inst, n, pv, strict, bc, as, sms, s)
    if (R == 3 && !bc) {
        if (n <= 12) TSB_LU_GO(4, 3);
        TSB_LU_GO(8, 3);
    }
    if (R == 3) R = 2;
    if (n <= 8) {
        if (R == 4) TSB_LU_GO(2, 4);
        if (R == 2) TSB_LU_GO(4, 2);
        TSB_LU_GO(8, 1);
    }
    if (n <= 16) {
        if (R == 4) TSB_LU_GO(4, 4);
        if (R == 2) TSB_LU_GO(8, 2);
        TSB_LU_GO(16, 1);
    }
    if (R == 2) TSB_LU_GO(16, 2);
    TSB_LU_GO(32, 1);
#undef TSB_LU_GO
}

}  // namespace tsb
