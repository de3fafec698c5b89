// coop.cuh — cooperative mapping of the transient analysis: ONE circuit instance is advanced by P threads that sit in P
// different warps of a block (hand-written, generic; the per-part code comes from codegen.cpp: emit_coop).  Two drivers:
// tsb_coop_tran_part (circuits without nonlinear devices: one solve per step attempt, described below) and
// tsb_coop_tran_nl_part (diodes / MOSFETs: the Newton loop around the same two phases, at the end of this file).
//
// Why: the thread-per-circuit mapping (skeleton.cuh) keeps a whole instance in one thread's registers.  Past n ~ 12-14
// unknowns the state no longer fits (255 registers), the kernels spill, and the per-thread statistics (32 bytes per result
// column) limit an SM to a handful of warps.  Here the netlist is cut into P sub-circuits that touch each other only
// through a small separator (plan.cpp: build_coop).  Warp w of a group of P warps runs the code of part w % P for the
// same 32 instances: each warp executes only its own straight-line code (no divergence inside a warp — lanes are
// instances, as before), holds 1/P of the parameters, device state, factors and statistics, and per step attempt the
// parts exchange a few doubles through shared memory around ONE named barrier:
//
//     phase_a   own stamps, elimination of the own interior, own forward substitution, own truncation-error decisions
//               -> contribution to the separator system, flags (LTE reject / small, zero pivot) [, result-store key: part 0]
//     bar.sync  (barrier id = 1 + group, 32 * P threads; the exchange buffer is double-buffered, so one barrier per attempt
//               suffices: a buffer is rewritten two attempts later, after the partner has passed the barrier in between)
//     phase_b   every part sums the contributions in part order (identical additions -> identical bits everywhere), solves
//               the separator system for itself, back-substitutes its interior, and takes the step decision — the same in
//               every part because it is computed from the same bits — then updates the state of its own devices and the
//               statistics / waveform rows of its own result columns.
//
// The elimination order is a nested-dissection order, not the reference's: like the condensed elimination of the fast
// build it is a re-association (results differ by rounding x conditioning), held to the parity contract by the tests.
// Step control, result-store de-duplication and counters follow tsb_tran_linear statement by statement (tran.go:96-151,
// anlysis.go:61-85), with the truncation-error test evaluated before the solve is known to be needed: a rejected attempt's
// phase_a is wasted work, the price of one barrier per attempt instead of two.
// The operating point is not run here: tsb_optran (thread-per-circuit, launched first with TsbArgs.coop_state set) runs it
// and hands over device state and solution.
#ifndef TSB_COOP_CUH
#define TSB_COOP_CUH

// Result sink of one part: the part's own columns of the waveform / statistics arrays (layout as TsbSink: per column two
// double2 slots {min, max}, {sum, last}, TSB_COOP_BLOCK entries apart).
template <class Part>
struct TsbCoopSink {
    const TsbArgs& a;
    long long inst;
    long long n_rows;
    bool overflow;
    double2* sm;
    __device__ __forceinline__ TsbCoopSink(const TsbArgs& a_, long long inst_) : a(a_), inst(inst_), n_rows(0), overflow(false) {
        sm = reinterpret_cast<double2*>(tsb_smem) + threadIdx.x;
    }
    __device__ __forceinline__ void begin() {
        if (a.out_flags & TSB_OUT_STATS) {
#pragma unroll
            for (int j = 0; j < Part::NOWN; ++j) {
                sm[(2 * j) * TSB_COOP_BLOCK] = make_double2(__longlong_as_double(0x7ff0000000000000LL), __longlong_as_double(0xfff0000000000000LL));
                sm[(2 * j + 1) * TSB_COOP_BLOCK] = make_double2(0.0, 0.0);
            }
        }
    }
    __device__ __forceinline__ void push(const Part& c, const double* row) {
        if (a.out_flags & TSB_OUT_WAVE) {
            if (n_rows < a.cap_rows) {
                double* w = a.wave + (n_rows * TSB_COOP_NCOL) * a.n_inst + inst;
#pragma unroll
                for (int j = 0; j < Part::NOWN; ++j) __stcs(w + (long long)c.col(j) * a.n_inst, row[j]);
            } else overflow = true;
        }
        if (a.out_flags & TSB_OUT_STATS) {
#pragma unroll
            for (int j = 0; j < Part::NOWN; ++j) {
                const double v = row[j];
                if (Part::PART == 0 && j == 0) {                 // TIME: monotone, extremes = first and last row
                    if (n_rows == 0) sm[0] = make_double2(v, v);
                } else {
                    const double2 mm = sm[(2 * j) * TSB_COOP_BLOCK];
                    double* slot = reinterpret_cast<double*>(&sm[(2 * j) * TSB_COOP_BLOCK]);
                    if (v < mm.x) slot[0] = v;
                    if (v > mm.y) slot[1] = v;
                }
                double2 sl = sm[(2 * j + 1) * TSB_COOP_BLOCK];
                sl.x += v;
                sl.y = v;
                sm[(2 * j + 1) * TSB_COOP_BLOCK] = sl;
            }
        }
        ++n_rows;
    }
    __device__ __forceinline__ void finish(const Part& c) {
        if (a.out_flags & TSB_OUT_STATS) {
#pragma unroll
            for (int j = 0; j < Part::NOWN; ++j) {
                double2 mm = sm[(2 * j) * TSB_COOP_BLOCK];
                const double2 sl = sm[(2 * j + 1) * TSB_COOP_BLOCK];
                if (Part::PART == 0 && j == 0 && n_rows > 0) {
                    const double f = mm.x, l = sl.y;
                    mm.x = f < l ? f : l; mm.y = f < l ? l : f;
                }
                const long long cj = c.col(j);
                a.stats[(0 * TSB_COOP_NCOL + cj) * a.n_inst + inst] = mm.x;
                a.stats[(1 * TSB_COOP_NCOL + cj) * a.n_inst + inst] = mm.y;
                a.stats[(2 * TSB_COOP_NCOL + cj) * a.n_inst + inst] = sl.x;
                a.stats[(3 * TSB_COOP_NCOL + cj) * a.n_inst + inst] = sl.y;
            }
        }
    }
};

__device__ __forceinline__ void tsb_coop_barrier(int id) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(32 * TSB_COOP_PARTS) : "memory");
}

// One part's view of the transient of one instance.  Entered by whole warps (lanes without a live instance only take part
// in the votes and barriers); the P warps of a group iterate in lock step because they take identical decisions.
// xg: the group's exchange area, [2 buffers][P parts][NX slots][32 lanes] doubles.
template <class Part>
__device__ __forceinline__ void tsb_coop_tran_part(const TsbArgs& a, long long inst, bool valid, double* xg, int bar_id) {
    constexpr int NP = TSB_COOP_PARTS, NX = TSB_COOP_NX;
    const int lane = threadIdx.x & 31;
    Part c;
    TsbCoopSink<Part> sink(a, inst);
    bool live = false;
    if (valid) {
        sink.begin();
        live = a.status[inst] == TSB_ST_OK;          // written by tsb_optran: an instance whose operating point failed has no transient
        if (live) c.load(a, inst);
    }
    double time = 0.0, dt = a.minstep;               // tran.go:93
    double last_key = -1.0;
    TsbTimeKeyer keyer; keyer.reset();
    int n_acc = 0, n_rej = 0, n_bad = 0;
    int status = TSB_ST_OK;
    double fail_at = 0.0;
    live = live && time < a.tstop;
    const bool ran = live;
    int buf = 0;
    while (__any_sync(0xffffffffu, live)) {
        double* xb = xg + (long long)buf * (NP * NX * 32) + lane;
        double next_time = time + dt;
        if (live) {
            if (next_time > a.tstop) { next_time = a.tstop; dt = next_time - time; }
            const double rdt = tsb_rcp_dt(dt);
            bool lte_gt, lte_small;
            c.lte_flags(dt, rdt, a.trtol, a.trtol / 100, lte_gt, lte_small);
            if (!c.eval_sources_nb(time)) c.eval_sources(time, 1.0);       // sources at the START of the step
            double* mine = xb + Part::PART * NX * 32;
            const bool ok = c.phase_a(time, dt, rdt, mine);
            mine[0] = __longlong_as_double((long long)((lte_gt ? 1 : 0) | (lte_small ? 2 : 0) | (ok ? 4 : 0)));
            if (Part::PART == 0) mine[32] = next_time >= a.tstart ? keyer.key_any(next_time) : -2.0;
        }
        tsb_coop_barrier(bar_id);
        if (live) {
            int f_or = 0, f_and = 7;
#pragma unroll
            for (int q = 0; q < NP; ++q) { const int f = (int)__double_as_longlong(xb[q * NX * 32]); f_or |= f; f_and &= f; }
            const bool lte_gt = (f_or & 1) != 0, lte_small = (f_and & 2) != 0, pivots_ok = (f_and & 4) != 0;
            if (lte_gt && dt > a.minstep) { dt /= 2; ++n_rej; }
            else {
                const bool solved = c.phase_b(xb) & pivots_ok;
                if (!solved) {
                    ++n_bad;
                    if (dt > a.minstep) { dt /= 2; ++n_rej; }
                    else { status = TSB_ST_TRAN_FAILED; fail_at = time; live = false; }
                } else {
                    // an accepted step (tsb_accept_step): LoadState, Update, advance, StoreTimeResult, step growth
                    c.load_state(dt);
                    c.update_state();
                    time = next_time;
                    if (time >= a.tstart) {
                        const double key = xb[32];
                        if (key != last_key) {
                            double row[Part::NOWN > 0 ? Part::NOWN : 1];
                            c.signals(time, row);
                            sink.push(c, row);
                            last_key = key;
                        }
                    }
                    if (time < a.tstop && dt < a.maxstep) {
                        const double grown = lte_small ? dt * 2 : dt * 1.1;
                        dt = grown < a.maxstep ? grown : a.maxstep;
                    }
                    ++n_acc;
                    live = time < a.tstop;
                }
            }
        }
        buf ^= 1;
    }
    tsb_coop_barrier(bar_id);          // the next instance's first attempt reuses buffer 0: everybody is done reading
    if (!valid) return;
    sink.finish(c);
    if (Part::PART == 0) {
        if (sink.overflow && status == TSB_ST_OK) status = TSB_ST_OVERFLOW;
        a.rows[inst] = sink.n_rows;
        if (status != TSB_ST_OK) a.status[inst] = status;
        if (ran) {
            const int failed = status == TSB_ST_TRAN_FAILED ? 1 : 0;
            a.counters[0 * a.n_inst + inst] = n_acc;
            a.counters[1 * a.n_inst + inst] = n_rej;
            a.counters[2 * a.n_inst + inst] = 2LL * (n_acc + n_rej) - n_bad + 2 * failed;      // as the reference counts its solves (tsb_tran_linear)
            a.counters[5 * a.n_inst + inst] = __double_as_longlong(fail_at);
            a.counters[6 * a.n_inst + inst] += n_acc + n_bad;                                   // passes executed (the operating point's are already in)
            a.counters[7 * a.n_inst + inst] = sink.n_rows;
        }
    }
}

// The same for circuits WITH nonlinear devices (diodes, MOSFETs; BJT circuits do not partition): the Newton loop of
// tran.go:157-216 around phase_a / phase_b.  Per iteration two barriers: contributions -> [barrier] -> separator solve,
// back substitution and the part's own convergence test -> [barrier] -> every part combines the convergence flags and takes
// the same decision.  Trip counts are uniform over the warp (votes, as in tsb_tran_nonlinear) and therefore over the group:
// every part of an instance holds the same Newton state.  The truncation-error predicates read only the state of the last
// accepted steps (not the solution being computed), so they travel with the first iteration's exchange, and so does the
// result-store key of the step's end.
template <class Part>
__device__ __forceinline__ void tsb_coop_tran_nl_part(const TsbArgs& a, long long inst, bool valid, double* xg, int bar_id) {
    constexpr int NP = TSB_COOP_PARTS, NX = TSB_COOP_NX;
    const int lane = threadIdx.x & 31;
    Part c;
    TsbCoopSink<Part> sink(a, inst);
    bool live = false;
    if (valid) {
        sink.begin();
        live = a.status[inst] == TSB_ST_OK;
        if (live) c.load(a, inst);
    }
    double time = 0.0, dt = a.minstep;
    double last_key = -1.0;
    TsbTimeKeyer keyer; keyer.reset();
    int n_acc = 0, n_rej = 0, n_sol = 0;
    int status = TSB_ST_OK;
    double fail_at = 0.0;
    live = live && time < a.tstop;
    const bool ran = live;
    int buf = 0;
    while (__any_sync(0xffffffffu, live)) {
        double next_time = time + dt;
        double rdt = 0.0, key = -2.0;
        bool lte_gt = false, lte_small = false;
        if (live) {
            if (next_time > a.tstop) { next_time = a.tstop; dt = next_time - time; }
            rdt = tsb_rcp_dt(dt);
            if (!c.eval_sources_nb(time)) c.eval_sources(time, 1.0);
        }
        int iter = 0;
        int nr = live ? 0 : 2;                        // 0 iterating, 1 converged, 2 failed / not live
        while (__any_sync(0xffffffffu, nr == 0)) {
            double* xb = xg + (long long)buf * (NP * NX * 32) + lane;
            double* mine = xb + Part::PART * NX * 32;
            if (nr == 0) {
                if (iter > 0) c.update_nl(c.xo);
                const bool ok = c.phase_a(time, dt, rdt, mine);
                int f = ok ? 4 : 0;
                if (iter == 0) {
                    bool g, sm;
                    c.lte_flags(dt, rdt, a.trtol, a.trtol / 100, g, sm);
                    f |= (g ? 1 : 0) | (sm ? 2 : 0);
                    if (Part::PART == 0) mine[32] = next_time >= a.tstart ? keyer.key_any(next_time) : -2.0;
                }
                mine[0] = __longlong_as_double((long long)f);
            }
            tsb_coop_barrier(bar_id);
            bool solved = false;
            if (nr == 0) {
                int f_or = 0, f_and = 7;
#pragma unroll
                for (int q = 0; q < NP; ++q) { const int f = (int)__double_as_longlong(xb[q * NX * 32]); f_or |= f; f_and &= f; }
                if (iter == 0) { lte_gt = (f_or & 1) != 0; lte_small = (f_and & 2) != 0; key = xb[32]; }
                solved = c.phase_b(xb) & ((f_and & 4) != 0);          // the same in every part: separator pivots and the AND of the own ones
                ++n_sol;
                const bool cv = solved && iter > 0 && c.converged_own(a.reltol, a.abstol);
                mine[64] = __longlong_as_double((long long)(cv ? 1 : 0));
            }
            tsb_coop_barrier(bar_id);
            if (nr == 0) {
                if (!solved) nr = 2;
                else {
                    int cv_and = 1;
#pragma unroll
                    for (int q = 0; q < NP; ++q) cv_and &= (int)__double_as_longlong(xb[(q * NX + 2) * 32]);
                    if (iter > 0 && cv_and) nr = 1;
                    else {
                        c.keep_old();
                        if (++iter >= a.max_iter) nr = 2;
                    }
                }
            }
            buf ^= 1;
        }
        if (live) {
            if (nr == 2) {                           // tran.go:113-120
                if (dt > a.minstep) { dt /= 2; ++n_rej; }
                else { status = TSB_ST_TRAN_FAILED; fail_at = time; live = false; }
            } else if (lte_gt && dt > a.minstep) { dt /= 2; ++n_rej; }
            else {
                c.load_state(dt);
                c.update_state();
                time = next_time;
                if (time >= a.tstart && key != last_key) {
                    double row[Part::NOWN > 0 ? Part::NOWN : 1];
                    c.signals(time, row);
                    sink.push(c, row);
                    last_key = key;
                }
                if (time < a.tstop && dt < a.maxstep) {
                    const double grown = lte_small ? dt * 2 : dt * 1.1;
                    dt = grown < a.maxstep ? grown : a.maxstep;
                }
                ++n_acc;
                live = time < a.tstop;
            }
        }
    }
    tsb_coop_barrier(bar_id);
    if (!valid) return;
    sink.finish(c);
    if (Part::PART == 0) {
        if (sink.overflow && status == TSB_ST_OK) status = TSB_ST_OVERFLOW;
        a.rows[inst] = sink.n_rows;
        if (status != TSB_ST_OK) a.status[inst] = status;
        if (ran) {
            a.counters[0 * a.n_inst + inst] = n_acc;
            a.counters[1 * a.n_inst + inst] = n_rej;
            a.counters[2 * a.n_inst + inst] = n_sol;
            a.counters[5 * a.n_inst + inst] = __double_as_longlong(fail_at);
            a.counters[6 * a.n_inst + inst] += n_sol;
            a.counters[7 * a.n_inst + inst] = sink.n_rows;
        }
    }
}

#endif  // TSB_COOP_CUH
