// kernel_template.cuh — hand-written sm_100a kernels of the batched interior-point MCP solver.
//
// This file is NOT compiled on its own: plan.cpp prepends a generated prologue (problem-size macros,
// assembly tables as __device__ const arrays, and the device functions mcp_eval_newton_par / mcp_eval_sens_par
// lowered from the traced G, H and Jacobian entries) and hands the result to NVRTC for sm_100a.
//
// One SUB-WARP (SUB = 16 lanes when the factorisation window has ≤ 16 rows, else the full warp) owns one
// problem instance for its whole solve (persistent over Newton iterations): the
// iterate (x, y, s), residuals, Jacobian entries and the active window of the banded factorisation live
// in shared memory; HBM is touched for θ/x₀/y₀/s₀ in, (x, y, s, …) out, and an L2-resident scratch that
// streams the banded condensed matrix into the window and the finished U rows out of it.
//
// Reference semantics restated here (file:line in /root/reference):
//   Newton / ϵ-homotopy loop            src/solver.jl:63-121
//   regularised KKT solve               src/solver.jl:79-83   (∇F + tol·I) δz = −F
//   fraction-to-the-boundary linesearch src/solver.jl:93-94,127-138
//   IFT sensitivities                   src/AutoDiff.jl:18-40,59-76,98
//
// Macros provided by the prologue:
//   NX NY NT NRED KL KU WC WR WS1 WSS NRHS_SENS NJV NJTV ND THETA_IN_SMEM SUB
//   SOLVE_INST SENS_INST (instances per CTA)  SOLVE_SMEM_DOUBLES SENS_SMEM_DOUBLES (per instance)
//   SHARED_TABLE_DOUBLES (per CTA)  SOLVE_SCRATCH SENS_SCRATCH CVAL_DOUBLES (doubles per instance)  HAS_JT
// Tables: D_ROWPTR D_CPOS D_TP (dests), T_COEF T_I (terms), R_GROW R_PTR R_CODE R_K R_COEF (rhs),
//   H_PTR H_CODE H_COL H_COEF (H_x rows), PERM, Q_PTR Q_ROW Q_CODE Q_COEF (θ-Jacobian by column)

#define FULLMASK 0xffffffffu
// State vectors are re-written during the kernel.  In shared memory `__restrict__` is a pure aliasing hint; once
// they live in global memory (LARGE_STATE) `const T* __restrict__` would licence the non-coherent read-only path
// (ld.global.nc) and stale reads, so the qualifier is dropped there.
#if LARGE_STATE || S_GLOBAL
#define RS
#else
#define RS __restrict__
#endif
#ifndef SUB
#define SUB 32  // lanes per instance
#endif
#define DBL_MAX_ 1.7976931348623157e308

struct SolveParams {
  long long B;
  const double* theta;
  const double* x0;
  const double* y0;
  const double* s0;
  double* x_out;
  double* y_out;
  double* s_out;
  double* kkt_out;
  double* eps_out;
  int* outer_out;
  int* status_out;
  int* steps_out;
  double* scratch;
  double* state;                 // LARGE_STATE: per-instance vectors (global, L2-resident)
  unsigned long long* counters;  // [0] work queue, [1] Σ newton steps, [2] # solved, [3] # deferred, [4] pass-1 queue
  int* deferred;                 // instance ids handed from pass 0 to pass 1
  double tol;
  double tightening_rate;
  double loosening_rate;
  double min_stepsize;
  int max_inner;
  int max_outer;
  int pass;         // 0: every instance, up to `step_budget` Newton steps; 1: resume the deferred ones
  int step_budget;  // <= 0: no deferral
};

struct SensParams {
  long long B;
  const double* theta;
  const double* x;
  const double* y;
  const double* s;
  double* dzdtheta;        // [n x NT x B] or null
  const double* zbar;      // [n x B] or null
  double* thetabar;        // [NT x B] or null
  const double* theta_p;   // [NT x P x B] or null
  double* z_p;             // [n x P x B] or null
  int* status_out;         // or null
  double* scratch;
  double* state;           // LARGE_STATE: per-instance vectors (global)
  unsigned long long* counters;
  int P;
};

__device__ __forceinline__ double opval(int code, const double* RS jv, const double* RS th) {
  return code >= 0 ? jv[code] : (code == -1 ? 1.0 : th[-2 - code]);
}

// lanes of my sub-warp (the whole warp when SUB == 32)
__device__ __forceinline__ unsigned sub_mask(int lane) {
  return (SUB == 32) ? FULLMASK : (0xffffu << (lane & 16));
}

__device__ __forceinline__ double sub_sum(double v, unsigned smask) {
#pragma unroll
  for (int o = SUB / 2; o > 0; o >>= 1) v += __shfl_xor_sync(smask, v, o, SUB);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) { return sub_sum(v, FULLMASK); }

// NaN-propagating max, like Julia's norm(F, Inf) (src/solver.jl:107)
__device__ __forceinline__ double nanmax(double a, double b) { return (a != a) ? a : ((b != b) ? b : fmax(a, b)); }

__device__ __forceinline__ double sub_nanmax(double v, unsigned smask) {
#pragma unroll
  for (int o = SUB / 2; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(smask, v, o, SUB));
  return v;
}

__device__ __forceinline__ double warp_nanmax(double v) { return sub_nanmax(v, FULLMASK); }

// ------------------------------------------------------------------------------------------------
// CTA-shared tables (first SHARED_TABLE_DOUBLES doubles of dynamic shared memory): ROWPTR_S[NRED+1]
// (first dest of each condensed row) and CPOS_S[ND] (circular window position of each dest).
// ------------------------------------------------------------------------------------------------
// Geometry of the register-resident window (see band_solve): NPART lanes per window row, PW positions each.
// BS_SPARE = 1: WR + 1 row slots — the spare one lets the row that enters the window be staged a whole pivot step before it
// is needed (takes the staging round trip off the pivot chain).  Measured neutral on the lane-change game, and it costs a
// fourth quarter-warp of broadcast loads (26 instead of 24 active lanes), so the plans emit BS_SPARE = 0.
#ifndef BS_SPARE
#define BS_SPARE 0
#endif
#ifndef BS_PRED_ACTIVE
#define BS_PRED_ACTIVE 2   // 0: every lane loads the pivot row; 1: branch around idle lanes (measured −8 %); 2: predicated loads (+3 %)
#endif
#ifndef BS_SELF_CLEAR
#define BS_SELF_CLEAR 0    // (measured −6 % in the bench) staging rows cleared by the lanes that wrote them (one wavefront) instead of zero-filled per step (two)
#endif
#ifndef BS_PIV_FROM_U
#define BS_PIV_FROM_U 1    // pivot from the broadcast load + one multiplier shuffle, instead of two shuffles
#endif
#define BS_RS (WR + BS_SPARE)
#define BS_NPART ((BS_RS <= SUB) ? ((SUB / BS_RS) >= 4 ? 4 : ((SUB / BS_RS) >= 2 ? 2 : 1)) : 1)
#define BS_PW ((((WC + BS_SPARE + BS_NPART - 1) / BS_NPART) + 1) & ~1)
#define BS_REGWIN (REGWIN && BS_RS <= SUB && BS_PW <= REGWIN_PW_MAX)
#define BS_PREFETCH_AT ((NRED > 8) ? 8 : 1)   // pivot steps before the end of the factorisation at which Uᵀ is prefetched into the L2

template <int NRHS = 1>
__device__ __forceinline__ void load_shared_tables(double* smem_base, bool transposed = false) {
  int* rowptr = reinterpret_cast<int*>(smem_base);
  unsigned short* cpos = reinterpret_cast<unsigned short*>(rowptr + NRED + 1);
  LOAD_HOT_TABLES();   // (plan.cpp, HotTables: the per-step index / coefficient tables → static shared memory, when they fit)
#if HAS_ADJOINT
  if (transposed) {   // the column-major view of the same non-zeros: the factorisation then sees Cᵀ
    for (int i = threadIdx.x; i <= NRED; i += blockDim.x) rowptr[i] = DT_ROWPTR[i];
    for (int i = threadIdx.x; i < ND; i += blockDim.x) cpos[i] = (unsigned short)DT_CPOS[i];
  } else
#endif
  {
    for (int i = threadIdx.x; i <= NRED; i += blockDim.x) rowptr[i] = D_ROWPTR[i];
    for (int i = threadIdx.x; i < ND; i += blockDim.x) cpos[i] = (unsigned short)D_CPOS[i];
  }
  __syncthreads();
#if BS_REGWIN
  // The register window wants each non-zero's position RELATIVE to the first column of the window at the moment its
  // row enters: row r enters relative to column max(0, r − KL), so rel = col − max(0, r − KL) ∈ [0, WC).  Stored in
  // place of the circular position (col mod WC), from which it follows.
  // Rows 0 … WR are in the initial window (relative to column 0); row r > WR is staged at the end of pivot step
  // r − WR − 1, relative to column r − WR, one step before its first column (r − WR + 1 = r − KL) is eliminated: its
  // positions are 1 … WC.  Row WR itself sits at positions 1 … WC of the initial window.
  // (without the spare slot: rows 0 … WR−1 initial, row r ≥ WR staged at the end of step r − WR relative to column
  // r − WR + 1 = r − KL, positions 0 … WC−1)
  for (int r = threadIdx.x; r < NRED; r += blockDim.x) {
    const int first = (r >= BS_RS) ? (r - BS_RS + 1) % WC : 0;
    for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      int d = (int)cpos[e] - first;
      if (d < 0) d += WC;
      if (BS_SPARE && r >= WR && d == 0) d = WC;   // (col mod WC cannot tell WC from 0; these rows have no entry at position 0)
      cpos[e] = (unsigned short)d;
    }
  }
  __syncthreads();
#endif
}

// Named barrier among the NW warps that share one instance (ids 1…15; id 0 is __syncthreads).
__device__ __forceinline__ void wide_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Assembly of the condensed matrix C = G_x + tol·I − G_y D⁻¹ H_x: one value per structural non-zero
// ("dest", sorted by row then column) into the compact L2-resident array Cval[ND].
// ------------------------------------------------------------------------------------------------
// With NW > 1 the NW warps of a cooperative instance share the work (lane index wrole·SUB + sl, stride NW·SUB) and
// the phases are separated by the instance's named barrier instead of __syncwarp.
template <int NW = 1>
__device__ __forceinline__ void assemble_matrix(double* RS Cval, double* RS tmp,
                                                const double* RS jv, const double* RS th,
                                                const double* RS dinv, double tol, int sl, unsigned smask,
                                                int wrole = 0, int bar_id = 0) {
  const int wl = wrole * SUB + sl;   // lane index within the instance
  constexpr int WL = NW * SUB;
  // Constant contributions are folded into D_BASE on the host; only the z/θ/D-dependent terms remain.
#if ASM_TWO_PHASE
  // Two phases per chunk (chunks = runs of dests whose terms fit the shared term buffer):
  //   A, term-parallel (all table loads independent and coalesced): tmp[t] = coef·val(a)·[D⁻¹_k·val(b)]
  //   B, dest-parallel: sum the (contiguous) terms of each dest
#pragma unroll 1
  for (int c = 0; c < ASM_NCHUNK; ++c) {
    const int tb = CH_T[c], te = CH_T[c + 1];
#pragma unroll 4
    for (int t = tb + wl; t < te; t += WL) {
      const TI_T ti = T_I[t];  // {a, b, k, -}
      double v = T_COEF_AT(t) * opval(ti.x, jv, th);
      if (ti.z >= 0) v *= dinv[ti.z] * opval(ti.y, jv, th);
      tmp[t - tb] = v;
    }
    if constexpr (NW > 1) wide_bar(bar_id, NW * 32); else __syncwarp(smask);
    for (int d = CH_D[c] + wl; d < CH_D[c + 1]; d += WL) {
      const int tp = D_TP[d];
      const int t1 = D_TP[d + 1] & D_TP_MASK;
      double acc = D_BASE_AT(d) + ((tp < 0) ? tol : 0.0);  // sign bit of D_TP marks a diagonal dest
      for (int t = tp & D_TP_MASK; t < t1; ++t) acc += tmp[t - tb];
      __stcg(Cval + d, acc);   // streaming scratch: keep L1 for the assembly tables
    }
    if constexpr (NW > 1) wide_bar(bar_id, NW * 32); else __syncwarp(smask);
  }
#else
  for (int d = wl; d < ND; d += WL) {
    const int tp = D_TP[d];
    const int t1 = D_TP[d + 1] & D_TP_MASK;
    double acc = D_BASE_AT(d) + ((tp < 0) ? tol : 0.0);
    for (int t = tp & D_TP_MASK; t < t1; ++t) {
      const TI_T ti = T_I[t];
      double v = T_COEF_AT(t) * opval(ti.x, jv, th);
      if (ti.z >= 0) v *= dinv[ti.z] * opval(ti.y, jv, th);
      acc += v;
    }
    __stcg(Cval + d, acc);
  }
  if constexpr (NW > 1) wide_bar(bar_id, NW * 32);
#endif
}

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");  // L2 only
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Banded LU with partial pivoting on a sliding window held in shared memory, forward substitution
// folded into the elimination (RHS columns ride along as extra window columns), then a column-sweep
// back substitution.  One warp.
//
//   W    : shared, WR slots × WS doubles (WS ≡ 2 mod 4 ⇒ row-strided 128-bit accesses are conflict free).  Column c
//          of a row lives at circular position c % WC, the RHS at WC … WC+NRHS-1.  Slots are never
//          swapped: the pivot's slot is retired and re-used by the row entering the window, so partial
//          pivoting moves no data.
//   Cval : global, the assembled non-zeros (row-sorted); rows are scattered into the window as they enter.
//   UT   : global scratch, NRED × UTS (UTS = WC rounded up to even): U stored TRANSPOSED, UT[c*UTS + (c-i)] = U[i][c], with the
//          reciprocal pivot at offset 0, so that the back substitution reads one contiguous row per
//          column and needs no warp reduction.
//   sol  : shared, NRHS × NRED: right-hand sides on entry, solution on exit (permuted ordering).
// Lanes own window ROWS during the elimination (all rows update in parallel, zero multipliers are
// predicated off) and consecutive rows during the back-substitution sweep.
// Returns 0, or 1 if a pivot is zero / non-finite (the reference's `:failed` retcode branch,
// src/solver.jl:84-88).
// ------------------------------------------------------------------------------------------------
// One warp's share of the elimination sweep of the shared-memory window: column batches bi ≡ wrole (mod NW), all rows.
template <int NRHS, int WS, int NW>
__device__ __forceinline__ void window_sweep(double* W, const double* Wp, const double (&m)[(WR + SUB - 1) / SUB], int sl, int wrole) {
  constexpr int RPL = (WR + SUB - 1) / SUB;
  constexpr int NP = (WC + NRHS + 1) / 2;
  constexpr int PB = (NP < 9) ? NP : 9;
  const double2* Wp2 = reinterpret_cast<const double2*>(Wp);
#pragma unroll
  for (int b0 = 0; b0 < NP; b0 += PB) {
    if (NW > 1 && (b0 / PB) % NW != wrole) continue;
    double2 u[PB];
#pragma unroll
    for (int i = 0; i < PB; ++i)
      if (b0 + i < NP) u[i] = Wp2[b0 + i];
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      const int r = sl + SUB * k;
      if (r < WR) {
        double2* Wr2 = reinterpret_cast<double2*>(W + r * WS);
        double2 a[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i)
          if (b0 + i < NP) a[i] = Wr2[b0 + i];
#pragma unroll
        for (int i = 0; i < PB; ++i)
          if (b0 + i < NP) {
            a[i].x = fma(m[k], u[i].x, a[i].x);
            a[i].y = fma(m[k], u[i].y, a[i].y);
            Wr2[b0 + i] = a[i];
          }
      }
    }
  }
}

// Mailbox of a cooperative (NWIDE > 1) instance, behind its window: WR multipliers, then the command word and one
// double argument.  Commands: p >= 0 sweep the window for pivot slot p; WIDE_EXIT; WIDE_ASSEMBLE; WIDE_EVAL.
#define WIDE_EXIT (-1)
#define WIDE_ASSEMBLE (-2)
#define WIDE_EVAL (-3)
#define WIDE_MBOX(W_, WS_) ((W_) + WR * (WS_))
#define WIDE_CMD(W_, WS_) reinterpret_cast<volatile int*>(WIDE_MBOX(W_, WS_) + ((WR + 1) & ~1))
#define WIDE_ARG(W_, WS_) (WIDE_MBOX(W_, WS_) + ((WR + 1) & ~1) + 1)

template <int NRHS, int WS, int NW = 1>
__device__ int band_solve(double* RS W, const double* RS Cval, double* RS UT,
                          double* RS sol, const int* RS rowptr,
                          const unsigned short* RS cpos, const double* RS jv,
                          const double* RS th, const double* RS dinv, double* RS stage,
                          int sl, unsigned smask, int bar_id = 0, const int* RS src = nullptr) {
  // `src` (adjoint solves): dest e of the row tables is Cval[src[e]] — with the column-major tables the window then
  // holds Cᵀ without re-assembling anything
  auto cval_at = [&](int e) -> double { return __ldcg(Cval + (src ? src[e] : e)); };
  constexpr int CPW = (WC + SUB - 1) / SUB;  // matrix positions per sl
  constexpr int RPL = (WR + SUB - 1) / SUB;  // row slots per sl
  constexpr int WPL = (WS + SUB - 1) / SUB;

  // 2-D register layout: NPART lanes per window row, PW (even) consecutive positions each
  constexpr int NPART = BS_NPART;
  constexpr int PW = BS_PW;
  constexpr int WCP = PW * NPART;               // padded matrix width (≥ WC + 1)
  constexpr int ES = (WCP + NRHS + 3) & ~1;     // stride of the publish / staging rows (even, one pair of slack)
  // (r2 experiment, removed: a COLUMN-owner register window — lane ℓ owns window column ≡ ℓ mod WC with all WR row slots
  // in registers, pivot search local to one lane, multipliers broadcast by shuffle, no pivot-row publish / reload — cut the
  // shared-memory + shuffle operations per pivot step from ≈62 to ≈35 but needs ≈186 instead of ≈134 issued instructions
  // per step (local search, two uniform switches on the pivot slot, per-slot multiplier shuffles); measured in the
  // fixed-work mode 178.3 vs 167.7 ms: the step is bound by the warp's serial instruction latency, not by the MIO pipe.)
  if constexpr (BS_REGWIN) {
    // ============ register-resident window ==========================================================
    // Window rows live in REGISTERS for their whole life in the window, each row split over NPART lanes:
    // lane (row, part) holds a[i] = the entry in column j + part·PW + i — a layout relative to the pivot
    // column, so eliminating column j and advancing the window are ONE operation,
    //     a[i] ← a[i+1] − m·u[i+1],
    // with the FMA's destination register doing the shift for free (position 0 is always the pivot column,
    // the loop body stays small and rolled).  The pivot row is published once to shared memory and read back
    // with aligned 128-bit broadcast loads; shared memory otherwise only stages the entering row.
    // Why: with the window itself in shared memory (variant below) the smem pipe sat at 60 % of peak
    // re-loading/re-storing rows, and broadcasting by shuffle costs the same crossbar as many wavefronts.
    // lane = row·NPART + part: the parts of one row sit in the same quarter-warp, so the two-lane 128-bit publish /
    // reload of a row costs one shared-memory wavefront per instruction instead of one per lane.
    //
    // r2 — the pivot step is a latency chain (ncu: ≈1150 cycles per step for ≈130 issued instructions, six dependent
    // shared-memory / shuffle round trips), so the loop is organised to take work OFF that chain:
    //   * BS_RS = WR + 1 row slots: the spare slot receives the next row a whole step before its first column is
    //     eliminated, so its staging (scatter → barrier → reload) is off the chain.  A freshly loaded row has a
    //     structural zero in the next pivot column; `head` (the lane's pivot-column entry, kept apart from a[0]) is set
    //     to zero for it, so neither the pivot search nor the multipliers wait for the reload;
    //   * the pivot row is published as soon as the pivot is known and its broadcast loads are issued before the
    //     reciprocal / multiplier arithmetic, which then runs in their shadow;
    //   * positions relative to the entering window are precomputed (load_shared_tables), the pivot's validity is read
    //     off the reduction key, the reciprocal is MUFU + two Newton steps, right-hand sides move straight between
    //     registers and `sol`, and every per-step address is a running pointer.
    const bool active = sl < BS_RS * NPART;
    const int row = active ? sl / NPART : BS_RS;
    const int part = active ? sl % NPART : 0;
    double a[PW], rh[NRHS];
    double* Pb = W;       // published pivot row: WCP matrix positions, then the RHS
    double* E = W + ES;   // two staging rows, alternating between steps
#pragma unroll
    for (int i = 0; i < PW; ++i) a[i] = 0.0;
#pragma unroll
    for (int q = 0; q < NRHS; ++q) rh[q] = 0.0;
    constexpr int NFILL = (BS_RS < NRED) ? BS_RS : NRED;
    for (int r = 0; r < NFILL; ++r) {  // rows 0 … WR, relative to column 0
      for (int q = sl; q < ES; q += SUB) E[q] = 0.0;
      __syncwarp(smask);
      for (int e = rowptr[r] + sl; e < rowptr[r + 1]; e += SUB) E[cpos[e]] = cval_at(e);
      if (sl < NRHS) E[WCP + sl] = sol[sl * NRED + r];
      __syncwarp(smask);
      if (active && row == r) {
        const double2* src = reinterpret_cast<const double2*>(E + part * PW);
#pragma unroll
        for (int k = 0; k < PW / 2; ++k) {
          const double2 v = src[k];
          a[2 * k] = v.x;
          a[2 * k + 1] = v.y;
        }
#pragma unroll
        for (int q = 0; q < NRHS; ++q) rh[q] = E[WCP + q];
      }
      __syncwarp(smask);
    }
    const bool lane_pub = active && part == 0;
    double head = a[0];                            // my row's entry in the pivot column (part-0 lanes)
    int e_lo = (BS_RS < NRED) ? rowptr[BS_RS] : 0;   // first non-zero of the row staged at the end of step 0 (row WR + 1)
    const int* rpj = rowptr + ((BS_RS < NRED) ? BS_RS + 1 : 0);   // → rowptr[(row being staged) + 1]
    double* UTj = UT + (size_t)sl * (UTS + 1);     // lane t writes U[j][j+t] to UT[(j+t)·UTS + t] = UTj[0], UTj += UTS
    double* solj = sol;                            // → sol[j]; the staged row's right-hand side is solj[BS_RS]
    double* Eb = E;
    double* Eo = E + ES;
    for (int q = sl; q < 2 * ES; q += SUB) E[q] = 0.0;   // both staging rows start (and are kept) all-zero
    int zrel[CPW];
#pragma unroll
    for (int k = 0; k < CPW; ++k) zrel[k] = -1;
    __syncwarp(smask);
    int left = NRED;                               // NRED − j
#if BS_PRED_ACTIVE == 2
    double u[PW + 2];
#pragma unroll
    for (int i = 0; i < PW + 2; ++i) u[i] = 0.0;
#endif
#pragma unroll 1
    for (; left > 0; --left) {
      if (left == BS_PREFETCH_AT) {
        // The back substitution reads Uᵀ from its END; the rows written first (one factorisation ago for the L2: all
        // resident instances stream ~100 MB through it in that time) are the ones most likely to have been evicted.
        // Ask for the whole array now, a few thousand cycles ahead, so the sweep finds it in the L2.
        for (int i = sl; i < (NRED * UTS + 15) / 16; i += SUB) prefetch_l2(UT + (size_t)i * 16);
      }
      const bool entering = left > BS_RS;           // row j + WR + 1 exists
      const int e_hi = entering ? *rpj : e_lo;
      double pre[CPW];
      int prel[CPW];
#pragma unroll
      for (int k = 0; k < CPW; ++k) {   // prefetch the staged row's non-zeros (consumed at the end of the step)
        const int e = e_lo + sl + SUB * k;
        prel[k] = -1;
        pre[k] = 0.0;
        if (e < e_hi) {
          pre[k] = cval_at(e);
          prel[k] = cpos[e];
        }
      }
#if !BS_SELF_CLEAR
      for (int q = sl; q < ES; q += SUB) Eb[q] = 0.0;   // (this buffer was last read two steps ago)
#endif
      // ---- pivot search over column j: max |head| on a 12-bit-truncated mantissa, row in the low byte ----
      unsigned key = 0;
      if (lane_pub) key = ((unsigned)__double2hiint(fabs(head)) & 0xffffff00u) | (unsigned)(255 - row);
      const unsigned best = __reduce_max_sync(smask, key);
      // zero / denormal (exponent field 0), Inf / NaN or ≥ 2^1022 (reciprocal not representable): the reference's
      // `:failed` retcode branch (src/solver.jl:84-88)
      if (best < 0x00100000u || best >= 0x7fd00000u) return 1;
      const int p = 255 - (int)(best & 0xffu);
      const bool mine = active && row == p;
      // ---- publish the pivot row at once; its right-hand side retires straight into sol -------------------
      if (mine) {
        double2* dst = reinterpret_cast<double2*>(Pb + part * PW);
#pragma unroll
        for (int k = 0; k < PW / 2; ++k) dst[k] = make_double2(a[2 * k], a[2 * k + 1]);
        if (part == NPART - 1) {   // the right-hand sides ride with the row's LAST part: Pb[WCP] is the pair it loads anyway
#pragma unroll
          for (int q = 0; q < NRHS; ++q) {
            Pb[WCP + q] = rh[q];
            solj[q * NRED] = rh[q];
          }
        }
      }
      __syncwarp(smask);
      // the staging row of the PREVIOUS step has been consumed (its reload precedes the barrier above): the lanes that
      // scattered into it clear their entries, so both staging rows are all-zero whenever they are scattered into
#if BS_SELF_CLEAR
#pragma unroll
      for (int k = 0; k < CPW; ++k)
        if (zrel[k] >= 0) Eo[zrel[k]] = 0.0;
#endif
      // ---- broadcast loads of the pivot row (in flight while the multipliers are computed) ------------------
      // Only the lanes that hold window rows load (the idle quarter-warp costs no shared-memory wavefronts).
#if BS_PRED_ACTIVE != 2
      double u[PW + 2];   // the pivot row as seen by my part
#endif
      double urh[NRHS];
      double piv = 1.0;
      if (BS_PRED_ACTIVE != 1 || active) {   // BS_PRED_ACTIVE == 1: only the lanes that hold rows load and update (a branch)
#if BS_PRED_ACTIVE == 2
        // predicated (not branched) 128-bit loads: the idle quarter-warp issues no shared-memory wavefronts, the warp
        // stays converged; u is loop-carried (idle lanes keep the zeros it was initialised with)
        const unsigned ua = (unsigned)__cvta_generic_to_shared(Pb + part * PW);
        const int act = active ? 1 : 0;
#pragma unroll
        for (int k = 0; k <= PW / 2; ++k)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.shared.v2.f64 {%0, %1}, [%3];\n\t}"
                       : "+d"(u[2 * k]), "+d"(u[2 * k + 1]) : "r"(act), "r"(ua + 16u * k) : "memory");
#else
        const double2* up = reinterpret_cast<const double2*>(Pb + part * PW);
#pragma unroll
        for (int k = 0; k <= PW / 2; ++k) {  // PW/2 + 1 aligned pairs: my part and the first entry of the next
          const double2 v = up[k];
          u[2 * k] = v.x;
          u[2 * k + 1] = v.y;
        }
#endif
        urh[0] = u[PW];   // last part: Pb[WCP], the pivot row's first right-hand side (other parts never publish theirs)
#pragma unroll
        for (int q = 1; q < NRHS; ++q) urh[q] = Pb[WCP + q];
        piv = u[0];       // part-0 lanes: Pb[0], the pivot itself (the other parts never use their rp)
      }
      double ut_out[CPW];
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int t = sl + SUB * k;
        ut_out[k] = (t < WC) ? Pb[t] : 0.0;
      }
      // ---- pivot, reciprocal, multipliers --------------------------------------------------------------------
      // (the multiplier is formed by the row's part-0 lane, which holds the pivot-column entry, and handed to the other
      // parts with ONE shuffle — shuffles share the shared-memory data pipe, which is what bounds this loop)
#if !BS_PIV_FROM_U
      piv = __shfl_sync(smask, head, p * NPART, SUB);
#endif
      double rp;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rp) : "d"(piv));   // ~20 bits, then two Newton steps: ≤ 1 ulp
      {
        double er = fma(-piv, rp, 1.0);
        rp = fma(rp, er, rp);
        er = fma(-piv, rp, 1.0);
        rp = fma(rp, er, rp);
      }
#if BS_PIV_FROM_U
      double m = (active && !mine) ? -(head * rp) : 0.0;
      if constexpr (NPART > 1) m = __shfl_sync(smask, m, sl - part, SUB);
#else
      double a0row = head;
      if constexpr (NPART > 1) a0row = __shfl_sync(smask, head, sl - part, SUB);   // my row's entry in the pivot column
      const double m = (active && !mine) ? -(a0row * rp) : 0.0;
#endif
      double nx = 0.0;
      if constexpr (NPART > 1) {
        nx = __shfl_sync(smask, a[0], (sl + 1) % SUB, SUB);     // first entry of my row's next part
        if (part == NPART - 1) nx = 0.0;
      }
      // ---- U row j goes out transposed (coalesced over the lanes) ------------------------------------------
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int t = sl + SUB * k;
#ifndef EXP_NO_UT_STORE
        if (t < WC && t < left) __stcg(UTj + (size_t)k * SUB * (UTS + 1), (t == 0) ? rp : ut_out[k]);
#endif
      }
      // ---- eliminate column j and slide the window: a[i] ← a[i+1] − m·u[i+1] ------------------------------
      if (BS_PRED_ACTIVE != 1 || active) {
#pragma unroll
        for (int i = 0; i + 1 < PW; ++i) a[i] = fma(m, u[i + 1], a[i + 1]);
        a[PW - 1] = (part == NPART - 1) ? 0.0 : fma(m, u[PW], nx);   // last part: column j+WCP enters, zero
#pragma unroll
        for (int q = 0; q < NRHS; ++q) rh[q] = fma(m, urh[q], rh[q]);
      }
      head = a[0];
      // ---- the next row (relative to column j+1, first needed at step j+2) takes over the retired row's lanes ----
#pragma unroll
      for (int k = 0; k < CPW; ++k)
        if (prel[k] >= 0) Eb[prel[k]] = pre[k];
      __syncwarp(smask);
      if (mine) {
        const double2* src = reinterpret_cast<const double2*>(Eb + part * PW);
#pragma unroll
        for (int k = 0; k < PW / 2; ++k) {
          const double2 v = src[k];
          a[2 * k] = v.x;
          a[2 * k + 1] = v.y;
        }
#pragma unroll
        for (int q = 0; q < NRHS; ++q) rh[q] = entering ? solj[q * NRED + BS_RS] : 0.0;
        head = BS_SPARE ? 0.0 : a[0];   // with the spare slot: structurally zero in column j+1, nothing waits for the reload
      }
#if BS_SELF_CLEAR
#pragma unroll
      for (int k = 0; k < CPW; ++k) zrel[k] = prel[k];
#endif
      e_lo = e_hi;
      if (entering) ++rpj;
      UTj += UTS;
      ++solj;
      {
        double* t_ = Eb;
        Eb = Eo;
        Eo = t_;
      }
    }
    __syncwarp(smask);
  } else {
  // ============ shared-memory window (any size) =====================================================
  // ---- initial window: rows 0 … WR-1 -----------------------------------------------------------
  for (int i = sl; i < WR * WS; i += SUB) W[i] = 0.0;
  __syncwarp(smask);
  for (int r = 0; r < WR; ++r) {
    for (int e = rowptr[r] + sl; e < rowptr[r + 1]; e += SUB) W[r * WS + cpos[e]] = cval_at(e);
    if (sl < NRHS) W[r * WS + WC + sl] = sol[sl * NRED + r];
  }
  __syncwarp(smask);
#if DENSE_SCHUR
  // Dense problems (the whole condensed matrix is resident: WR == NRED, WC == NRED): the Schur complement
  // −G_y D⁻¹ H_x is accumulated as one rank-1 update per constraint k, (G_y[:,k] D⁻¹_k) ⊗ H_x[k,:], from two
  // staged dense vectors — a term-by-term table would cost nx²·ny lookups (10⁶ per Newton step for the
  // 100×100 QP).  Rows whose G_y entry is numerically zero are skipped.
  {
    double* hb = stage;            // H_x[k, :]   (new column ordering; WC == NRED so position == column)
    double* gb = stage + STAGE_N;  // G_y[:, k]·D⁻¹_k (new row ordering)
    constexpr int NPR = (NRED + 1) / 2;
    for (int k = 0; k < NY; ++k) {
      for (int i = sl; i < 2 * STAGE_N; i += SUB) stage[i] = 0.0;
      __syncwarp(smask);
#pragma unroll 4
      for (int e = H_PTR[k] + sl; e < H_PTR[k + 1]; e += SUB) hb[H_COL[e]] = H_COEF_AT(e) * opval(H_CODE[e], jv, th);
      const double dk = dinv[k];
#pragma unroll 4
      for (int e = GK_PTR[k] + sl; e < GK_PTR[k + 1]; e += SUB) gb[GK_ROW[e]] = GK_COEF_AT(e) * opval(GK_CODE[e], jv, th) * dk;
      __syncwarp(smask);
      const double2* hb2 = reinterpret_cast<const double2*>(hb);
#pragma unroll 1
      for (int kk = 0; kk < RPL; ++kk) {
        const int r = sl + SUB * kk;
        const double gr = (r < NRED) ? gb[r] : 0.0;
        if (gr != 0.0) {
          double2* Wr2 = reinterpret_cast<double2*>(W + r * WS);
#pragma unroll 4
          for (int c2 = 0; c2 < NPR; ++c2) {
            const double2 hv = hb2[c2];
            double2 v = Wr2[c2];
            v.x = fma(-gr, hv.x, v.x);
            v.y = fma(-gr, hv.y, v.y);
            Wr2[c2] = v;
          }
        }
      }
      __syncwarp(smask);
    }
  }
#endif

  constexpr int NP = (WC + NRHS + 1) / 2;  // position pairs swept per row (matrix + rhs columns)
  constexpr int PB = (NP < 9) ? NP : 9;    // pairs per register batch
  static_assert(2 * NP <= WS, "row stride must cover the padded position pairs");

  int cj = 0;  // j % WC
  for (int j = 0; j < NRED; ++j) {
    // prefetch the non-zeros of the row that enters the window at the end of this step
    const int ienter = j + WR;
    int e0 = 0, e1 = 0;
    if (ienter < NRED) {
      e0 = rowptr[ienter];
      e1 = rowptr[ienter + 1];
    }
    double pre[CPW];
#pragma unroll
    for (int k = 0; k < CPW; ++k) {
      const int e = e0 + sl + SUB * k;
      pre[k] = (e < e1) ? cval_at(e) : 0.0;   // written with st.cg: must not be served from a stale L1 line
    }

    // ---- pivot search over column j: max |a| on a 12-bit-truncated mantissa, slot in the low byte ----
    unsigned best = 0;
    double m[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      const int r = sl + SUB * k;
      m[k] = (r < WR) ? W[r * WS + cj] : 0.0;
      if (r < WR) {
        const unsigned key = ((unsigned)__double2hiint(fabs(m[k])) & 0xffffff00u) | (unsigned)(255 - r);
        best = max(best, key);
      }
    }
    best = __reduce_max_sync(smask, best);
    const int p = 255 - (int)(best & 0xffu);
    double piv = m[0];
#pragma unroll
    for (int k = 1; k < RPL; ++k)
      if ((p / SUB) == k) piv = m[k];
    piv = __shfl_sync(smask, piv, p % SUB, SUB);
    if (!(fabs(piv) > 0.0) || !(fabs(piv) <= DBL_MAX_)) return 1;  // zero, NaN or Inf pivot
    const double rp = 1.0 / piv;
    double* Wp = W + p * WS;
    // Invariant: every entry of a window row outside its structural extent is exactly zero, so the
    // update can sweep ALL positions with static code; only the pivot's own position must read as zero.
    if (sl == 0) Wp[cj] = 0.0;
    __syncwarp(smask);

    // ---- retire the pivot row: U row j goes out transposed, its RHS into sol ------------------------
    {
      const int tmax = min(WC - 1, NRED - 1 - j);
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int q = sl + SUB * k;
        if (q < WC) {
          int d = q - cj;
          if (d < 0) d += WC;
          if (d <= tmax) __stcg(UT + (size_t)(j + d) * UTS + d, (d == 0) ? rp : Wp[q]);
        }
      }
      if (sl < NRHS) sol[sl * NRED + j] = Wp[WC + sl];
    }

    // ---- eliminate column j: every row (lanes own rows) minus multiplier × pivot row -----------------
    // Rows with a zero multiplier (and the pivot row itself, whose multiplier is forced to zero) are
    // rewritten unchanged: no per-sl branches, 128-bit conflict-free accesses (WS ≡ 2 mod 4).
    {
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int r = sl + SUB * k;
        m[k] = (r < WR && r != p) ? -(m[k] * rp) : 0.0;
      }
      if constexpr (NW > 1) {
        // cooperative sweep: publish the pivot slot and the multipliers, all NW warps of the instance sweep their
        // column batches between the two barriers (the helpers wait at barrier A whenever the leader is elsewhere)
        double* mbox = WIDE_MBOX(W, WS);
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const int r = sl + SUB * k;
          if (r < WR) mbox[r] = m[k];
        }
        if (sl == 0) WIDE_CMD(W, WS)[0] = p;
        wide_bar(bar_id, NW * 32);
        window_sweep<NRHS, WS, NW>(W, Wp, m, sl, 0);
        wide_bar(bar_id, NW * 32);
      } else {
        window_sweep<NRHS, WS, 1>(W, Wp, m, sl, 0);
      }
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int r = sl + SUB * k;
        if (r < WR && r != p) W[r * WS + cj] = 0.0;  // the eliminated entry (exactly zero by construction)
      }
    }
    __syncwarp(smask);

    // ---- the entering row takes the retired slot ---------------------------------------------------
    {
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        const int q = sl + SUB * k;
        if (q < WS) Wp[q] = 0.0;
      }
      __syncwarp(smask);
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int e = e0 + sl + SUB * k;
        if (e < e1) Wp[cpos[e]] = pre[k];
      }
      if (sl < NRHS && ienter < NRED) Wp[WC + sl] = sol[sl * NRED + ienter];
    }
    __syncwarp(smask);
    cj = (cj + 1 == WC) ? 0 : cj + 1;
  }

  }

  // ---- back substitution: column sweep, x_j = rhs_j / u_jj then rhs_i −= U[i][j] x_j for i < j.
  // The right-hand sides live in REGISTERS while they are being updated: lane L owns the rows i ≡ L (mod SUB) of the
  // dependency window (NACC of them), so one column costs one shared load (the lane's entry of Uᵀ's row j), one
  // multiply + broadcast by the row's owner and one FMA per accumulator — `sol` is read and written once per row
  // instead of once per (row, column).  The columns of U (rows of UT) stream back through two cp.async buffers of
  // BS_CH rows each that re-use the window's shared memory.
#ifndef EXP_NO_BACKSUB
  {
    // Each lane needs ONE entry of Uᵀ's row j per accumulator (its own row's coefficient in column j), so the rows of
    // Uᵀ are read straight from the L2 into registers, a block of BS_BD columns ahead of their use (two register
    // blocks, ping-pong): ≈ BS_BD·50 cycles of lead cover the L2 latency, nothing goes through shared memory, there
    // is no barrier inside the sweep.  (r2 measurement, fixed-work mode: the cp.async ring this replaces kept only
    // 4–8 rows in flight and the sweep ran at ≈400 cycles per column — 19 % of the whole Newton step.)
    constexpr int NACC = (WC - 1) / SUB + 1;
#ifndef BS_BD_MAX
#define BS_BD_MAX 8
#endif
    constexpr int BS_BD = (BS_BD_MAX / (NACC * NRHS) >= 4) ? BS_BD_MAX / (NACC * NRHS) : 4;
    double acc[NRHS][NACC];
    const int t0s = (((NRED - 1 - sl) % SUB) + SUB) % SUB;   // (column − lane) mod SUB at column NRED − 1
#pragma unroll
    for (int q = 0; q < NRHS; ++q)
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        const int i = NRED - 1 - t0s - SUB * k;
        acc[q][k] = (i >= 0) ? sol[q * NRED + i] : 0.0;
      }
    // my entries of Uᵀ's rows jhi, jhi − 1, …: row j − t of column j with t = ((j − lane) mod SUB) + SUB·k
    auto load_block = [&](int jhi, double (&dst)[BS_BD][NACC]) {
#pragma unroll
      for (int jj = 0; jj < BS_BD; ++jj) {
        const int j = jhi - jj;
        const int t0 = (j - sl) & (SUB - 1);
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
          const int t = t0 + SUB * k;
          dst[jj][k] = (j >= 0 && t < WC) ? __ldcg(UT + (size_t)j * UTS + t) : 0.0;   // t == 0: the reciprocal pivot
        }
      }
    };
    // columns jhi … jhi − BS_BD + 1: x_j from its owner (lane j mod SUB), then every lane updates the rows it owns
    // inside the dependency window.  Columns below 0 (last block) are virtual: they only touch accumulators of rows
    // that do not exist; the control flow stays uniform, so the shuffles need no re-convergence code.
    auto sweep_block = [&](int jhi, const double (&ub)[BS_BD][NACC]) {
#pragma unroll
      for (int jj = 0; jj < BS_BD; ++jj) {
        const int j = jhi - jj;
        const int t0 = (j - sl) & (SUB - 1);
#pragma unroll
        for (int q = 0; q < NRHS; ++q) {
          const double xj = __shfl_sync(smask, acc[q][0] * ub[jj][0], j & (SUB - 1), SUB);
#pragma unroll
          for (int k = 0; k < NACC; ++k) {   // rows j − t, t = t0 + SUB·k ∈ [1, WC)
            const int t = t0 + SUB * k;
            if (t >= 1 && t < WC) acc[q][k] = fma(-ub[jj][k], xj, acc[q][k]);
          }
          if (t0 == 0 && j >= 0) {   // row j is finished: store it, my accumulators move on to rows j − SUB, j − 2·SUB, …
            sol[q * NRED + j] = xj;
#pragma unroll
            for (int k = 0; k + 1 < NACC; ++k) acc[q][k] = acc[q][k + 1];
            const int inext = j - SUB * NACC;
            acc[q][NACC - 1] = (inext >= 0) ? sol[q * NRED + inext] : 0.0;
          }
        }
      }
    };
    __syncwarp(smask);   // the factorisation's stores to UT / sol (other lanes) are visible from here
    double ua[BS_BD][NACC], ub[BS_BD][NACC];
    load_block(NRED - 1, ua);
#pragma unroll 1
    for (int jhi = NRED - 1; jhi >= 0; jhi -= 2 * BS_BD) {
      load_block(jhi - BS_BD, ub);
      sweep_block(jhi, ua);
      load_block(jhi - 2 * BS_BD, ua);
      sweep_block(jhi - BS_BD, ub);
    }
    __syncwarp(smask);
  }
#endif
  return 0;
}

// `fraction_to_the_boundary_linesearch` — src/solver.jl:127-138, literally (τ = 0.995, decay = 0.5).
__device__ __forceinline__ double ftb_linesearch(const double* RS v, const double* RS d,
                                                 double min_step, int sl, unsigned smask) {
  const double c = 1.0 - 0.995;
  double alpha = 1.0;
  for (int it = 0; it < 1200; ++it) {
    bool viol = false;
    for (int k = sl; k < NY; k += SUB) viol = viol || (v[k] + alpha * d[k] < c * v[k]);  // :129
    if (!__any_sync(smask, viol)) return alpha;
    if (alpha < min_step) break;  // :130 — tested before halving
    alpha *= 0.5;                 // :134
  }
  return __longlong_as_double(0x7ff8000000000000LL);  // NaN (:131)
}

#if DENSE_KERNEL
// ------------------------------------------------------------------------------------------------
// Dense solve kernel (plans whose whole condensed matrix is resident and whose Schur product is dense,
// e.g. the QP configs): ONE CTA of 256 threads per instance, two CTAs per SM.
//   * G/H evaluation split over up to 256 generated parts (thread i evaluates part i);
//   * C = G_x + tol·I − G_y D⁻¹ H_x : the direct part is written dest-parallel, the Schur part is a
//     register-tiled rank-k update — thread (ti, tj) of a 16×16 grid owns a TR×TR tile of C in registers
//     and consumes constraints in blocks of DKB, staged (scaled by D⁻¹) into shared memory;
//   * LU with partial pivoting on the shared-memory matrix, threads mapped (row, column half), RHS as an
//     extra column; pivot rows stay in place (no scratch), a pivot order array drives the column-sweep back
//     substitution.
// The solver logic is the same literal restatement of src/solver.jl:63-121 as in mcp_solve_kernel.
// ------------------------------------------------------------------------------------------------
#if DENSE_KERNEL == 3
#define DT 512
#else
#define DT 256
#endif
#define DKB 8
#define DTR ((NRED + 15) / 16)  // register tile edge
#define DNP (16 * DTR)          // padded dimension of the staged vectors

__device__ __forceinline__ double blk_nanmax(double v, double* red, int t) {
  v = sub_nanmax(v, FULLMASK);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = v;
  __syncthreads();
  double r = red[0];
#pragma unroll
  for (int i = 1; i < DT / 32; ++i) r = nanmax(r, red[i]);
  return r;
}

// `fraction_to_the_boundary_linesearch` (src/solver.jl:127-138) with the whole CTA
__device__ __forceinline__ double blk_ftb_linesearch(const double* RS v, const double* RS d,
                                                     double min_step, int t) {
  const double c = 1.0 - 0.995;
  double alpha = 1.0;
  for (int it = 0; it < 1200; ++it) {
    bool viol = false;
    for (int k = t; k < NY; k += DT) viol = viol || (v[k] + alpha * d[k] < c * v[k]);  // :129
    if (!__syncthreads_or(viol)) return alpha;
    if (alpha < min_step) break;  // :130
    alpha *= 0.5;                 // :134
  }
  return __longlong_as_double(0x7ff8000000000000LL);
}

#if DENSE_KERNEL == 3
// ------------------------------------------------------------------------------------------------
// Dense kernel v3: the condensed matrix never touches shared memory.  One CTA of 16 warps per instance;
// lane l owns rows {l, l+32, …} (D3_RCH chunks), warp w owns columns {w, w+16, …} (D3_CPW of them, the
// right-hand side is column NRED), so every thread keeps a D3_RCH × D3_CPW tile of C in registers from the
// Schur accumulation through the whole factorisation:
//   * (G_x + tol·I)ᵀ (θ-only for these plans) is built once per solve into an L2-resident global block and
//     loaded coalesced into the tiles each Newton step;  + H_xᵀ D⁻¹ H_x is accumulated on top from the
//     cached H_x (4 conflict-free + 7 broadcast shared loads per 28 DFMAs);
//   * LU with partial pivoting, one __syncthreads per column: column j lives entirely in warp j mod 16, which
//     finds the pivot with one warp-wide max, and publishes the pivot row index and the multipliers; after the
//     barrier lane (pivot row mod 32) of every warp stores its part of the pivot row into Uᵀ (transposed, pivot
//     order) and clears it in the tile, the warp reads it back as broadcasts and updates its columns > j
//     (finished columns are skipped warp-uniformly);
//   * back substitution is a column sweep by warp 0 alone with the right-hand side in registers (no barriers).
// The solver logic is the same literal restatement of src/solver.jl:63-121 as in mcp_solve_kernel.
// ------------------------------------------------------------------------------------------------
#define D3_RCH ((NRED + 31) / 32)
#define D3_CPW ((NRED + 1 + 15) / 16)
#define D3_CPN ((NRED + 15) / 16)
#define D3_GLD 128
#ifndef D3_SCHUR_UNROLL
#define D3_SCHUR_UNROLL 2
#endif
constexpr int kSchurUnroll = D3_SCHUR_UNROLL;
#ifndef D3_LOOKAHEAD
#define D3_LOOKAHEAD 0
#endif

// Pivot search for column j (tile column A of the calling warp, which holds the whole column) and publication of
// the pivot row index, the reciprocal pivot and the multipliers of every row.  Partial pivoting with the same key
// as the other kernels: truncated |value| first, lowest row index among ties.  Rows that have been pivots (and the
// padding rows) are all-zero in the tiles — the update phase clears a row once it is stored in Uᵀ — so they need
// no masking: their key is below that of any usable pivot and their multiplier is ∓0.  The reciprocal is taken
// per lane on the lane's own best candidate, in the shadow of the warp-wide max.  This runs on ONE warp while the
// other fifteen wait at the barrier, so it is written for the shortest dependent instruction chain.
template <int A>
__device__ __forceinline__ void d3_search(const double (&acc)[D3_RCH][D3_CPW], const int lane, const int j, double* mbuf,
                                          int* sh_pr, double* rd) {
  const int par = j & 1;
  unsigned k[D3_RCH];
  unsigned key = 0;
#pragma unroll
  for (int b = 0; b < D3_RCH; ++b) {
    k[b] = ((unsigned)__double2hiint(acc[b][A]) & 0x7fffff00u) | (unsigned)(255 - (lane + 32 * b));
    key = max(key, k[b]);
  }
  double vbest = acc[0][A];
#pragma unroll
  for (int b = 1; b < D3_RCH; ++b) vbest = (k[b] == key) ? acc[b][A] : vbest;
  const double rpl = 1.0 / vbest;
  const unsigned best = __reduce_max_sync(FULLMASK, key);
  const int prr = 255 - (int)(best & 0xffu);
  const double piv = __shfl_sync(FULLMASK, vbest, prr & 31);   // keys are unique per row: that lane's best is the pivot
  const double rp = __shfl_sync(FULLMASK, rpl, prr & 31);
  const bool bad = !(fabs(piv) > 0.0) || !(fabs(piv) <= DBL_MAX_);  // :84-88 (a finished or padding row can only win with 0)
  double* mb = mbuf + par * 128 + lane;
#pragma unroll
  for (int b = 0; b < D3_RCH; ++b) mb[32 * b] = -(acc[b][A] * rp);
  if (lane == (prr & 31)) mbuf[par * 128 + prr] = 0.0;   // the pivot row itself (same thread as the store above)
  if (lane == 0) {
    sh_pr[par] = bad ? -1 : prr;
    rd[j] = rp;
  }
}

// 16 elimination steps (columns 16·AJ … 16·AJ+15) of the register-resident LU; AJ is a template parameter so
// that every tile index is static.  With D3_LOOKAHEAD the owner of the NEXT column updates that column first and
// runs its search while the other warps are still in their trailing updates (measured slower on B200: the owner's
// extra work lengthens the step more than the overlap saves, so it is off by default).
// Returns true when a pivot was rejected (src/solver.jl:84-88).
template <int AJ>
__device__ __forceinline__ bool d3_lu_blocks(double (&acc)[D3_RCH][D3_CPW], const int wid, const int lane, double* mbuf,
                                             int* sh_pr, double* rd, double* UT) {
  constexpr int UTLD = DENSE_UTLD;
  if constexpr (AJ >= D3_CPN) {
    return false;
  } else {
#pragma unroll 1
    for (int wj = 0; wj < 16; ++wj) {
      const int j = AJ * 16 + wj;
      if (j >= NRED) break;
      const int par = j & 1;
#if !D3_LOOKAHEAD
      if (wid == wj) d3_search<AJ>(acc, lane, j, mbuf, sh_pr, rd);
#endif
      __syncthreads();
      const int pr = sh_pr[par];
      if (pr < 0) return true;
      double m[D3_RCH];
#pragma unroll
      for (int b = 0; b < D3_RCH; ++b) m[b] = mbuf[par * 128 + lane + 32 * b];
      const int src = pr & 31, pb = pr >> 5;
      // Pivot row → U: lane src holds this warp's part of it in row chunk pb; it stores the entries straight into Uᵀ
      // (transposed, pivot order — what the back substitution reads), clears them in the tile (a finished row must
      // be all-zero, see d3_search) and the warp reads them back as broadcasts.
      // pb is CTA-uniform, so the chunk is chosen by a branch around plain stores: no register selects.
      double* ut = UT + wid * UTLD + j;
      if (lane == src) {
#define D3_PUT(B_)                                                                                          \
  if (D3_RCH > B_ && pb == B_) {                                                                            \
    _Pragma("unroll") for (int a = AJ; a < D3_CPW; ++a)                                                      \
      if (16 * a + 15 <= NRED || wid + 16 * a <= NRED) ut[16 * a * UTLD] = acc[B_ < D3_RCH ? B_ : 0][a];     \
    _Pragma("unroll") for (int a = AJ; a < D3_CPW; ++a) acc[B_ < D3_RCH ? B_ : 0][a] = 0.0;                  \
  }
        D3_PUT(0) else D3_PUT(1) else D3_PUT(2) else D3_PUT(3)
#undef D3_PUT
      }
      __syncwarp();
      double u[D3_CPW];
#pragma unroll
      for (int a = AJ; a < D3_CPW; ++a) u[a] = (16 * a + 15 <= NRED || wid + 16 * a <= NRED) ? ut[16 * a * UTLD] : 0.0;
      // look-ahead: column j+1
      int la = -1;   // tile column already updated by the look-ahead
#if D3_LOOKAHEAD
      if (j + 1 < NRED) {
        if (wj < 15) {
          if (wid == wj + 1) {
#pragma unroll
            for (int b = 0; b < D3_RCH; ++b) acc[b][AJ] = fma(m[b], u[AJ], acc[b][AJ]);
            d3_search<AJ>(acc, lane, j + 1, mbuf, sh_pr, rd);
            la = AJ;
          }
        } else if constexpr (AJ + 1 < D3_CPN) {
          if (wid == 0) {
#pragma unroll
            for (int b = 0; b < D3_RCH; ++b) acc[b][AJ + 1] = fma(m[b], u[AJ + 1], acc[b][AJ + 1]);
            d3_search<AJ + 1>(acc, lane, j + 1, mbuf, sh_pr, rd);
            la = AJ + 1;
          }
        }
      }
#endif
#pragma unroll
      for (int a = AJ; a < D3_CPW; ++a)
        if ((a > AJ || wid > wj) && a != la) {
#pragma unroll
          for (int b = 0; b < D3_RCH; ++b) acc[b][a] = fma(m[b], u[a], acc[b][a]);
        }
    }
    return d3_lu_blocks<AJ + 1>(acc, wid, lane, mbuf, sh_pr, rd, UT);
  }
}

// ---- symmetric path (r2) -------------------------------------------------------------------------------------------
// The condensed matrix of these plans is C = G_x + tol·I + H_xᵀ D⁻¹ H_x (G_y = −H_xᵀ is a plan-time condition of the
// dense kernels), symmetric whenever G_x is — every QP — and positive definite whenever G_x is positive semi-definite
// (D⁻¹ > 0 in the interior).  Then the partial-pivoted LU above is replaced by LDLᵀ WITHOUT pivoting, in a layout of
// its own: lane l owns rows {l, l+32, …} as before, warp w owns the column PAIRS {2w, 2w+1} + 32a, and only the tile
// blocks on or below the diagonal exist (row chunk b ≥ column block a: 20 doubles per thread for n = 100 instead
// of 28, in the Schur accumulation too).  One synchronisation per PANEL of two columns: the owner warp eliminates
// column j₀ from column j₁ locally (the panel's diagonal 2×2 block sits in chunk a, lanes 2w and 2w+1 of its tiles, and
// is mirrored in every lane: d3p_diag_update) and publishes both columns straight into Uᵀ (by symmetry row j of U is
// column j) with the two reciprocal pivots; everybody then applies a rank-2 update (two dependent FMAs per entry, same
// rounding as column by column).  No pivot search, no pivot-row store→load, finished row chunks and column blocks are
// skipped statically, and the owner of the NEXT panel updates and factorises it first (look-ahead).  The right-hand
// side (forward substitution) lives in four registers of warp D3P_RW and never crosses warps.  Entries above the
// diagonal, padding rows and padding columns hold garbage that is never read.
// Whether G_x is symmetric is checked once per solve on the values (it comes from θ); a pivot that is not positive
// and finite sends that Newton step — and the rest of the instance — back to the pivoted LU, so nothing the
// reference solves is lost (src/solver.jl:81-88: UMFPACK factorises any non-singular matrix).  Same Uᵀ / rd layout
// as the LU, so the back substitution is shared.
#ifndef D3_SYM
#define D3_SYM 1
#endif
#define D3P_RW 15   // warp that carries the right-hand side

// One mbarrier per panel instead of a CTA-wide barrier per panel (D3_SYM_MBAR): the owner arrives once its panel is
// in shared memory, everybody else waits for THAT — not for the other fourteen warps — so a warp that is late (the one
// carrying the right-hand side, the next owner) delays nobody who does not need its data (measured: −2.7 % kernel time;
// the warps still move nearly in lockstep).  Every cell of Uᵀ / rd is written once per factorisation, so
// there is nothing to protect against overwriting.  The barriers are initialised ONCE per CTA (re-initialising a live
// mbarrier is undefined) and each completes exactly once per factorisation — an abandoned factorisation completes the
// remaining ones by hand — so the waiters' phase parity is the parity of the CTA's factorisation count.
#ifndef D3_SYM_MBAR
#define D3_SYM_MBAR 1
#endif
#define D3P_NP ((NRED + 1) / 2)
__device__ __forceinline__ void d3_mbar_init(unsigned long long* bar, const unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void d3_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void d3_mbar_wait(unsigned long long* bar, const unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2, 2000;\n\t"   // (suspend-time hint, ns)
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done) {   // a logic error must not hang the device: give up (launch failure) after ~1 s
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 2000000000LL) __trap();
    }
  } while (!done);
}

__device__ __forceinline__ double d3_rcp(const double d) {
  double rp;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rp) : "d"(d));   // ~20 bits, then two Newton steps: ≤ 1 ulp
  double er = fma(-d, rp, 1.0);
  rp = fma(rp, er, rp);
  er = fma(-d, rp, 1.0);
  return fma(rp, er, rp);
}

// The 2×2 diagonal block of the calling warp's NEXT panel, kept redundantly by all its lanes: dg = (a00, a10, a11).
// It receives the same two FMAs per panel as the tile entries it mirrors (lanes 2w, 2w+1 of chunk AD), from the
// same operands — the pivot-row entries of the warp's own columns, which are also the multipliers of those two rows —
// so it is bit-identical to them, and the panel's factorisation needs no shuffle: the owner's critical path is
// wait → loads → arithmetic → stores → arrive.
__device__ __forceinline__ void d3p_diag_update(double (&dg)[3], const double2 v0, const double2 v1, const double rp0,
                                                const double rp1, const bool two) {
  const double ma0 = -(v0.x * rp0), mb0 = -(v1.x * rp0);          // multipliers of rows c0, c0+1 for step j0
  const double ma1 = two ? -(v0.y * rp1) : 0.0, mb1 = two ? -(v1.y * rp1) : 0.0;   // … for step j0+1
  dg[0] = fma(ma1, v0.y, fma(ma0, v0.x, dg[0]));
  dg[1] = fma(mb1, v0.y, fma(mb0, v0.x, dg[1]));
  dg[2] = fma(mb1, v1.y, fma(mb0, v1.x, dg[2]));
}

template <int AD>
__device__ __forceinline__ void d3p_diag_seed(double (&dg)[3], const double (&T)[D3_RCH][D3_RCH][2], const int wid) {
  if constexpr (AD < D3_RCH) {
    dg[0] = __shfl_sync(FULLMASK, T[AD][AD][0], 2 * wid);
    dg[1] = __shfl_sync(FULLMASK, T[AD][AD][0], 2 * wid + 1);
    dg[2] = __shfl_sync(FULLMASK, T[AD][AD][1], 2 * wid + 1);
  }
}

// Factorise the panel (columns j0, j0+1) held by the calling warp in column block A, whose entries (and dg) are up to
// date with every earlier panel, and publish it: UT[r·UTLD + j] = entry (r, j) for the rows of chunks ≥ A (rows
// above the diagonal land in cells nothing reads), rd[j] = 1/d_j — NaN when d_j is not positive and finite.
template <int A>
__device__ __forceinline__ void d3p_panel(double (&T)[D3_RCH][D3_RCH][2], const double (&dg)[3], const int lane,
                                          const int j0, double* UT, double* rd, const bool lastrow,
                                          unsigned long long* bars) {
  constexpr int UTLD = DENSE_UTLD;
  const bool two = j0 + 1 < NRED;
  const double d0 = dg[0], u0 = dg[1];   // u0 = U[j0][j0+1]
  const double rp0 = d3_rcp(d0);
  const double d1 = fma(-(u0 * rp0), u0, dg[2]);
  const double rp1 = d3_rcp(d1);
  double2* q = reinterpret_cast<double2*>(UT + lane * UTLD + j0);   // (even row stride, even j0: 16-byte aligned)
#pragma unroll
  for (int b = A; b < D3_RCH; ++b) {
    T[b][A][1] = fma(-(T[b][A][0] * rp0), u0, T[b][A][1]);
    if (b < D3_RCH - 1 || lastrow) q[16 * b * UTLD] = make_double2(T[b][A][0], T[b][A][1]);   // (cell j0+1 = NRED is unused)
  }
  if (lane == 0) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    rd[j0] = (d0 > 0.0 && d0 <= DBL_MAX_ && rp0 == rp0) ? rp0 : qnan;
    if (two) rd[j0 + 1] = (d1 > 0.0 && d1 <= DBL_MAX_ && rp1 == rp1) ? rp1 : qnan;
  }
#if D3_SYM_MBAR
  __syncwarp();   // every lane's cells before the one arrival
  if (lane == 0) d3_mbar_arrive(bars + (j0 >> 1));
#endif
}

// the pivot-row entries of the calling warp's two columns of block AC for panel (j0, j0+1)
template <int AC>
__device__ __forceinline__ void d3p_load_u(double2& v0, double2& v1, const double* RS UT, const int wid, const int j0) {
  constexpr int UTLD = DENSE_UTLD;
  const int c0 = 32 * AC + 2 * wid;
  // (only the last column block can run past the matrix: padding columns read the right-hand-side row, finite or
  // not — their tile entries are never read)
  v0 = *reinterpret_cast<const double2*>(UT + (AC == D3_RCH - 1 ? min(c0, NRED) : c0) * UTLD + j0);
  v1 = *reinterpret_cast<const double2*>(UT + (AC == D3_RCH - 1 ? min(c0 + 1, NRED) : c0 + 1) * UTLD + j0);
}

// rank-2 update of the calling thread's tile block (b ≥ AC, column block AC) with panel (j0, j0+1)
template <int AC>
__device__ __forceinline__ void d3p_apply(double (&T)[D3_RCH][D3_RCH][2], const double (&m0)[D3_RCH],
                                          const double (&m1)[D3_RCH], const double2 v0, const double2 v1) {
#pragma unroll
  for (int b = AC; b < D3_RCH; ++b) {
    T[b][AC][0] = fma(m1[b], v0.y, fma(m0[b], v0.x, T[b][AC][0]));
    T[b][AC][1] = fma(m1[b], v1.y, fma(m0[b], v1.x, T[b][AC][1]));
  }
}

template <int AC>
__device__ __forceinline__ void d3p_update(double (&T)[D3_RCH][D3_RCH][2], const double (&m0)[D3_RCH],
                                           const double (&m1)[D3_RCH], const double* RS UT, const int wid, const int j0) {
  if constexpr (AC < D3_RCH) {
    double2 v0, v1;
    d3p_load_u<AC>(v0, v1, UT, wid, j0);
    d3p_apply<AC>(T, m0, m1, v0, v1);
  }
}

// The panels of column block A (columns 32·A … 32·A+31, owner warps 0 … 15 in turn); returns −1, or the index of the
// panel whose pivot was not positive (the caller falls back to the pivoted LU; nothing beyond that panel has been
// published or signalled, because the look-ahead comes after the test).  On entry the panel (32·A, 32·A+1) has
// been published by warp 0 and dg mirrors the diagonal block of the calling warp's next panel (block A for warps
// whose turn is still to come, block A+1 for warp 0).
template <int A>
__device__ __forceinline__ int d3p_blocks(double (&T)[D3_RCH][D3_RCH][2], double (&rv)[D3_RCH], double (&dg)[3],
                                          const int wid, const int lane, double* rd, double* UT, const bool lastrow,
                                          unsigned long long* bars, const unsigned parity) {
  constexpr int UTLD = DENSE_UTLD;
  if constexpr (A >= D3_RCH) {
    return -1;
  } else {
    // my rows' cells of Uᵀ (rows of chunks ≥ A are still live; padding rows read the right-hand-side row)
    const double* mq = UT + lane * UTLD;
    const double* mql = UT + min(lane + 32 * (D3_RCH - 1), NRED) * UTLD;
#pragma unroll 1
    for (int wo = 0; wo < 16; ++wo) {
      const int j0 = 32 * A + 2 * wo;
      if (j0 >= NRED) break;
      const bool two = j0 + 1 < NRED;
      const bool mine_ahead = wid > wo;   // my panel of block A is still to come: dg mirrors block A, else block A+1
#if D3_SYM_MBAR
      d3_mbar_wait(bars + (j0 >> 1), parity);
#else
      __syncthreads();
#endif
      // every load of the step up front (they do not depend on the pivot test)
      const double rp0 = rd[j0];
      const double rp1 = two ? rd[j0 + 1] : 0.0;
      double2 mv[D3_RCH], vd0, vd1;
#pragma unroll
      for (int b = A; b < D3_RCH; ++b)
        mv[b] = *reinterpret_cast<const double2*>((b == D3_RCH - 1 ? mql : mq + 32 * b * UTLD) + j0);
      if (mine_ahead) {
        d3p_load_u<A>(vd0, vd1, UT, wid, j0);
      } else if constexpr (A + 1 < D3_RCH) {
        d3p_load_u<A + 1>(vd0, vd1, UT, wid, j0);
      }
      if (!(rp0 == rp0) || !(rp1 == rp1)) return j0 >> 1;
      if (mine_ahead || A + 1 < D3_RCH) d3p_diag_update(dg, vd0, vd1, rp0, rp1, two);
      double m0[D3_RCH], m1[D3_RCH];
#pragma unroll
      for (int b = A; b < D3_RCH; ++b) {
        m0[b] = -(mv[b].x * rp0);
        m1[b] = two ? -(mv[b].y * rp1) : 0.0;
      }
      // look-ahead: the owner of the next panel brings its block up to date, factorises and publishes it first
      bool reseed = false;
      if (mine_ahead) {
        d3p_apply<A>(T, m0, m1, vd0, vd1);
        if (wid == wo + 1 && j0 + 2 < NRED) {
          d3p_panel<A>(T, dg, lane, j0 + 2, UT, rd, lastrow, bars);
          reseed = true;
        }
      } else if constexpr (A + 1 < D3_RCH) {
        d3p_apply<A + 1>(T, m0, m1, vd0, vd1);
        if (wo == 15 && wid == 0 && j0 + 2 < NRED) {
          d3p_panel<A + 1>(T, dg, lane, j0 + 2, UT, rd, lastrow, bars);
          reseed = true;
        }
      }
      if (mine_ahead) d3p_update<A + 1>(T, m0, m1, UT, wid, j0);
      d3p_update<A + 2>(T, m0, m1, UT, wid, j0);
      d3p_update<A + 3>(T, m0, m1, UT, wid, j0);
      if (reseed) {   // my next panel is one column block further on
        if (wo == 15) d3p_diag_seed<A + 2>(dg, T, wid);
        else d3p_diag_seed<A + 1>(dg, T, wid);
      }
      if (wid == D3P_RW) {   // forward substitution of the right-hand side (rows j0, j0+1 are in chunk A, lanes 2wo, 2wo+1)
        const double r0 = __shfl_sync(FULLMASK, rv[A], 2 * wo);
        if (lane == 0) UT[NRED * UTLD + j0] = r0;
#pragma unroll
        for (int b = A; b < D3_RCH; ++b) rv[b] = fma(m0[b], r0, rv[b]);
        if (two) {
          const double r1 = __shfl_sync(FULLMASK, rv[A], 2 * wo + 1);
          if (lane == 0) UT[NRED * UTLD + j0 + 1] = r1;
#pragma unroll
          for (int b = A; b < D3_RCH; ++b) rv[b] = fma(m1[b], r1, rv[b]);
        }
      }
    }
    return d3p_blocks<A + 1>(T, rv, dg, wid, lane, rd, UT, lastrow, bars, parity);
  }
}

// tile of the symmetric layout: the direct part (G_x + tol·I) from the L2-resident block, then + H_xᵀ D⁻¹ H_x, one
// rank-1 update per constraint from the cached H_x (4 conflict-free + 8 broadcast shared loads per 20 DFMAs)
__device__ __forceinline__ void d3p_assemble(double (&T)[D3_RCH][D3_RCH][2], const double* Gc, const double* RS Hc,
                                             const double* RS dinv, const int wid, const int lane) {
  constexpr int HCS = DENSE_HCS;
#pragma unroll
  for (int b = 0; b < D3_RCH; ++b)
#pragma unroll
    for (int a = 0; a <= b; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = lane + 32 * b, c = 32 * a + 2 * wid + e;
        T[b][a][e] = (r < NRED && c < NRED) ? __ldcg(Gc + c * D3_GLD + r) : 0.0;
      }
  int ri[D3_RCH], ci[D3_RCH];
#pragma unroll
  for (int b = 0; b < D3_RCH; ++b) {
    ri[b] = min(lane + 32 * b, NRED);   // column NRED of the cached H_x stays 0
    ci[b] = min(32 * b + 2 * wid, NRED - 1);   // (ci+1 ≤ NRED)
  }
#pragma unroll kSchurUnroll
  for (int k = 0; k < NY; ++k) {
    const double dk = dinv[k];
    const double* hr = Hc + k * HCS;
    double av[D3_RCH], bv[D3_RCH][2];
#pragma unroll
    for (int b = 0; b < D3_RCH; ++b) av[b] = hr[ri[b]] * dk;
#pragma unroll
    for (int a = 0; a < D3_RCH; ++a) {
      bv[a][0] = hr[ci[a]];
      bv[a][1] = hr[ci[a] + 1];
    }
#pragma unroll
    for (int b = 0; b < D3_RCH; ++b)
#pragma unroll
      for (int a = 0; a <= b; ++a) {
        T[b][a][0] = fma(av[b], bv[a][0], T[b][a][0]);
        T[b][a][1] = fma(av[b], bv[a][1], T[b][a][1]);
      }
  }
}

// The same tile when H_x is sparse IN VALUE (the benchmark's A has 10 % non-zeros; the plan only knows that every
// entry is a parameter): row i of C = (G_x + tol·I) + Σ_k (H_x[k,i]·D⁻¹_k)·H_x[k,:] only sums the constraints k with
// H_x[k,i] ≠ 0 — listed once per solve (KLIST / kcnt) — in the same order, so the values are those of the dense
// accumulation (the skipped terms are ±0).  Warp w builds rows i ≡ w (mod 16) in registers, lanes over the columns,
// into S (the Uᵀ array, free at this point); the tiles are then read from S with 128-bit loads.
__device__ __forceinline__ void d3p_assemble_sparse(double (&T)[D3_RCH][D3_RCH][2], const double* Gc, const double* RS Hc,
                                                    const double* RS dinv, const unsigned char* RS KLIST, const int* RS kcnt,
                                                    double* S, const int wid, const int lane) {
  constexpr int HCS = DENSE_HCS, UTLD = DENSE_UTLD, KLS = DENSE_KLS;
  int ji[D3_RCH];
#pragma unroll
  for (int q = 0; q < D3_RCH; ++q) ji[q] = min(lane + 32 * q, NRED);   // column NRED of the cached H_x stays 0
  for (int i = wid; i < NRED; i += DT / 32) {
    // (only the column chunks up to the one holding the diagonal: the tiles read the lower triangle, whole diagonal
    // blocks included)
    double sa[D3_RCH];
#pragma unroll
    for (int q = 0; q < D3_RCH; ++q)
      sa[q] = (32 * q <= i && lane + 32 * q < NRED) ? __ldcg(Gc + i * D3_GLD + lane + 32 * q) : 0.0;
    const unsigned char* kl = KLIST + i * KLS;
    const int n = kcnt[i];
#pragma unroll 4
    for (int e = 0; e < n; ++e) {
      const int k = kl[e];
      const double* hr = Hc + k * HCS;
      const double av = hr[i] * dinv[k];
#pragma unroll
      for (int q = 0; q < D3_RCH; ++q)
        if (32 * q <= i) sa[q] = fma(av, hr[ji[q]], sa[q]);
    }
#pragma unroll
    for (int q = 0; q < D3_RCH; ++q)
      if (32 * q <= i && lane + 32 * q < NRED) S[i * UTLD + lane + 32 * q] = sa[q];
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < D3_RCH; ++b)
#pragma unroll
    for (int a = 0; a <= b; ++a) {
      const int r = min(lane + 32 * b, NRED - 1), c = min(32 * a + 2 * wid, UTLD - 2);   // (padding: any finite cell)
      const double2 v = *reinterpret_cast<const double2*>(S + r * UTLD + c);
      T[b][a][0] = v.x;
      T[b][a][1] = v.y;
    }
}

// + H_xᵀ D⁻¹ H_x on the register tiles of the LU layout, one rank-1 update per constraint from the cached H_x (4
// conflict-free + 7 broadcast shared loads per 28 DFMAs)
__device__ __forceinline__ void d3_schur(double (&acc)[D3_RCH][D3_CPW], const double* RS Hc, const double* RS dinv,
                                         const int wid, const int lane) {
  constexpr int HCS = DENSE_HCS;
#pragma unroll kSchurUnroll
  for (int k = 0; k < NY; ++k) {
    const double dk = dinv[k];
    const double* hr = Hc + k * HCS;
    double av[D3_RCH], bv[D3_CPW];
#pragma unroll
    for (int b = 0; b < D3_RCH; ++b) av[b] = hr[min(lane + 32 * b, NRED)] * dk;
#pragma unroll
    for (int a = 0; a < D3_CPW; ++a) bv[a] = hr[min(wid + 16 * a, NRED)];
#pragma unroll
    for (int b = 0; b < D3_RCH; ++b)
#pragma unroll
      for (int a = 0; a < D3_CPW; ++a) acc[b][a] = fma(av[b], bv[a], acc[b][a]);
  }
}

// the register tile of the direct part (G_x + tol·I)ᵀ, built once per solve in the L2-resident block Gc
__device__ __forceinline__ void d3_load_tile(double (&acc)[D3_RCH][D3_CPW], const double* Gc, const int wid, const int lane) {
#pragma unroll
  for (int b = 0; b < D3_RCH; ++b)
#pragma unroll
    for (int a = 0; a < D3_CPW; ++a) {
      const int r = lane + 32 * b, c = wid + 16 * a;
      acc[b][a] = (r < NRED && c < NRED) ? __ldcg(Gc + c * D3_GLD + r) : 0.0;
    }
}

extern "C" __global__ void __launch_bounds__(DT, 1) mcp_solve_kernel(const SolveParams p) {
  extern __shared__ double smem[];
  const int t = threadIdx.x;
  const int wid = t >> 5, lane = t & 31;
  double* x = smem + DENSE_OFF_X;
  double* y = smem + DENSE_OFF_Y;
  double* s = smem + DENSE_OFF_S;
  double* g = smem + DENSE_OFF_G;
  double* w = smem + DENSE_OFF_W;        // H rows, then w, then δy
  double* dinv = smem + DENSE_OFF_DINV;  // D⁻¹, then δs
  double* sol = smem + DENSE_OFF_SOL;    // right-hand side, then δx (new ordering)
  double* jv = smem + DENSE_OFF_JV;      // computed Jacobian entries (none for these plans)
  double* red = smem + DENSE_OFF_RED;    // 16 doubles of reduction scratch
  double* rd = smem + DENSE_OFF_RD;      // reciprocal pivots
  double* Hc = smem + DENSE_OFF_HC;      // H_x cached for the whole solve: NY rows × DENSE_HCS (column NRED stays 0)
  double* xt = smem + DENSE_OFF_XT;      // x in the new ordering
  double* gh0 = smem + DENSE_OFF_G0;     // G(0;θ), H(0;θ)
  double* UT = smem + DENSE_OFF_UT;      // Uᵀ in pivot order: UT[c·UTLD + step]; row NRED = forward-substituted rhs
  double* mbuf = smem + DENSE_OFF_MBUF;  // 2 × 128 multipliers of the current column (double-buffered by parity)
  unsigned char* KLIST = reinterpret_cast<unsigned char*>(smem + DENSE_OFF_KL);   // non-zeros of H_x by column
  int* kcnt = reinterpret_cast<int*>(smem + DENSE_OFF_KCNT);
  double* Gc = p.scratch + (size_t)blockIdx.x * SOLVE_SCRATCH;  // (G_x + tol·I)ᵀ, NRED columns × D3_GLD
  constexpr int HCS = DENSE_HCS;
  constexpr int UTLD = DENSE_UTLD;
  __shared__ unsigned long long sh_q;
  __shared__ int sh_pr[2];
  __shared__ int sh_nnz;
#if D3_SYM
  __shared__ unsigned long long d3_bars[D3P_NP];
#endif
#if THETA_IN_SMEM
  double* th = smem + DENSE_OFF_TH;
#else
  const double* th = p.theta;
#endif
  const double tol = p.tol;
  const unsigned long long n_deferred = p.pass ? p.counters[3] : 0ULL;
#if D3_SYM
  unsigned nfact = 0;   // symmetric factorisations of this CTA (phase parity of the panel barriers)
#if D3_SYM_MBAR
  if (t < D3P_NP) d3_mbar_init(d3_bars + t, 1);   // one arrival (the panel's owner) completes a phase
#endif
#endif

  for (;;) {
    __syncthreads();
    if (t == 0) {
      sh_q = atomicAdd(p.counters + (p.pass ? 4 : 0), 1ULL);
      sh_nnz = 0;
    }
    __syncthreads();
    unsigned long long inst = sh_q;
    if (p.pass) {
      if (inst >= n_deferred) break;
      inst = (unsigned long long)p.deferred[inst];
    } else if (inst >= (unsigned long long)p.B) {
      break;
    }
#if THETA_IN_SMEM
    for (int i = t; i < NT; i += DT) th[i] = p.theta[inst * NT + i];
#else
    th = p.theta + inst * NT;
#endif
    double eps = 1.0;                                        // :67
    double kkt = __longlong_as_double(0x7ff0000000000000LL);  // :68
    int status = 0, outer = 1, steps = 0;                    // :69-70
    if (p.pass) {
      for (int i = t; i < NX; i += DT) x[i] = p.x_out[inst * NX + i];
      for (int i = t; i < NY; i += DT) {
        y[i] = p.y_out[inst * NY + i];
        s[i] = p.s_out[inst * NY + i];
      }
      eps = p.eps_out[inst];
      kkt = p.kkt_out[inst];
      outer = p.outer_out[inst];
      steps = p.steps_out[inst];
    } else {
      for (int i = t; i < NX; i += DT) x[i] = p.x0 ? p.x0[inst * NX + i] : 0.0;
      for (int i = t; i < NY; i += DT) {
        y[i] = p.y0 ? p.y0[inst * NY + i] : 1.0;
        s[i] = p.s0 ? p.s0[inst * NY + i] : 1.0;
      }
    }
    __syncthreads();
    // ---- once per solve: cache H_x and G_x + tol·I (θ-only), evaluate the constant part of the residual ------------
    for (int i = t; i < NY * HCS; i += DT) Hc[i] = 0.0;
    for (int i = t; i < NRED * D3_GLD; i += DT) __stcg(Gc + i, 0.0);
    if (t < 256) mcp_eval_const_par(t, x, y, th, gh0);
    __syncthreads();
    for (int k = wid; k < NY; k += DT / 32)
      for (int e = H_PTR[k] + lane; e < H_PTR[k + 1]; e += 32) Hc[k * HCS + H_COL[e]] = H_COEF_AT(e) * opval(H_CODE[e], jv, th);
    for (int d = t; d < ND; d += DT) {
      const int tp = D_TP[d], t1 = D_TP[d + 1] & D_TP_MASK;
      double a0 = D_BASE_AT(d) + ((tp < 0) ? tol : 0.0);
      for (int q = tp & D_TP_MASK; q < t1; ++q) a0 += T_COEF_AT(q) * opval(T_I[q].x, jv, th);
      __stcg(Gc + D_CPOS[d] * D3_GLD + D_ROW[d], a0);
    }
    __syncthreads();
    // is G_x + tol·I symmetric (to 1e-14 relative, far inside the factorisation's own backward error)?  NaN ⇒ no.
    bool use_sym = false;
#if D3_SYM
    {
      bool symok = true;
      for (int idx = t; idx < NRED * NRED; idx += DT) {
        const int r = idx / NRED, c = idx - r * NRED;
        if (r > c) {
          const double lo = __ldcg(Gc + c * D3_GLD + r), up = __ldcg(Gc + r * D3_GLD + c);
          symok = symok && (fabs(lo - up) <= 1e-14 * (fabs(lo) + fabs(up)));
        }
      }
      use_sym = __syncthreads_and(symok);
    }
    // the non-zeros of H_x by column, for the sparse form of the Schur product (d3p_assemble_sparse)
    bool sparse_hx = false;
#ifndef D3_SPARSE_SCHUR
#define D3_SPARSE_SCHUR 1
#endif
#if D3_SPARSE_SCHUR
    if (use_sym) {
      for (int i = wid; i < NRED; i += DT / 32) {
        int base = 0;
        for (int kb = 0; kb < NY; kb += 32) {
          const int k = kb + lane;
          const bool nz = k < NY && Hc[k * HCS + i] != 0.0;   // (a NaN counts as a non-zero)
          const unsigned bal = __ballot_sync(FULLMASK, nz);
          if (nz) KLIST[i * DENSE_KLS + base + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)k;
          base += __popc(bal);
        }
        if (lane == 0) {
          kcnt[i] = base;
          atomicAdd(&sh_nnz, base);
        }
      }
      __syncthreads();
      sparse_hx = 3 * sh_nnz <= NY * NRED;   // denser than a third: the dense accumulation is the faster one
    }
#endif
#endif
    bool parked = false;
    while (kkt > tol && eps > tol && outer < p.max_outer) {  // :71
      if (p.pass == 0 && p.step_budget > 0 && steps >= p.step_budget) {
        parked = true;
        break;
      }
      int inner = 1;  // :72
      status = 0;     // :73
      while (kkt > eps && inner < p.max_inner) {  // :75
        for (int c = t; c < NRED; c += DT) xt[c] = x[PERM[c]];
        // ---- tile of the direct part G_x + tol·I ------------------------------------------------------------------
        double acc[D3_RCH][D3_CPW];
        d3_load_tile(acc, Gc, wid, lane);
        __syncthreads();
        // ---- F (:79) from the affine structure: H = H(0) + H_x x,  G = G(0) + G_x x + G_y y,  G_y = −H_xᵀ ----------
        {
          double* stage = UT;   // 16 × 128 partial row sums of (G_x + tol·I)·x, one slice per warp
#pragma unroll
          for (int b = 0; b < D3_RCH; ++b) {
            double sa = 0.0;
#pragma unroll
            for (int a = 0; a < D3_CPN; ++a) {
              const int c = wid + 16 * a;
              if (c < NRED) sa = fma(acc[b][a], xt[c], sa);
            }
            stage[wid * 128 + lane + 32 * b] = sa;
          }
        }
        if (t < NY) {
          double a0 = 0.0, a1 = 0.0;
          const double* hr = Hc + t * HCS;
#pragma unroll 4
          for (int c = 0; c + 1 < NRED; c += 2) {
            a0 = fma(hr[c], xt[c], a0);
            a1 = fma(hr[c + 1], xt[c + 1], a1);
          }
          if (NRED & 1) a0 = fma(hr[NRED - 1], xt[NRED - 1], a0);
          w[t] = gh0[NX + t] + a0 + a1;
        }
        __syncthreads();
        if (t >= 128 && t < 128 + NRED) {  // G in the NEW row ordering (row i ↔ old row PERM[i])
          const int i = t - 128;
          double a1 = 0.0, a2 = 0.0, gx = 0.0;
#pragma unroll
          for (int q = 0; q < 16; ++q) gx += UT[q * 128 + i];
#pragma unroll 4
          for (int k = 0; k + 1 < NY; k += 2) {
            a1 = fma(Hc[k * HCS + i], y[k], a1);
            a2 = fma(Hc[(k + 1) * HCS + i], y[k + 1], a2);
          }
          if (NY & 1) a1 = fma(Hc[(NY - 1) * HCS + i], y[NY - 1], a1);
          g[i] = (gh0[PERM[i]] - tol * xt[i] + gx) - (a1 + a2);
        }
        __syncthreads();
        double fmax_ = 0.0;
        for (int i = t; i < NX; i += DT) fmax_ = nanmax(fmax_, fabs(g[i]));
        for (int k = t; k < NY; k += DT) {
          const double f2 = w[k] - s[k];
          const double f3 = s[k] * y[k] - eps;
          const double yt = y[k] + tol;
          const double di = 1.0 / (tol + s[k] / yt);
          dinv[k] = di;
          w[k] = di * (-f2 - f3 / yt);
          fmax_ = nanmax(fmax_, nanmax(fabs(f2), fabs(f3)));
        }
        const double kkt_new = blk_nanmax(fmax_, red, t);  // :107 (also orders the writes above)
        // ---- right-hand side: −G − G_y w = −G + H_xᵀ w, into column NRED of the tiles -----------------------------
        if (t < NRED) {
          double a0 = -g[t], a1 = 0.0;
#pragma unroll 4
          for (int k = 0; k + 1 < NY; k += 2) {
            a0 = fma(Hc[k * HCS + t], w[k], a0);
            a1 = fma(Hc[(k + 1) * HCS + t], w[k + 1], a1);
          }
          if (NY & 1) a0 = fma(Hc[(NY - 1) * HCS + t], w[NY - 1], a0);
          sol[t] = a0 + a1;
        }
        // ---- Schur part: C −= G_y D⁻¹ H_x = + H_xᵀ D⁻¹ H_x, accumulated on the register tiles; factorisation with the
        //      forward substitution riding in column NRED -------------------------------------------------------------
        bool failed = false;
        do {
#if D3_SYM
        if (use_sym) {
          double T[D3_RCH][D3_RCH][2], rv[D3_RCH];
          const bool lastrow = lane + 32 * (D3_RCH - 1) < NRED;
          if (sparse_hx) {
            __syncthreads();   // the staging use of UT is over, D⁻¹ is complete
            d3p_assemble_sparse(T, Gc, Hc, dinv, KLIST, kcnt, UT, wid, lane);
          } else {
            d3p_assemble(T, Gc, Hc, dinv, wid, lane);
          }
          __syncthreads();   // sol (the right-hand side) is complete; every tile is loaded before Uᵀ is written
#pragma unroll
          for (int b = 0; b < D3_RCH; ++b) rv[b] = (lane + 32 * b < NRED) ? sol[lane + 32 * b] : 0.0;
          double dg[3];
          d3p_diag_seed<0>(dg, T, wid);
          if (wid == 0) {
            d3p_panel<0>(T, dg, lane, 0, UT, rd, lastrow, d3_bars);
            d3p_diag_seed<1>(dg, T, wid);
          }
          const unsigned par = nfact & 1u;
          ++nfact;
          const int pfail = d3p_blocks<0>(T, rv, dg, wid, lane, rd, UT, lastrow, d3_bars, par);
          if (pfail < 0) break;
          use_sym = false;   // a pivot was not positive: pivoted LU from here on
          __syncthreads();
#if D3_SYM_MBAR
          if (t > pfail && t < D3P_NP) d3_mbar_arrive(d3_bars + t);   // every barrier completes once per factorisation
#endif
          d3_load_tile(acc, Gc, wid, lane);
        }
#endif
        {
          d3_schur(acc, Hc, dinv, wid, lane);
          __syncthreads();
          if (wid == (NRED & 15)) {
#pragma unroll
            for (int b = 0; b < D3_RCH; ++b) acc[b][NRED >> 4] = (lane + 32 * b < NRED) ? sol[lane + 32 * b] : 0.0;
          }
#if D3_LOOKAHEAD
          if (wid == 0) d3_search<0>(acc, lane, 0, mbuf, sh_pr, rd);
#endif
          failed = d3_lu_blocks<0>(acc, wid, lane, mbuf, sh_pr, rd, UT);
        }
        } while (false);
        __syncthreads();   // UT, rd complete (failed is CTA-uniform: every thread read the same key and pivot)
        double a_s = 1.0, a_y = 1.0;
        if (!failed) {
          // ---- back substitution: column sweep over Uᵀ by one warp, right-hand side in registers ---------------------
          if (wid == 0) {
            double rhs[D3_RCH];
#pragma unroll
            for (int b = 0; b < D3_RCH; ++b) rhs[b] = (lane + 32 * b < NRED) ? UT[NRED * UTLD + lane + 32 * b] : 0.0;
            // (two columns per trip — one round of shuffles, the 2×2 block solved in registers; same operations in the
            // same order as column by column.  Everything a trip reads from shared memory — two reciprocal pivots, the
            // cross entry, the two columns of Uᵀ — is loaded one trip ahead, so only shuffle → multiply → FMA is on the
            // dependent chain)
            constexpr int NPAIR = (NRED + 1) / 2;
            double rdc0, rdc1, crc, uc0c[D3_RCH], uc1c[D3_RCH];
            auto bs_load = [&](const int pp, double& r0_, double& r1_, double& cr_, double (&u0_)[D3_RCH], double (&u1_)[D3_RCH]) {
              const int q0 = 2 * pp, q1 = min(q0 + 1, NRED - 1);   // (an odd NRED: the last pair's second column is a dummy)
              r0_ = rd[q0];
              r1_ = rd[q1];
              cr_ = UT[q1 * UTLD + q0];
#pragma unroll
              for (int b = 0; b < D3_RCH; ++b) {
                u0_[b] = UT[q0 * UTLD + min(lane + 32 * b, NRED - 1)];
                u1_[b] = UT[q1 * UTLD + min(lane + 32 * b, NRED - 1)];
              }
            };
            bs_load(NPAIR - 1, rdc0, rdc1, crc, uc0c, uc1c);
#pragma unroll 2
            for (int pp = NPAIR - 1; pp >= 0; --pp) {
              const int j0 = 2 * pp, j1 = j0 + 1, bj = j0 >> 5, l0 = j0 & 31;
              double rdn0, rdn1, crn, uc0n[D3_RCH], uc1n[D3_RCH];
              bs_load(pp > 0 ? pp - 1 : 0, rdn0, rdn1, crn, uc0n, uc1n);
              double rj = rhs[0];
#pragma unroll
              for (int b = 1; b < D3_RCH; ++b) rj = (bj == b) ? rhs[b] : rj;
              const double r0 = __shfl_sync(FULLMASK, rj, l0), r1 = __shfl_sync(FULLMASK, rj, l0 + 1);
              if (j1 < NRED) {
                const double x1 = r1 * rdc1;
                const double x0 = fma(-crc, x1, r0) * rdc0;
                if (lane == 0) {
                  sol[j0] = x0;
                  sol[j1] = x1;
                }
#pragma unroll
                for (int b = 0; b < D3_RCH; ++b)
                  if (lane + 32 * b < j0) rhs[b] = fma(-uc0c[b], x0, fma(-uc1c[b], x1, rhs[b]));
              } else {
                const double x0 = r0 * rdc0;
                if (lane == 0) sol[j0] = x0;
#pragma unroll
                for (int b = 0; b < D3_RCH; ++b)
                  if (lane + 32 * b < j0) rhs[b] = fma(-uc0c[b], x0, rhs[b]);
              }
              rdc0 = rdn0;
              rdc1 = rdn1;
              crc = crn;
#pragma unroll
              for (int b = 0; b < D3_RCH; ++b) {
                uc0c[b] = uc0n[b];
                uc1c[b] = uc1n[b];
              }
            }
          }
          __syncthreads();
          // ---- δy = w − D⁻¹ H_x δx ;  δs = −(F₃ + s δy)/(y + tol) ---------------------------------------------------
          if (t < NY) {
            double a0 = 0.0, a1 = 0.0;
            const double* hr = Hc + t * HCS;
#pragma unroll 4
            for (int c = 0; c + 1 < NRED; c += 2) {
              a0 = fma(hr[c], sol[c], a0);
              a1 = fma(hr[c + 1], sol[c + 1], a1);
            }
            if (NRED & 1) a0 = fma(hr[NRED - 1], sol[NRED - 1], a0);
            const double dy = w[t] - dinv[t] * (a0 + a1);
            const double f3 = s[t] * y[t] - eps;
            w[t] = dy;
            dinv[t] = -(f3 + s[t] * dy) / (y[t] + tol);
          }
          __syncthreads();
          a_s = blk_ftb_linesearch(s, dinv, p.min_stepsize, t);  // :93
          a_y = blk_ftb_linesearch(y, w, p.min_stepsize, t);     // :94
          failed = (a_s != a_s) || (a_y != a_y);                 // :96-100
        }
        if (failed) {
          status = 1;
          break;
        }
        for (int c = t; c < NRED; c += DT) x[PERM[c]] += a_s * sol[c];  // :103
        for (int k = t; k < NY; k += DT) {
          s[k] += a_s * dinv[k];                                        // :104
          y[k] += a_y * w[k];                                           // :105
        }
        __syncthreads();
        kkt = kkt_new;  // :107
        ++inner;        // :108
        ++steps;
      }
      eps *= (status == 0) ? 1.0 - exp(-p.tightening_rate * inner) : 1.0 + exp(-p.loosening_rate * inner);  // :111-113
      ++outer;                                                                                              // :114
    }
    if (!parked && outer == p.max_outer) status = 1;  // :117-119
    for (int i = t; i < NX; i += DT) p.x_out[inst * NX + i] = x[i];
    for (int i = t; i < NY; i += DT) {
      p.y_out[inst * NY + i] = y[i];
      p.s_out[inst * NY + i] = s[i];
    }
    if (t == 0) {
      p.kkt_out[inst] = kkt;
      p.eps_out[inst] = eps;
      p.outer_out[inst] = outer;
      p.status_out[inst] = status;
      p.steps_out[inst] = steps;
      if (parked) {
        p.deferred[atomicAdd(p.counters + 3, 1ULL)] = (int)inst;
      } else {
        atomicAdd(p.counters + 1, (unsigned long long)steps);
        if (status == 0) atomicAdd(p.counters + 2, 1ULL);
      }
    }
  }
}
#elif DENSE_KERNEL == 2
extern "C" __global__ void __launch_bounds__(DT, 2) mcp_solve_kernel(const SolveParams p) {
  extern __shared__ double smem[];
  const int t = threadIdx.x;
  double* x = smem + DENSE_OFF_X;
  double* y = smem + DENSE_OFF_Y;
  double* s = smem + DENSE_OFF_S;
  double* g = smem + DENSE_OFF_G;
  double* w = smem + DENSE_OFF_W;        // H rows, then w, then δy
  double* dinv = smem + DENSE_OFF_DINV;  // D⁻¹, then δs
  double* sol = smem + DENSE_OFF_SOL;
  double* jv = smem + DENSE_OFF_JV;      // computed Jacobian entries (none for the QP configs)
  double* W = smem + DENSE_OFF_WIN;      // NRED rows × WS1 (column NRED = right-hand side)
  double* gbs = smem + DENSE_OFF_STG;    // DKB × DNP : G_y[:, k]·D⁻¹_k
  double* hbs = gbs + DKB * DNP;         // DKB × DNP : H_x[k, :]
  double* part = smem + DENSE_OFF_PART;  // 2 × 128 partial sums
  double* red = smem + DENSE_OFF_RED;    // 16 doubles of reduction scratch
  double* rd = smem + DENSE_OFF_RD;      // reciprocal pivots
  double* Hc = smem + DENSE_OFF_HC;      // H_x cached for the whole solve: NY rows × DENSE_HCS (new column ordering)
  double* xt = smem + DENSE_OFF_XT;      // x in the new ordering
  double* gh0 = smem + DENSE_OFF_G0;     // G(0;θ), H(0;θ)
  constexpr int HCS = DENSE_HCS;
  int* ord = reinterpret_cast<int*>(smem + DENSE_OFF_ORD);  // pivot row of each elimination step
  __shared__ unsigned long long sh_q;
  __shared__ unsigned sh_key[DT / 32];
#if THETA_IN_SMEM
  double* th = smem + DENSE_OFF_TH;
#else
  const double* th = p.theta;
#endif
  const double tol = p.tol;
  constexpr int WS = WS1;
  const unsigned long long n_deferred = p.pass ? p.counters[3] : 0ULL;
  const int r_ = t & 127, hf = t >> 7;   // (row, column half) mapping of the factorisation
  const int ti = t >> 4, tj = t & 15;    // 16×16 tile grid of the Schur update

  for (;;) {
    __syncthreads();
    if (t == 0) sh_q = atomicAdd(p.counters + (p.pass ? 4 : 0), 1ULL);
    __syncthreads();
    unsigned long long inst = sh_q;
    if (p.pass) {
      if (inst >= n_deferred) break;
      inst = (unsigned long long)p.deferred[inst];
    } else if (inst >= (unsigned long long)p.B) {
      break;
    }
#if THETA_IN_SMEM
    for (int i = t; i < NT; i += DT) th[i] = p.theta[inst * NT + i];
#else
    th = p.theta + inst * NT;
#endif
    double eps = 1.0;                                        // :67
    double kkt = __longlong_as_double(0x7ff0000000000000LL);  // :68
    int status = 0, outer = 1, steps = 0;                    // :69-70
    if (p.pass) {
      for (int i = t; i < NX; i += DT) x[i] = p.x_out[inst * NX + i];
      for (int i = t; i < NY; i += DT) {
        y[i] = p.y_out[inst * NY + i];
        s[i] = p.s_out[inst * NY + i];
      }
      eps = p.eps_out[inst];
      kkt = p.kkt_out[inst];
      outer = p.outer_out[inst];
      steps = p.steps_out[inst];
    } else {
      for (int i = t; i < NX; i += DT) x[i] = p.x0 ? p.x0[inst * NX + i] : 0.0;
      for (int i = t; i < NY; i += DT) {
        y[i] = p.y0 ? p.y0[inst * NY + i] : 1.0;
        s[i] = p.s0 ? p.s0[inst * NY + i] : 1.0;
      }
    }
    __syncthreads();
    // ---- once per solve: cache H_x (θ-only) in shared memory and evaluate the constant part of the residual ----
    for (int i = t; i < NY * HCS; i += DT) Hc[i] = 0.0;
    mcp_eval_const_par(t, x, y, th, gh0);
    __syncthreads();
    for (int k = (t >> 5); k < NY; k += DT / 32)
      for (int e = H_PTR[k] + (t & 31); e < H_PTR[k + 1]; e += 32) Hc[k * HCS + H_COL[e]] = H_COEF_AT(e) * opval(H_CODE[e], jv, th);
    __syncthreads();
    bool parked = false;
    while (kkt > tol && eps > tol && outer < p.max_outer) {  // :71
      if (p.pass == 0 && p.step_budget > 0 && steps >= p.step_budget) {
        parked = true;
        break;
      }
      int inner = 1;  // :72
      status = 0;     // :73
      while (kkt > eps && inner < p.max_inner) {  // :75
        // ---- direct part of C: G_x + tol·I, streamed from θ (the only per-step θ traffic) ----------------------
        for (int c = t; c < NRED; c += DT) xt[c] = x[PERM[c]];
        for (int d0 = t; d0 < ND; d0 += 4 * DT) {
          int tp[4], t1[4], rc[4];
          double acc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int d = min(d0 + u * DT, ND - 1);
            tp[u] = D_TP[d];
            t1[u] = D_TP[d + 1] & D_TP_MASK;
            acc[u] = D_BASE_AT(d) + ((tp[u] < 0) ? tol : 0.0);
            rc[u] = D_ROW[d] * WS + D_CPOS[d];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            for (int q = tp[u] & D_TP_MASK; q < t1[u]; ++q) acc[u] += T_COEF_AT(q) * opval(T_I[q].x, jv, th);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (d0 + u * DT < ND) W[rc[u]] = acc[u];
        }
        __syncthreads();
        // ---- F (:79) from the affine structure: H = H(0) + H_x x,  G = G(0) + G_x x + G_y y,  G_y = −H_xᵀ ----------
        if (t < NY) {
          double a0 = 0.0, a1 = 0.0;
          const double* hr = Hc + t * HCS;
#pragma unroll 4
          for (int c = 0; c + 1 < NRED; c += 2) {
            a0 = fma(hr[c], xt[c], a0);
            a1 = fma(hr[c + 1], xt[c + 1], a1);
          }
          if (NRED & 1) a0 = fma(hr[NRED - 1], xt[NRED - 1], a0);
          w[t] = gh0[NX + t] + a0 + a1;
        }
        if (t >= 128 && t < 128 + NRED) {  // G in the NEW row ordering (row i ↔ old row PERM[i]); second half of the CTA
          const int i = t - 128;
          double a0 = gh0[PERM[i]] - tol * xt[i], a1 = 0.0;
          const double* wr = W + i * WS;
#pragma unroll 4
          for (int c = 0; c < NRED; ++c) a0 = fma(wr[c], xt[c], a0);
#pragma unroll 4
          for (int k = 0; k < NY; ++k) a1 = fma(Hc[k * HCS + i], y[k], a1);
          g[i] = a0 - a1;
        }
        __syncthreads();
        double fmax_ = 0.0;
        for (int i = t; i < NX; i += DT) fmax_ = nanmax(fmax_, fabs(g[i]));
        for (int k = t; k < NY; k += DT) {
          const double f2 = w[k] - s[k];
          const double f3 = s[k] * y[k] - eps;
          const double yt = y[k] + tol;
          const double di = 1.0 / (tol + s[k] / yt);
          dinv[k] = di;
          w[k] = di * (-f2 - f3 / yt);
          fmax_ = nanmax(fmax_, nanmax(fabs(f2), fabs(f3)));
        }
        const double kkt_new = blk_nanmax(fmax_, red, t);  // :107 (also orders the writes above)
        // ---- right-hand side: −G − G_y w = −G + H_xᵀ w, into column NRED of W -----------------------------------
        if (t < NRED) {
          double a0 = -g[t];
#pragma unroll 4
          for (int k = 0; k < NY; ++k) a0 = fma(Hc[k * HCS + t], w[k], a0);
          W[t * WS + NRED] = a0;
        }
        // ---- Schur part: C −= G_y D⁻¹ H_x = + H_xᵀ D⁻¹ H_x, register-tiled straight from the cached H_x -------------
        {
          double acc[DTR][DTR];
#pragma unroll
          for (int i = 0; i < DTR; ++i)
#pragma unroll
            for (int j = 0; j < DTR; ++j) acc[i][j] = 0.0;
#pragma unroll 2
          for (int k = 0; k < NY; ++k) {
            const double dk = dinv[k];
            const double* hr = Hc + k * HCS;
            double av[DTR], bv[DTR];
#pragma unroll
            for (int i = 0; i < DTR; ++i) av[i] = hr[ti * DTR + i] * dk;
#pragma unroll
            for (int j = 0; j < DTR; ++j) bv[j] = hr[tj * DTR + j];
#pragma unroll
            for (int i = 0; i < DTR; ++i)
#pragma unroll
              for (int j = 0; j < DTR; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
          }
          __syncthreads();
#pragma unroll
          for (int i = 0; i < DTR; ++i)
#pragma unroll
            for (int j = 0; j < DTR; ++j) {
              const int r = ti * DTR + i, c = tj * DTR + j;
              if (r < NRED && c < NRED) W[r * WS + c] += acc[i][j];
            }
          __syncthreads();
        }
        // ---- LU with partial pivoting, rows stay in place; forward substitution rides in column NRED -----------
        bool failed = false;
        {
          bool done_row = false;  // my row has already been a pivot
          constexpr int NPR = (NRED + 2) / 2;         // double2 pairs per row incl. the rhs column
          for (int j = 0; j < NRED; ++j) {
            const double vj = (r_ < NRED) ? W[r_ * WS + j] : 0.0;
            unsigned key = 0;
            if (hf == 0 && r_ < NRED && !done_row)
              key = ((unsigned)__double2hiint(fabs(vj)) & 0xffffff00u) | (unsigned)(255 - r_);
            key = __reduce_max_sync(FULLMASK, key);
            if ((t & 31) == 0) sh_key[t >> 5] = key;
            __syncthreads();
            unsigned best = sh_key[0];
#pragma unroll
            for (int i = 1; i < DT / 32; ++i) best = max(best, sh_key[i]);
            const int pr = 255 - (int)(best & 0xffu);
            const double piv = W[pr * WS + j];
            if (best == 0 || !(fabs(piv) > 0.0) || !(fabs(piv) <= DBL_MAX_)) {  // :84-88
              failed = true;
              break;
            }
            const double rp = 1.0 / piv;
            if (t == 0) {
              ord[j] = pr;
              rd[j] = rp;
            }
            const double m = (r_ < NRED && !done_row && r_ != pr) ? -(vj * rp) : 0.0;
            if (m != 0.0) {
              const double2* Wp2 = reinterpret_cast<const double2*>(W + pr * WS);
              double2* Wr2 = reinterpret_cast<double2*>(W + r_ * WS);
              // columns < j are already zero in every unpivoted row: start at pair j/2; the two thread halves
              // take alternate pairs so both stay busy as the active part shrinks
#pragma unroll 4
              for (int c = (j >> 1) + hf; c < NPR; c += 2) {
                const double2 u = Wp2[c];
                double2 a = Wr2[c];
                a.x = fma(m, u.x, a.x);
                a.y = fma(m, u.y, a.y);
                Wr2[c] = a;
              }
            }
            if (r_ == pr) done_row = true;
            __syncthreads();
          }
        }
        failed = __syncthreads_or(failed);
        double a_s = 1.0, a_y = 1.0;
        if (!failed) {
          // ---- back substitution, column sweep over the pivot order ------------------------------------------------
          for (int j = NRED - 1; j >= 0; --j) {
            const int pr = ord[j];
            const double xj = W[pr * WS + NRED] * rd[j];
            if (t == 0) sol[j] = xj;
            if (t < j) {
              const int ri = ord[t];
              W[ri * WS + NRED] = fma(-W[ri * WS + j], xj, W[ri * WS + NRED]);
            }
            __syncthreads();
          }
          // ---- δy = w − D⁻¹ H_x δx ;  δs = −(F₃ + s δy)/(y + tol) ---------------------------------------------------
          if (t < NY) {
            double a0 = 0.0, a1 = 0.0;
            const double* hr = Hc + t * HCS;
#pragma unroll 4
            for (int c = 0; c + 1 < NRED; c += 2) {
              a0 = fma(hr[c], sol[c], a0);
              a1 = fma(hr[c + 1], sol[c + 1], a1);
            }
            if (NRED & 1) a0 = fma(hr[NRED - 1], sol[NRED - 1], a0);
            const double dy = w[t] - dinv[t] * (a0 + a1);
            const double f3 = s[t] * y[t] - eps;
            w[t] = dy;
            dinv[t] = -(f3 + s[t] * dy) / (y[t] + tol);
          }
          __syncthreads();
          a_s = blk_ftb_linesearch(s, dinv, p.min_stepsize, t);  // :93
          a_y = blk_ftb_linesearch(y, w, p.min_stepsize, t);     // :94
          failed = (a_s != a_s) || (a_y != a_y);                 // :96-100
        }
        if (failed) {
          status = 1;
          break;
        }
        for (int c = t; c < NRED; c += DT) x[PERM[c]] += a_s * sol[c];  // :103
        for (int k = t; k < NY; k += DT) {
          s[k] += a_s * dinv[k];                                        // :104
          y[k] += a_y * w[k];                                           // :105
        }
        __syncthreads();
        kkt = kkt_new;  // :107
        ++inner;        // :108
        ++steps;
      }
      eps *= (status == 0) ? 1.0 - exp(-p.tightening_rate * inner) : 1.0 + exp(-p.loosening_rate * inner);  // :111-113
      ++outer;                                                                                              // :114
    }
    if (!parked && outer == p.max_outer) status = 1;  // :117-119
    for (int i = t; i < NX; i += DT) p.x_out[inst * NX + i] = x[i];
    for (int i = t; i < NY; i += DT) {
      p.y_out[inst * NY + i] = y[i];
      p.s_out[inst * NY + i] = s[i];
    }
    if (t == 0) {
      p.kkt_out[inst] = kkt;
      p.eps_out[inst] = eps;
      p.outer_out[inst] = outer;
      p.status_out[inst] = status;
      p.steps_out[inst] = steps;
      if (parked) {
        p.deferred[atomicAdd(p.counters + 3, 1ULL)] = (int)inst;
      } else {
        atomicAdd(p.counters + 1, (unsigned long long)steps);
        if (status == 0) atomicAdd(p.counters + 2, 1ULL);
      }
    }
  }
}
#else
extern "C" __global__ void __launch_bounds__(DT, 2) mcp_solve_kernel(const SolveParams p) {
  extern __shared__ double smem[];
  const int t = threadIdx.x;
  double* x = smem + DENSE_OFF_X;
  double* y = smem + DENSE_OFF_Y;
  double* s = smem + DENSE_OFF_S;
  double* g = smem + DENSE_OFF_G;
  double* w = smem + DENSE_OFF_W;        // H rows, then w, then δy
  double* dinv = smem + DENSE_OFF_DINV;  // D⁻¹, then δs
  double* sol = smem + DENSE_OFF_SOL;
  double* jv = smem + DENSE_OFF_JV;      // computed Jacobian entries (none for the QP configs)
  double* W = smem + DENSE_OFF_WIN;      // NRED rows × WS1 (column NRED = right-hand side)
  double* gbs = smem + DENSE_OFF_STG;    // DKB × DNP : G_y[:, k]·D⁻¹_k
  double* hbs = gbs + DKB * DNP;         // DKB × DNP : H_x[k, :]
  double* part = smem + DENSE_OFF_PART;  // 2 × 128 partial sums
  double* red = smem + DENSE_OFF_RED;    // 16 doubles of reduction scratch
  double* rd = smem + DENSE_OFF_RD;      // reciprocal pivots
  int* ord = reinterpret_cast<int*>(smem + DENSE_OFF_ORD);  // pivot row of each elimination step
  __shared__ unsigned long long sh_q;
  __shared__ unsigned sh_key[DT / 32];
#if THETA_IN_SMEM
  double* th = smem + DENSE_OFF_TH;
#else
  const double* th = p.theta;
#endif
  const double tol = p.tol;
  constexpr int WS = WS1;
  const unsigned long long n_deferred = p.pass ? p.counters[3] : 0ULL;
  const int r_ = t & 127, hf = t >> 7;   // (row, column half) mapping of the factorisation
  const int ti = t >> 4, tj = t & 15;    // 16×16 tile grid of the Schur update

  for (;;) {
    __syncthreads();
    if (t == 0) sh_q = atomicAdd(p.counters + (p.pass ? 4 : 0), 1ULL);
    __syncthreads();
    unsigned long long inst = sh_q;
    if (p.pass) {
      if (inst >= n_deferred) break;
      inst = (unsigned long long)p.deferred[inst];
    } else if (inst >= (unsigned long long)p.B) {
      break;
    }
#if THETA_IN_SMEM
    for (int i = t; i < NT; i += DT) th[i] = p.theta[inst * NT + i];
#else
    th = p.theta + inst * NT;
#endif
    double eps = 1.0;                                        // :67
    double kkt = __longlong_as_double(0x7ff0000000000000LL);  // :68
    int status = 0, outer = 1, steps = 0;                    // :69-70
    if (p.pass) {
      for (int i = t; i < NX; i += DT) x[i] = p.x_out[inst * NX + i];
      for (int i = t; i < NY; i += DT) {
        y[i] = p.y_out[inst * NY + i];
        s[i] = p.s_out[inst * NY + i];
      }
      eps = p.eps_out[inst];
      kkt = p.kkt_out[inst];
      outer = p.outer_out[inst];
      steps = p.steps_out[inst];
    } else {
      for (int i = t; i < NX; i += DT) x[i] = p.x0 ? p.x0[inst * NX + i] : 0.0;
      for (int i = t; i < NY; i += DT) {
        y[i] = p.y0 ? p.y0[inst * NY + i] : 1.0;
        s[i] = p.s0 ? p.s0[inst * NY + i] : 1.0;
      }
    }
    __syncthreads();
    bool parked = false;
    while (kkt > tol && eps > tol && outer < p.max_outer) {  // :71
      if (p.pass == 0 && p.step_budget > 0 && steps >= p.step_budget) {
        parked = true;
        break;
      }
      int inner = 1;  // :72
      status = 0;     // :73
      while (kkt > eps && inner < p.max_inner) {  // :75
        // ---- F (:79): thread i evaluates generated part i; H lands in w -----------------------------------
        mcp_eval_newton_par(t, x, y, th, g, w, jv);
        __syncthreads();
        double fmax_ = 0.0;
        for (int i = t; i < NX; i += DT) fmax_ = nanmax(fmax_, fabs(g[i]));
        for (int k = t; k < NY; k += DT) {
          const double f2 = w[k] - s[k];
          const double f3 = s[k] * y[k] - eps;
          const double yt = y[k] + tol;
          const double di = 1.0 / (tol + s[k] / yt);
          dinv[k] = di;
          w[k] = di * (-f2 - f3 / yt);
          fmax_ = nanmax(fmax_, nanmax(fabs(f2), fabs(f3)));
        }
        const double kkt_new = blk_nanmax(fmax_, red, t);  // :107 (also orders the writes above)
        // ---- right-hand side of the condensed system into column NRED of W ----------------------------------
        {
          double acc = 0.0;
          if (r_ < NRED) {
            const int e0 = R_PTR[r_], e1 = R_PTR[r_ + 1], mid = e0 + (e1 - e0 + 1) / 2;
            for (int e = hf ? mid : e0; e < (hf ? e1 : mid); ++e) acc += R_COEF_AT(e) * opval(R_CODE[e], jv, th) * w[R_K[e]];
          }
          part[hf * 128 + r_] = acc;
          __syncthreads();
          if (t < NRED) W[t * WS + NRED] = -g[R_GROW[t]] - part[t] - part[128 + t];
        }
        // ---- direct part of C: G_x + tol·I (one dest per matrix entry, row-sorted) ---------------------------
        for (int d = t; d < ND; d += DT) {
          const int tp = D_TP[d];
          const int t1 = D_TP[d + 1] & D_TP_MASK;
          double acc = D_BASE_AT(d) + ((tp < 0) ? tol : 0.0);
          for (int q = tp & D_TP_MASK; q < t1; ++q) acc += T_COEF_AT(q) * opval(T_I[q].x, jv, th);
          W[D_ROW[d] * WS + D_CPOS[d]] = acc;
        }
        // ---- Schur part: C −= Σ_k (G_y[:,k] D⁻¹_k) ⊗ H_x[k,:], register-tiled over blocks of DKB constraints ----
        {
          double acc[DTR][DTR];
#pragma unroll
          for (int i = 0; i < DTR; ++i)
#pragma unroll
            for (int j = 0; j < DTR; ++j) acc[i][j] = 0.0;
          for (int k0 = 0; k0 < NY; k0 += DKB) {
            __syncthreads();
            for (int i = t; i < 2 * DKB * DNP; i += DT) gbs[i] = 0.0;
            __syncthreads();
            {
              const int kk = t >> 5, k = k0 + kk, ln = t & 31;  // warp kk stages constraint k0+kk
              if (k < NY) {
                const double dk = dinv[k];
                for (int e = H_PTR[k] + ln; e < H_PTR[k + 1]; e += 32) hbs[kk * DNP + H_COL[e]] = H_COEF_AT(e) * opval(H_CODE[e], jv, th);
                for (int e = GK_PTR[k] + ln; e < GK_PTR[k + 1]; e += 32) gbs[kk * DNP + GK_ROW[e]] = GK_COEF_AT(e) * opval(GK_CODE[e], jv, th) * dk;
              }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < DKB; ++kk) {
              double av[DTR], bv[DTR];
#pragma unroll
              for (int i = 0; i < DTR; ++i) av[i] = gbs[kk * DNP + ti * DTR + i];
#pragma unroll
              for (int j = 0; j < DTR; ++j) bv[j] = hbs[kk * DNP + tj * DTR + j];
#pragma unroll
              for (int i = 0; i < DTR; ++i)
#pragma unroll
                for (int j = 0; j < DTR; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
            }
          }
          __syncthreads();
#pragma unroll
          for (int i = 0; i < DTR; ++i)
#pragma unroll
            for (int j = 0; j < DTR; ++j) {
              const int r = ti * DTR + i, c = tj * DTR + j;
              if (r < NRED && c < NRED) W[r * WS + c] -= acc[i][j];
            }
          __syncthreads();
        }
        // ---- LU with partial pivoting, rows stay in place; forward substitution rides in column NRED -----------
        bool failed = false;
        {
          bool done_row = false;  // my row has already been a pivot
          constexpr int NPR = (NRED + 2) / 2;         // double2 pairs per row incl. the rhs column
          for (int j = 0; j < NRED; ++j) {
            const double vj = (r_ < NRED) ? W[r_ * WS + j] : 0.0;
            unsigned key = 0;
            if (hf == 0 && r_ < NRED && !done_row)
              key = ((unsigned)__double2hiint(fabs(vj)) & 0xffffff00u) | (unsigned)(255 - r_);
            key = __reduce_max_sync(FULLMASK, key);
            if ((t & 31) == 0) sh_key[t >> 5] = key;
            __syncthreads();
            unsigned best = sh_key[0];
#pragma unroll
            for (int i = 1; i < DT / 32; ++i) best = max(best, sh_key[i]);
            const int pr = 255 - (int)(best & 0xffu);
            const double piv = W[pr * WS + j];
            if (best == 0 || !(fabs(piv) > 0.0) || !(fabs(piv) <= DBL_MAX_)) {  // :84-88
              failed = true;
              break;
            }
            const double rp = 1.0 / piv;
            if (t == 0) {
              ord[j] = pr;
              rd[j] = rp;
            }
            const double m = (r_ < NRED && !done_row && r_ != pr) ? -(vj * rp) : 0.0;
            if (m != 0.0) {
              const double2* Wp2 = reinterpret_cast<const double2*>(W + pr * WS);
              double2* Wr2 = reinterpret_cast<double2*>(W + r_ * WS);
              // columns < j are already zero in every unpivoted row: start at pair j/2; the two thread halves
              // take alternate pairs so both stay busy as the active part shrinks
#pragma unroll 4
              for (int c = (j >> 1) + hf; c < NPR; c += 2) {
                const double2 u = Wp2[c];
                double2 a = Wr2[c];
                a.x = fma(m, u.x, a.x);
                a.y = fma(m, u.y, a.y);
                Wr2[c] = a;
              }
            }
            if (r_ == pr) done_row = true;
            __syncthreads();
          }
        }
        failed = __syncthreads_or(failed);
        double a_s = 1.0, a_y = 1.0;
        if (!failed) {
          // ---- back substitution, column sweep over the pivot order ------------------------------------------------
          for (int j = NRED - 1; j >= 0; --j) {
            const int pr = ord[j];
            const double xj = W[pr * WS + NRED] * rd[j];
            if (t == 0) sol[j] = xj;
            if (t < j) {
              const int ri = ord[t];
              W[ri * WS + NRED] = fma(-W[ri * WS + j], xj, W[ri * WS + NRED]);
            }
            __syncthreads();
          }
          // ---- δy = w − D⁻¹ H_x δx ;  δs = −(F₃ + s δy)/(y + tol) ---------------------------------------------------
          {
            double acc = 0.0;
            if (r_ < NY) {
              const int e0 = H_PTR[r_], e1 = H_PTR[r_ + 1], mid = e0 + (e1 - e0 + 1) / 2;
              for (int e = hf ? mid : e0; e < (hf ? e1 : mid); ++e) acc += H_COEF_AT(e) * opval(H_CODE[e], jv, th) * sol[H_COL[e]];
            }
            part[hf * 128 + r_] = acc;
            __syncthreads();
            for (int k = t; k < NY; k += DT) {
              const double dy = w[k] - dinv[k] * (part[k] + part[128 + k]);
              const double f3 = s[k] * y[k] - eps;
              w[k] = dy;
              dinv[k] = -(f3 + s[k] * dy) / (y[k] + tol);
            }
            __syncthreads();
          }
          a_s = blk_ftb_linesearch(s, dinv, p.min_stepsize, t);  // :93
          a_y = blk_ftb_linesearch(y, w, p.min_stepsize, t);     // :94
          failed = (a_s != a_s) || (a_y != a_y);                 // :96-100
        }
        if (failed) {
          status = 1;
          break;
        }
        for (int c = t; c < NRED; c += DT) x[PERM[c]] += a_s * sol[c];  // :103
        for (int k = t; k < NY; k += DT) {
          s[k] += a_s * dinv[k];                                        // :104
          y[k] += a_y * w[k];                                           // :105
        }
        __syncthreads();
        kkt = kkt_new;  // :107
        ++inner;        // :108
        ++steps;
      }
      eps *= (status == 0) ? 1.0 - exp(-p.tightening_rate * inner) : 1.0 + exp(-p.loosening_rate * inner);  // :111-113
      ++outer;                                                                                              // :114
    }
    if (!parked && outer == p.max_outer) status = 1;  // :117-119
    for (int i = t; i < NX; i += DT) p.x_out[inst * NX + i] = x[i];
    for (int i = t; i < NY; i += DT) {
      p.y_out[inst * NY + i] = y[i];
      p.s_out[inst * NY + i] = s[i];
    }
    if (t == 0) {
      p.kkt_out[inst] = kkt;
      p.eps_out[inst] = eps;
      p.outer_out[inst] = outer;
      p.status_out[inst] = status;
      p.steps_out[inst] = steps;
      if (parked) {
        p.deferred[atomicAdd(p.counters + 3, 1ULL)] = (int)inst;
      } else {
        atomicAdd(p.counters + 1, (unsigned long long)steps);
        if (status == 0) atomicAdd(p.counters + 2, 1ULL);
      }
    }
  }
}
#endif  // DENSE_KERNEL == 2
#elif TINY_KERNEL
// ------------------------------------------------------------------------------------------------
// Thread-per-instance solve kernel for problems of a few unknowns (NRED ≤ 6, NY ≤ 8: the README QP).  The iterate,
// the condensed matrix and every intermediate live in registers; residual, Jacobian entries, assembly, right-hand
// side and H_x products are straight-line code generated from the plan's tables (tiny_* functions); the
// factorisation is a fully unrolled dense LU with partial pivoting.  The loop is the same literal restatement of
// src/solver.jl:63-121 as in the other kernels; the two-pass budget parks long runs exactly like there (one slow
// thread would otherwise hold its warp).
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(128) mcp_solve_kernel(const SolveParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_work = p.pass ? (long long)p.counters[3] : p.B;
  const bool valid = idx < n_work;
  int my_steps = 0, my_solved = 0;
  if (valid) {
    const long long inst = p.pass ? (long long)p.deferred[idx] : idx;
    double th[NT > 0 ? NT : 1], x[NX], y[NY], s[NY];
#pragma unroll
    for (int i = 0; i < NT; ++i) th[i] = p.theta[inst * NT + i];
    double eps = 1.0;                                        // :67
    double kkt = __longlong_as_double(0x7ff0000000000000LL);  // :68
    int status = 0, outer = 1, steps = 0;                    // :69-70
    if (p.pass) {
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = p.x_out[inst * NX + i];
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        y[i] = p.y_out[inst * NY + i];
        s[i] = p.s_out[inst * NY + i];
      }
      eps = p.eps_out[inst];
      kkt = p.kkt_out[inst];
      outer = p.outer_out[inst];
      steps = p.steps_out[inst];
    } else {
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = p.x0 ? p.x0[inst * NX + i] : 0.0;
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        y[i] = p.y0 ? p.y0[inst * NY + i] : 1.0;
        s[i] = p.s0 ? p.s0[inst * NY + i] : 1.0;
      }
    }
    const double tol = p.tol;
    bool parked = false;
    while (kkt > tol && eps > tol && outer < p.max_outer) {  // :71
      if (p.pass == 0 && p.step_budget > 0 && steps >= p.step_budget) {
        parked = true;
        break;
      }
      int inner = 1;  // :72
      status = 0;     // :73
      while (kkt > eps && inner < p.max_inner) {  // :75
        double g[NX], h[NY], jv[NJV > 0 ? NJV : 1], dinv[NY], w[NY], C[NRED][NRED + 1];
        tiny_eval(x, y, th, g, h, jv);  // :79-80
        double fmax_ = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) fmax_ = nanmax(fmax_, fabs(g[i]));
#pragma unroll
        for (int k = 0; k < NY; ++k) {
          const double f2 = h[k] - s[k];
          const double f3 = s[k] * y[k] - eps;
          const double yt = y[k] + tol;
          const double di = 1.0 / (tol + s[k] / yt);
          dinv[k] = di;
          w[k] = di * (-f2 - f3 / yt);
          fmax_ = nanmax(fmax_, nanmax(fabs(f2), fabs(f3)));
        }
        const double kkt_new = fmax_;  // :107
        tiny_assemble(jv, th, dinv, tol, C);
        tiny_rhs(g, jv, th, w, C);
        // ---- dense LU with partial pivoting, rows swapped in registers; forward substitution in column NRED ----------
        bool failed = false;
        double rpv[NRED];
#pragma unroll
        for (int j = 0; j < NRED; ++j) {
          int pr = j;
          unsigned best = (unsigned)__double2hiint(fabs(C[j][j])) & 0xffffff00u;
#pragma unroll
          for (int i = j + 1; i < NRED; ++i) {
            const unsigned k = (unsigned)__double2hiint(fabs(C[i][j])) & 0xffffff00u;
            if (k > best) {
              best = k;
              pr = i;
            }
          }
#pragma unroll
          for (int i = j + 1; i < NRED; ++i)
            if (pr == i) {
#pragma unroll
              for (int c = j; c <= NRED; ++c) {
                const double t = C[i][c];
                C[i][c] = C[j][c];
                C[j][c] = t;
              }
            }
          const double piv = C[j][j];
          if (!(fabs(piv) > 0.0) || !(fabs(piv) <= DBL_MAX_)) failed = true;  // :84-88
          const double rp = 1.0 / piv;
          rpv[j] = rp;
#pragma unroll
          for (int i = j + 1; i < NRED; ++i) {
            const double m = -(C[i][j] * rp);
#pragma unroll
            for (int c = j + 1; c <= NRED; ++c) C[i][c] = fma(m, C[j][c], C[i][c]);
          }
        }
        double a_s = 1.0, a_y = 1.0;
        double sol[NRED], ds[NY], dy[NY];
        if (!failed) {
#pragma unroll
          for (int j = NRED - 1; j >= 0; --j) {
            const double xj = C[j][NRED] * rpv[j];
            sol[j] = xj;
#pragma unroll
            for (int i = 0; i < j; ++i) C[i][NRED] = fma(-C[i][j], xj, C[i][NRED]);
          }
          // δy = w − D⁻¹ H_x δx ;  δs = −(F₃ + s δy)/(y + tol)
          double hx[NY];
          tiny_hx(jv, th, sol, hx);
#pragma unroll
          for (int k = 0; k < NY; ++k) {
            dy[k] = w[k] - dinv[k] * hx[k];
            const double f3 = s[k] * y[k] - eps;
            ds[k] = -(f3 + s[k] * dy[k]) / (y[k] + tol);
          }
          // `fraction_to_the_boundary_linesearch` (src/solver.jl:127-138), once for (s, δs), once for (y, δy)
          const double c995 = 1.0 - 0.995;
#pragma unroll 1
          for (int which = 0; which < 2; ++which) {
            double alpha = 1.0;
            bool ok = false;
            for (int it = 0; it < 1200; ++it) {
              bool viol = false;
#pragma unroll
              for (int k = 0; k < NY; ++k) {
                const double v = which ? y[k] : s[k], d = which ? dy[k] : ds[k];
                viol = viol || (v + alpha * d < c995 * v);  // :129
              }
              if (!viol) {
                ok = true;
                break;
              }
              if (alpha < p.min_stepsize) break;  // :130
              alpha *= 0.5;                       // :134
            }
            if (!ok) alpha = __longlong_as_double(0x7ff8000000000000LL);
            if (which) a_y = alpha; else a_s = alpha;
          }
          failed = (a_s != a_s) || (a_y != a_y);  // :96-100
        }
        if (failed) {
          status = 1;
          break;
        }
        tiny_update_x(x, sol, a_s);  // :103
#pragma unroll
        for (int k = 0; k < NY; ++k) {
          s[k] += a_s * ds[k];  // :104
          y[k] += a_y * dy[k];  // :105
        }
        kkt = kkt_new;  // :107
        ++inner;        // :108
        ++steps;
      }
      eps *= (status == 0) ? 1.0 - exp(-p.tightening_rate * inner) : 1.0 + exp(-p.loosening_rate * inner);  // :111-113
      ++outer;                                                                                              // :114
    }
    if (!parked && outer == p.max_outer) status = 1;  // :117-119
#pragma unroll
    for (int i = 0; i < NX; ++i) p.x_out[inst * NX + i] = x[i];
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      p.y_out[inst * NY + i] = y[i];
      p.s_out[inst * NY + i] = s[i];
    }
    p.kkt_out[inst] = kkt;
    p.eps_out[inst] = eps;
    p.outer_out[inst] = outer;
    p.status_out[inst] = status;
    p.steps_out[inst] = steps;
    if (parked) {
      p.deferred[atomicAdd(p.counters + 3, 1ULL)] = (int)inst;
    } else {
      my_steps = steps;
      my_solved = (status == 0);
    }
  }
  // warp-aggregated statistics (every lane of the warp arrives here)
  __syncwarp();
  const int ws = __reduce_add_sync(FULLMASK, my_steps), wv = __reduce_add_sync(FULLMASK, my_solved);
  if ((threadIdx.x & 31) == 0) {
    if (ws) atomicAdd(p.counters + 1, (unsigned long long)ws);
    if (wv) atomicAdd(p.counters + 2, (unsigned long long)wv);
  }
}
#else   // !DENSE_KERNEL
// ------------------------------------------------------------------------------------------------
// The solve kernel: persistent CTAs; every sub-warp (SUB lanes) pulls instances from a global queue and
// runs the reference's loop (src/solver.jl:63-121) as a small state machine, one Newton step per trip
// of the main loop.  The sub-warps of a warp re-align once per Newton step (the full-warp vote below),
// so the heavy phases of their two instances execute as the same instructions.
// ------------------------------------------------------------------------------------------------
#ifndef SOLVE_LB_THREADS
#define SOLVE_LB_THREADS (SUB * SOLVE_INST * NWIDE)
#endif
extern "C" __global__ void __launch_bounds__(SOLVE_LB_THREADS, 1) mcp_solve_kernel(const SolveParams p) {
  extern __shared__ double smem[];
#if NWIDE > 1
  // cooperative instances: NWIDE consecutive warps per instance, the first is the leader (runs everything below),
  // the others only help with the window sweep of the factorisation (wide_helper)
  const int sl = threadIdx.x & 31;
  const int slot = (threadIdx.x >> 5) / NWIDE;
  const int wrole = (threadIdx.x >> 5) % NWIDE;
#else
  const int sl = threadIdx.x % SUB;
  const int slot = threadIdx.x / SUB;  // instance slot of this sub-warp within the CTA
#endif
  const unsigned smask = sub_mask(threadIdx.x & 31);
  load_shared_tables<1>(smem);
  const int* rowptr = reinterpret_cast<const int*>(smem);
  const unsigned short* cpos = reinterpret_cast<const unsigned short*>(rowptr + NRED + 1);
  double* V = smem + SHARED_TABLE_DOUBLES + (size_t)slot * SOLVE_SMEM_DOUBLES;  // my shared-memory block
#if LARGE_STATE
  double* S = p.state + ((size_t)blockIdx.x * SOLVE_INST + slot) * SOLVE_STATE_DOUBLES;  // vectors in global memory
#if WIN_GLOBAL
  double* W = S + SOLVE_OFF_WIN;   // not even one window fits shared memory: it follows the vectors in the global block
#else
  double* W = V;                   // only the window is shared
#endif
#else
  double* S = V;
  double* W = V + SOLVE_OFF_WIN;
#endif
  double* x = S + SOLVE_OFF_X;
  double* y = S + SOLVE_OFF_Y;
#if S_GLOBAL && !LARGE_STATE
  // `s` is touched only by lane-strided sweeps: it lives in the instance's global block (coalesced, L1/L2-resident), which
  // is what lets 20 instead of 16 instances share an SM's shared memory for the lane-change game
  double* s = p.state + ((size_t)blockIdx.x * SOLVE_INST + slot) * SOLVE_STATE_DOUBLES;
#else
  double* s = S + SOLVE_OFF_S;
#endif
  // G rows: in `sol` at their permuted positions (SOLVE_G_ON_SOL), else aliasing the window region when they fit
  double* hh = S + SOLVE_OFF_H;       // H rows (alias w: H[k] is consumed where w[k] is produced)
  double* jv = S + SOLVE_OFF_JV;
  double* dinv = S + SOLVE_OFF_DINV;  // D⁻¹, later δs
  double* w = S + SOLVE_OFF_W;        // w, later δy
#if LARGE_STATE && SOL_IN_SMEM
  double* sol = V + SOLVE_SOL_SMEM_OFF;   // δx in the permuted ordering: behind the window, in shared memory
#else
  double* sol = S + SOLVE_OFF_SOL;    // δx in the permuted ordering
#endif
  double* g = SOLVE_G_ON_SOL ? sol : (SOLVE_G_IN_WIN ? W : S + SOLVE_OFF_G);
#if THETA_IN_SMEM
  double* th = S + SOLVE_OFF_TH;
#else
  const double* th = p.theta;
#endif
  double* Cval = p.scratch + ((size_t)blockIdx.x * SOLVE_INST + slot) * SOLVE_SCRATCH;
  double* UT = Cval + CVAL_DOUBLES;
  const double tol = p.tol;
#if NWIDE > 1
  static_assert(THETA_IN_SMEM, "cooperative instances need θ at a fixed address");
  if (wrole != 0) {
    // Helper warps: wait for the leader's command, do their share, report back (see WIDE_* above).  They sit at
    // barrier A whenever the leader is in a phase that is not shared.
    constexpr int RPL = (WR + SUB - 1) / SUB;
    const int bar_id = 1 + slot;
    for (;;) {
      wide_bar(bar_id, NWIDE * 32);   // A: command published
      const int c = WIDE_CMD(W, WS1)[0];
      if (c == WIDE_EXIT) return;
      if (c == WIDE_ASSEMBLE) {
        assemble_matrix<NWIDE>(Cval, W, jv, th, dinv, *WIDE_ARG(W, WS1), sl, smask, wrole, bar_id);   // ends with a barrier
      } else if (c == WIDE_EVAL) {
        mcp_eval_newton_par(wrole * 32 + sl, x, y, th, g, hh, jv);
        wide_bar(bar_id, NWIDE * 32);   // B
      } else {
        double m[RPL];
        const double* mbox = WIDE_MBOX(W, WS1);
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const int r = sl + SUB * k;
          m[k] = (r < WR) ? mbox[r] : 0.0;
        }
        window_sweep<1, WS1, NWIDE>(W, W + c * WS1, m, sl, wrole);
        wide_bar(bar_id, NWIDE * 32);   // B: sweep complete
      }
    }
  }
#endif

  // Scheduling (semantics-neutral): instances that never converge run ~30x longer than the rest (up to
  // (max_outer-1)(max_inner-1) Newton steps) and would leave most of the GPU idle behind a long tail.
  // Pass 0 therefore runs every instance only up to `step_budget` steps (checked at outer-iteration
  // boundaries, where the whole solver state is (x, y, s, ϵ, kkt_error, outer_iters)), parks the rest in
  // the output arrays and a deferred list; pass 1 (a second launch) resumes them, all long, together.
  // r2: pass 1 runs them in SLICES of `step_budget` steps too and re-queues an unfinished instance at the end of the same
  // list, so the work stays evenly spread over the slots until the very end: with whole instances (≈ 870 steps each,
  // ≈ 5 per slot for the bench batch) the last round ran at half occupancy on average — pass 1 ran at 84 % of pass 0's
  // Newton-step rate.  List slots ≥ n_deferred are filled during the pass (−1 = not yet): a slot whose sub-warp has
  // claimed such an entry polls it; everyone leaves when all n_deferred instances have finished.
  const unsigned long long n_deferred = p.pass ? p.counters[3] : 0ULL;
  int steps_entry = 0;

  // solver state of my sub-warp's instance (replicated in each of its lanes)
  bool have = false, done = false, head = true, brk = false;
  unsigned long long inst = 0;
  double eps = 1.0, kkt = 0.0;
  int status = 0, outer = 1, inner = 1, steps = 0;

  for (;;) {
    // ---- control: advance my instance to its next Newton step (or fetch another, or run dry) -------------
    bool step = false;
    while (!done && !step) {
      if (!have) {
        unsigned long long q = 0;
        if (sl == 0) q = atomicAdd(p.counters + (p.pass ? 4 : 0), 1ULL);
        q = __shfl_sync(smask, q, 0, SUB);
        if (p.pass) {
          long long got = -1;
          if (q < n_deferred) {
            got = p.deferred[q];
          } else {
            // an entry that is (or may still be) produced during this pass
            if (sl == 0) {
              const long long t0 = clock64();
              for (;;) {
                got = *reinterpret_cast<volatile int*>(p.deferred + q);
                if (got >= 0) break;
                if (*reinterpret_cast<volatile unsigned long long*>(p.counters + 6) >= n_deferred) break;   // all finished
                if (clock64() - t0 > 20000000000LL) break;   // (≈ 10 s: never hang the device on a logic error)
                __nanosleep(1000);
              }
            }
            got = __shfl_sync(smask, got, 0, SUB);
          }
          if (got < 0) {
            done = true;
            break;
          }
          inst = (unsigned long long)got;
        } else {
          if (q >= (unsigned long long)p.B) {
            done = true;
            break;
          }
          inst = q;
        }
        // load θ and the initial point (src/solver.jl:39-41,64-66) — or the parked state in pass 1
#if THETA_IN_SMEM
        for (int i = sl; i < NT; i += SUB) th[i] = p.theta[inst * NT + i];
#else
        th = p.theta + inst * NT;
#endif
        eps = 1.0;                                         // :67
        kkt = __longlong_as_double(0x7ff0000000000000LL);  // Inf, :68
        status = 0;                                        // :69
        outer = 1;                                         // :70
        steps = 0;
        if (p.pass) {   // (L2 loads: the state may have been parked by another SM during this very launch)
          for (int i = sl; i < NX; i += SUB) x[i] = __ldcg(p.x_out + inst * NX + i);
          for (int i = sl; i < NY; i += SUB) {
            y[i] = __ldcg(p.y_out + inst * NY + i);
            s[i] = __ldcg(p.s_out + inst * NY + i);
          }
          eps = __ldcg(p.eps_out + inst);
          kkt = __ldcg(p.kkt_out + inst);
          outer = __ldcg(p.outer_out + inst);
          steps = __ldcg(p.steps_out + inst);
        } else {
          for (int i = sl; i < NX; i += SUB) x[i] = p.x0 ? p.x0[inst * NX + i] : 0.0;
          for (int i = sl; i < NY; i += SUB) {
            y[i] = p.y0 ? p.y0[inst * NY + i] : 1.0;
            s[i] = p.s0 ? p.s0[inst * NY + i] : 1.0;
          }
        }
        __syncwarp(smask);
        steps_entry = steps;
        have = true;
        head = true;
        brk = false;
      }
      if (head) {  // top of the outer (ϵ-homotopy) loop, :71
#ifdef EXP_FIXED_STEPS   // timing experiments (MCPB200_DEFS): exactly EXP_FIXED_STEPS identical Newton steps per instance
        const bool go = steps < EXP_FIXED_STEPS;
#else
        const bool go = kkt > tol && eps > tol && outer < p.max_outer;
#endif
        const bool park = go && p.step_budget > 0 && steps - steps_entry >= p.step_budget;
        if (!go || park) {
          if (!park && outer == p.max_outer) status = 1;  // :117-119
          for (int i = sl; i < NX; i += SUB) p.x_out[inst * NX + i] = x[i];
          for (int i = sl; i < NY; i += SUB) {
            p.y_out[inst * NY + i] = y[i];
            p.s_out[inst * NY + i] = s[i];
          }
          if (sl == 0) {
            p.kkt_out[inst] = kkt;
            p.eps_out[inst] = eps;
            p.outer_out[inst] = outer;
            p.status_out[inst] = status;
            p.steps_out[inst] = steps;
          }
          if (park && p.pass) __threadfence();   // every lane's part of the parked state, before the entry is published
          __syncwarp(smask);
          if (sl == 0) {
            if (park && p.pass == 0) {
              p.deferred[atomicAdd(p.counters + 3, 1ULL)] = (int)inst;
            } else if (park) {
              // re-queue behind everything already in the list; the parked state must be visible device-wide before
              // the entry is
              __threadfence();
              const unsigned long long slot_q = n_deferred + atomicAdd(p.counters + 5, 1ULL);
              atomicExch(p.deferred + slot_q, (int)inst);
            } else {
              atomicAdd(p.counters + 1, (unsigned long long)steps);
              if (status == 0) atomicAdd(p.counters + 2, 1ULL);
              if (p.pass) {
                __threadfence();
                atomicAdd(p.counters + 6, 1ULL);
              }
            }
          }
          __syncwarp(smask);
          have = false;
          continue;
        }
        inner = 1;   // :72
        status = 0;  // :73
        head = false;
      }
#ifdef EXP_FIXED_STEPS
      if (steps < EXP_FIXED_STEPS) {
#else
      if (!brk && kkt > eps && inner < p.max_inner) {  // :75
#endif
        step = true;
      } else {  // the inner loop is over: ϵ update, :111-114
        eps *= (status == 0) ? 1.0 - exp(-p.tightening_rate * inner) : 1.0 + exp(-p.loosening_rate * inner);
        ++outer;
        head = true;
        brk = false;
      }
    }
    if (__all_sync(FULLMASK, done)) break;  // also re-aligns the warp's sub-warps once per Newton step
    if (!step) continue;

    // ---- one Newton step (src/solver.jl:76-108) -------------------------------------------------------------
    // F and the Jacobian entries at the current iterate (:79-80); lane i evaluates output group i
#if NWIDE > 1
    if (sl == 0) WIDE_CMD(W, WS1)[0] = WIDE_EVAL;
    wide_bar(1 + slot, NWIDE * 32);
    mcp_eval_newton_par(sl, x, y, th, g, hh, jv);
    wide_bar(1 + slot, NWIDE * 32);
#else
#ifndef EXP_NO_EVAL
    mcp_eval_newton_par(sl, x, y, th, g, hh, jv);
#endif
    __syncwarp(smask);
#endif
    double fmax_ = 0.0;
#if SOLVE_G_ON_SOL && FULL_Y
    for (int i = sl; i < NRED; i += SUB)
      if (R_GROW[i] < NX) fmax_ = nanmax(fmax_, fabs(g[i]));   // (the y rows of `sol` hold stale data here)
#else
    for (int i = sl; i < NX; i += SUB) fmax_ = nanmax(fmax_, fabs(g[i]));   // (a permutation of the G rows when G lives in sol)
#endif
    for (int k = sl; k < NY; k += SUB) {
      const double f2 = hh[k] - s[k];             // H − s        (src/mcp.jl:78)
      const double f3 = s[k] * y[k] - eps;        // s∘y − ϵ      (src/mcp.jl:79)
      const double yt = y[k] + tol;               // (3,3) block diag(y) + tol·I  (:81)
#if FULL_Y
      // mode B (∇_y H ≠ 0): only δs is eliminated.  The per-constraint array holds s/(y+tol), which the assembly adds
      // (with tol) to the diagonal of the H rows; w is the right-hand side of row nx+k
      dinv[k] = s[k] / yt;
      w[k] = -f2 - f3 / yt;
#else
      const double di = 1.0 / (tol + s[k] / yt);  // D⁻¹, D = (2,2) block tol·I + S (Y+tol)⁻¹
      dinv[k] = di;
      w[k] = di * (-f2 - f3 / yt);
#endif
      fmax_ = nanmax(fmax_, nanmax(fabs(f2), fabs(f3)));
    }
    const double kkt_new = sub_nanmax(fmax_, smask);  // ‖F‖∞ of the pre-step residual (:107)
    __syncwarp(smask);
    // (∇F + tol·I) δz = −F, condensed to NRED unknowns (:81-83)
    for (int i = sl; i < NRED; i += SUB) {
#if FULL_Y
      const int old = R_GROW[i];
      sol[i] = (old < NX) ? -g[SOLVE_G_ON_SOL ? i : old] : w[old - NX];
#else
      double r = -g[SOLVE_G_ON_SOL ? i : R_GROW[i]];   // (in place when G lives in sol: row i reads and writes sol[i])
#pragma unroll 4
      for (int e = R_PTR[i]; e < R_PTR[i + 1]; ++e) r -= R_COEF_AT(e) * opval(R_CODE[e], jv, th) * w[R_K[e]];
      sol[i] = r;
#endif
    }
    __syncwarp(smask);  // G (aliased onto the window) is dead from here on: the window becomes scratch
#if NWIDE > 1
    if (sl == 0) {
      *WIDE_ARG(W, WS1) = tol;
      WIDE_CMD(W, WS1)[0] = WIDE_ASSEMBLE;
    }
    wide_bar(1 + slot, NWIDE * 32);
    assemble_matrix<NWIDE>(Cval, W, jv, th, dinv, tol, sl, smask, 0, 1 + slot);
#else
#ifndef EXP_NO_ASM
    assemble_matrix(Cval, W, jv, th, dinv, tol, sl, smask);
#endif
    __syncwarp(smask);
#endif
#ifdef EXP_NO_LU
    bool failed = false;
#else
    bool failed = band_solve<1, WS1, NWIDE>(W, Cval, UT, sol, rowptr, cpos, jv, th, dinv, S + SOLVE_OFF_STAGE, sl, smask, 1 + slot) != 0;  // :84-88
#endif
    double a_s = 1.0, a_y = 1.0;
    if (!failed) {
      // δy = w − D⁻¹ H_x δx ;  δs = −(F₃ + s δy)/(y + tol)
      for (int k = sl; k < NY; k += SUB) {
#if FULL_Y
        const double dy = sol[IPERM[NX + k]];
#else
        double hx = 0.0;
#pragma unroll 4
        for (int e = H_PTR[k]; e < H_PTR[k + 1]; ++e) hx += H_COEF_AT(e) * opval(H_CODE[e], jv, th) * sol[H_COL[e]];
        const double dy = w[k] - dinv[k] * hx;
#endif
        const double f3 = s[k] * y[k] - eps;
        w[k] = dy;
        dinv[k] = -(f3 + s[k] * dy) / (y[k] + tol);
      }
      __syncwarp(smask);
      a_s = ftb_linesearch(s, dinv, p.min_stepsize, sl, smask);  // :93
      a_y = ftb_linesearch(y, w, p.min_stepsize, sl, smask);     // :94
      failed = (a_s != a_s) || (a_y != a_y);                     // :96-100
    }
#ifdef EXP_FIXED_STEPS
    failed = false;   // the iterate stays where it is: every step does the same work
    a_s = 0.0;
    a_y = 0.0;
#endif
    if (failed) {
      status = 1;
      brk = true;
    } else {
#ifdef EXP_FIXED_STEPS
      if (a_s != 0.0)
#endif
#if FULL_Y
      for (int c = sl; c < NRED; c += SUB)
        if (PERM[c] < NX) x[PERM[c]] += a_s * sol[c];
#else
      for (int c = sl; c < NRED; c += SUB) x[PERM[c]] += a_s * sol[c];  // :103 (x uses α_s)
#endif
#ifdef EXP_FIXED_STEPS
      if (a_s != 0.0)
#endif
      for (int k = sl; k < NY; k += SUB) {
        s[k] += a_s * dinv[k];                                          // :104
        y[k] += a_y * w[k];                                             // :105
      }
      kkt = kkt_new;                                                    // :107
      ++inner;                                                          // :108
      ++steps;
    }
    __syncwarp(smask);
  }
#if NWIDE > 1
  {   // release this instance slot's helper warps
    if (sl == 0) WIDE_CMD(W, WS1)[0] = WIDE_EXIT;
    wide_bar(1 + slot, NWIDE * 32);
  }
#endif
}

#endif  // DENSE_KERNEL

// ------------------------------------------------------------------------------------------------
// Sensitivity kernel: ∂z/∂θ = (−∇F_z)⁻¹ ∇F_θ at the returned point, no tol·I (src/AutoDiff.jl:18-40),
// through the same condensation with D = S Y⁻¹, NRHS_SENS right-hand sides per factorisation pass.
// ------------------------------------------------------------------------------------------------
#if HAS_JT
#if FULL_Y   // mode B (∇_y H ≠ 0): the generic forward loop below is the only sensitivity path
#undef USE_DIRECT_JVP
#define USE_DIRECT_JVP 0
#endif
#ifndef USE_DIRECT_JVP
#define USE_DIRECT_JVP 1
#endif
// Per-instance memory layouts of the sensitivity kernels (offsets in doubles, generated by plan.cpp): the forward
// kernel carries NRHS_SENS right-hand sides per pass; the adjoint kernel needs one, so many more instances fit an SM.
struct SensLayout {
  static constexpr bool ADJ = false;
  static constexpr int NRHS = NRHS_SENS, INST = SENS_INST, WS = WSS;
  static constexpr long long SMEM = SENS_SMEM_DOUBLES, STATE = SENS_STATE_DOUBLES;
  static constexpr int OFF_X = SENS_OFF_X, OFF_Y = SENS_OFF_Y, OFF_S = SENS_OFF_S, OFF_JV = SENS_OFF_JV, OFF_JTV = SENS_OFF_JTV,
                       OFF_DINV = SENS_OFF_DINV, OFF_WQ = SENS_OFF_WQ, OFF_SOL = SENS_OFF_SOL, OFF_STAGE = SENS_OFF_STAGE,
                       OFF_WIN = SENS_OFF_WIN;
#if THETA_IN_SMEM
  static constexpr int OFF_TH = SENS_OFF_TH;
#endif
};
#if HAS_ADJOINT
struct AdjLayout {
  static constexpr bool ADJ = true;
  static constexpr int NRHS = 1, INST = ADJ_INST, WS = WS1;
  static constexpr long long SMEM = ADJ_SMEM_DOUBLES, STATE = ADJ_STATE_DOUBLES;
  static constexpr int OFF_X = ADJ_OFF_X, OFF_Y = ADJ_OFF_Y, OFF_S = ADJ_OFF_S, OFF_JV = ADJ_OFF_JV, OFF_JTV = ADJ_OFF_JTV,
                       OFF_DINV = ADJ_OFF_DINV, OFF_WQ = ADJ_OFF_WQ, OFF_SOL = ADJ_OFF_SOL, OFF_STAGE = ADJ_OFF_STAGE,
                       OFF_WIN = ADJ_OFF_WIN;
#if THETA_IN_SMEM
  static constexpr int OFF_TH = ADJ_OFF_TH;
#endif
};
#endif

#define SENS_NAN __longlong_as_double(0x7ff8000000000000LL)
template <class L>
__device__ __forceinline__ void sens_body(const SensParams& p) {
  extern __shared__ double smem[];
  const int sl = threadIdx.x % SUB;
  const int slot = threadIdx.x / SUB;
  const unsigned smask = sub_mask(threadIdx.x & 31);
  // Pullback only (z̄ → θ̄) runs in mcp_adj_kernel: adjoint mode — ONE solve with Cᵀ per instance instead of one solve
  // of C per column of ∇F_θ (nθ = 160 for the masked game at N = 10).  The Jacobian and the pushforward keep the
  // forward solves (mcp_sens_kernel).
  constexpr bool adjoint = L::ADJ;
  constexpr int NRHS_L = L::NRHS;
  load_shared_tables<NRHS_L>(smem, adjoint);
  const int* rowptr = reinterpret_cast<const int*>(smem);
  const unsigned short* cpos = reinterpret_cast<const unsigned short*>(rowptr + NRED + 1);
  double* V = smem + SHARED_TABLE_DOUBLES + (size_t)slot * L::SMEM;
#if LARGE_STATE
  double* S = p.state + ((size_t)blockIdx.x * L::INST + slot) * L::STATE;
#if WIN_GLOBAL
  double* W = S + L::OFF_WIN;
#else
  double* W = V;
#endif
#else
  double* S = V;
  double* W = V + L::OFF_WIN;
#endif
  double* x = S + L::OFF_X;
  double* y = S + L::OFF_Y;
  double* s = S + L::OFF_S;
  double* jv = S + L::OFF_JV;
  double* jtv = S + L::OFF_JTV;
  double* dinv = S + L::OFF_DINV;
  double* wq = S + L::OFF_WQ;    // [NRHS_L][NY]
  double* sol = S + L::OFF_SOL;  // [NRHS_L][NRED]
#if THETA_IN_SMEM
  double* th = S + L::OFF_TH;
#endif
  double* Cval = p.scratch + ((size_t)blockIdx.x * L::INST + slot) * SENS_SCRATCH;
  double* UT = Cval + CVAL_DOUBLES;
  constexpr int NZ = NX + 2 * NY;

  for (;;) {
    unsigned long long inst = 0;
    if (sl == 0) inst = atomicAdd(p.counters, 1ULL);
    inst = __shfl_sync(smask, inst, 0, SUB);
    if (inst >= (unsigned long long)p.B) break;
#if THETA_IN_SMEM
    for (int i = sl; i < NT; i += SUB) th[i] = p.theta[inst * NT + i];
#else
    const double* th = p.theta + inst * NT;
#endif
    for (int i = sl; i < NX; i += SUB) x[i] = p.x[inst * NX + i];
    for (int i = sl; i < NY; i += SUB) {
      y[i] = p.y[inst * NY + i];
      s[i] = p.s[inst * NY + i];
    }
    __syncwarp(smask);
    mcp_eval_sens_par(sl, x, y, th, jv, jtv);
    __syncwarp(smask);
#if FULL_Y
    // mode B: J_B [Z_x; Z_y] = −[∇_θG; ∇_θH] with J_B = [G_x, G_y; H_x, H_y + diag(s/y)] (the solve kernel's system at
    // tol = 0: row 3 of src/AutoDiff.jl:27-39 gives Z_s = −(s/y) Z_y); the per-constraint array holds s/y
    for (int k = sl; k < NY; k += SUB) dinv[k] = s[k] / y[k];
#else
    for (int k = sl; k < NY; k += SUB) dinv[k] = y[k] / s[k];  // D⁻¹ with D = s/y (tol = 0)
#endif
    if (p.z_p)
      for (int i = sl; i < NZ * p.P; i += SUB) p.z_p[inst * NZ * p.P + i] = 0.0;
    __syncwarp(smask);
    int bad = 0;
#if HAS_ADJOINT
    if constexpr (adjoint) {
      // Jᵀλ = z̄ condensed like the forward system:  Cᵀ λ₁ = x̄ − H_xᵀ v,  v = D⁻¹ȳ − s̄,
      // λ₂ = v − D⁻¹ G_yᵀ λ₁,  θ̄ = −∇F_θᵀ [λ₁; λ₂]   (the complementarity rows do not depend on θ)
      const double* zb = p.zbar + inst * NZ;
      for (int k = sl; k < NY; k += SUB) wq[k] = dinv[k] * zb[NX + k] - zb[NX + NY + k];
      for (int i = sl; i < NRED; i += SUB) sol[i] = zb[PERM[i]];
      __syncwarp(smask);
      for (int k = sl; k < NY; k += SUB) {
        const double vk = wq[k];
        for (int e = H_PTR[k]; e < H_PTR[k + 1]; ++e) atomicAdd(&sol[H_COL[e]], -H_COEF_AT(e) * opval(H_CODE[e], jv, th) * vk);
      }
      __syncwarp(smask);
      assemble_matrix(Cval, W, jv, th, dinv, 0.0, sl, smask);
      __syncwarp(smask);
      bad = band_solve<1, L::WS>(W, Cval, UT, sol, rowptr, cpos, jv, th, dinv, S + L::OFF_STAGE, sl, smask, 0, DT_SRC);
      if (!bad) {
        for (int i = sl; i < NRED; i += SUB) {
          const double li = sol[i];
          for (int e = R_PTR[i]; e < R_PTR[i + 1]; ++e)
            atomicAdd(&wq[R_K[e]], -dinv[R_K[e]] * R_COEF_AT(e) * opval(R_CODE[e], jv, th) * li);
        }
        __syncwarp(smask);
        for (int q = sl; q < NT; q += SUB) {
          double acc = 0.0;
          for (int e = Q_PTR[q]; e < Q_PTR[q + 1]; ++e) {
            const int row = Q_ROW[e];
            const double lam = (row < NX) ? sol[IPERM[row]] : wq[row - NX];
            acc -= Q_COEF_AT(e) * opval(Q_CODE[e], jtv, th) * lam;
          }
          p.thetabar[inst * NT + q] = acc;
        }
      } else {
        // singular / non-finite pivot: the reference's QR would hand back Inf/NaN — never leave the caller's
        // (uninitialised) output untouched
        for (int q = sl; q < NT; q += SUB) p.thetabar[inst * NT + q] = SENS_NAN;
      }
      if (sl == 0 && p.status_out) p.status_out[inst] = bad;
      __syncwarp(smask);
      continue;
    }
#endif
#if USE_DIRECT_JVP
    if (p.z_p && !p.dzdtheta && !p.zbar) {
      // Pushforward only (θ_p → z_p, the Dual overload): the tangents are the right-hand sides — r = −∇F_θ θ_p, one
      // solve of C per tangent (NRHS_L at a time) instead of one per column of ∇F_θ followed by a contraction
      for (int p0 = 0; p0 < p.P; p0 += NRHS_L) {
        const int np = min(NRHS_L, p.P - p0);
        for (int i = sl; i < NRHS_L * NY; i += SUB) wq[i] = 0.0;
        for (int i = sl; i < NRHS_L * NRED; i += SUB) sol[i] = 0.0;
        assemble_matrix(Cval, W, jv, th, dinv, 0.0, sl, smask);
        __syncwarp(smask);
        for (int q = 0; q < NT; ++q) {
          for (int e = Q_PTR[q] + sl; e < Q_PTR[q + 1]; e += SUB) {   // rows within one column are distinct: no atomics
            const double f = -Q_COEF_AT(e) * opval(Q_CODE[e], jtv, th);
            const int row = Q_ROW[e];
            for (int rp = 0; rp < np; ++rp) {
              const double v = f * p.theta_p[(inst * p.P + p0 + rp) * NT + q];
              if (row < NX) sol[rp * NRED + IPERM[row]] += v;
              else wq[rp * NY + (row - NX)] += dinv[row - NX] * v;
            }
          }
          __syncwarp(smask);   // the next column may touch the same rows from other lanes
        }
        for (int i = sl; i < NRED; i += SUB)
          for (int rp = 0; rp < np; ++rp) {
            double r = sol[rp * NRED + i];
            for (int e = R_PTR[i]; e < R_PTR[i + 1]; ++e) r -= R_COEF_AT(e) * opval(R_CODE[e], jv, th) * wq[rp * NY + R_K[e]];
            sol[rp * NRED + i] = r;
          }
        __syncwarp(smask);
        if (band_solve<NRHS_L, L::WS>(W, Cval, UT, sol, rowptr, cpos, jv, th, dinv, S + L::OFF_STAGE, sl, smask)) {
          bad = 1;
          break;
        }
        for (int rp = 0; rp < np; ++rp) {
          const double* so = sol + rp * NRED;
          double* zp = p.z_p + (inst * p.P + p0 + rp) * NZ;
          for (int c = sl; c < NRED; c += SUB) zp[PERM[c]] = so[c];
          for (int k = sl; k < NY; k += SUB) {
            double hx = 0.0;
            for (int e = H_PTR[k]; e < H_PTR[k + 1]; ++e) hx += H_COEF_AT(e) * opval(H_CODE[e], jv, th) * so[H_COL[e]];
            const double zy = wq[rp * NY + k] - dinv[k] * hx;
            zp[NX + k] = zy;
            zp[NX + NY + k] = -s[k] * zy / y[k];
          }
        }
        __syncwarp(smask);
      }
      if (bad)
        for (int i = sl; i < NZ * p.P; i += SUB) p.z_p[inst * NZ * p.P + i] = SENS_NAN;
      if (sl == 0 && p.status_out) p.status_out[inst] = bad;
      __syncwarp(smask);
      continue;
    }
#endif
    for (int q0 = 0; q0 < NT; q0 += NRHS_L) {
      const int nq = min(NRHS_L, NT - q0);
      for (int i = sl; i < NRHS_L * NY; i += SUB) wq[i] = 0.0;
      for (int i = sl; i < NRHS_L * NRED; i += SUB) sol[i] = 0.0;
      assemble_matrix(Cval, W, jv, th, dinv, 0.0, sl, smask);
      __syncwarp(smask);
      // right-hand sides r = −∇F_θ[:, q]:  G rows go to the reduced rhs, H rows to w = D⁻¹ r₂
      for (int rq = 0; rq < nq; ++rq) {
        const int q = q0 + rq;
        for (int e = Q_PTR[q] + sl; e < Q_PTR[q + 1]; e += SUB) {
          const double v = -Q_COEF_AT(e) * opval(Q_CODE[e], jtv, th);
          const int row = Q_ROW[e];
#if FULL_Y
          sol[rq * NRED + IPERM[row]] = v;   // (the G and the H rows are both unknowns' rows; row < NX + NY)
#else
          if (row < NX) sol[rq * NRED + IPERM[row]] = v;
          else wq[rq * NY + (row - NX)] = dinv[row - NX] * v;
#endif
        }
      }
      __syncwarp(smask);
#if !FULL_Y
      for (int i = sl; i < NRED; i += SUB) {
        for (int rq = 0; rq < nq; ++rq) {
          double r = sol[rq * NRED + i];
          for (int e = R_PTR[i]; e < R_PTR[i + 1]; ++e)
            r -= R_COEF_AT(e) * opval(R_CODE[e], jv, th) * wq[rq * NY + R_K[e]];
          sol[rq * NRED + i] = r;
        }
      }
      __syncwarp(smask);
#endif
      if (band_solve<NRHS_L, L::WS>(W, Cval, UT, sol, rowptr, cpos, jv, th, dinv, S + L::OFF_STAGE, sl, smask)) {
        bad = 1;
        break;
      }
      // recover the y and s rows:  Z_y = w − D⁻¹ H_x Z_x ,  Z_s = −s Z_y / y
      for (int rq = 0; rq < nq; ++rq) {
        const int q = q0 + rq;
        const double* so = sol + rq * NRED;
        double tb = 0.0;
        for (int c = sl; c < NRED; c += SUB) {
          const double zx = so[c];
          const int row = PERM[c];
          if (p.dzdtheta) p.dzdtheta[(inst * NT + q) * NZ + row] = zx;
          if (p.zbar) tb += p.zbar[inst * NZ + row] * zx;
          if (p.z_p)
            for (int pp = 0; pp < p.P; ++pp)
              p.z_p[(inst * p.P + pp) * NZ + row] += zx * p.theta_p[(inst * p.P + pp) * NT + q];
#if FULL_Y
          if (row >= NX) {   // a y unknown: its s row follows from the complementarity row
            const int k = row - NX;
            const double zs = -s[k] * zx / y[k];
            if (p.dzdtheta) p.dzdtheta[(inst * NT + q) * NZ + NX + NY + k] = zs;
            if (p.zbar) tb += p.zbar[inst * NZ + NX + NY + k] * zs;
            if (p.z_p)
              for (int pp = 0; pp < p.P; ++pp)
                p.z_p[(inst * p.P + pp) * NZ + NX + NY + k] += zs * p.theta_p[(inst * p.P + pp) * NT + q];
          }
#endif
        }
#if !FULL_Y
        for (int k = sl; k < NY; k += SUB) {
          double hx = 0.0;
          for (int e = H_PTR[k]; e < H_PTR[k + 1]; ++e) hx += H_COEF_AT(e) * opval(H_CODE[e], jv, th) * so[H_COL[e]];
          const double zy = wq[rq * NY + k] - dinv[k] * hx;
          const double zs = -s[k] * zy / y[k];
          if (p.dzdtheta) {
            p.dzdtheta[(inst * NT + q) * NZ + NX + k] = zy;
            p.dzdtheta[(inst * NT + q) * NZ + NX + NY + k] = zs;
          }
          if (p.zbar) tb += p.zbar[inst * NZ + NX + k] * zy + p.zbar[inst * NZ + NX + NY + k] * zs;
          if (p.z_p)
            for (int pp = 0; pp < p.P; ++pp) {
              const double tp = p.theta_p[(inst * p.P + pp) * NT + q];
              p.z_p[(inst * p.P + pp) * NZ + NX + k] += zy * tp;
              p.z_p[(inst * p.P + pp) * NZ + NX + NY + k] += zs * tp;
            }
        }
#endif
        if (p.thetabar) {
          tb = sub_sum(tb, smask);
          if (sl == 0) p.thetabar[inst * NT + q] = tb;
        }
      }
      __syncwarp(smask);
    }
    if (bad) {   // a failed factorisation leaves NaN in every requested output of this instance
      if (p.dzdtheta)
        for (long long i = sl; i < (long long)NZ * NT; i += SUB) p.dzdtheta[inst * NT * NZ + i] = SENS_NAN;
      if (p.thetabar)
        for (int q = sl; q < NT; q += SUB) p.thetabar[inst * NT + q] = SENS_NAN;
      if (p.z_p)
        for (int i = sl; i < NZ * p.P; i += SUB) p.z_p[inst * NZ * p.P + i] = SENS_NAN;
    }
    if (sl == 0 && p.status_out) p.status_out[inst] = bad;
    __syncwarp(smask);
  }
}

extern "C" __global__ void __launch_bounds__(SUB * SENS_INST, 1) mcp_sens_kernel(const SensParams p) { sens_body<SensLayout>(p); }
#if HAS_ADJOINT
extern "C" __global__ void __launch_bounds__(SUB * ADJ_INST, 1) mcp_adj_kernel(const SensParams p) { sens_body<AdjLayout>(p); }
#endif
#endif  // HAS_JT
