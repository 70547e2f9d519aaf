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
  unsigned long long* counters;
  int P;
};

__device__ __forceinline__ double opval(int code, const double* __restrict__ jv, const double* __restrict__ th) {
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

// NaN-propagating max, like Julia's norm(F, Inf) (src/solver.jl:107)
__device__ __forceinline__ double nanmax(double a, double b) { return (a != a) ? a : ((b != b) ? b : fmax(a, b)); }

__device__ __forceinline__ double sub_nanmax(double v, unsigned smask) {
#pragma unroll
  for (int o = SUB / 2; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(smask, v, o, SUB));
  return v;
}

// ------------------------------------------------------------------------------------------------
// CTA-shared tables (first SHARED_TABLE_DOUBLES doubles of dynamic shared memory): ROWPTR_S[NRED+1]
// (first dest of each condensed row) and CPOS_S[ND] (circular window position of each dest).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_shared_tables(double* smem_base) {
  int* rowptr = reinterpret_cast<int*>(smem_base);
  unsigned short* cpos = reinterpret_cast<unsigned short*>(rowptr + NRED + 1);
  for (int i = threadIdx.x; i <= NRED; i += blockDim.x) rowptr[i] = D_ROWPTR[i];
  for (int i = threadIdx.x; i < ND; i += blockDim.x) cpos[i] = (unsigned short)D_CPOS[i];
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Assembly of the condensed matrix C = G_x + tol·I − G_y D⁻¹ H_x: one value per structural non-zero
// ("dest", sorted by row then column) into the compact L2-resident array Cval[ND].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void assemble_matrix(double* __restrict__ Cval, double* __restrict__ tmp,
                                                const double* __restrict__ jv, const double* __restrict__ th,
                                                const double* __restrict__ dinv, double tol, int lane) {
  // Constant contributions are folded into D_BASE on the host; only the z/θ/D-dependent terms remain.
#if ASM_TWO_PHASE
  // phase A, term-parallel (all table loads independent and coalesced): tmp[t] = coef·val(a)·[D⁻¹_k·val(b)]
  for (int t = lane; t < NTERMS; t += 32) {
    const int4 ti = T_I[t];  // {a, b, k, -}
    double v = T_COEF[t] * opval(ti.x, jv, th);
    if (ti.z >= 0) v *= dinv[ti.z] * opval(ti.y, jv, th);
    tmp[t] = v;
  }
  __syncwarp();
  // phase B, dest-parallel: sum the (contiguous) terms of each dest
  for (int d = lane; d < ND; d += 32) {
    const int tp = D_TP[d];
    const int t1 = D_TP[d + 1] & 0x7fffffff;
    double acc = D_BASE[d] + ((tp < 0) ? tol : 0.0);  // sign bit of D_TP marks a diagonal dest
    for (int t = tp & 0x7fffffff; t < t1; ++t) acc += tmp[t];
    Cval[d] = acc;
  }
  __syncwarp();
#else
  for (int d = lane; d < ND; d += 32) {
    const int tp = D_TP[d];
    const int t1 = D_TP[d + 1] & 0x7fffffff;
    double acc = D_BASE[d] + ((tp < 0) ? tol : 0.0);
    for (int t = tp & 0x7fffffff; t < t1; ++t) {
      const int4 ti = T_I[t];
      double v = T_COEF[t] * opval(ti.x, jv, th);
      if (ti.z >= 0) v *= dinv[ti.z] * opval(ti.y, jv, th);
      acc += v;
    }
    Cval[d] = acc;
  }
#endif
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
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
//   UT   : global scratch, NRED × WC: U stored TRANSPOSED, UT[c*WC + (c-i)] = U[i][c], with the
//          reciprocal pivot at offset 0, so that the back substitution reads one contiguous row per
//          column and needs no warp reduction.
//   sol  : shared, NRHS × NRED: right-hand sides on entry, solution on exit (permuted ordering).
// Lanes own window ROWS during the elimination (all rows update in parallel, zero multipliers are
// predicated off) and consecutive rows during the back-substitution sweep.
// Returns 0, or 1 if a pivot is zero / non-finite (the reference's `:failed` retcode branch,
// src/solver.jl:84-88).
// ------------------------------------------------------------------------------------------------
template <int NRHS, int WS>
__device__ int band_solve(double* __restrict__ W, const double* __restrict__ Cval, double* __restrict__ UT,
                          double* __restrict__ sol, const int* __restrict__ rowptr,
                          const unsigned short* __restrict__ cpos, int lane) {
  constexpr int CPW = (WC + 31) / 32;  // matrix positions per lane
  constexpr int RPL = (WR + 31) / 32;  // row slots per lane
  constexpr int WPL = (WS + 31) / 32;

  // ---- initial window: rows 0 … WR-1 -----------------------------------------------------------
  for (int i = lane; i < WR * WS; i += 32) W[i] = 0.0;
  __syncwarp();
  for (int r = 0; r < WR; ++r) {
    for (int e = rowptr[r] + lane; e < rowptr[r + 1]; e += 32) W[r * WS + cpos[e]] = Cval[e];
    if (lane < NRHS) W[r * WS + WC + lane] = sol[lane * NRED + r];
  }
  __syncwarp();

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
      const int e = e0 + lane + 32 * k;
      pre[k] = (e < e1) ? Cval[e] : 0.0;
    }

    // ---- pivot search over column j: max |a| on a 12-bit-truncated mantissa, slot in the low byte ----
    unsigned best = 0;
    double m[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      const int r = lane + 32 * k;
      m[k] = (r < WR) ? W[r * WS + cj] : 0.0;
      if (r < WR) {
        const unsigned key = ((unsigned)__double2hiint(fabs(m[k])) & 0xffffff00u) | (unsigned)(255 - r);
        best = max(best, key);
      }
    }
    best = __reduce_max_sync(FULLMASK, best);
    const int p = 255 - (int)(best & 0xffu);
    double piv = m[0];
#pragma unroll
    for (int k = 1; k < RPL; ++k)
      if ((p >> 5) == k) piv = m[k];
    piv = __shfl_sync(FULLMASK, piv, p & 31);
    if (!(fabs(piv) > 0.0) || !(fabs(piv) < 1.0e300 * 1.0e300)) return 1;  // zero, NaN or Inf pivot
    const double rp = 1.0 / piv;
    double* Wp = W + p * WS;
    // Invariant: every entry of a window row outside its structural extent is exactly zero, so the
    // update can sweep ALL positions with static code; only the pivot's own position must read as zero.
    if (lane == 0) Wp[cj] = 0.0;
    __syncwarp();

    // ---- retire the pivot row: U row j goes out transposed, its RHS into sol ------------------------
    {
      const int tmax = min(WC - 1, NRED - 1 - j);
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int q = lane + 32 * k;
        if (q < WC) {
          int d = q - cj;
          if (d < 0) d += WC;
          if (d <= tmax) UT[(j + d) * WC + d] = (d == 0) ? rp : Wp[q];
        }
      }
      if (lane < NRHS) sol[lane * NRED + j] = Wp[WC + lane];
    }

    // ---- eliminate column j: every row (lanes own rows) minus multiplier × pivot row -----------------
    // Rows with a zero multiplier (and the pivot row itself, whose multiplier is forced to zero) are
    // rewritten unchanged: no per-lane branches, 128-bit conflict-free accesses (WS ≡ 2 mod 4).
    {
      const double2* Wp2 = reinterpret_cast<const double2*>(Wp);
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int r = lane + 32 * k;
        m[k] = (r < WR && r != p) ? -(m[k] * rp) : 0.0;
      }
#pragma unroll
      for (int b0 = 0; b0 < NP; b0 += PB) {
        double2 u[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i)
          if (b0 + i < NP) u[i] = Wp2[b0 + i];
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const int r = lane + 32 * k;
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
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int r = lane + 32 * k;
        if (r < WR && r != p) W[r * WS + cj] = 0.0;  // the eliminated entry (exactly zero by construction)
      }
    }
    __syncwarp();

    // ---- the entering row takes the retired slot ---------------------------------------------------
    {
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        const int q = lane + 32 * k;
        if (q < WS) Wp[q] = 0.0;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int e = e0 + lane + 32 * k;
        if (e < e1) Wp[cpos[e]] = pre[k];
      }
      if (lane < NRHS && ienter < NRED) Wp[WC + lane] = sol[lane * NRED + ienter];
    }
    __syncwarp();
    cj = (cj + 1 == WC) ? 0 : cj + 1;
  }

  // ---- back substitution: column sweep, x_j = rhs_j / u_jj then rhs_i −= U[i][j] x_j for i < j.
  // The columns of U (rows of UT) stream back through a RING_D-deep cp.async ring that re-uses the
  // window's shared memory, so the L2/HBM latency of the scratch is off the critical path.
  {
    double* ring = W;
    auto issue = [&](int col) {
      if (col >= 0) {
        const double* src = UT + (size_t)col * WC;
        double* dst = ring + ((NRED - 1 - col) % RING_D) * WC;
#pragma unroll
        for (int k = 0; k < CPW; ++k) {
          const int t = lane + 32 * k;
          if (t < WC) cp_async8(dst + t, src + t);
        }
      }
      cp_async_commit();  // (possibly empty) group: keeps the group count uniform
    };
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < RING_D - 1; ++i) issue(NRED - 1 - i);
#pragma unroll 1
    for (int j = NRED - 1; j >= 0; --j) {
      issue(j - (RING_D - 1));
      cp_async_wait<RING_D - 1>();
      __syncwarp();
      const double* Uj = ring + ((NRED - 1 - j) % RING_D) * WC;
      const double rd = Uj[0];
      double ut[CPW];
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int t = lane + 1 + 32 * k;
        ut[k] = (t < WC) ? Uj[t] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < NRHS; ++q) {
        double xj = 0.0;
        if (lane == 0) {
          xj = sol[q * NRED + j] * rd;
          sol[q * NRED + j] = xj;
        }
        xj = __shfl_sync(FULLMASK, xj, 0);
#pragma unroll
        for (int k = 0; k < CPW; ++k) {
          const int t = lane + 1 + 32 * k;
          if (t < WC && t <= j) sol[q * NRED + j - t] = fma(-ut[k], xj, sol[q * NRED + j - t]);
        }
      }
      __syncwarp();
    }
  }
  return 0;
}

// `fraction_to_the_boundary_linesearch` — src/solver.jl:127-138, literally (τ = 0.995, decay = 0.5).
__device__ __forceinline__ double ftb_linesearch(const double* __restrict__ v, const double* __restrict__ d,
                                                 double min_step, int lane) {
  const double c = 1.0 - 0.995;
  double alpha = 1.0;
  for (int it = 0; it < 1200; ++it) {
    bool viol = false;
    for (int k = lane; k < NY; k += 32) viol = viol || (v[k] + alpha * d[k] < c * v[k]);  // :129
    if (!__any_sync(FULLMASK, viol)) return alpha;
    if (alpha < min_step) break;  // :130 — tested before halving
    alpha *= 0.5;                 // :134
  }
  return __longlong_as_double(0x7ff8000000000000LL);  // NaN (:131)
}

// ------------------------------------------------------------------------------------------------
// The solve kernel: persistent warps pull instances from a global queue.
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(32 * SOLVE_WARPS, 1) mcp_solve_kernel(const SolveParams p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  load_shared_tables(smem);
  const int* rowptr = reinterpret_cast<const int*>(smem);
  const unsigned short* cpos = reinterpret_cast<const unsigned short*>(rowptr + NRED + 1);
  double* S = smem + SHARED_TABLE_DOUBLES + (size_t)warp * SOLVE_SMEM_DOUBLES;
  double* x = S + SOLVE_OFF_X;
  double* y = S + SOLVE_OFF_Y;
  double* s = S + SOLVE_OFF_S;
  double* g = S + SOLVE_OFF_G;        // G rows (aliases the window when it fits)
  double* hh = S + SOLVE_OFF_H;       // H rows (alias w: H[k] is consumed where w[k] is produced)
  double* jv = S + SOLVE_OFF_JV;
  double* dinv = S + SOLVE_OFF_DINV;  // D⁻¹, later δs
  double* w = S + SOLVE_OFF_W;        // w, later δy
  double* sol = S + SOLVE_OFF_SOL;    // δx in the permuted ordering
  double* W = S + SOLVE_OFF_WIN;
#if THETA_IN_SMEM
  double* th = S + SOLVE_OFF_TH;
#endif
  double* Cval = p.scratch + ((size_t)blockIdx.x * SOLVE_WARPS + warp) * SOLVE_SCRATCH;
  double* UT = Cval + CVAL_DOUBLES;
  const double tol = p.tol;

  // Scheduling (semantics-neutral): instances that never converge run ~30x longer than the rest (up to
  // (max_outer-1)(max_inner-1) Newton steps) and would leave most of the GPU idle behind a long tail.
  // Pass 0 therefore runs every instance only up to `step_budget` steps (checked at outer-iteration
  // boundaries, where the whole solver state is (x, y, s, ϵ, kkt_error, outer_iters)), parks the rest in
  // the output arrays and a deferred list; pass 1 (a second launch) resumes them, all long, together.
  const unsigned long long n_deferred = p.pass ? p.counters[3] : 0ULL;
  for (;;) {
    unsigned long long inst = 0;
    if (lane == 0) inst = atomicAdd(p.counters + (p.pass ? 4 : 0), 1ULL);
    inst = __shfl_sync(FULLMASK, inst, 0);
    if (p.pass) {
      if (inst >= n_deferred) break;
      inst = (unsigned long long)p.deferred[inst];
    } else if (inst >= (unsigned long long)p.B) {
      break;
    }

    // ---- load θ and the initial point (src/solver.jl:39-41,64-66) — or the parked state in pass 1 ----
#if THETA_IN_SMEM
    for (int i = lane; i < NT; i += 32) th[i] = p.theta[inst * NT + i];
#else
    const double* th = p.theta + inst * NT;
#endif
    double eps = 1.0;                                        // :67
    double kkt = __longlong_as_double(0x7ff0000000000000LL);  // Inf, :68
    int status = 0;                                          // :69
    int outer = 1;                                           // :70
    int steps = 0;
    if (p.pass) {
      for (int i = lane; i < NX; i += 32) x[i] = p.x_out[inst * NX + i];
      for (int i = lane; i < NY; i += 32) {
        y[i] = p.y_out[inst * NY + i];
        s[i] = p.s_out[inst * NY + i];
      }
      eps = p.eps_out[inst];
      kkt = p.kkt_out[inst];
      outer = p.outer_out[inst];
      steps = p.steps_out[inst];
    } else {
      for (int i = lane; i < NX; i += 32) x[i] = p.x0 ? p.x0[inst * NX + i] : 0.0;
      for (int i = lane; i < NY; i += 32) {
        y[i] = p.y0 ? p.y0[inst * NY + i] : 1.0;
        s[i] = p.s0 ? p.s0[inst * NY + i] : 1.0;
      }
    }
    __syncwarp();
    bool parked = false;
    while (kkt > tol && eps > tol && outer < p.max_outer) {  // :71
      if (p.pass == 0 && p.step_budget > 0 && steps >= p.step_budget) {
        parked = true;
        break;
      }
      int inner = 1;                                         // :72
      status = 0;                                            // :73
      while (kkt > eps && inner < p.max_inner) {             // :75
        // F and the Jacobian entries at the current iterate (:79-80)
        mcp_eval_newton_par(lane, x, y, th, g, hh, jv);   // lane i evaluates output group i
        __syncwarp();
        double fmax_ = 0.0;
        for (int i = lane; i < NX; i += 32) fmax_ = nanmax(fmax_, fabs(g[i]));
        for (int k = lane; k < NY; k += 32) {
          const double f2 = hh[k] - s[k];         // H − s        (src/mcp.jl:78)
          const double f3 = s[k] * y[k] - eps;         // s∘y − ϵ      (src/mcp.jl:79)
          const double yt = y[k] + tol;                // (3,3) block diag(y) + tol·I  (:81)
          const double di = 1.0 / (tol + s[k] / yt);   // D⁻¹, D = (2,2) block tol·I + S (Y+tol)⁻¹
          dinv[k] = di;
          w[k] = di * (-f2 - f3 / yt);
          fmax_ = nanmax(fmax_, nanmax(fabs(f2), fabs(f3)));
        }
        const double kkt_new = warp_nanmax(fmax_);           // ‖F‖∞ of the pre-step residual (:107)
        __syncwarp();

        // (∇F + tol·I) δz = −F, condensed to NRED unknowns (:81-83)
        for (int i = lane; i < NRED; i += 32) {
          double r = -g[R_GROW[i]];
          for (int e = R_PTR[i]; e < R_PTR[i + 1]; ++e) r -= R_COEF[e] * opval(R_CODE[e], jv, th) * w[R_K[e]];
          sol[i] = r;
        }
        __syncwarp();   // G (aliased onto the window) is dead from here on: the window becomes scratch
        assemble_matrix(Cval, W, jv, th, dinv, tol, lane);
        __syncwarp();
        if (band_solve<1, WS1>(W, Cval, UT, sol, rowptr, cpos, lane)) {          // :84-88
          status = 1;
          break;
        }
        // δy = w − D⁻¹ H_x δx ;  δs = −(F₃ + s δy)/(y + tol)
        for (int k = lane; k < NY; k += 32) {
          double hx = 0.0;
          for (int e = H_PTR[k]; e < H_PTR[k + 1]; ++e) hx += H_COEF[e] * opval(H_CODE[e], jv, th) * sol[H_COL[e]];
          const double dy = w[k] - dinv[k] * hx;
          const double f3 = s[k] * y[k] - eps;
          w[k] = dy;
          dinv[k] = -(f3 + s[k] * dy) / (y[k] + tol);
        }
        __syncwarp();
        const double a_s = ftb_linesearch(s, dinv, p.min_stepsize, lane);  // :93
        const double a_y = ftb_linesearch(y, w, p.min_stepsize, lane);     // :94
        if (a_s != a_s || a_y != a_y) {                                    // :96-100
          status = 1;
          break;
        }
        for (int c = lane; c < NRED; c += 32) x[PERM[c]] += a_s * sol[c];  // :103 (x uses α_s)
        for (int k = lane; k < NY; k += 32) {
          s[k] += a_s * dinv[k];                                           // :104
          y[k] += a_y * w[k];                                              // :105
        }
        __syncwarp();
        kkt = kkt_new;                                                     // :107
        ++inner;                                                           // :108
        ++steps;
      }
      eps *= (status == 0) ? 1.0 - exp(-p.tightening_rate * inner) : 1.0 + exp(-p.loosening_rate * inner);  // :111-113
      ++outer;                                                             // :114
    }
    if (!parked && outer == p.max_outer) status = 1;                       // :117-119

    for (int i = lane; i < NX; i += 32) p.x_out[inst * NX + i] = x[i];
    for (int i = lane; i < NY; i += 32) {
      p.y_out[inst * NY + i] = y[i];
      p.s_out[inst * NY + i] = s[i];
    }
    if (lane == 0) {
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
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Sensitivity kernel: ∂z/∂θ = (−∇F_z)⁻¹ ∇F_θ at the returned point, no tol·I (src/AutoDiff.jl:18-40),
// through the same condensation with D = S Y⁻¹, NRHS_SENS right-hand sides per factorisation pass.
// ------------------------------------------------------------------------------------------------
#if HAS_JT
extern "C" __global__ void __launch_bounds__(32 * SENS_WARPS, 1) mcp_sens_kernel(const SensParams p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  load_shared_tables(smem);
  const int* rowptr = reinterpret_cast<const int*>(smem);
  const unsigned short* cpos = reinterpret_cast<const unsigned short*>(rowptr + NRED + 1);
  double* S = smem + SHARED_TABLE_DOUBLES + (size_t)warp * SENS_SMEM_DOUBLES;
  double* x = S + SENS_OFF_X;
  double* y = S + SENS_OFF_Y;
  double* s = S + SENS_OFF_S;
  double* jv = S + SENS_OFF_JV;
  double* jtv = S + SENS_OFF_JTV;
  double* dinv = S + SENS_OFF_DINV;
  double* wq = S + SENS_OFF_WQ;    // [NRHS_SENS][NY]
  double* sol = S + SENS_OFF_SOL;  // [NRHS_SENS][NRED]
  double* W = S + SENS_OFF_WIN;
#if THETA_IN_SMEM
  double* th = S + SENS_OFF_TH;
#endif
  double* Cval = p.scratch + ((size_t)blockIdx.x * SENS_WARPS + warp) * SENS_SCRATCH;
  double* UT = Cval + CVAL_DOUBLES;
  constexpr int NZ = NX + 2 * NY;

  for (;;) {
    unsigned long long inst = 0;
    if (lane == 0) inst = atomicAdd(p.counters, 1ULL);
    inst = __shfl_sync(FULLMASK, inst, 0);
    if (inst >= (unsigned long long)p.B) break;
#if THETA_IN_SMEM
    for (int i = lane; i < NT; i += 32) th[i] = p.theta[inst * NT + i];
#else
    const double* th = p.theta + inst * NT;
#endif
    for (int i = lane; i < NX; i += 32) x[i] = p.x[inst * NX + i];
    for (int i = lane; i < NY; i += 32) {
      y[i] = p.y[inst * NY + i];
      s[i] = p.s[inst * NY + i];
    }
    __syncwarp();
    mcp_eval_sens_par(lane, x, y, th, jv, jtv);
    __syncwarp();
    for (int k = lane; k < NY; k += 32) dinv[k] = y[k] / s[k];  // D⁻¹ with D = s/y (tol = 0)
    if (p.z_p)
      for (int i = lane; i < NZ * p.P; i += 32) p.z_p[inst * NZ * p.P + i] = 0.0;
    __syncwarp();
    int bad = 0;
    for (int q0 = 0; q0 < NT; q0 += NRHS_SENS) {
      const int nq = min(NRHS_SENS, NT - q0);
      for (int i = lane; i < NRHS_SENS * NY; i += 32) wq[i] = 0.0;
      for (int i = lane; i < NRHS_SENS * NRED; i += 32) sol[i] = 0.0;
      assemble_matrix(Cval, W, jv, th, dinv, 0.0, lane);
      __syncwarp();
      // right-hand sides r = −∇F_θ[:, q]:  G rows go to the reduced rhs, H rows to w = D⁻¹ r₂
      for (int rq = 0; rq < nq; ++rq) {
        const int q = q0 + rq;
        for (int e = Q_PTR[q] + lane; e < Q_PTR[q + 1]; e += 32) {
          const double v = -Q_COEF[e] * opval(Q_CODE[e], jtv, th);
          const int row = Q_ROW[e];
          if (row < NX) sol[rq * NRED + IPERM[row]] = v;
          else wq[rq * NY + (row - NX)] = dinv[row - NX] * v;
        }
      }
      __syncwarp();
      for (int i = lane; i < NRED; i += 32) {
        for (int rq = 0; rq < nq; ++rq) {
          double r = sol[rq * NRED + i];
          for (int e = R_PTR[i]; e < R_PTR[i + 1]; ++e)
            r -= R_COEF[e] * opval(R_CODE[e], jv, th) * wq[rq * NY + R_K[e]];
          sol[rq * NRED + i] = r;
        }
      }
      __syncwarp();
      if (band_solve<NRHS_SENS, WSS>(W, Cval, UT, sol, rowptr, cpos, lane)) {
        bad = 1;
        break;
      }
      // recover the y and s rows:  Z_y = w − D⁻¹ H_x Z_x ,  Z_s = −s Z_y / y
      for (int rq = 0; rq < nq; ++rq) {
        const int q = q0 + rq;
        const double* so = sol + rq * NRED;
        double tb = 0.0;
        for (int c = lane; c < NRED; c += 32) {
          const double zx = so[c];
          const int row = PERM[c];
          if (p.dzdtheta) p.dzdtheta[(inst * NT + q) * NZ + row] = zx;
          if (p.zbar) tb += p.zbar[inst * NZ + row] * zx;
          if (p.z_p)
            for (int pp = 0; pp < p.P; ++pp)
              p.z_p[(inst * p.P + pp) * NZ + row] += zx * p.theta_p[(inst * p.P + pp) * NT + q];
        }
        for (int k = lane; k < NY; k += 32) {
          double hx = 0.0;
          for (int e = H_PTR[k]; e < H_PTR[k + 1]; ++e) hx += H_COEF[e] * opval(H_CODE[e], jv, th) * so[H_COL[e]];
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
        if (p.thetabar) {
          tb = warp_sum(tb);
          if (lane == 0) p.thetabar[inst * NT + q] = tb;
        }
      }
      __syncwarp();
    }
    if (lane == 0 && p.status_out) p.status_out[inst] = bad;
    __syncwarp();
  }
}
#endif  // HAS_JT
