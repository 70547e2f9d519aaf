// plan.h — host-side "MCP compiler": turns the MCP-IR (include/mcpb200.h) into
//   (1) the structure of the condensed KKT system and its banded ordering,
//   (2) table-driven assembly programs for that system,
//   (3) CUDA source: generated device functions for G, H and the Jacobian entries, spliced into the
//       sm_100a kernel template (kernel_template.cuh).
//
// This replaces the symbolic → compiled-closure step of the reference's PrimalDualMCP constructor
// (/root/reference/src/mcp.jl:82-148) and the symbolic analysis UMFPACK does once per sparsity pattern
// (/root/reference/src/solver.jl:61).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mcpb200.h"

namespace mcpb200 {

// A Jacobian entry (or any scalar operand of the assembly programs) is  coef * base  with
//   code == -1      : base = 1            (entry is a numeric constant, folded into coef)
//   code >=  0      : base = jv[code]     (computed per Newton step by the generated device function)
//   code <= -2      : base = theta[-2-code]   (entry is ±θ_i: read straight from the parameter vector)
struct Operand {
  double coef = 0.0;
  int32_t code = -1;
};

struct Plan {
  // ---- problem ------------------------------------------------------------------------------
  int nx = 0, ny = 0, nt = 0;
  std::vector<int32_t> op, a, b;
  std::vector<double> consts;
  std::vector<int32_t> gh_nodes;
  std::vector<int32_t> jz_rows, jz_cols, jz_nodes;
  bool has_jt = false;
  std::vector<int32_t> jt_rows, jt_cols, jt_nodes;

  // ---- Jacobian entry classification -----------------------------------------------------------
  std::vector<Operand> jz_opnd;       // per Jz entry
  std::vector<int32_t> jv_nodes;      // tape node of each computed entry (slot → node)
  std::vector<Operand> jt_opnd;       // per Jθ entry (code space: jtv slots)
  std::vector<int32_t> jtv_nodes;
  int n_const_entries = 0;

  // ---- condensed system -------------------------------------------------------------------------
  // Mode A (every config of the reference: ∇_y H ≡ 0): eliminate δs and δy, factorise the nx×nx matrix
  //   C = G_x + tol·I − G_y D⁻¹ H_x ,  D = tol + s/(y+tol)        
  // Mode B (∇_y H ≠ 0, e.g. LCP-style H(x, y)): only δs is eliminated; the (nx+ny)-dimensional system
  //   [G_x + tol·I, G_y; H_x, H_y + tol·I + diag(s/(y+tol))] [δx; δy] = [−F1; −F2 − F3/(y+tol)]
  // is factorised by the same banded machinery (the kernel's per-constraint array then holds s/(y+tol) instead of D⁻¹).
  int full_y = 0;
  int N = 0;                          // reduced dimension
  std::vector<int32_t> perm, iperm;   // perm[new] = old, iperm[old] = new  (fill-reducing ordering)
  int kl = 0, ku = 0;                 // bandwidths in the new ordering
  int WC = 0, R = 0;                  // window columns (circular) and rows (slots)
  int nrhs_sens = 0;                  // RHS columns per pass of the sensitivity kernel
  int WS1 = 0, WSS = 0;               // window row stride (doubles) for the solve / sensitivity kernels

  // dest d: row d_row[d] (new ordering), circular column position d_cpos[d]; terms [d_tptr[d], d_tptr[d+1])
  std::vector<int32_t> d_row, d_col, d_cpos, d_tptr, d_diag;
  std::vector<double> d_base;        // numeric-constant part of each dest, folded on the host
  // term t: value = t_coef · val(t_a) · (t_k >= 0 ? dinv[t_k] · val(t_b) : 1)
  std::vector<double> t_coef;
  std::vector<int32_t> t_a, t_b, t_k;
  // rhs of reduced row i (new ordering): −G[r_grow[i]] − Σ_e coef·val(code)·w[k]
  std::vector<int32_t> r_grow, r_ptr, r_code, r_k;
  std::vector<double> r_coef;
  // (H_x v)_k = Σ_e coef·val(code)·v[col]  with col in the new ordering
  std::vector<int32_t> h_ptr, h_code, h_col;
  std::vector<double> h_coef;
  // G_y by column k (rows in the new ordering): dense Schur accumulation
  int large_state = 0;                // per-instance vectors in a global block, only the window in shared memory
  int64_t state_doubles_solve = 0, state_doubles_sens = 0;
  int dense_schur = 0;
  int tiny_kernel = 0;                // thread-per-instance solve kernel for problems of a few unknowns (README QP)
  int dense_kernel = 0;               // CTA-per-instance dense solve kernel (kernel_template.cuh, DENSE_KERNEL)
  int dense_ctas_per_sm = 1;
  int dense_threads = 256;            // CTA size of the dense kernel (v3: 512)
  bool gy_is_mhxt = false, hx_zconst = false, affine = false;   // structure flags (dense kernel v2)
  std::vector<int32_t> gk_ptr, gk_row, gk_code;
  std::vector<double> gk_coef;
  // θ-Jacobian, by column q: entries (row in [G;H], operand)
  std::vector<int32_t> q_ptr, q_row, q_code;
  std::vector<double> q_coef;

  // ---- launch configuration -----------------------------------------------------------------------
  int sub = 32;                       // lanes per instance (16: two instances per warp)
  int regwin = 0;                     // window rows in registers (shuffle broadcast) instead of shared memory
  int ipc_solve = 1, ipc_sens = 1;    // instances (warps) per CTA
  int hot_smem_bytes = 0;             // static shared memory of the hot tables (0: they stay in global memory)
  int has_adjoint = 0, ipc_adj = 1;   // adjoint-mode pullback kernel (one right-hand side: its own, denser layout)
  int64_t smem_adj = 0, state_doubles_adj = 0;
  int nwide = 1;                      // warps cooperating on one instance's window sweep (solve kernel, shared-memory window)
  int theta_in_smem = 1;
  int64_t smem_solve = 0, smem_sens = 0;
  int64_t scratch_doubles_solve = 0, scratch_doubles_sens = 0;  // per warp, global memory
  double flops_band = 0.0;

  std::string source;                 // generated CUDA translation unit (tables, dispatchers, kernels)
  std::vector<std::string> units;     // big problems: the generated evaluation parts, split into separately compiled units
  std::string error;
};

// Returns MCPB200_OK or an error code (message in plan.error).
int build_plan(const mcpb200_problem_desc& desc, const std::string& kernel_template, Plan& plan);

}  // namespace mcpb200
