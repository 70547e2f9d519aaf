/*
 * mcpb200.h — C ABI of libmcpb200.so, the B200-native batched interior-point MCP solver.
 *
 * The reference (TianyuQ/MCP = MixedComplementarityProblems.jl v0.1.9) has no FFI of its own: its
 * boundary is the Julia API exported at src/MixedComplementarityProblems.jl:16 plus the AD rules.
 * Each entry point below names the reference interface it replaces; the Julia `ccall` stubs a
 * maintainer would add are in INTEGRATION.md (and julia/MCPB200.jl), the Python ctypes mirror in
 * mcp_b200/capi.py.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns an int status (0 = MCPB200_OK,
 *     negative = error, message via mcpb200_last_error / mcpb200_global_error).  No exception
 *     crosses the ABI.
 *   - all floating point data is IEEE double; matrices are COLUMN-MAJOR like Julia's, so a
 *     parameter batch is `theta[ntheta x B]` with each instance's θ contiguous.
 *   - host entry points take HOST pointers and block until results are on the host; the
 *     `_device` variants take DEVICE pointers on the current CUDA device and enqueue on `stream`.
 *   - per-instance numerical outcome is in `status_out` (the reference's `status` Symbol,
 *     src/solver.jl:69,86,98,118): 0 = :solved, 1 = :failed.
 *   - calls on one handle are serialised internally; distinct handles are independent.
 *   - there is NO CPU fallback: compute entry points fail with MCPB200_ERR_CUDA when no CUDA
 *     device / driver is usable.
 */
#ifndef MCPB200_H
#define MCPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCPB200_VERSION 100  /* 0.1.0 */

/* ---- status codes ---------------------------------------------------------------------- */
#define MCPB200_OK 0
#define MCPB200_ERR_INVALID_ARGUMENT (-1)
#define MCPB200_ERR_UNSUPPORTED (-2)       /* problem structure outside what the kernels handle */
#define MCPB200_ERR_COMPILE (-3)           /* NVRTC rejected the generated source */
#define MCPB200_ERR_CUDA (-4)              /* CUDA runtime / driver failure, or no GPU */
#define MCPB200_ERR_NO_SENSITIVITIES (-5)  /* handle built without the θ-Jacobian; the wrapper turns this into
                                              the reference's ArgumentError (src/AutoDiff.jl:19-23) */
#define MCPB200_ERR_INTERNAL (-6)

/* ---- MCP-IR op-codes (tape of the traced G, H; mirrors mcp_b200/trace.py) ------------------ */
enum mcpb200_op {
  MCPB200_OP_CONST = 0, /* a = index into consts */
  MCPB200_OP_X = 1,     /* a = index into x (unconstrained variable) */
  MCPB200_OP_Y = 2,     /* a = index into y (constrained variable)   */
  MCPB200_OP_THETA = 3, /* a = index into theta */
  MCPB200_OP_ADD = 4,
  MCPB200_OP_SUB = 5,
  MCPB200_OP_MUL = 6,
  MCPB200_OP_DIV = 7,
  MCPB200_OP_NEG = 8,
  MCPB200_OP_SQRT = 9,
  MCPB200_OP_EXP = 10,
  MCPB200_OP_LOG = 11,
  MCPB200_OP_SIN = 12,
  MCPB200_OP_COS = 13,
  MCPB200_OP_POWI = 14 /* a ** b, b an integer literal */
};

/*
 * Problem description = what `PrimalDualMCP(G_symbolic, H_symbolic, x, y, θ; compute_sensitivities)`
 * (src/mcp.jl:55-150) derives from the traced expressions: the residual rows G (nx) and H (ny), the
 * sparse Jacobian of [G; H] w.r.t. [x; y] in CSC order (src/mcp.jl:97-120) and, optionally, w.r.t. θ
 * (src/mcp.jl:122-148).  The slack/barrier rows  H - s,  s∘y - ϵ  (src/mcp.jl:76-80) and their Jacobian
 * blocks -I, diag(s), diag(y) are structural and added by the library.
 * Nodes are in topological order (operands precede users).  All arrays are only read during
 * mcpb200_create.
 */
typedef struct mcpb200_problem_desc {
  int32_t nx;      /* unconstrained_dimension (src/mcp.jl:21) */
  int32_t ny;      /* constrained_dimension   (src/mcp.jl:23) */
  int32_t ntheta;  /* parameter_dimension */
  int32_t n_nodes;
  const int32_t* op; /* [n_nodes] enum mcpb200_op */
  const int32_t* a;  /* [n_nodes] */
  const int32_t* b;  /* [n_nodes] */
  int32_t n_consts;
  const double* consts;    /* [n_consts] */
  const int32_t* gh_nodes; /* [nx+ny] tape nodes of [G; H] */
  int32_t jz_nnz;
  const int32_t* jz_rows;  /* [jz_nnz] row in [G; H], 0-based, CSC order */
  const int32_t* jz_cols;  /* [jz_nnz] column in [x; y] */
  const int32_t* jz_nodes; /* [jz_nnz] */
  int32_t jt_nnz;          /* -1: built with compute_sensitivities = false (∇F_θ! === nothing, src/mcp.jl:123) */
  const int32_t* jt_rows;
  const int32_t* jt_cols;  /* column = θ index */
  const int32_t* jt_nodes;
} mcpb200_problem_desc;

/* Keyword arguments of `solve(::InteriorPoint, mcp, θ; …)`, src/solver.jl:42-49 (same defaults). */
typedef struct mcpb200_solver_opts {
  double tol;             /* 1e-4 */
  int32_t max_inner_iters; /* 20 */
  int32_t max_outer_iters; /* 50 */
  double tightening_rate; /* 0.1 */
  double loosening_rate;  /* 0.5 */
  double min_stepsize;    /* 1e-4 (src/solver.jl:48; the docstring's 1e-2 is stale) */
} mcpb200_solver_opts;

/* create flags */
#define MCPB200_COMPILE_ONLY 1u /* run the IR compiler + NVRTC for sm_100a but never touch a GPU (CPU build check) */
#define MCPB200_NO_CACHE 2u     /* neither read nor write the on-disk cubin cache */

typedef struct mcpb200_info {
  int32_t nx, ny, ntheta;
  int32_t n_reduced;      /* dimension of the condensed KKT system that is factorised */
  int32_t kl, ku;         /* its lower / upper bandwidth after the fill-reducing (RCM) ordering */
  int32_t window_rows, window_cols, row_stride;
  int32_t n_jac_computed; /* Jacobian entries evaluated per Newton step (z- or θ-dependent) */
  int32_t n_jac_constant; /* entries folded into constant tables */
  int32_t n_assembly_dests, n_assembly_terms;
  int32_t threads_per_instance, instances_per_cta, ctas_per_sm;
  int32_t smem_bytes_per_cta;
  int32_t regs_solve, regs_sens; /* registers/thread reported by the driver (0 in COMPILE_ONLY mode) */
  int32_t has_sensitivities;
  int32_t cache_hit;      /* 1 if the cubin came from the on-disk cache */
  double flops_per_newton_step_band; /* algorithmic flops of one banded factor+solve */
} mcpb200_info;

typedef struct mcpb200_timing {
  double kernel_ms;       /* CUDA-event time of the last solve/sensitivity kernel(s), max over devices */
  double h2d_ms, d2h_ms;  /* host-entry-point copies (0 for _device calls) */
  int64_t launches;       /* kernels of this library launched by the last call */
  int64_t newton_steps;   /* total Newton steps taken by the last solve call (sum over instances) */
  int64_t solved;         /* instances with status 0 in the last solve call */
  double pass0_ms;        /* part of kernel_ms spent in the first scheduling pass */
  int64_t deferred;       /* instances parked by pass 0 and finished by pass 1 */
} mcpb200_timing;

typedef struct mcpb200_problem* mcpb200_handle;

/* ---- lifecycle --------------------------------------------------------------------------- */

/* Replaces the code-generation half of the PrimalDualMCP constructors (src/mcp.jl:55-150): analyses the
 * IR, builds the condensed-system assembly tables, generates CUDA source and compiles it for sm_100a.
 * Any H(x, y; θ) the reference accepts is accepted (∇_y H ≠ 0 switches to the (nx+ny)-dimensional system,
 * sensitivities included); a factorisation window too large for one SM's shared memory moves to global
 * memory (slow, not refused).  MCPB200_ERR_UNSUPPORTED remains for one case: a condensed system whose
 * ordered bandwidth needs more than 256 window rows. */
int mcpb200_create(const mcpb200_problem_desc* desc, uint32_t flags, mcpb200_handle* out);
int mcpb200_destroy(mcpb200_handle h);
const char* mcpb200_last_error(mcpb200_handle h);
const char* mcpb200_global_error(void); /* errors that happened before a handle existed */
int mcpb200_get_info(mcpb200_handle h, mcpb200_info* info);
int mcpb200_get_timing(mcpb200_handle h, mcpb200_timing* t);
/* generated CUDA source (NUL-terminated, owned by the handle) */
int mcpb200_get_source(mcpb200_handle h, const char** src, int64_t* len);
void mcpb200_default_opts(mcpb200_solver_opts* opts); /* src/solver.jl:42-49 */

/* Devices used by the HOST entry points; the θ batch is split into contiguous column blocks, one per
 * device, no collective (instances are independent).  Default: device 0. */
int mcpb200_set_devices(mcpb200_handle h, const int32_t* device_ids, int32_t count);

/* ---- the hot path -------------------------------------------------------------------------- */

/* Batched `solve(InteriorPoint(), mcp, θ; x₀, y₀, s₀, tol, …)` (src/solver.jl:35-122), one solve per column
 * of theta.  x0/y0/s0 may be NULL (defaults zeros/ones/ones, src/solver.jl:39-41) or [nx|ny|ny x B].
 * Outputs (caller-allocated): x_out [nx x B], y_out, s_out [ny x B], kkt_error_out, eps_out [B]
 * (ϵ after the last update, src/solver.jl:111-113,121), outer_iters_out, status_out [B];
 * newton_steps_out [B] may be NULL.  Outputs may alias the corresponding x0/y0/s0 (the reference mutates
 * and returns its x₀, src/solver.jl:64-66). */
int mcpb200_solve_batched(mcpb200_handle h, int64_t B, const double* theta, const double* x0, const double* y0,
                          const double* s0, const mcpb200_solver_opts* opts, double* x_out, double* y_out,
                          double* s_out, double* kkt_error_out, double* eps_out, int32_t* outer_iters_out,
                          int32_t* status_out, int32_t* newton_steps_out);

/* Same, all pointers are device memory on the CURRENT device; work is enqueued on `stream`
 * (a cudaStream_t passed as void*).  Kernel time is recorded with CUDA events on that stream and
 * readable through mcpb200_get_timing after the stream is synchronised. */
int mcpb200_solve_batched_device(mcpb200_handle h, int64_t B, const double* theta, const double* x0,
                                 const double* y0, const double* s0, const mcpb200_solver_opts* opts,
                                 double* x_out, double* y_out, double* s_out, double* kkt_error_out,
                                 double* eps_out, int32_t* outer_iters_out, int32_t* status_out,
                                 int32_t* newton_steps_out, void* stream);

/* ---- sensitivities --------------------------------------------------------------------------- */

/* Batched `_solve_jacobian_θ` (src/AutoDiff.jl:18-40): ∂z/∂θ = (-∇F_z)⁻¹ ∇F_θ at the given solutions
 * (x, y, s, eps as returned by the solve); NO tol·I regularisation, like the reference.
 *   dzdtheta_out  [n x ntheta x B] (n = nx+2ny, rows ordered x, y, s), or NULL
 *   zbar [n x B] + thetabar_out [ntheta x B]: the pullback of the rrule, src/AutoDiff.jl:59-76, or NULL/NULL
 *   theta_p [ntheta x P x B] + z_p_out [n x P x B]: forward rule z_p = ∂z∂θ·θ_p, src/AutoDiff.jl:98, or NULL/NULL
 * sens_status_out [B] (may be NULL): 0 ok, 1 singular system.
 * Returns MCPB200_ERR_NO_SENSITIVITIES when the handle has no θ-Jacobian. */
int mcpb200_sensitivities(mcpb200_handle h, int64_t B, const double* theta, const double* x, const double* y,
                          const double* s, const double* eps, double* dzdtheta_out, const double* zbar,
                          double* thetabar_out, int32_t P, const double* theta_p, double* z_p_out,
                          int32_t* sens_status_out);
int mcpb200_sensitivities_device(mcpb200_handle h, int64_t B, const double* theta, const double* x,
                                 const double* y, const double* s, const double* eps, double* dzdtheta_out,
                                 const double* zbar, double* thetabar_out, int32_t P, const double* theta_p,
                                 double* z_p_out, int32_t* sens_status_out, void* stream);

/* ---- measurement helpers (used by bench.py; not part of the reference surface) ---------------- */
/* Runs a dependent-FMA FP64 kernel on the current device and returns achieved TFLOP/s: the
 * FP64-pipe roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
int mcpb200_measure_fp64_peak(double* tflops_out);
/* Writes `bytes` of device memory to evict L2 between timed iterations. */
int mcpb200_flush_l2(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCPB200_H */
