"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol the header
declares, compiles problems for sm_100a in COMPILE_ONLY mode, reports API misuse through error codes,
and refuses — loudly — to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mcp_b200 import capi, problems
from mcp_b200.mcp import PrimalDualMCP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mcpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcpb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = capi.load_library()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mcpb200.h but not exported"
    assert sorted(capi.EXPORTS) == syms


def test_default_opts_match_reference():
    """src/solver.jl:42-49."""
    o = capi.default_opts()
    assert (o.tol, o.max_inner_iters, o.max_outer_iters) == (1e-4, 20, 50)
    assert (o.tightening_rate, o.loosening_rate, o.min_stepsize) == (0.1, 0.5, 1e-4)


@pytest.mark.parametrize("name", ["readme", "clamp", "lane", "qp12"])
def test_compile_only_for_sm100a(name):
    mcp = {"readme": problems.readme_qp, "clamp": lambda: problems.clamp_game().mcp,
           "lane": lambda: problems.lane_change_game().mcp, "qp12": lambda: problems.random_qp(12, 10)}[name]()
    h = capi.Handle(mcp.ir, capi.COMPILE_ONLY)
    info = h.info()
    assert info["nx"] == mcp.unconstrained_dimension and info["ny"] == mcp.constrained_dimension
    assert info["n_reduced"] == info["nx"]                     # condensed to the unconstrained block
    assert 1 <= info["window_rows"] <= 256 and info["window_cols"] <= info["n_reduced"]
    assert info["smem_bytes_per_cta"] <= 227 * 1024 and info["instances_per_cta"] >= 1
    src = h.source()
    assert "mcp_eval_newton" in src and "mcp_solve_kernel" in src
    if name == "lane":
        # the fill-reducing ordering must find the stage structure: SURVEY.md App. B measured bandwidth 32
        # for the stage-major ordering; RCM does better
        assert info["kl"] <= 20 and info["ku"] <= 20
        assert info["n_jac_constant"] == 1080 and info["has_sensitivities"] == 1
    if name == "qp12":
        assert info["has_sensitivities"] == 0 and info["n_jac_computed"] == 0   # every entry is ±θ_i: read in place
    h.close()


def test_cubin_is_sm100a(tmp_path, monkeypatch):
    import subprocess
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    h = capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY)
    h.close()
    cubins = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cubins) == 1
    out = subprocess.run(["cuobjdump", "-lelf", str(tmp_path / cubins[0])], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # second creation must hit the cache
    h2 = capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY)
    assert h2.info()["cache_hit"] == 1
    h2.close()


def test_invalid_ir_is_rejected():
    ir = problems.readme_qp().ir
    bad = type(ir)(**{**ir.__dict__})
    bad.a = ir.a.copy()
    k = int(np.nonzero(ir.op == 4)[0][0])     # an ADD node: point its operand at a later node
    bad.a[k] = len(ir.op) - 1
    with pytest.raises(capi.MCPB200Error) as e:
        capi.Handle(bad, capi.COMPILE_ONLY)
    assert e.value.code == capi.ERR_INVALID_ARGUMENT


def test_h_depending_on_y_compiles_in_full_y_mode():
    """∇_y H ≠ 0 (the reference accepts any H(x, y; θ), `src/mcp.jl:27-52,76-80`): the plan switches to the
    (nx+ny)-dimensional system with only δs eliminated; sensitivities (r2) are forward solves of the same system — no
    adjoint kernel in that mode."""
    mcp = PrimalDualMCP(lambda x, y, θ: x - θ - y, lambda x, y, θ: x + 0.5 * y, unconstrained_dimension=1,
                        constrained_dimension=1, parameter_dimension=1)
    h = capi.Handle(mcp.ir, capi.COMPILE_ONLY)
    src, info = h.source(), h.info()
    assert _macros(src)["FULL_Y"] == "1" and info["n_reduced"] == 2 and info["has_sensitivities"] == 1
    assert _macros(src)["HAS_ADJOINT"] == "0"
    h.close()
    # a problem of the reference's own structure stays in the condensed mode
    h = capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY)
    assert _macros(h.source())["FULL_Y"] == "0" and h.info()["n_reduced"] == 2
    h.close()


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail with an error code — never produce numbers."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from mcp_b200 import InteriorPoint, solve
    with pytest.raises(capi.MCPB200Error) as e:
        solve(InteriorPoint(), problems.readme_qp(), np.array([-0.5, 0.5]))
    assert e.value.code == capi.ERR_CUDA
    # a COMPILE_ONLY handle refuses device work as well
    h = capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY)
    out = np.zeros(8)
    rc = h._lib.mcpb200_solve_batched(h.raw, 1, out.ctypes.data, None, None, None, None, out.ctypes.data,
                                      out.ctypes.data, out.ctypes.data, out.ctypes.data, out.ctypes.data,
                                      out.ctypes.data, out.ctypes.data, None)
    assert rc == capi.ERR_CUDA


def test_missing_sensitivities_error_code():
    """The handle-level analogue of the ArgumentError at src/AutoDiff.jl:19-23."""
    h = capi.Handle(problems.readme_qp(compute_sensitivities=False).ir, capi.COMPILE_ONLY)
    z = np.zeros(8)
    rc = h._lib.mcpb200_sensitivities(h.raw, 1, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data,
                                      z.ctypes.data, z.ctypes.data, None, None, 0, None, None, None)
    assert rc == capi.ERR_NO_SENSITIVITIES
    assert b"Missing sensitivities" in h._lib.mcpb200_last_error(h.raw)


def test_solve_argument_validation(readme_mcp):
    from mcp_b200 import InteriorPoint, solve
    with pytest.raises(ValueError):
        solve(InteriorPoint(), readme_mcp, np.zeros(3))                  # wrong θ length
    with pytest.raises(ValueError):
        solve(InteriorPoint(), readme_mcp, np.zeros((2, 4)), x0=np.zeros((2, 3)))
    with pytest.raises(ValueError):
        solve(InteriorPoint(), readme_mcp, np.zeros(2), linear_solve_algorithm="KLU")
    with pytest.raises(TypeError):
        solve(object(), readme_mcp, np.zeros(2))


def _macros(src):
    import re
    return dict(re.findall(r"^#define (\w+) (-?\d+)$", src, flags=re.M))


def test_plan_variants_are_selected_and_compile(monkeypatch, tmp_path):
    """The planner picks a kernel family per problem structure; every family must compile for sm_100a."""
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    # banded trajectory game: register-resident window, one warp per instance
    m = _macros(capi.Handle(problems.lane_change_game().mcp.ir, capi.COMPILE_ONLY).source())
    assert (m["REGWIN"], m["SUB"], m["DENSE_KERNEL"], m["LARGE_STATE"]) == ("1", "32", "0", "0")
    assert int(m["KL"]) <= 12 and int(m["WR"]) <= 13          # annealed ordering (RCM alone gives 17)
    # tiny QP: one THREAD per instance for the solve (straight-line generated code on register arrays); the
    # sensitivity kernels keep two instances per warp
    h = capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY)
    m = _macros(h.source())
    assert (m["TINY_KERNEL"], m["SUB"], m["DENSE_KERNEL"]) == ("1", "16", "0") and h.info()["threads_per_instance"] == 1
    assert "tiny_assemble" in h.source() and "C[0][0] = " in h.source()
    monkeypatch.setenv("MCPB200_TINY", "0")
    h = capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY)
    assert _macros(h.source())["TINY_KERNEL"] == "0" and h.info()["threads_per_instance"] == 16
    monkeypatch.delenv("MCPB200_TINY")
    # dense QP with G_y = −H_xᵀ and an affine residual: CTA-per-instance kernel, matrix in register tiles (v3)
    h = capi.Handle(problems.random_qp(12, 10).ir, capi.COMPILE_ONLY)
    m = _macros(h.source())
    assert (m["DENSE_KERNEL"], m["DENSE_SCHUR"]) == ("3", "1") and h.info()["threads_per_instance"] == 512
    assert "mcp_eval_const_par" in h.source() and "mcp_eval_newton_p0" not in h.source()
    # v2 keeps the matrix in shared memory (256 threads), H_x cached per solve
    monkeypatch.setenv("MCPB200_DENSE_KERNEL", "2")
    h = capi.Handle(problems.random_qp(12, 10).ir, capi.COMPILE_ONLY)
    assert _macros(h.source())["DENSE_KERNEL"] == "2" and h.info()["threads_per_instance"] == 256
    # the same problem with the v2 structure requirements switched off falls back to v1, then to the window kernel
    monkeypatch.setenv("MCPB200_DENSE_KERNEL", "1")
    assert _macros(capi.Handle(problems.random_qp(12, 10).ir, capi.COMPILE_ONLY).source())["DENSE_KERNEL"] == "1"
    monkeypatch.setenv("MCPB200_DENSE_KERNEL", "0")
    assert _macros(capi.Handle(problems.random_qp(12, 10).ir, capi.COMPILE_ONLY).source())["DENSE_KERNEL"] == "0"
    monkeypatch.delenv("MCPB200_DENSE_KERNEL")
    # large-state mode (vectors in global memory) can be forced on any banded problem
    monkeypatch.setenv("MCPB200_LARGE_STATE", "1")
    h = capi.Handle(problems.lane_change_game().mcp.ir, capi.COMPILE_ONLY)
    assert _macros(h.source())["LARGE_STATE"] == "1" and h.info()["smem_bytes_per_cta"] < 64 * 1024


def test_shape_grouped_evaluation_and_split_compilation(monkeypatch, tmp_path):
    """Code generation: outputs that repeat an expression shape go into lock-step shape groups (index tables), the
    rest into lane-partitioned functions; large sets of those are compiled as separate units and linked."""
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    ir = problems.lane_change_game().mcp.ir
    src = capi.Handle(ir, capi.COMPILE_ONLY).source()
    m = re.search(r"// (\d+) outputs: (\d+) in (\d+) shape groups, (\d+) lane-partitioned", src)
    total, shaped, groups, rest = map(int, m.groups())
    assert total == 200 + 250 + 140 and shaped + rest == total
    assert shaped >= 0.9 * total and groups <= 40          # 10 stages x 2 players repeat a few dozen shapes
    assert "mcp_eval_newton_s" in src and "_I[" in src
    # every shape group stores through its index table into one of the target arrays
    assert len(re.findall(r"^\s+(?:g|h|jv)\[mcp_eval_newton_s\d+_I\[\d+ \+ q\]\] = ", src, flags=re.M)) == groups
    # shapes off: everything lane-partitioned, one translation unit
    monkeypatch.setenv("MCPB200_SHAPES", "0")
    h = capi.Handle(ir, capi.COMPILE_ONLY)
    assert "mcp_eval_newton_s0" not in h.source() and "mcp_eval_newton_rest_p31" in h.source()
    assert "extern __device__ void mcp_eval_newton_rest_p0" not in h.source()
    # ... and forced into separately compiled units: the main unit only declares the parts, nvJitLink resolves them
    monkeypatch.setenv("MCPB200_SPLIT_COMPILE", "1")
    h = capi.Handle(ir, capi.COMPILE_ONLY | capi.NO_CACHE)
    assert "extern __device__ void mcp_eval_newton_rest_p0" in h.source()
    assert "void mcp_eval_newton_rest_p0(const double* __restrict__ x" not in h.source().replace("extern __device__ void mcp_eval_newton_rest_p0(", "")


def test_cooperative_instances_are_planned_for_big_windows(monkeypatch, tmp_path):
    """Plans whose window stays in shared memory (too wide for the register layout) get NWIDE warps per instance;
    register-window plans and tiny problems keep one (sub-)warp per instance."""
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    # r2: the masked game at N = 4 (46 window entries per lane) is inside the register-window limit of 48 — one warp
    # per instance, at most 256 threads per CTA (≈ 250 registers per thread)
    h = capi.Handle(problems.masked_game(4, 30).mcp.ir, capi.COMPILE_ONLY)
    m = _macros(h.source())
    assert (m["NWIDE"], m["LARGE_STATE"], m["SUB"]) == ("1", "1", "32")
    assert h.info()["threads_per_instance"] == 32 and h.info()["instances_per_cta"] == 8
    # with the r1 limit its window stays in shared memory and gets helper warps
    monkeypatch.setenv("MCPB200_REGWIN_PW", "40")
    h = capi.Handle(problems.masked_game(4, 30).mcp.ir, capi.COMPILE_ONLY)
    m = _macros(h.source())
    assert (m["NWIDE"], m["LARGE_STATE"], m["SUB"]) == ("2", "1", "32")
    assert h.info()["threads_per_instance"] == 64 and h.info()["instances_per_cta"] == 8
    assert "WIDE_ASSEMBLE" in h.source()
    assert _macros(capi.Handle(problems.lane_change_game().mcp.ir, capi.COMPILE_ONLY).source())["NWIDE"] == "1"
    assert _macros(capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY).source())["NWIDE"] == "1"
    monkeypatch.setenv("MCPB200_NWIDE", "1")
    h = capi.Handle(problems.masked_game(4, 30).mcp.ir, capi.COMPILE_ONLY)
    assert _macros(h.source())["NWIDE"] == "1" and h.info()["threads_per_instance"] == 32


def test_hot_tables_move_to_shared_memory_when_it_costs_no_instance(monkeypatch, tmp_path):
    """r2: the per-step index / coefficient tables become `__shared__` arrays (filled from a global image `NAME_G` by
    LOAD_HOT_TABLES) when the plan keeps its resident instances with them: the lane-change plan (16 instances + 14.5 KB
    of tables); not the masked game (tables beyond the 48 KB static limit), not the dense kernels."""
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    h = capi.Handle(problems.lane_change_game().mcp.ir, capi.COMPILE_ONLY)
    src, info = h.source(), h.info()
    assert _macros(src)["HOT_SMEM"] == "1" and info["instances_per_cta"] == 16
    for name in ("D_TP", "T_COEF", "R_PTR", "R_COEF", "H_COL", "PERM"):
        assert re.search(r"^__shared__ \w[\w ]* %s\[\d+\];$" % name, src, flags=re.M), name
        assert re.search(r"^__device__ const \w[\w ]* %s_G\[\d+\] = " % name, src, flags=re.M), name
        assert "%s[i_] = %s_G[i_];" % (name, name) in src
    assert re.search(r"^__shared__ TI_T T_I\[\d+\];$", src, flags=re.M)
    assert "__shared__ short D_CPOS" not in src           # (already copied into the dynamic block by load_shared_tables)
    monkeypatch.setenv("MCPB200_HOT_SMEM", "0")
    h = capi.Handle(problems.lane_change_game().mcp.ir, capi.COMPILE_ONLY)
    assert _macros(h.source())["HOT_SMEM"] == "0" and "__shared__ short D_TP" not in h.source()
    assert h.info()["instances_per_cta"] == 16
    monkeypatch.delenv("MCPB200_HOT_SMEM")
    assert _macros(capi.Handle(problems.random_qp(12, 10).ir, capi.COMPILE_ONLY).source())["HOT_SMEM"] == "0"
    assert _macros(capi.Handle(problems.readme_qp().ir, capi.COMPILE_ONLY).source())["HOT_SMEM"] == "0"


@pytest.mark.parametrize("mk", [lambda: problems.lane_change_game(horizon=4).mcp, lambda: problems.lane_change_game(horizon=16).mcp,
                                lambda: problems.masked_game(3, 6).mcp, lambda: problems.masked_game(2, 12).mcp])
def test_hot_tables_never_cost_a_build(mk, monkeypatch, tmp_path):
    """Whatever the planner decides about the shared-memory tables for a banded problem of another size, the module must
    still compile (ptxas rejects > 48 KB of static shared memory) and keep the instances per CTA it has without them."""
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    mcp = mk()
    h = capi.Handle(mcp.ir, capi.COMPILE_ONLY)
    with_hot, hot = h.info()["instances_per_cta"], _macros(h.source())["HOT_SMEM"]
    h.close()
    monkeypatch.setenv("MCPB200_HOT_SMEM", "0")
    h = capi.Handle(mcp.ir, capi.COMPILE_ONLY)
    assert _macros(h.source())["HOT_SMEM"] == "0"
    assert h.info()["instances_per_cta"] == with_hot, hot
    h.close()


def test_split_units_respect_the_kernels_register_budget(monkeypatch, tmp_path):
    """Separately compiled evaluation units do not see the kernels' launch bounds; the runtime caps their registers
    so that nvJitLink accepts them (a 256-thread, 2-CTA/SM dense kernel may call only ≤128-register functions)."""
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    h = capi.Handle(problems.random_qp(111, 128).ir, capi.COMPILE_ONLY)     # too big for the register-tiled kernel
    assert h.info()["threads_per_instance"] == 256
    assert "extern __device__ void mcp_eval_newton_rest_p0" in h.source()    # its evaluation is split into units


def test_adjoint_tables_are_a_column_major_view_of_the_dests(monkeypatch, tmp_path):
    """Adjoint pullback: DT_SRC permutes the assembled non-zeros into column-major order, DT_ROWPTR/DT_CPOS describe
    the transposed matrix to the factorisation (same band: the plan requires kl == ku)."""
    monkeypatch.setenv("MCPB200_CACHE_DIR", str(tmp_path))
    src = capi.Handle(problems.lane_change_game().mcp.ir, capi.COMPILE_ONLY).source()
    m = _macros(src)
    assert m["HAS_ADJOINT"] == "1" and m["KL"] == m["KU"]

    def table(name):
        body = re.search(r"__device__ const (?:int|short) %s\[\d+\] = \{([^}]*)\}" % name, src).group(1)
        return np.array([int(v) for v in body.replace("\n", "").split(",")])

    nd, n, wc = int(m["ND"]), int(m["NRED"]), int(m["WC"])
    rowptr, cpos, tptr, tcpos, tsrc = (table(t) for t in ("D_ROWPTR", "D_CPOS", "DT_ROWPTR", "DT_CPOS", "DT_SRC"))
    assert sorted(tsrc.tolist()) == list(range(nd)) and tptr[0] == 0 and tptr[-1] == nd and len(tptr) == n + 1
    row_of = np.repeat(np.arange(n), np.diff(rowptr))              # row of every dest in row-major order
    for c in range(n):                                             # column c of C = row c of Cᵀ
        e = np.arange(tptr[c], tptr[c + 1])
        assert np.all(cpos[tsrc[e]] == c % wc)                     # ... really are the entries of column c
        assert np.all(tcpos[e] == row_of[tsrc[e]] % wc)            # ... placed at their row's window position
        assert np.all(np.diff(row_of[tsrc[e]]) > 0)
    # no sensitivities compiled in => no adjoint tables
    src2 = capi.Handle(problems.random_qp(12, 10).ir, capi.COMPILE_ONLY).source()
    assert _macros(src2)["HAS_ADJOINT"] == "0" and "const int DT_SRC[" not in src2
