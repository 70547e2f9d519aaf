"""TEST INFRASTRUCTURE ONLY — IR tape → straight-line C for [G; H] and the ∇F_z entries, compiled `-O2 -march=native`.

BASELINE.md §3 asks for the CPU baseline to evaluate F / ∇F from GENERATED code, as the reference does (its
`build_function` output, `/root/reference/src/mcp.jl:82-120`), not through a tape interpreter.  The emitted function is
what `oracle/c/mcp_oracle.c` calls when `oracle_problem.eval_fn` is set; same operations in the same order as the tape,
so results are bit-identical to the interpreted path.  Used by `oracle/c_oracle.py`; never by the product.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
MAX_NODES = 400000     # beyond this the interpreter is used instead

_BIN = {4: "+", 5: "-", 6: "*", 7: "/"}
_FUN = {9: "sqrt", 10: "exp", 11: "log", 12: "sin", 13: "cos"}


def _lit(v: float) -> str:
    if np.isnan(v):
        return "NAN"
    if np.isinf(v):
        return "INFINITY" if v > 0 else "-INFINITY"
    return float(v).hex()          # exact


def emit_source(ir) -> str:
    need = np.zeros(len(ir.op), dtype=bool)
    stack = [int(n) for n in ir.gh_nodes] + [int(n) for n in ir.jz_nodes]
    while stack:
        n = stack.pop()
        if need[n]:
            continue
        need[n] = True
        op = int(ir.op[n])
        if op in _BIN:
            stack += [int(ir.a[n]), int(ir.b[n])]
        elif op >= 8:
            stack.append(int(ir.a[n]))
    # Values live in a caller-provided scratch array and the code is split into functions of ≤ CHUNK statements:
    # gcc's compile time is super-linear in the size of one straight-line function (a 16 k-statement QP evaluator
    # took 570 s at -O2 as a single function, ≈ 2 s this way).
    CHUNK = 200
    head = ["#include <math.h>",
            "static inline double powi_(double a, int n) { double r = 1.0; int neg = n < 0; if (neg) n = -n; "
            "while (n) { if (n & 1) r *= a; a *= a; n >>= 1; } return neg ? 1.0 / r : r; }"]
    body, funcs = [], []

    def flush():
        if body:
            name = f"part{len(funcs)}"
            funcs.append(name)
            head.append(f"static void {name}(const double* restrict x, const double* restrict y, const double* restrict th, "
                        "double* restrict v, double* restrict gh, double* restrict jz) {")
            head.extend(body)
            head.append("}")
            body.clear()

    for n in np.nonzero(need)[0]:
        op, a, b = int(ir.op[n]), int(ir.a[n]), int(ir.b[n])
        if op == 0:
            e = _lit(float(ir.consts[a]))
        elif op == 1:
            e = f"x[{a}]"
        elif op == 2:
            e = f"y[{a}]"
        elif op == 3:
            e = f"th[{a}]"
        elif op in _BIN:
            e = f"v[{a}] {_BIN[op]} v[{b}]"
        elif op == 8:
            e = f"-v[{a}]"
        elif op in _FUN:
            e = f"{_FUN[op]}(v[{a}])"
        elif op == 14:
            e = f"powi_(v[{a}], {b})"
        else:
            e = "NAN"
        body.append(f"  v[{n}] = {e};")
        if len(body) >= CHUNK:
            flush()
    for i, n in enumerate(ir.gh_nodes):
        body.append(f"  gh[{i}] = v[{int(n)}];")
        if len(body) >= CHUNK:
            flush()
    for k, n in enumerate(ir.jz_nodes):
        body.append(f"  jz[{k}] = v[{int(n)}];")
        if len(body) >= CHUNK:
            flush()
    flush()
    head.append(f"int mcp_eval_scratch(void) {{ return {len(ir.op)}; }}")
    head.append("void mcp_eval(const double* x, const double* y, const double* th, double* gh, double* jz, double* v) {")
    head.extend(f"  {f}(x, y, th, v, gh, jz);" for f in funcs)
    head.append("}")
    return "\n".join(head) + "\n"


_cache = {}


def compiled_eval(ir, cpu_tag: str = ""):
    """ctypes function pointer (as c_void_p value) of the compiled evaluator, or None if the tape is too large."""
    if len(ir.op) > MAX_NODES:
        return None
    key = id(ir)
    if key in _cache:
        return _cache[key][1]
    src = emit_source(ir)
    h = hashlib.sha1((src + cpu_tag).encode()).hexdigest()[:16]
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, f"eval_{h}.so")
    if not os.path.exists(so):
        cpath = os.path.join(BUILD, f"eval_{h}.c")
        with open(cpath, "w") as f:
            f.write(src)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        # -ffp-contract=off: no FMA contraction, so the compiled code rounds exactly like the interpreted tape
        res = subprocess.run([cc, "-O2", "-march=native", "-ffp-contract=off", "-fPIC", "-shared", "-o", so + ".tmp", cpath, "-lm"],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("gcc failed on the generated evaluator:\n" + res.stderr[-2000:])
        os.replace(so + ".tmp", so)
    lib = C.CDLL(so)
    fn = C.cast(lib.mcp_eval, C.c_void_p)
    _cache[key] = (lib, fn)
    return fn
