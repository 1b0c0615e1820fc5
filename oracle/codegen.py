"""TEST / BASELINE INFRASTRUCTURE: a code-generating CPU baseline for the batch Taylor stepper.

The reference does not interpret anything at run time: taylor_add_adaptive_step() (src/taylor_00.cpp:712-865) emits,
through LLVM, ONE straight-line function for the whole decomposition, unrolled over the Taylor orders, on SIMD
vectors of `batch_size` doubles (default, non-compact mode: src/taylor_02.cpp:1306-1419), with the default-mode
(pairwise) summation of the recurrences (src/math/*.cpp taylor_diff_*_impl) and contraction allowed
(src/llvm_state.cpp:842-845). LLVM is not available in this image, so this module does the same thing with gcc: it
walks the lowered program (the same hy_program the CUDA kernels run), emits C with GCC vector types (W = 4 or 8
lanes), one function per order, every u variable's order-n coefficient as one expression with literal constants,
and compiles it with -O2 -march=native -ffp-contract=fast. The generated jet plugs into the oracle's driver
(oracle/taylor_oracle.c: step size, state update, propagate loop) through oracle_set_jet_w{4,8}().

It is what bench.py's `--impl reference` arm and `cpu_baseline` time (kind "codegen"); tests/test_codegen_cpu.py pins
it to the same golden fixtures as the interpreting oracle. Only tests/, bench.py's CPU legs and
__graft_entry__.build() may import this.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")

# opcodes (include/heyoka_b200.h)
(OP_SUM, OP_SUM_SQ, OP_SUB_VV, OP_SUB_VN, OP_SUB_NV, OP_SUB_VP, OP_SUB_PV, OP_NEG, OP_MUL_VV, OP_MUL_NV, OP_MUL_PV,
 OP_DIV_VV, OP_DIV_NV, OP_DIV_PV, OP_DIV_VN, OP_DIV_VP, OP_SQUARE, OP_SQRT, OP_POW_VN, OP_POW_VP, OP_SIN, OP_COS,
 OP_TANH, OP_EXP, OP_LOG, OP_TIME, OP_CFUNC, OP_SIGMOID, OP_RELU, OP_RELUP) = range(30)
REF_VAR, REF_NUM, REF_PAR = 0, 1, 2
POW_GENERAL, POW_POS_INT, POW_NEG_INT, POW_POS_HALF, POW_NEG_HALF = range(5)
CF = ["identity", "sum", "prod", "sub", "div", "pow", "sum_sq", "sin", "cos", "tanh", "exp", "log", "sigmoid", "relu",
      "relup"]


def _kind(r):
    return r >> 30


def _idx(r):
    return r & 0x3fffffff


def _lit(x):
    """Exact C literal of a double."""
    x = float(x)
    if x != x:
        return "NAN"
    if x in (float("inf"), float("-inf")):
        return "INFINITY" if x > 0 else "-INFINITY"
    return "S(%s)" % x.hex()


def _pairwise(terms):
    """pairwise_reduce (src/detail/llvm_helpers_algo.cpp:271-308) as one expression."""
    terms = list(terms)
    if not terms:
        return "ZERO"
    while len(terms) > 1:
        nxt = []
        for i in range(0, len(terms), 2):
            nxt.append(terms[i] if i + 1 == len(terms) else "(%s + %s)" % (terms[i], terms[i + 1]))
        terms = nxt
    return terms[0]


class _Gen:
    def __init__(self, P):
        d = P.desc
        self.n_eq, self.n_uvars, self.n_pars, self.order = d.n_eq, d.n_uvars, d.n_pars, d.order
        self.ops = [] if d.n_uvars == d.n_eq else [tuple(int(v) for v in row) for row in P.ops_array()]
        def arr(ptr, ctype, n):
            return [] if n == 0 else np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).tolist()

        self.args = arr(d.args, C.c_uint32, d.n_args)
        self.consts = arr(d.consts, C.c_double, d.n_consts)
        self.sv_defs = arr(d.sv_defs, C.c_uint32, d.n_eq)

    def T(self, o, u):
        return "T[%d]" % (o * self.n_uvars + u)

    def numpar(self, ref):
        return _lit(self.consts[_idx(ref)]) if _kind(ref) == REF_NUM else "par[%d]" % _idx(ref)

    def conv(self, n, ua, ub, j0, j1, weight=0, alpha=None):
        """sum_{j=j0..j1} w(j) A^[n-j] B^[j], pairwise (default mode)."""
        terms = []
        for j in range(j0, j1 + 1):
            t = "%s * %s" % (self.T(n - j, ua), self.T(j, ub))
            if weight == 1:
                t = "%s * (%s)" % (_lit(j), t)
            elif weight == 2:
                t = "%s * (%s)" % (alpha(n, j), t)
            terms.append("(%s)" % t)
        return _pairwise(terms)

    def pow_ebs(self, base, e):
        if e == 0:
            return "ONE"
        if e == 1:
            return base
        if e % 2 == 0:
            return self.pow_ebs("(%s * %s)" % (base, base), e // 2)
        return "(%s * %s)" % (base, self.pow_ebs("(%s * %s)" % (base, base), (e - 1) // 2))

    def pow_eval(self, algo, x, expo):
        typ, n = algo >> 8, algo & 0xff
        if typ == POW_POS_INT:
            return self.pow_ebs(x, n)
        if typ == POW_NEG_INT:
            return "(ONE / %s)" % self.pow_ebs(x, n)
        if typ == POW_POS_HALF:
            return self.pow_ebs("v_sqrt(%s)" % x, n)
        if typ == POW_NEG_HALF:
            return "(ONE / %s)" % self.pow_ebs("v_sqrt(%s)" % x, n)
        return "v_pow(%s, %s)" % (x, expo)

    def cfunc(self, op):
        _, fn, off, cnt = op
        v = [self.numpar(self.args[off + k]) for k in range(cnt)]
        name = CF[fn]
        if name == "identity":
            return v[0]
        if name == "sum":
            return _pairwise(v)
        if name in ("prod", "sub", "div"):
            return "(%s %s %s)" % (v[0], {"prod": "*", "sub": "-", "div": "/"}[name], v[1])
        if name == "pow":
            eref = self.args[off + 1]
            algo = _pow_algo_of(self.consts[_idx(eref)]) if _kind(eref) == REF_NUM else (POW_GENERAL << 8)
            return self.pow_eval(algo, v[0], v[1])
        if name == "sum_sq":
            return _pairwise(["(%s * %s)" % (x, x) for x in v])
        if name == "relu":
            return "v_relu(%s, %s, %s)" % (v[0], v[0], v[1])
        if name == "relup":
            return "v_relup(%s, %s)" % (v[0], v[1])
        return "v_%s(%s)" % (name, v[0])

    def diff(self, k, n):
        """Expression of the order-n coefficient of u variable n_eq + k (default-mode recurrences)."""
        op = self.ops[k]
        oc, a, b, c = op
        u = self.n_eq + k
        T = self.T
        if oc == OP_SUM:
            v = []
            for i in range(b):
                ref = self.args[a + i]
                v.append(T(n, _idx(ref)) if _kind(ref) == REF_VAR else (self.numpar(ref) if n == 0 else "ZERO"))
            return _pairwise(v)
        if oc == OP_SUM_SQ:
            tmp = []
            for i in range(b):
                ref = self.args[a + i]
                isv = _kind(ref) == REF_VAR
                if n % 2 == 1:
                    tmp.append(self.conv(n, _idx(ref), _idx(ref), 0, (n - 1) // 2) if isv else "ZERO")
                    continue
                if isv:
                    sq = "(%s * %s)" % (T(n // 2, _idx(ref)), T(n // 2, _idx(ref)))
                elif n == 0:
                    sq = "(%s * %s)" % (self.numpar(ref), self.numpar(ref))
                else:
                    sq = "ZERO"
                if n > 0:
                    acc = self.conv(n, _idx(ref), _idx(ref), 0, (n - 2) // 2) if isv else "ZERO"
                    tmp.append("(v_dbl(%s) + %s)" % (acc, sq))
                else:
                    tmp.append(sq)
            r = _pairwise(tmp)
            return "v_dbl(%s)" % r if n % 2 == 1 else r
        if oc == OP_SUB_VV:
            return "%s - %s" % (T(n, a), T(n, b))
        if oc in (OP_SUB_VN, OP_SUB_VP):
            ref = (REF_NUM if oc == OP_SUB_VN else REF_PAR) << 30 | b
            return "%s - %s" % (T(n, a), self.numpar(ref)) if n == 0 else T(n, a)
        if oc in (OP_SUB_NV, OP_SUB_PV):
            ref = (REF_NUM if oc == OP_SUB_NV else REF_PAR) << 30 | a
            return "%s - %s" % (self.numpar(ref), T(n, b)) if n == 0 else "-%s" % T(n, b)
        if oc == OP_NEG:
            return "-%s" % T(n, a)
        if oc == OP_MUL_NV:
            return "%s * %s" % (_lit(self.consts[a]), T(n, b))
        if oc == OP_MUL_PV:
            return "par[%d] * %s" % (a, T(n, b))
        if oc == OP_MUL_VV:
            return self.conv(n, a, b, 0, n)
        if oc in (OP_DIV_VV, OP_DIV_NV, OP_DIV_PV):
            c0 = T(0, b)
            if n == 0:
                num = T(0, a) if oc == OP_DIV_VV else self.numpar((REF_NUM if oc == OP_DIV_NV else REF_PAR) << 30 | a)
                return "%s / %s" % (num, c0)
            acc = self.conv(n, u, b, 1, n)
            return "(%s - %s) / %s" % (T(n, a), acc, c0) if oc == OP_DIV_VV else "(-%s) / %s" % (acc, c0)
        if oc == OP_DIV_VN:
            return "%s / %s" % (T(n, a), _lit(self.consts[b]))
        if oc == OP_DIV_VP:
            return "%s / par[%d]" % (T(n, a), b)
        if oc == OP_SQUARE:
            if n == 0:
                return "%s * %s" % (T(0, a), T(0, a))
            if n % 2 == 1:
                return "v_dbl(%s)" % self.conv(n, a, a, 0, (n - 1) // 2)
            return "v_dbl(%s) + (%s * %s)" % (self.conv(n, a, a, 0, (n - 2) // 2), T(n // 2, a), T(n // 2, a))
        if oc == OP_SQRT:
            if n == 0:
                return "v_sqrt(%s)" % T(0, a)
            even = n % 2 == 0
            upper = (n - (2 if even else 1)) // 2
            fac = T(n, a)
            if even:
                fac = "(%s - %s * %s)" % (fac, T(n // 2, u), T(n // 2, u))
            if upper >= 1:
                fac = "(%s - v_dbl(%s))" % (fac, self.conv(n, u, u, 1, upper))
            return "%s / v_dbl(%s)" % (fac, T(0, u))
        if oc in (OP_POW_VN, OP_POW_VP):
            if oc == OP_POW_VN:
                al = self.consts[b]
                expo = _lit(al)
                fac = lambda nn, j: _lit(float(nn) * al - float(j) * (al + 1.))  # noqa: E731
                algo = c
            else:
                expo = "par[%d]" % b
                fac = lambda nn, j: "(%s * %s - %s * (%s + ONE))" % (_lit(nn), expo, _lit(j), expo)  # noqa: E731
                algo = POW_GENERAL << 8
            if n == 0:
                return self.pow_eval(algo, T(0, a), expo)
            return "%s / (%s * %s)" % (self.conv(n, a, u, 0, n - 1, 2, fac), _lit(n), T(0, a))
        if oc in (OP_SIN, OP_COS, OP_TANH, OP_EXP, OP_LOG):
            name = {OP_SIN: "sin", OP_COS: "cos", OP_TANH: "tanh", OP_EXP: "exp", OP_LOG: "log"}[oc]
            if n == 0:
                return "v_%s(%s)" % (name, T(0, a))
            if oc == OP_SIN:
                return "%s / %s" % (self.conv(n, c, a, 1, n, 1), _lit(n))
            if oc == OP_COS:
                return "%s / %s" % (self.conv(n, c, a, 1, n, 1), _lit(-n))
            if oc == OP_TANH:
                return "%s - %s / %s" % (T(n, a), self.conv(n, c, a, 1, n, 1), _lit(n))
            if oc == OP_EXP:
                return "%s / %s" % (self.conv(n, u, a, 1, n, 1), _lit(n))
            ret = "%s * %s" % (_lit(n), T(n, a))
            if n > 1:
                ret = "(%s - %s)" % (ret, self.conv(n, a, u, 1, n - 1, 1))
            return "%s / (%s * %s)" % (ret, _lit(n), T(0, a))
        if oc == OP_TIME:
            return "tm" if n == 0 else ("ONE" if n == 1 else "ZERO")
        if oc == OP_CFUNC:
            return self.cfunc(op) if n == 0 else "ZERO"
        if oc == OP_SIGMOID:
            if n == 0:
                return "v_sigmoid(%s)" % T(0, a)
            terms = ["(((%s - %s) * %s) * %s)" % (T(n - j, u), T(n - j, c), T(j, a), _lit(j)) for j in range(1, n + 1)]
            return "%s / %s" % (_pairwise(terms), _lit(n))
        if oc == OP_RELUP:
            return "v_relup(%s, %s)" % (T(0, a), _lit(self.consts[b])) if n == 0 else "ZERO"
        if oc == OP_RELU:
            return "v_relu(%s, %s, %s)" % (T(0, a), T(n, a), _lit(self.consts[b]))
        raise NotImplementedError("opcode %d" % oc)

    def sv(self, i, n):
        ref = self.sv_defs[i]
        if _kind(ref) == REF_VAR:
            return "%s / %s" % (self.T(n - 1, _idx(ref)), _lit(n))
        return self.numpar(ref) if n == 1 else "ZERO"

    def source(self, W):
        out = [_PRELUDE % {"W": W, "B": W * 8}]
        n_ops = self.n_uvars - self.n_eq
        for n in range(self.order + 1):
            out.append("static void cg_order_%d(v *restrict T, const v *restrict par, const v tm)\n{\n" % n)
            out.append("    (void)par; (void)tm;\n")
            if n > 0:
                for i in range(self.n_eq):
                    out.append("    %s = %s;\n" % (self.T(n, i), self.sv(i, n)))
            if n < self.order:
                for k in range(n_ops):
                    out.append("    %s = %s;\n" % (self.T(n, self.n_eq + k), self.diff(k, n)))
            out.append("}\n\n")
        out.append("/* Orders 0..p-1 of every u variable, order p of the state variables; T[0 .. n_eq) holds the state. */\n")
        out.append("void cg_jet(v *T, const v *par, const v *tm)\n{\n")
        for n in range(self.order + 1):
            out.append("    cg_order_%d(T, par, *tm);\n" % n)
        out.append("}\n")
        out.append("unsigned cg_width(void) { return %d; }\nunsigned cg_n_uvars(void) { return %d; }\n"
                   "unsigned cg_order(void) { return %d; }\n" % (W, self.n_uvars, self.order))
        return "".join(out)


def _pow_algo_of(e):
    if e == int(e) if abs(e) < 1e9 else False:
        if 0 <= e <= 16:
            return (POW_POS_INT << 8) | int(e)
        if e < 0 and -e <= 16:
            return (POW_NEG_INT << 8) | int(-e)
    elif abs(e) < 1e9:
        y = 2 * e
        if y == int(y):
            if 0 <= y <= 16:
                return (POW_POS_HALF << 8) | int(y)
            if y < 0 and -y <= 16:
                return (POW_NEG_HALF << 8) | int(-y)
    return POW_GENERAL << 8


_PRELUDE = r"""/* GENERATED by oracle/codegen.py -- the batch Taylor jet of one ODE system as straight-line SIMD code. */
#include <math.h>
typedef double v __attribute__((vector_size(%(B)d), aligned(8)));
#define W %(W)d
#define S(x) (((v){0}) + (x))
#define ZERO ((v){0})
#define ONE S(1.0)
static inline v v_dbl(v x) { return x + x; }
#define MAP1(name, fn) static inline v name(v x) { v r; for (int l = 0; l < W; ++l) r[l] = fn(x[l]); return r; }
static inline double sigmoid_d(double x) { return 1. / (1. + exp(-x)); }
MAP1(v_sqrt, sqrt) MAP1(v_sin, sin) MAP1(v_cos, cos) MAP1(v_tanh, tanh) MAP1(v_exp, exp) MAP1(v_log, log)
MAP1(v_sigmoid, sigmoid_d)
static inline v v_pow(v x, v y) { v r; for (int l = 0; l < W; ++l) r[l] = pow(x[l], y[l]); return r; }
static inline v v_relu(v x0, v val, v slope)
{
    v r;
    for (int l = 0; l < W; ++l) r[l] = x0[l] > 0. ? val[l] : (slope[l] == 0. ? 0. : slope[l] * val[l]);
    return r;
}
static inline v v_relup(v x, v slope) { v r; for (int l = 0; l < W; ++l) r[l] = x[l] > 0. ? 1. : slope[l]; return r; }

"""


def _isa_flag():
    """-march of the host: the generated code is compiled where it runs."""
    return "-march=native"


def _host_id():
    """Identity of the host CPU (model + ISA flags): part of the cache key, because -march=native objects built in one
    container must not be picked up on a different machine (oracle/_build/ travels with the repository snapshot)."""
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        keep = [ln for ln in txt.splitlines() if ln.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("\n".join(keep).encode()).hexdigest()[:8]
    except OSError:
        return "unknown"


def build(P, W=8, march=None, opt="-O2", verbose=False):
    """Generate + compile the jet of program P for W lanes. Returns the path of the shared object (cached by content)."""
    src = _Gen(P).source(W)
    march = march or _isa_flag()
    tag = hashlib.sha1((src + march + opt + _host_id()).encode()).hexdigest()[:16]
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "cg_%s_w%d.so" % (tag, W))
    if not os.path.exists(so):
        cfile = os.path.join(BUILD, "cg_%s_w%d.c" % (tag, W))
        with open(cfile, "w") as f:
            f.write(src)
        cmd = ["gcc", "-std=gnu11", opt, march, "-ffp-contract=fast", "-fno-math-errno", "-fPIC", "-shared", cfile, "-o",
               so + ".tmp", "-lm"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        os.replace(so + ".tmp", so)
    return so


class Jet:
    """A compiled jet, installed into the oracle's driver (oracle_set_jet_w{W})."""

    def __init__(self, P, W=8, **kw):
        self.P, self.W = P, W
        self.path = build(P, W, **kw)
        self.lib = C.CDLL(self.path)
        assert self.lib.cg_width() == W and self.lib.cg_n_uvars() == P.n_uvars and self.lib.cg_order() == P.order
        self.fn = C.cast(self.lib.cg_jet, C.c_void_p)

    def install(self, oracle_lib):
        getattr(oracle_lib, "oracle_set_jet_w%d" % self.W)(self.fn)

    @staticmethod
    def uninstall(oracle_lib, W):
        getattr(oracle_lib, "oracle_set_jet_w%d" % W)(C.c_void_p(0))
