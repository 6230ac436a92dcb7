"""Differential test of the emitter's algebra (odexp_b200/cudagen.py: function inlining, pow / division / reciprocal-power
peepholes, ternary -> select, the guarded-product form odx_zmul, hoisting of particle-only factors out of the %C sum).

Those rewrites work on expression text.  What pins them here: a few hundred random model expressions (nested calls, mixed
precedence, unary minus, chained divisions, guards) are emitted twice -- with every peephole on (what the GPU runs) and with all of
them off (the expression as the reference's generator would paste it into model.c, src/odexp:905-1043) -- and BOTH generated
OdxModel structs are compiled for the host with g++ and evaluated on the same random particles, including pairs with equal
coordinates (the guards).  The two must agree to a few ulp: a peephole that mis-parses precedence or captures the wrong operand
shows up as an O(1) difference.  No GPU, no CUDA: the generated model code is `__host__ __device__` by construction."""
import os, random, re, subprocess, sys, tempfile
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEXPR_C, NEXPR_A = 160, 120

FUNS = """fun f1(u,v) = (u-v)/pow(u*u+v*v+0.3,1.5);
fun g1(u) = u*u+1.0;
fun nc(a,b) = pow(a*a+b*b+0.2,1.5);
fun gr(p,q,a,b) = (p == q) ? 0.0 : (p-q)/nc(a,b);
fun h2(u,v) = (u > v) ? u-v : v*0.5;
"""


class Gen:
    """random C expressions over the atoms of one model section; denominators and pow bases are positive by construction"""

    def __init__(self, rng, atoms):
        self.r, self.atoms = rng, atoms

    def num(self):
        return self.r.choice(["0.5", "2.0", "1.25", "3", "0.1", "1e-1", "2.5e0", "7"])

    def atom(self):
        return self.r.choice(self.atoms + [self.num()])

    def sq(self, d):
        e = self.expr(d - 1) if d > 0 else self.atom()
        return "(%s)*(%s)" % (e, e)

    def posden(self, d):
        k = self.r.randrange(6)
        d = d - 1
        if d < 0:
            return "(1.0+%s)" % self.sq(0)
        if k == 0:
            return "pow(%s+%s,1.5)" % (self.sq(d), self.num())
        if k == 1:
            return "g1(%s)" % self.expr(d)
        if k == 2:
            return "(1.0+%s)" % self.sq(d)
        if k == 3:
            return "sqrt(%s+%s)" % (self.sq(d), self.num())
        if k == 4:
            return "nc(%s,%s)" % (self.expr(d), self.expr(d))
        return "pow(%s+%s,0.5)" % (self.sq(d), self.num())

    def factor(self, d):
        k = self.r.randrange(12) if d > 0 else self.r.randrange(3)
        if k <= 1:
            return self.atom()
        if k == 2:
            return "(-%s)" % self.atom()
        if k == 3:
            return "(%s)" % self.expr(d - 1)
        if k == 4:
            return "f1(%s,%s)" % (self.expr(d - 1), self.expr(d - 1))
        if k == 5:
            a, b = self.atom(), self.atom()
            return "((%s == %s) ? 0.0 : (%s-%s)/%s)" % (a, b, a, b, self.posden(d - 1))
        if k == 6:
            return "gr(%s,%s,%s,%s)" % (self.atom(), self.atom(), self.expr(d - 1), self.expr(d - 1))
        if k == 7:
            return "pow(%s,%s)" % (self.sq(d - 1) + "+0.5", self.r.choice(["2", "3", "1.5", "0.5", "2.7"]))
        if k == 8:
            return "h2(%s,%s)" % (self.expr(d - 1), self.expr(d - 1))
        if k == 10:                                          # sin / cos of one argument: hoisted into one odx_sincos by the emitter
            # arguments of moderate size: at 1e9 a rounding error of the argument alone is worth 1e-7 in the cosine
            a = "%s*%s+%s" % (self.atom(), self.atom(), self.atom())
            b = "%s-%s" % (self.atom(), self.atom())
            return self.r.choice(["sin(%s)" % a, "cos(%s)" % a, "(sin(%s)*cos(%s))" % (a, a), "(cos(%s)-sin(%s))" % (a, b)])
        if k == 11:
            return "exp(-(%s))" % self.sq(d - 1)
        return "((%s > %s) ? %s : %s)" % (self.expr(d - 1), self.expr(d - 1), self.expr(d - 1), self.expr(d - 1))

    def term(self, d):
        fs = [self.factor(d) for _ in range(self.r.randrange(1, 4))]
        t = fs[0]
        for f in fs[1:]:
            t += self.r.choice(["*", " * "]) + f
        for _ in range(self.r.randrange(0, 3) if d > 0 else 0):              # chained divisions: (a / f) / g
            t += self.r.choice(["/", " / "]) + self.posden(d)
            if self.r.random() < 0.3:
                t += "*" + self.factor(d)
        return t

    def expr(self, d):
        if d < 0:
            return self.atom()
        n = self.r.randrange(1, 3) if d > 0 else 1
        e = ("-" if self.r.random() < 0.2 else "") + self.term(d)
        for _ in range(n - 1):
            e += self.r.choice(["+", "-", " + ", " - "]) + self.term(d)
        return e

    def int_division(self):
        """integer divisions of the reference's C (ID and POP_SIZE are ints): a reciprocal peephole must leave them alone"""
        return self.r.choice(["ID/(2)", "7/(ID+1)", "1/POP_SIZE", "(ID+3)/(POP_SIZE)", "ID/(POP_SIZE-1)", "(2*ID+1)/(2)", "ID/(1+ID/(2))",
                              "7/(n)", "(ID+1)/(n)", "(n+4)/(2)", "7/(n+p)",
                              "((ID > 2) ? 3 : 1)/2", "((ID > 4) ? ID : 1)/(2)", "((n > 2) ? 7 : p)/2",      # a conditional of two ints is an int          # `par n 3 {int}` is an int in the generated code, p a double
                              "3/(2.0+ID)", "ID/(%s+1.5)" % self.r.choice(self.atoms), "POP_SIZE/(ID+1)*%s" % self.atom()])

    def gravity_like(self):
        """the exact shape the N-body example has: -c*MU(p)*gr(MY(v), OY(v), differences...)"""
        v = self.r.choice(["x", "y"])
        return "%s*MU(p)*gr(MY(%s),OY(%s),MY(x)-OY(x),MY(y)-OY(y))" % (self.r.choice(["-2.0", "0.5", "-MU(q)"]), v, v)


def _model_text(seed):
    rng = random.Random(seed)
    gc = Gen(rng, ["MY(x)", "OY(x)", "MY(y)", "OY(y)", "MU(p)", "MU(q)", "ME(w)", "OE(w)"])
    ga = Gen(rng, ["x", "y", "p", "q", "w", "t"])
    L = ["par p 0.7", "par q 1.3", "par n 3 {int}", FUNS, "expr w 0.5+rand01()"]
    cexprs, aexprs = [], []
    for k in range(NEXPR_C):
        e = gc.gravity_like() if k % 8 == 0 else gc.expr(2)
        cexprs.append(e)
        L.append("%%C C%d %s" % (k, e))
    for k in range(NEXPR_A):
        e = ga.expr(2) if k % 6 else "%s*(%s) + %s" % (ga.atom(), ga.int_division(), ga.int_division())
        aexprs.append(e)
        L.append("aux A%d %s" % (k, e))
    L += ["dx/dt = -x + C0", "dy/dt = -y", "init x rand01()", "init y rand01()", "timespan 0 1", "opt popmode population", "opt popsize 4"]
    return "\n".join(L) + "\n", cexprs, aexprs


HARNESS = r"""
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "odx.h"
#include "odx_rng.h"
#include "odx_model_prelude.cuh"
struct Ctx { const double *mu; double pop_size, h; int id, oid, atbirth, atrepli, ist0, has_sister, stop, slow;
             const double *sis_y, *sis_e, *sis_aux; };
namespace fastm {
%(fast)s
}
namespace refm {
%(ref)s
}
int main() {
  const int NC = %(nc)d, NA = %(na)d, NPT = 240;
  double mu[3] = {0.7, 1.3, 3.0};
  srand(12345);
  for (int it = 0; it < NPT; it++) {
    double y[2], oy[2], e[1], oe[1], aux[NA + 1], oaux[NA + 1];
    for (int k = 0; k < 2; k++) { y[k] = 4.0 * rand() / RAND_MAX - 2.0; oy[k] = 4.0 * rand() / RAND_MAX - 2.0; }
    e[0] = 0.5 + rand() / (double)RAND_MAX; oe[0] = 0.5 + rand() / (double)RAND_MAX;
    if (it %% 5 == 1) oy[0] = y[0];                  /* guards: equal coordinates, then the self pair */
    if (it %% 5 == 2) { oy[0] = y[0]; oy[1] = y[1]; oe[0] = e[0]; }
    if (it %% 7 == 3) oy[1] = y[1];
    const double t = 0.01 * it;
    Ctx c = {mu, 4.0, 0.1, 3 + it %% 7, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double pf[NC], pr[NC], sc[NC];
    for (int k = 0; k < NC; k++) { pf[k] = 0.0; pr[k] = 0.0; sc[k] = 1.0; }
    fastm::OdxModel::cpl_terms<true>(c, t, y, e, aux, oy, oe, oaux, pf);
    fastm::OdxModel::cpl_scale<true>(c, t, y, e, aux, sc);
    refm::OdxModel::cpl_terms<true>(c, t, y, e, aux, oy, oe, oaux, pr);
    for (int k = 0; k < NC; k++) printf("C %%d %%d %%.17g %%.17g\n", it, k, pf[k] * sc[k], pr[k]);
    double psi[NC], mfd[1] = {0.0}, f[2], af[NA + 1], ar[NA + 1], d = 0.0, r = 0.0;
    for (int k = 0; k < NC; k++) psi[k] = 0.0;
    fastm::OdxModel::rhs<true>(c, t, y, e, psi, mfd, f, af, d, r);
    refm::OdxModel::rhs<true>(c, t, y, e, psi, mfd, f, ar, d, r);
    for (int k = 0; k < NA; k++) printf("A %%d %%d %%.17g %%.17g\n", it, k, af[k], ar[k]);
  }
  return 0;
}
"""


def _strip(cu):
    """the generated translation unit without its includes and without the engine at its end"""
    return "\n".join(ln for ln in cu.splitlines() if not ln.startswith("#include"))


@pytest.mark.parametrize("seed", [1, 2])
def test_peepholes_preserve_the_value_of_random_expressions(seed):
    from odexp_b200 import popfile, cudagen
    text, cexprs, aexprs = _model_text(seed)
    model = popfile.parse(text, name="fuzz")
    fast = cudagen.Emitter(model, fast_pow=True, fast_div=True).emit()
    ref = cudagen.Emitter(model, fast_pow=False, fast_div=False).emit()
    assert fast.count("odx_zmul<") >= NEXPR_C // 8          # the pattern the N-body example relies on did fire
    assert "odx_rcp(" in fast and "odx_sel(" in fast and "odx_pow15(" in fast
    assert "odx_rcp(" not in ref and "odx_zmul<" not in ref
    src = HARNESS % dict(fast=_strip(fast), ref=_strip(ref), nc=NEXPR_C, na=NEXPR_A)
    with tempfile.TemporaryDirectory() as td:
        cpp, exe = os.path.join(td, "fuzz.cpp"), os.path.join(td, "fuzz")
        open(cpp, "w").write(src)
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "include"),
                            "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), "-x", "c++", cpp, "-o", exe, "-lm"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-4000:]
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    bad, loose, total = [], 0, 0
    for ln in out.splitlines():
        kind, it, k, a, b = ln.split()
        a, b = float(a), float(b)
        total += 1
        if (np.isnan(a) and np.isnan(b)) or a == b:
            continue
        rel = abs(a - b) / max(abs(a), abs(b), 1e-300)
        # a peephole changes roundings by an ulp or two (x * sqrt(x) for pow(x, 1.5), a reciprocal for a quotient); differences of
        # nearly equal sub-expressions amplify that, so the bound is 1e-8 for every value and 1e-12 for all but a few per thousand.
        # A mis-parsed expression is wrong by O(1).
        loose += rel > 1e-12
        if not rel <= 1e-8:
            bad.append((kind, int(it), (cexprs if kind == "C" else aexprs)[int(k)], a, b))
    assert total == 240 * (NEXPR_C + NEXPR_A)
    assert not bad, "peepholes changed %d values, e.g. %r" % (len(bad), bad[:3])
    assert loose <= 0.003 * total, "%d of %d values differ by more than 1e-12 relative" % (loose, total)
