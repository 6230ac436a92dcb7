"""The stage-to-stage forms the emitter generates for models whose RHS and %M terms share a sincos argument (`rhs_inc`,
`mfd_terms_inc`: sin / cos of a stage state from the recorded triple of a nearby state by angle addition, odx_sincos_inc) against the
plain forms (`rhs`, `mfd_terms`), on the host: the generated OdxModel of the Kuramoto fixture is `__host__ __device__`, so g++
compiles it.  The lean RK4 step kernel recomputes stages through these forms; on the GPU they are covered by the
parity tests, here the arithmetic itself is pinned: same value to a few ulp for increments below the bound, the slow flag above it.
Reference: the model's expressions are those of examples/kuramoto.pop (src/odexp:981-1043 lays them out as ode_rhs)."""
import os, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

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
%(model)s
int main() {
  double mu[1] = {0.1};
  srand(7);
  for (int it = 0; it < 4000; it++) {
    const double y0 = 200.0 * rand() / RAND_MAX - 100.0, om = 0.6 + 0.2 * rand() / RAND_MAX;
    const double d = (it %% 10 == 9 ? 0.3 : 0.12) * (2.0 * rand() / RAND_MAX - 1.0);          /* every tenth increment beyond the bound */
    const double mfd[3] = {0.3 * rand() / RAND_MAX, -0.2 * rand() / RAND_MAX, 0.62}, psi[1] = {0.0};
    double ya[1] = {y0}, yb[1] = {y0 + d}, e[1] = {om}, f0[1], f1[1], f2[1], aux0[2], aux1[2], aux2[2], dr = 0, rr = 0;
    Ctx c = {mu, 1000.0, 0.05, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    OdxHint h0, h1, h2;
    OdxModel::rhs<false>(c, 0.0, ya, e, psi, mfd, f0, aux0, dr, rr, &h0);                     /* records (y0, sin y0, cos y0) */
    OdxModel::rhs<false>(c, 0.0, yb, e, psi, mfd, f1, aux1, dr, rr, &h1);                     /* plain form at the shifted state */
    Ctx ci = c;
    OdxModel::rhs_inc<false>(ci, 0.0, yb, e, psi, mfd, f2, aux2, dr, rr, h0, &h2);            /* the same from the triple of y0 */
    double m1[3], m2[3];
    Ctx cm = c;
    OdxModel::mfd_terms<false>(c, 0.0, yb, e, aux1, m1);
    OdxModel::mfd_terms_inc<false>(cm, 0.0, yb, e, aux1, m2, h0);
    printf("%%.17g %%d %%d %%.17g %%.17g %%.17g %%.17g %%.17g %%.17g %%.17g %%.17g\n", d, ci.slow, cm.slow, f1[0], f2[0], m1[0], m2[0], m1[1], m2[1], h1.s[0], h2.s[0]);
  }
  return 0;
}
"""


def test_increment_forms_equal_plain_forms(tmp_path):
    from odexp_b200 import popfile, cudagen
    cu = cudagen.emit_cuda(popfile.parse_file(os.path.join(ROOT, "models", "kuramoto.pop")))
    assert "rhs_inc" in cu and "mfd_terms_inc" in cu and "odx_sincos_inc<" in cu
    model = "\n".join(ln for ln in cu.splitlines() if not ln.startswith("#include"))
    src, exe = tmp_path / "inc.cpp", tmp_path / "inc"
    src.write_text(HARNESS % dict(model=model))
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), "-x", "c++", str(src), "-o", str(exe), "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    rows = np.array([[float(x) for x in ln.split()] for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()])
    d, slow_rhs, slow_mfd = rows[:, 0], rows[:, 1], rows[:, 2]
    inside = np.abs(d) < 0.125
    assert inside.sum() > 3000 and (~inside).sum() > 100
    # beyond the bound the increment form only reports it (the engine then redoes the particle with the library path)
    assert np.all(slow_rhs[~inside] == 1) and np.all(slow_mfd[~inside] == 1)
    assert np.all(slow_rhs[inside] == 0) and np.all(slow_mfd[inside] == 0)
    for a, b in ((3, 4), (5, 6), (7, 8), (9, 10)):            # f, sin term, cos term, the recorded sine of the new triple
        assert np.abs(rows[inside, a] - rows[inside, b]).max() <= 4e-16 * max(1.0, np.abs(rows[inside, a]).max())
