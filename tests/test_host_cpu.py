"""CPU-side tests: model-file parser, CUDA emitter, C-ABI export table, libc rand restatement."""
import ctypes, os, re, subprocess, sys
import numpy as np
import pytest
from conftest import ROOT, fixture_path


def test_parse_kuramoto():
    from odexp_b200 import popfile
    m = popfile.parse_file(fixture_path("kuramoto"))
    assert (m.nvar, m.nexpr, m.naux, m.npsi, m.nmfd, m.npar) == (1, 1, 2, 0, 3, 1)
    assert m.options["solver"] == "rkck" and m.options["popsize"] == "1000" and m.options["res"] == "801"
    assert m.timespan == [0.0, 160.0]
    assert [e.name for e in m.mfds] == ["A1", "A2", "OMEGA"]


def test_parse_nbody_and_sde():
    from odexp_b200 import popfile
    m = popfile.parse_file(fixture_path("nbody"))
    assert (m.nvar, m.npsi, m.npar) == (6, 3, 1) and len(m.funs) == 2
    assert "impl" in m.pars[0].attr
    s = popfile.parse_file(fixture_path("vdp_sde"))
    assert s.birth == "0.5" and s.death == "MU(d)" and s.repli == "b*(x>0)"
    assert [p.name for p in s.pars] == ["mu", "sigma", "d", "b"]


def test_parse_features():
    from odexp_b200 import popfile
    txt = """
    let R 4
    a 1.5 {impl}     # naked parameter
    par b 2; c 3
    expr U[i=0:R] i*rand01()
    dx/dt = -a*x {init = 1.0}
    opt ps 7; meth dde
    times 0 2
    """
    m = popfile.parse(txt)
    assert [p.name for p in m.pars] == ["a", "b", "c"] and m.exprs[0].length == 4
    assert m.inits[0].rhs == "1.0" and m.options["popsize"] == "7" and m.options["solver"] == "dde"
    with pytest.raises(popfile.PopError):
        popfile.parse("dx/dt = 1\n")           # no initial condition
    with pytest.raises(popfile.PopError):
        popfile.parse("const a[3] 1 2 3\ndx/dt=1\ninit x 0\n")


def test_emit_translates_macros():
    from odexp_b200 import popfile, cudagen
    cu = cudagen.emit_cuda(popfile.parse_file(fixture_path("kuramoto")))
    # sin / cos of the same argument are hoisted into one odx_sincos
    assert "odx_sincos<ODX_EXACT>(odx_c, (odx_y[0]), &odx_sn0, &odx_cs0);" in cu and "odx_m[0] = odx_sn0;" in cu and "odx_m[1] = odx_cs0;" in cu
    assert "odx_m[2] = (odx_e[0]);" in cu and cu.count("odx_sincos<ODX_EXACT>(") == 2
    assert "NVAR = 1, NEXPR = 1, NAUX = 2, NPSI = 0, NMFD = 3" in cu
    nb = cudagen.emit_cuda(popfile.parse_file(fixture_path("nbody")))
    # pow(x,1.5) -> x*sqrt(x) in the emitted helper function; inlined + 1/pow(x,1.5) -> rsqrt(x)^3 in the %C terms
    assert "odx_pow15(" in nb and "odx_sel(" in nb
    # %C terms: guard + product -> odx_zmul, 1/pow(x,1.5) -> rsqrt-seeded x^(-3/2), particle-independent factors hoisted out of the sum
    assert "odx_zmul<ODX_EXACT>(odx_c, " in nb and "odx_rpow15_fast<ODX_EXACT>(" in nb
    assert "CPL_SCALED = true" in nb and "CPL_POISON = true" in nb and "odx_s[0] = (-G) * ((odx_c.mu[0]));" in nb
    assert "CPL_SCALED = false" in cu and "CPL_POISON = false" in cu
    assert "pow(" not in nb.replace("odx_pow", "").replace("odx_rpow", "")
    sde = cudagen.emit_cuda(popfile.parse_file(fixture_path("vdp_sde")))
    assert "odx_c.dwdt()" in sde and "odx_c.atbirth" in sde and "HAS_BIRTH = true" in sde


def test_emit_vect_meanfields():
    """%M name rhs {vect}: assigned once per RHS evaluation (src/odexp:886-889), never summed over the particles"""
    from odexp_b200 import popfile, cudagen
    m = popfile.parse_file(fixture_path("vect_meanfield"))
    assert [cudagen.is_vect(e) for e in m.mfds] == [False, True, True]
    cu = cudagen.emit_cuda(m)
    assert "HAS_VECT_MFD = true" in cu and "odx_mfd[1] = " in cu and "odx_mfd[2] = " in cu
    assert "odx_m[1] = 0.0;" in cu and "odx_m[2] = 0.0;" in cu          # no per-particle term
    assert "HAS_VECT_MFD = false" in cudagen.emit_cuda(popfile.parse_file(fixture_path("kuramoto")))
    bad = popfile.parse("%M T MY(x) {vect}\ndx/dt = -x\ninit x 1\n")
    with pytest.raises(popfile.PopError):
        cudagen.emit_cuda(bad)


def test_const_arrays():
    """`decl const double a[] = {..}` (src/odexp:660-674) and the shipped example's older `const a[] {..}` spelling"""
    from odexp_b200 import popfile, cudagen
    m = popfile.parse_file(fixture_path("nbody_const"))
    assert [c[1] for c in m.consts] == ["x_ic", "y_ic", "z_ic", "vx0", "vy0", "vz0"] and m.consts[0][0] == "double"
    cu = cudagen.emit_cuda(m)
    assert "constexpr double x_ic[] = { 0.0, 1.0, 0.5 };" in cu and "odx_y[3] = v0*vx0[(odx_c.id)];" in cu
    assert "const double v0 = odx_c.mu[0];" in cu                   # {initial}: attribute prefixes count (src/odexp:1269)
    old = popfile.parse("const a[] { 1, 2 }\ndecl const int k = 3\ndx/dt = a[ID] + k\ninit x 0\n")
    assert old.consts == [("double", "a", "[]", "{ 1, 2 }"), ("int", "k", "", "3")]
    with pytest.raises(popfile.PopError):
        popfile.parse("decl double buffer[10]\ndx/dt = 1\ninit x 0\n")        # only constants
    tab = popfile.parse_file(fixture_path("table_drive"))           # FILE table: 4 rows x 3 columns after one skipped line
    assert tab.consts == [("double", "tab", "[4][3]", "{{0.0,1.0,0.1}, {0.5,0.8,0.2}, {1.0,0.6,0.3}, {1.5,0.4,0.4}}")]
    assert "constexpr double tab[4][3] = " in cudagen.emit_cuda(tab)
    with pytest.raises(popfile.PopError):
        popfile.parse("file tab no_such_file.csv\ndx/dt = 1\ninit x 0\n")
    lag = popfile.parse_file(fixture_path("lag_chain"))              # LAG -> chain of hidden variables + aux + inits, in place
    assert lag.flat_names(lag.eqs) == ["xd_chain[%d]" % i for i in range(5)] + ["x"] and [a.name for a in lag.auxs] == ["xd"]
    assert ("xd_end", "5-1") in lag.macros and "linchaindelay(x,xd_chain,i,tau,5)" in cudagen.emit_cuda(lag)
    arr = popfile.parse("let NN 3\ndX[i=0:NN]/dt = -X[i] {init = 0.5*(i+1)}\n")      # {init = ...} on an array equation
    assert arr.nvar == 3 and arr.inits[0].rhs == "0.5*(i+1)" and arr.inits[0].length == 3
    assert "odx_y[0 + i - 0] = 0.5*(i+1);" in cudagen.emit_cuda(arr)
    eq = popfile.parse("decoli/dt = -ecoli\ninit ecoli 1\n")                    # a variable whose d/dt keyword starts with `dec`
    assert [e.name for e in eq.eqs] == ["ecoli"] and not eq.consts


def test_function_blocks_and_arrays():
    from odexp_b200 import popfile, cudagen
    m = popfile.parse_file(fixture_path("fun_block"))
    assert [(f.name, f.block) for f in m.funs] == [("mean", False), ("wmean", True), ("clip", True), ("fill", True)]
    cu = cudagen.emit_cuda(m)
    assert "ODX_HD double wmean(double *x, double *w) {" in cu and "for (j = 0; j < 3; j++)" in cu
    assert "for (size_t i = 0; i < 3; ++i) odx_f[0 + i - 0] = " in cu and "clip(MBAR, -b, b)" in cu          # called, not inlined
    assert "double pw[3]; fill(pw, m, 3);" in cu and "double c[3]; fill(c, 0.25, 3);" in cu                 # {vect}: filled by a function
    c = popfile.parse_file(fixture_path("chain_array"))
    assert (c.nvar, c.naux, c.nexpr) == (6, 6, 4) and c.flat_names(c.eqs) == ["z", "X[0]", "X[1]", "X[2]", "X[3]", "u"]
    with pytest.raises(popfile.PopError):
        popfile.parse("fun f(x)\n return x;\ndx/dt = f(x)\ninit x 0\n")                     # no `end`
    with pytest.raises(popfile.PopError):
        popfile.parse("dX[i=0:3]/dt = -X[i]\ninit X[i=0:2] 1\n")                             # ranges differ


def test_emit_lowrank_variants():
    from odexp_b200 import popfile, cudagen
    cu = cudagen.emit_cuda(popfile.parse_file(fixture_path("kuramoto_lrexp")))
    assert "LR0_RANK = 1" in cu and "LR0_UNWEIGHTED = true" in cu and "LR0_VAR_KIND = 0" in cu
    cu = cudagen.emit_cuda(popfile.parse_file(fixture_path("kuramoto_lowrank_aux")))
    assert "LR0_RANK = 3" in cu and "LR0_UNWEIGHTED = false" in cu and "LR0_VAR = 1" in cu and "LR0_VAR_KIND = 1" in cu
    with pytest.raises(popfile.PopError):
        cudagen.emit_cuda(popfile.parse("expr U[i=0:2] 1\nexpr V[i=0:2] 1\n%C P lrwexp {var = nope, fun = sin, U = U, V = V, rank = 2}\ndx/dt = P\ninit x 0\n"))


@pytest.mark.parametrize("name", ["kuramoto", "kuramoto_lowrank", "Nbody", "vanDerPol"])
def test_shipped_examples_emit_as_shipped(name):
    """every example the reference ships parses and emits unchanged (build container only: /root/reference does not travel)"""
    import os
    from odexp_b200 import popfile, cudagen
    path = os.path.join(os.environ.get("ODX_REFERENCE", "/root/reference"), "examples", name + ".pop")
    if not os.path.isfile(path):
        pytest.skip("reference tree not present")
    m = popfile.parse_file(path)
    cu = cudagen.emit_cuda(m)
    assert "struct OdxModel" in cu and "NVAR = %d" % m.nvar in cu
    if name == "Nbody":
        assert len(m.consts) == 6 and "x_ic[(odx_c.id)]" in cu and m.macros == [("G", "6.674e-11")]      # comment on the macro line dropped
    if name == "kuramoto_lowrank":
        assert "LR0_VAR_KIND = 1" in cu and "LR0_RANK = 1" in cu                                          # var = theta is an AUX; U = U[0]


def test_every_fixture_emits_and_is_registered():
    """every .pop under models/ parses, emits, has a restatement in oracle/models.c (port.DIMS agrees with the parser's
    dimensions) and is built by __graft_entry__.build()"""
    import glob, os
    from odexp_b200 import popfile, cudagen
    from oracle import port
    import __graft_entry__ as ge
    from conftest import FIXTURES
    names = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(FIXTURES, "*.pop")))
    assert len(names) >= 18
    for n in names:
        m = popfile.parse_file(fixture_path(n))
        assert "struct OdxModel" in cudagen.emit_cuda(m)
        assert n in port.DIMS, "no oracle dimensions for fixture %s" % n
        assert port.DIMS[n] == dict(nvar=m.nvar, naux=m.naux, npsi=m.npsi, nmfd=m.nmfd, nexpr=m.nexpr), n
        assert n in ge.MODELS, "fixture %s is not built by __graft_entry__.build()" % n


def test_ctypes_mirror_matches_the_c_structs(tmp_path):
    """the ctypes mirror of the C-ABI structs (odexp_b200/capi.py) has the layout gcc gives include/odx.h"""
    import ctypes as C, subprocess
    from conftest import ROOT
    from odexp_b200 import capi
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "odx.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(odx_solver_opts), offsetof(odx_solver_opts, outdir), offsetof(odx_solver_opts, mu), offsetof(odx_solver_opts, quick), '
                   'offsetof(odx_solver_opts, particle), sizeof(odx_solver_report), offsetof(odx_solver_report, final_n)); '
                   'printf("%zu %zu\\n", offsetof(odx_solver_report, hmin_reached), offsetof(odx_solver_report, hmin_t)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout.split()]
    O, R = capi.SolverOpts, capi.SolverReport
    assert got == [C.sizeof(O), O.outdir.offset, O.mu.offset, O.quick.offset, O.particle.offset, C.sizeof(R), R.final_n.offset,
                   R.hmin_reached.offset, R.hmin_t.offset]


def test_emit_rejects_unsupported():
    from odexp_b200 import popfile, cudagen
    # OA() of an auxiliary that reads a mean field (directly or through another auxiliary) or a coupling: its value for the other
    # particle depends on the list order in the reference
    for txt in ("%M XM MY(x)\naux a x*XM\n%C P OA(a)\ndx/dt = P\ninit x 0\n",
                "%M XM MY(x)\naux a x*XM\naux b 2*a\n%C P OA(b)\ndx/dt = P\ninit x 0\n",
                "aux a P\n%C P OA(a)\ndx/dt = P\ninit x 0\n",
                "aux a x\n%M XM OA(a)\ndx/dt = XM\ninit x 0\n",
                "aux a x\n%C P SA(a)\ndx/dt = P\ninit x 0\n"):
        with pytest.raises(popfile.PopError):
            cudagen.emit_cuda(popfile.parse(txt))
    # draws inside the population sums: serial libc draws in list order in the reference, no random stream in the parallel sums here
    for txt in ("%C P rand01()*sin(OY(x)-MY(x))\ndx/dt = P\ninit x 0\n", "%M XM MY(x)*randN(0,1)\ndx/dt = XM\ninit x 0\n",
                "%BIRTH 0.1*exprand(1.0)\ndx/dt = -x\ninit x 0\n"):
        with pytest.raises(popfile.PopError, match="random numbers inside"):
            cudagen.emit_cuda(popfile.parse(txt))
    # MA() inside AUX / D/DT: the previous evaluation's value in the reference (myself_->aux[] is rewritten at the end of ode_rhs)
    for txt in ("aux a0 x\naux a1 MA(a0)\ndx/dt = -x\ninit x 1\n", "aux a0 x\ndx/dt = -x + MA(a0)\ninit x 1\n"):
        with pytest.raises(popfile.PopError, match="read by its name"):
            cudagen.emit_cuda(popfile.parse(txt))
    cudagen.emit_cuda(popfile.parse("aux a0 x\ndx/dt = -x\ninit x 1\n%DEATH 0.1*MA(a0)\n%M AM MA(a0)\n"))      # rates and sums: the current values
    ok = cudagen.emit_cuda(popfile.parse("expr q 1\naux a q*x\n%C P OA(a)*MA(a)\ndx/dt = P\ninit x 0\n"))
    assert "CPL_USES_OAUX = true" in ok and "CPL_USES_AUX = true" in ok and "CPL_USES_EXPR = true" in ok


def test_unicode_identifiers_are_mangled():
    from odexp_b200 import popfile, cudagen
    cu = cudagen.emit_cuda(popfile.parse("par ω 1\ndθ/dt = ω\ninit θ 0\n"))
    assert "ω" not in cu.split("struct OdxModel")[1].split("info()")[0]


def test_library_exports_full_abi(model_lib):
    """the C-ABI library loads without a GPU and exports every symbol include/odx.h declares"""
    from odexp_b200 import capi
    lib = model_lib("kuramoto")
    hdr = open(os.path.join(ROOT, "include", "odx.h")).read()
    declared = set(re.findall(r"\b(odx_[a-z0-9_]+)\s*\(", hdr)) - {"odx_handle"}
    assert declared == set(capi.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(lib.lib, s), s
    assert lib.dims() == dict(nvar=1, nexpr=1, naux=2, npsi=0, nmfd=3)


def test_library_exports_reference_entry_points(model_lib):
    """every model library also exports the reference's own stepper entry points by name (src/methods_odexp.h:47-63), as declared
    in include/odx_compat.h; ncumsum needs no device and is checked against its definition (src/methods_odexp.c:1764-1777)"""
    lib = model_lib("kuramoto")
    hdr = open(os.path.join(ROOT, "include", "odx_compat.h")).read()
    decl = hdr.split("} odx_gsl_system;")[1].split("\n", 2)[2]        # the prototypes after the two typedefs
    decl = re.sub(r"/\*.*?\*/", "", decl, flags=re.S)
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", decl)) - {"defined", "visibility"}
    assert declared == {"odx_compat_bind", "odx_compat_last_rows", "fe_apply", "iteration_apply", "dde_apply", "SSA_timestep",
                        "SSA_apply_birthdeath", "tau_leaping_apply_birthdeath", "ncumsum"}
    for s_ in declared:
        assert hasattr(lib.lib, s_), s_
    x = (ctypes.c_double * 5)(0.5, 1.5, 0.0, 2.0, 1.0)
    tot = ctypes.c_double()
    assert lib.lib.ncumsum(x, 5, ctypes.byref(tot)) == 0
    assert tot.value == 5.0 and list(x) == [0.1, 0.4, 0.4, 0.8, 1.0]


def test_no_gpu_fails_loudly(model_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from odexp_b200 import capi
    with pytest.raises(capi.OdxError):
        capi.Population(model_lib("kuramoto"), 128)


def test_libc_rand_restatement_matches_this_libc():
    """oracle/rand_hook.c and csrc/odx_rng.h restate glibc rand(); check against the libc we run on"""
    libc = ctypes.CDLL("libc.so.6")
    src = r'''
    #include <stdio.h>
    #include "odx_rng.h"
    int main() { odx::HostRand h; h.seed(3141592u); for (int i = 0; i < 1000; i++) printf("%d\n", h.next()); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(src)
        subprocess.run(["g++", "-O1", "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), os.path.join(d, "t.cpp"), "-o", os.path.join(d, "t")], check=True)
        mine = [int(x) for x in subprocess.run([os.path.join(d, "t")], stdout=subprocess.PIPE, check=True).stdout.split()]
    libc.srand(3141592)
    ref = [libc.rand() for _ in range(1000)]
    assert mine == ref


def test_two_model_libraries_coexist(model_lib):
    """regression: inline statics / same-named symbols of two model libraries must not be merged by the loader"""
    a, b = model_lib("kuramoto"), model_lib("nbody")
    assert a.name == "kuramoto" and b.name == "nbody"
    assert (a.nvar, a.nmfd, a.npsi) == (1, 3, 0) and (b.nvar, b.nmfd, b.npsi) == (6, 0, 3)


def test_antisymmetric_coupling_detection():
    """OdxModel::CPL_SYM (the symmetric all-pairs kernel evaluates a term once and adds it to both particles) is claimed only when the
    emitter can prove term(j, i) = -term(i, j) bit for bit: guarded difference times a factor that reads the particles through squares
    of differences only.  Anything else keeps the ordered kernel."""
    from odexp_b200 import popfile, cudagen
    head = "par m 2.0\npar e 0.1\nexpr w rand01()\nfun nc(a,b) = pow(a*a+b*b+0.01,1.5);\n"
    tail = "dx/dt = P\ndy/dt = -y\ninit x rand01()\ninit y rand01()\n"
    yes = ["%C P MU(m)*((MY(x) == OY(x)) ? 0.0 : (MY(x)-OY(x))/nc(MY(x)-OY(x),MY(y)-OY(y)))",
           "%C P -3.0*((MY(x) == OY(x)) ? 0.0 : (MY(x)-OY(x))/pow((MY(y)-OY(y))*(MY(y)-OY(y))+MU(e)*MU(e),1.5))",
           "%C P ((MY(y) == OY(y)) ? 0.0 : (MY(y)-OY(y))/nc(OY(x)-MY(x),MY(y)-OY(y)))"]        # (oy - y)^2 is even as well
    no = ["%C P OE(w)*((MY(x) == OY(x)) ? 0.0 : (MY(x)-OY(x))/nc(MY(x)-OY(x),MY(y)-OY(y)))",     # the other particle's weight
          "%C P (MY(x) == OY(x)) ? 0.0 : (MY(x)-OY(x))/nc(MY(x)-OY(x),MY(y))",                   # not a function of the difference
          "%C P (MY(x) == OY(x)) ? 0.0 : (MY(x)-OY(x))/nc(MY(x)-OY(x),MY(y)-OY(y))*(MY(y)-OY(y))",  # odd factor: symmetric, not antisymmetric
          "%C P sin(OY(x)-MY(x))",                                                                # antisymmetric, but not in the proven form
          # a guard that is the whole expression stays a C conditional (only guards inside functions / parentheses become selects: a bare
          # `(i == 0) ? a : X[i-1]` must not evaluate its other arm), so it is not recognised either
          "%C P (MY(y) == OY(y)) ? 0.0 : (MY(y)-OY(y))/nc(OY(x)-MY(x),MY(y)-OY(y))"]
    with pytest.raises(popfile.PopError, match="random numbers inside"):       # a draw per ordered pair: refused altogether
        cudagen.emit_cuda(popfile.parse(head + "%C P ((MY(x) == OY(x)) ? 0.0 : (MY(x)-OY(x))/nc(MY(x)-OY(x),MY(y)-OY(y)))*rand01()\n" + tail, name="sym"))
    for line in yes:
        cu = cudagen.emit_cuda(popfile.parse(head + line + "\n" + tail, name="sym"))
        assert "CPL_SYM = true" in cu and "#define ODX_MODEL_CPL_SYM 1" in cu, line
    for line in no:
        cu = cudagen.emit_cuda(popfile.parse(head + line + "\n" + tail, name="sym"))
        assert "CPL_SYM = false" in cu and "ODX_MODEL_CPL_SYM" not in cu, line
    # two couplings: both must qualify
    two = "%C P (MY(x) == OY(x)) ? 0.0 : (MY(x)-OY(x))/nc(MY(x)-OY(x),MY(y)-OY(y))\n%C Q OY(y)-MY(y)\n"
    cu = cudagen.emit_cuda(popfile.parse(head + two + "dx/dt = P + Q\ndy/dt = -y\ninit x rand01()\ninit y rand01()\n", name="sym"))
    assert "CPL_SYM = false" in cu


def test_guarded_array_access_stays_a_conditional():
    """ternary -> select evaluates both arms; an arm that indexes an array with a computed index may rely on the guard to stay in range
    (`(i == 0) ? a : X[i-1]`), so such a ternary is left alone even inside parentheses"""
    from odexp_b200 import popfile, cudagen
    txt = "par k 1\ndX[i=0:3]/dt = ((i == 0) ? -k*X[0] : k*(X[i-1] - X[i]))\ninit X[i=0:3] 1.0\n"
    cu = cudagen.emit_cuda(popfile.parse(txt, name="g"))
    line = [l for l in cu.splitlines() if "odx_f[" in l][0]
    assert "?" in line and "odx_sel" not in line
    txt = "par k 1\nfun pick(a,b) = (a > b) ? a : b;\naux m pick(x, k)\ndx/dt = -m\ninit x 1.0\n"
    assert "odx_sel(" in cudagen.emit_cuda(popfile.parse(txt, name="g2"))         # pure arms: still a select


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the unchanged reference binary on the host cores; needs no GPU): one JSON line with the same metric /
    config as the GPU arm, `impl`, a `cpu_baseline` describing the sample and an `e2e` equal to the value"""
    import json, subprocess, sys
    from conftest import ROOT
    from oracle import refrun
    if not refrun.have("kuramoto"):
        pytest.skip("oracle/_ref not built here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stderr[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("particle-steps/sec") and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 1e5
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "n_particles" in d["config"] and "workload" in d["config"]


def test_model_macros_do_not_reach_the_engine(tmp_path):
    """MACRO lines are #defines (as in the reference's model.c); the engine is compiled in the same translation unit, so a model that
    names a macro T, h or n must not rewrite the engine's identifiers: the generated file #undefs them in front of the engine"""
    import shutil
    from odexp_b200 import popfile, cudagen
    txt = open(fixture_path("kuramoto")).read() + "\nmacro T 1.0\nmacro h 0.5\nmacro n 3\n"
    cu = cudagen.emit_cuda(popfile.parse(txt, name="mac"))
    tail = cu[cu.index("end of the model's macros"):]
    for name in ("PI", "T", "h", "n"):
        assert "#undef %s\n" % name in tail
    assert tail.rstrip().endswith('#include "odx_engine.cuh"')
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not here")
    src = tmp_path / "mac.cu"
    src.write_text(cu)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O0", "-std=c++17", "--expt-relaxed-constexpr", "-diag-suppress", "177,550,2361",
                        "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "mac.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_comment_on_a_parameter_line_ends_at_the_next_semicolon():
    """src/odexp:397 breaks PAR / OPT lines at every ';' before comments are looked at (the manual: "this comment ends at the semi-colon: b is
    another parameter!"); on every other keyword a comment runs to the end of the line"""
    from odexp_b200 import popfile
    m = popfile.parse("let K 4\npar a 1 {int} # an integer; b 2.3; c K # the LET is substituted; d 5\n"
                      "opt h0 0.05 # step; res 3\naux z x # no; split here\ndx/dt = -a*x # neither; here\ninit x 1\n")
    assert [(p.name, p.rhs, p.attr) for p in m.pars] == [("a", "1", "int"), ("b", "2.3", ""), ("c", "4", ""), ("d", "5", "")]
    assert m.options["h0"] == "0.05" and m.options["res"] == "3"
    assert m.auxs[0].rhs == "x" and m.eqs[0].rhs == "-a*x"


def test_mean_fields_cannot_be_read_inside_population_sums():
    """MF() is SIM->meanfield[k] in the reference (src/odexp.h:357-364): inside %M / %C / %BIRTH it would read the state of pop_rhs's own
    loops; rejected with a message.  {vect} mean fields (evaluated once, after the loops, in file order) may read the ones before them."""
    from odexp_b200 import popfile, cudagen
    for line in ("%C P cos(OY(x) - MY(x)) + MF(XM)", "%M X2 MY(x)*MF(XM)", "%BIRTH 0.1*MF(XM)"):
        with pytest.raises(popfile.PopError, match="cannot be read inside"):
            cudagen.emit_cuda(popfile.parse("%M XM MY(x)\n" + line + "\ndx/dt = -x\ninit x 1\n"))
    cu = cudagen.emit_cuda(popfile.parse("%M XM MY(x)\n%M TWICE 2*MF(XM) {vect}\naux a MF(TWICE)\ndx/dt = -x\ninit x 1\n"))
    assert "odx_mfd[1] = 2*(odx_mfd[0]);" in cu


def test_multi_index_arrays(tmp_path):
    """`name[i=a:b][j=c:d]`: stored row-major over the iterated ranges, declared with the end of every range (src/odexp:1288-1340), column
    names `name[i][j]`; the generated file compiles for sm_100a (its results are pinned to the reference binary in
    tests/test_emitter_vs_reference.py::test_documented_syntax_forms_match_the_reference_binary[multi_index])"""
    import shutil
    from odexp_b200 import popfile, cudagen
    txt = ("expr w[i=0:2][j=1:3] 0.1*i + j\naux a[p=1:3][q=0:2][r=0:2] w[p-1][1]*U[p-1][q] + r\ndx/dt = -x + a[2][1][0]\n"
           "dU[i=0:2][j=0:3]/dt = -U[i][j] + ((j > 0) ? U[i][j-1] : x) {init = 0.1*i + j}\ninit x 1\n")
    with pytest.raises(popfile.PopError, match="order of"):
        popfile.parse(txt)                                        # {init = ..} appends in file order: x must come first, as in the reference
    txt = txt.replace("dx/dt = -x + a[2][1][0]\n", "dx/dt = -x + a[2][1][0] {init = 1}\n").replace("init x 1\n", "")
    m = popfile.parse(txt, name="nd")
    assert (m.nvar, m.naux, m.nexpr) == (7, 8, 4)
    assert m.flat_names(m.exprs) == ["w[0][1]", "w[0][2]", "w[1][1]", "w[1][2]"]
    assert m.flat_names(m.eqs)[:4] == ["x", "U[0][0]", "U[0][1]", "U[0][2]"]
    cu = cudagen.emit_cuda(m)
    assert "double w[2][3]; for (size_t i = 0; i < 2; ++i) for (size_t j = 1; j < 3; ++j) w[i][j] = odx_e[0 + ((i - 0)*2 + (j - 1))];" in cu
    assert "odx_aux[0 + (((p - 1)*2 + (q - 0))*2 + (r - 0))] = a[p][q][r];" in cu
    assert "? U[i][j-1] : x" in cu                               # the guard of a computed index stays a conditional
    for bad in ("expr w[i=0:2][i=0:3] 1\ndx/dt = -x\ninit x 1\n", "expr w[i=0:2][3] 1\ndx/dt = -x\ninit x 1\n", "expr w[i=0:2][j=3:3] 1\ndx/dt = -x\ninit x 1\n"):
        with pytest.raises(popfile.PopError):
            popfile.parse(bad)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not here")
    src = tmp_path / "nd.cu"
    src.write_text(cu)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O0", "-std=c++17", "--expt-relaxed-constexpr", "-diag-suppress", "177,550,2361",
                        "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "nd.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_reserved_names_and_minmax_macros(tmp_path):
    """identifiers that are legal in the reference's C model but not in generated C++ / CUDA are renamed consistently (`C` is the template
    parameter of the generated functions; `new`, `class`, `this`, `or` are C++ keywords; a FUN called rsqrt would redefine CUDA's); max(a,b) /
    min(a,b) are the reference's type-generic macros (src/odexp.h:204-212) -- a value-returning template here, a variable called `min` stays
    a variable; the list-pointer macros are refused by name.  Results vs the reference binary: test_documented_syntax_forms...[reserved_names]"""
    import shutil
    from odexp_b200 import popfile, cudagen
    txt = ("par C 2\npar new 0.5\npar class 2\npar min 0.25\nfun rsqrt(a) = 1.0/sqrt(a);\nfun mix(C, this)\n  double or = C*this;\n  return max(or, 1) + 1.0;\nend\n"
           "aux a max(x, C)*new + rsqrt(1.0 + x)*class + max (ID, 1)/2 + min + mix(x, C) + abs(x - 2.6)\ndx/dt = -x + a\ninit x 1\n%M XM min(MY(x), 2)\n")
    cu = cudagen.emit_cuda(popfile.parse(txt, name="resv"))
    assert "const double odx_v_C = (double)odx_c.mu[0];" in cu and "odx_v_new" in cu and "odx_v_class" in cu
    assert "ODX_HD double odx_v_rsqrt(double a)" in cu and "double odx_v_or = odx_v_C*odx_v_this; return odx_max(odx_v_or, 1) + 1.0;" in cu
    assert "odx_max ((odx_c.id), 1)/2 + min + " in cu and "odx_min((odx_y[0]), 2)" in cu
    assert "odx_iabs(x - 2.6)" in cu and "int odx_iabs(int v)" in cu       # C's abs(): abs(-1.6) is 1 in the reference's model.c
    assert "-> decltype((a > b ? a : b) + 0)" in cu                       # a value: decltype of the bare conditional is a reference to an argument
    assert '"C", "new", "class", "min"' in cu                              # the names a user sees stay the model's
    assert "odx_max" not in cudagen.emit_cuda(popfile.parse("dx/dt = -x\ninit x 1\n"))
    for macro in ("NEXT", "FIRST", "EVENT_TYPE", "PRINT"):
        with pytest.raises(popfile.PopError, match="do not exist on the device"):
            cudagen.emit_cuda(popfile.parse("aux a %s\ndx/dt = -x\ninit x 1\n" % (macro + ("(\"x\")" if macro == "PRINT" else ""))))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not here")
    src = tmp_path / "resv.cu"
    src.write_text(cu)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O0", "-std=c++17", "--expt-relaxed-constexpr", "-diag-suppress", "177,550,2361",
                        "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "resv.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_fast_sincos_error_bounds(tmp_path):
    """the two fast sincos forms of the step kernels (csrc/odx_model_prelude.cuh) over their whole range |x| < 105615, against long-double
    libm: the polynomial form (host branch of the same function) and the table-assisted form of the persistent kernels, whose device
    branch tests/aux/sincos_bounds.cpp restates operation for operation on the product's own constants (fma = DFMA; table nodes correctly
    rounded, the device's sincospi gives them to <= 1 ulp).  3 million arguments: small, spread over the range, and next to multiples of
    pi/2 where the reduction cancels.  Both stay below 1 ulp of 1 in absolute error."""
    exe = tmp_path / "sc"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "odexp_b200", "csrc"),
                        os.path.join(ROOT, "tests", "aux", "sincos_bounds.cpp"), "-o", str(exe), "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe), "3000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    table, poly = (float(v) for v in out.stdout.split())
    assert table <= 2.22e-16 and poly <= 2.22e-16, (table, poly)


def test_fast_inverse_cube_error_bound(tmp_path):
    """odx_rpow15_fast (x^(-3/2) of the pair kernels' guarded product, csrc/odx_model_prelude.cuh): one approximate reciprocal square root and a
    second-order correction.  The device branch restated on the CPU (tests/aux/rpow15_bounds.cpp) with the hardware seed replaced by
    x^(-1/2) (1 + d): for any |d| <= 2^-20 the result is within 3e-16 relative (1.4 ulp) of the long-double value over 36 decades of x;
    the bound degrades gracefully (1.3e-15 at 18 bits), so the parity tolerances of the N-body tests do not hinge on the exact seed quality."""
    exe = tmp_path / "rp"
    r = subprocess.run(["g++", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tests", "aux", "rpow15_bounds.cpp"), "-o", str(exe), "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    e20 = float(subprocess.run([str(exe), "20", "1000000"], capture_output=True, text=True, check=True).stdout)
    e18 = float(subprocess.run([str(exe), "18", "1000000"], capture_output=True, text=True, check=True).stdout)
    assert e20 <= 3.0e-16 and e18 <= 1.5e-15, (e20, e18)


def test_device_model_builds_from_the_prescanned_file():
    """INTEGRATION.md section 1 builds the device model from .odexp/equations.pop, the copy of the model file the reference's prescan writes
    (src/odexp:311-402: LET values substituted, PAR / OPT lines broken at ';', LAG lines expanded, `{init = ..}` turned into INIT lines,
    `fun f(a) = return expr;`, multi-line functions glued into one line without `end`).  For every fixture that file must give the same
    generated code as the model file itself (oracle/_ref/<fixture>/.odexp/ is written by the reference's own generator)."""
    import glob
    from odexp_b200 import popfile, cudagen
    dirs = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "*", ".odexp")))
    if not dirs:
        pytest.skip("oracle/_ref not built")
    seen = 0
    for d in dirs:
        name = os.path.basename(os.path.dirname(d))
        src = fixture_path(name) if os.path.exists(os.path.join(ROOT, "models", name + ".pop")) else None
        if src is None:
            continue
        pre = popfile.parse(open(os.path.join(d, "equations.pop")).read(), name=name, base_dir=os.path.join(ROOT, "models"))
        assert cudagen.emit_cuda(pre) == cudagen.emit_cuda(popfile.parse_file(src)), name
        seen += 1
    assert seen >= 20
    # the two prescanned FUN spellings on their own
    m = popfile.parse("fun f(a,b) = return (a > b) ? a : b;\nfun g(*v, n) = int j;double s = 0;for (j = 0; j < n; j++) { s += v[j]; }return s;\n"
                      "dx/dt = -f(x, 1)\ninit x 1\n")
    assert (m.funs[0].body, m.funs[0].block) == ("(a > b) ? a : b", False)
    assert m.funs[1].block and m.funs[1].body == "int j; double s = 0; for (j = 0; j < n; j++) { s += v[j]; } return s;"


def test_sincos_with_a_draw_in_its_argument_is_not_hoisted():
    """sin(A) / cos(A) are hoisted into one odx_sincos(A) in front of the statement; not when A draws a random number: out of a conditional
    arm the draw would be consumed even when the reference does not evaluate the arm, and sin(randN(..))*cos(randN(..)) are two draws"""
    from odexp_b200 import popfile, cudagen
    cu = cudagen.emit_cuda(popfile.parse("aux a (x > 0.5) ? sin(rand01()) : 0.0\naux b cos(x)*sin(x) + sin(randN(0,1))*cos(randN(0,1))\n"
                                         "dx/dt = -x + a + b + sin(DWDT)\ninit x 1\n"))
    assert "const double a = (x > 0.5) ? sin(odx_c.rand01()) : 0.0;" in cu
    assert "const double b = odx_cs0*odx_sn0 + sin(odx_c.randN(0,1))*cos(odx_c.randN(0,1));" in cu
    assert "sin(odx_c.dwdt())" in cu and cu.count("odx_sincos<ODX_EXACT>(odx_c, x,") == 2          # rhs and rhs_inc


def test_function_calls_with_draws_are_not_inlined():
    """one-line model functions are substituted textually so that the peepholes see through them -- not when an argument draws a random
    number: `sq(rand01())` is u*u of ONE draw (call by value), the substituted text would be the product of two"""
    from odexp_b200 import popfile, cudagen
    cu = cudagen.emit_cuda(popfile.parse("fun sq(u) = u*u;\nfun g(p,q) = (p == q) ? 0.0 : (p-q)/sq(q);\n"
                                         "aux a sq(rand01()) + sq(x) + g(x, DWDT) + g(x, 2.0)\ndx/dt = -x + a\ninit x 1\n"))
    line = next(ln for ln in cu.splitlines() if "const double a" in ln)
    assert "sq(odx_c.rand01())" in line and "g(x, odx_c.dwdt())" in line and "((x)*(x))" in line and line.count("rand01") == 1


def test_model_functions_that_read_parameters_or_draw(tmp_path):
    """MU(par), POP_SIZE and rand01() ... inside a FUN: globals / plain C functions in the reference, the evaluation context here -- such a
    function takes the context as its first argument and every call site passes its own; a function that draws (directly or through
    another) makes the model's RHS a drawing one (RHS_RNG), is neither inlined nor evaluated eagerly inside a guard, and stays out of the
    population sums.  Results vs the reference binary: test_documented_syntax_forms...[context_functions]; compiles for sm_100a."""
    import shutil
    from odexp_b200 import popfile, cudagen
    txt = ("par k 2 {impl}\nfun f(u) = u*MU(k)/POP_SIZE;\nfun noise(s) = s*randN(0.0,1.0);\nfun two(u)\n  double a = noise(u);\n  return a + f(u);\nend\n"
           "fun plain(u) = u*u;\naux a (x > 0) ? noise(x) : 0.0\naux b sin(noise(x)) + plain(x)\ndx/dt = -f(x) + two(x) + a + b\ninit x f(1.0)\n%M XM MY(x)*f(2.0)\n")
    cu = cudagen.emit_cuda(popfile.parse(txt, name="ctxf"))
    assert "template <class OdxC> ODX_HD double f(OdxC &odx_c, double u) { (void)odx_c; ODX_CONSTS return u*(odx_c.mu[0])/((int)odx_c.pop_size); }" in cu
    assert "double a = noise(odx_c, u); return a + f(odx_c, u);" in cu and "ODX_HD double plain(double u)" in cu
    assert "const double a = (x > 0) ? noise(odx_c, x) : 0.0;" in cu and "const double b = sin(noise(odx_c, x)) + ((x)*(x));" in cu
    assert "RHS_RNG = true" in cu and "odx_y[0] = f(odx_c, 1.0);" in cu and "(odx_y[0])*f(odx_c, 2.0)" in cu
    with pytest.raises(popfile.PopError, match="random numbers inside"):
        cudagen.emit_cuda(popfile.parse(txt + "%C P noise(1.0)*sin(OY(x)-MY(x))\n"))
    with pytest.raises(popfile.PopError, match="not available inside a model function"):
        cudagen.emit_cuda(popfile.parse("fun g(u) = u*MY(x);\ndx/dt = -g(x)\ninit x 1\n"))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not here")
    src = tmp_path / "ctxf.cu"
    src.write_text(cu)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O0", "-std=c++17", "--expt-relaxed-constexpr", "-diag-suppress", "177,550,2361",
                        "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "ctxf.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_run_arguments_from_the_model_file(model_lib):
    """capi.solver_args / run_pop: a run started from the OPT and TIMESPAN lines of the model file, like `odexp model.pop -a key=value`
    (defaults of src/options.c:6-79, abbreviations ps / meth / r); without a GPU run_pop gets as far as odx_create and fails loudly"""
    from odexp_b200 import capi, popfile
    m = popfile.parse_file(fixture_path("reset_decay"))
    a = capi.solver_args(m)
    assert (a["tspan"], a["meth"], a["res"], a["h0"], a["popsize"], a["abstol"], a["seed"]) == ((0.0, 1.5, 3.25, 6.0), "dde", 25, 0.05, 20, 1e-6, 3141592)
    b = capi.solver_args(m, ps=50, meth="fe", r=11, reltol=1e-9)
    assert (b["popsize"], b["meth"], b["res"], b["reltol"]) == (50, "fe", 11, 1e-9)
    assert capi.solver_args(popfile.parse("dx/dt = -x\ninit x 1\n"))["meth"] == "rkck"            # the reference's default stepper
    with pytest.raises(capi.OdxError, match="popmode single"):
        capi.solver_args(m, popmode="single")
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(capi.OdxError, match="odx_create"):
            capi.run_pop(fixture_path("reset_decay"), ps=8)
