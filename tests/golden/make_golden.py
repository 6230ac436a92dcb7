#!/usr/bin/env python3
"""Generate the committed golden vectors from the UNCHANGED reference binary (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run in the build container only; the .npz files travel.

    python tests/golden/make_golden.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KUR = dict(nvar=1, naux=2, npsi=0, nmfd=3, nexpr=1)
NBD = dict(nvar=6, naux=0, npsi=3, nmfd=0, nexpr=0)
VDP = dict(nvar=2, naux=0, npsi=0, nmfd=0, nexpr=0)
RING = dict(nvar=2, naux=1, npsi=0, nmfd=1, nexpr=1)

FULL = dict(meth="rkck", res=201, hmin=1e-5, h0=0.1, abstol=1e-6, reltol=1e-6, ssahmin=0.0, aleap=0.01, seed=3141592)
CASES = {
    # name: (fixture, opts, timespan, dims)
    "kuramoto_dde_n64": ("kuramoto", {"meth": "dde", "h0": 0.1, "res": 5, "ps": 64}, (0, 0.8), KUR),
    "kuramoto_dde_n1000": ("kuramoto", {"meth": "dde", "h0": 0.1, "res": 11, "ps": 1000}, (0, 2.0), KUR),
    "kuramoto_dde_h003_n257": ("kuramoto", {"meth": "dde", "h0": 0.03, "res": 4, "ps": 257}, (0, 0.5), KUR),  # ragged: h does not divide dt
    "kuramoto_fe_n100": ("kuramoto", {"meth": "fe", "h0": 0.05, "res": 3, "ps": 100}, (0, 0.5), KUR),
    "kuramoto_rkck_n200": ("kuramoto", {"meth": "rkck", "res": 6, "ps": 200}, (0, 5.0), KUR),
    "kuramoto_rkf45_n200": ("kuramoto", {"meth": "rkf45", "res": 6, "ps": 200, "abstol": 1e-9, "reltol": 1e-9}, (0, 5.0), KUR),
    # BASELINE config 1: examples/kuramoto.pop exactly as shipped (N=1000, meth rkck, res 801, t in [0,160]), no overrides
    "kuramoto_shipped_rkck": ("kuramoto_shipped", {}, None, KUR),
    "kuramoto_rk4_n50": ("kuramoto", {"meth": "rk4", "res": 6, "ps": 50}, (0, 5.0), KUR),      # GSL step-doubling rk4
    "nbody_dde_n64": ("nbody", {"meth": "dde", "h0": 0.1, "res": 3, "ps": 64}, (0, 0.4), NBD),
    # short horizon: at T=0.4 a close encounter makes the reference itself change by 7e-6 under a 1e-7 relative change of the tolerance
    # KY() list neighbours: ring with nearest / next-nearest coupling; n = 3 wraps around more than once for degree 2
    "ring_dde_n50": ("ring_diffusion", {"meth": "dde", "h0": 0.05, "res": 11, "ps": 50}, (0, 2.0), RING),
    "ring_fe_n3": ("ring_diffusion", {"meth": "fe", "h0": 0.01, "res": 3, "ps": 3}, (0, 0.2), RING),
    # model-language coverage: every option spelled out
    "chain_array_dde_n20": ("chain_array", dict(FULL, meth="dde", h0=0.05, res=4, ps=20), (0, 3), dict(nvar=6, naux=6, npsi=0, nmfd=2, nexpr=4)),
    "fun_block_rkf45_n21": ("fun_block", dict(FULL, meth="rkf45", h0=0.05, res=4, ps=21), (0, 3), dict(nvar=3, naux=5, npsi=0, nmfd=1, nexpr=6)),
    "helper_calls_dde_n12": ("helper_calls", dict(FULL, meth="dde", h0=0.05, res=5, ps=12), (0, 2), dict(nvar=4, naux=13, npsi=0, nmfd=1, nexpr=6)),
    "helper_calls_rkck_n9": ("helper_calls", dict(FULL, meth="rkck", h0=0.05, res=4, ps=9), (0, 2), dict(nvar=4, naux=13, npsi=0, nmfd=1, nexpr=6)),
    "delay_pop_dde_n8": ("delay_pop", dict(FULL, meth="dde", h0=0.05, res=5, ps=8), (0, 4), dict(nvar=1, naux=1, npsi=0, nmfd=1, nexpr=1)),
    "delay_pop_dde_stop_n5": ("delay_pop", dict(FULL, meth="dde", h0=0.05, res=4, ps=5), (0, 1.0, 2.5), dict(nvar=1, naux=1, npsi=0, nmfd=1, nexpr=1)),
    "pair_aux_dde_n30": ("pair_aux", dict(FULL, meth="dde", h0=0.05, res=3, ps=30), (0, 2), dict(nvar=2, naux=3, npsi=2, nmfd=1, nexpr=1)),
    "interp_drive_dde_n12": ("interp_drive", dict(FULL, meth="dde", h0=0.05, res=5, ps=12), (0, 5), dict(nvar=1, naux=2, npsi=0, nmfd=1, nexpr=1)),
    "interp_drive_rkck_n7": ("interp_drive", dict(FULL, meth="rkck", h0=0.05, res=5, ps=7), (0, 5), dict(nvar=1, naux=2, npsi=0, nmfd=1, nexpr=1)),
    "repel2d_dde_n50": ("repel2d", dict(FULL, meth="dde", h0=0.05, res=3, ps=50), (0, 1), dict(nvar=4, naux=0, npsi=2, nmfd=2, nexpr=0)),
    "repel2d_rkf45_n33": ("repel2d", dict(FULL, meth="rkf45", h0=0.05, res=3, ps=33, abstol=1e-9, reltol=1e-9), (0, 1), dict(nvar=4, naux=0, npsi=2, nmfd=2, nexpr=0)),
    "nbody_const_rkf45": ("nbody_const", dict(FULL, meth="rkf45", h0=0.1, res=11, ps=3, abstol=1e-8, reltol=1e-8, hmin=0), (0, 20),
                          dict(nvar=6, naux=0, npsi=3, nmfd=0, nexpr=0)),
    "lowrank_aux_dde_n100": ("kuramoto_lowrank_aux", dict(FULL, meth="dde", h0=0.1, res=5, ps=100), (0, 2.0), dict(nvar=1, naux=2, npsi=1, nmfd=2, nexpr=7)),
    "lrexp_dde_n200": ("kuramoto_lrexp", dict(FULL, meth="dde", h0=0.1, res=3, ps=200), (0, 1.0), dict(nvar=1, naux=1, npsi=1, nmfd=2, nexpr=1)),
    "vect_meanfield_tauleap": ("vect_meanfield", dict(FULL, meth="fe", h0=0.01, res=11, ps=40, ssahmin=1.0), (0, 2), dict(nvar=1, naux=1, npsi=0, nmfd=3, nexpr=0)),
    "lineage_tauleap_n20": ("lineage", dict(FULL, meth="fe", h0=0.01, res=31, ps=20, ssahmin=1.0), (0, 6), dict(nvar=1, naux=1, npsi=0, nmfd=2, nexpr=4)),
    "lineage_ssa_n20": ("lineage", dict(FULL, meth="fe", h0=0.01, res=31, ps=20, ssahmin=0.0), (0, 6), dict(nvar=1, naux=1, npsi=0, nmfd=2, nexpr=4)),
    "nbody_rkf45_n96": ("nbody", {"meth": "rkf45", "h0": 0.02, "res": 3, "ps": 96, "abstol": 1e-8, "reltol": 1e-8, "hmin": 0}, (0, 0.2), NBD),
}


def main():
    only = set(sys.argv[1:])
    for name, (fx, opts, ts, dims) in CASES.items():
        if only and name not in only:
            continue
        r = refrun.run(fx, opts, timespan=ts, dims=dims)
        if name == "kuramoto_shipped_rkck":
            opts, ts = {"meth": "rkck", "res": 801, "ps": 1000}, (0, 160)     # what the shipped file sets itself
        st, tr = r["stats"], r["traj"]
        last = tr[tr["step"] == tr["step"].max()]
        first = tr[tr["step"] == 0]
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            stats_t=st["t"], stats_mfd=st["mfd"], stats_n=st["n"], stats_flag=st["stop_flag"], stats_event=st["event"],
                            first_id=first["id"], first_cols=first["cols"], last_id=last["id"], last_cols=last["cols"], last_t=last["t"],
                            nframes=np.array([tr["step"].max() + 1]), quick=r.get("quick", np.zeros((0, 3))),
                            meta=np.array([repr(dict(fixture=fx, opts=opts, timespan=ts, dims=dims))]))
        print(name, "stats rows", len(st), "frames", tr["step"].max() + 1)


if __name__ == "__main__":
    main()
