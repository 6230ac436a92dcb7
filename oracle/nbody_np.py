"""TEST INFRASTRUCTURE — numpy restatement of the all-pairs %C walk of the generated pop_rhs for the N-body fixture.

Only tests/ may import this.  Follows src/odexp:905-954 (generated code: for every particle `myself_` one walk over the
whole list `other_`, `psi_k += expr_k(myself_, other_)`, then `psi_k /= POP_SIZE`, self term included and guarded by the
model) with the expressions of models/nbody.pop (= examples/Nbody.pop:12-18):
    grav(a, b, ...) = (a == b) ? 0 : (a - b) / pow(dx*dx + dy*dy + dz*dz, 1.5),   GRAVITY_k = -G * mass * grav(...)
Evaluated for a SUBSET of i-particles, so that populations far larger than the C restatement can walk in seconds are
checkable (O(len(subset) * N)).  numpy sums pairwise, the reference sequentially: the two differ by O(1e-16) relative.
Pinned by tests/test_oracle_port.py::test_nbody_numpy_restatement_matches_port against oracle/odx_oracle.c, which is
bit-identical to the unchanged reference binary for this fixture.
"""
import numpy as np

G = 6.674e-11


def psi_subset(pos, idx, mass=1e9):
    """pos: (N, 3) positions in list order; idx: indices of the i-particles; returns (len(idx), 3) couplings"""
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[0]
    out = np.empty((len(idx), 3))
    for q, i in enumerate(idx):
        d = pos[i][None, :] - pos                      # x_i - x_j
        r2 = (d * d).sum(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = d / np.power(r2, 1.5)[:, None]
        g[d == 0.0] = 0.0                              # (numx == numxi) ? 0.0 : ...  per component
        out[q] = (-G * mass) * g.sum(axis=0) / n
    return out
