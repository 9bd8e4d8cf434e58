"""DEGENSAC's host-side seven-point H-degeneracy test (pxb_h_degenerate_sample, no GPU needed) against the numpy
restatement of fundamental_estimator.h:352-476 in oracle/oracle.py."""
import numpy as np

from pyprogressivex import _native
from pyprogressivex.synthetic import plane_dominated_pair


def test_h_degenerate_sample_matches_oracle(oracle):
    rows, lab, _ = plane_dominated_pair(60, 60, 0.2, 5)
    rng = np.random.default_rng(0)
    plane, off = np.flatnonzero(lab == 0), np.flatnonzero(lab == 1)
    checked = degenerate_seen = general_seen = 0
    for trial in range(200):
        n_on = [7, 6, 5, 3, 2, 0][trial % 6]
        sample = np.concatenate([rng.choice(plane, n_on, replace=False), rng.choice(off, 7 - n_on, replace=False)])
        sample = rng.permutation(sample).astype(np.int64)
        models, n, _, _ = oracle.solve_minimal(oracle.MODEL_F, rows, sample[None])
        for j in range(int(n[0])):
            F = models[0, j].reshape(3, 3)
            want, H_o, margins = oracle.h_degenerate_sample(rows, sample, F)
            if any(abs(e - 4.0) < 0.05 for m in margins for e in m):
                continue  # on the 2 px decision boundary: the two SVD routes may round differently
            got, H = _native.h_degenerate_sample(rows, sample, F)
            assert got == want, (trial, j, margins)
            checked += 1
            if got:
                degenerate_seen += 1
                np.testing.assert_allclose(H / np.linalg.norm(H), H_o / np.linalg.norm(H_o), rtol=0, atol=1e-7)
            else:
                general_seen += 1
    assert checked > 100 and degenerate_seen > 20 and general_seen > 20, (checked, degenerate_seen, general_seen)


def test_h_degenerate_sample_rejects_null_arguments():
    import ctypes
    lib = _native.load_library()
    assert lib.pxb_h_degenerate_sample(None, None, None, None, None) < 0
    assert b"null" in lib.pxb_last_error()
