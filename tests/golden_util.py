"""Helpers to read tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def compare(z, group, full, atol, rtol=0.0, what=""):
    """Compare `full` (array-like, any float dtype, last dim = row width) with packed group `group` of npz `z`.

    Returns the max-abs error over the stored rows; asserts rows and the full-tensor checksums."""
    a = np.asarray(full, dtype=np.float64)
    shape = tuple(int(s) for s in z[f"{group}.shape"])
    assert a.shape == shape, f"{what or group}: shape {a.shape} != golden {shape}"
    stride = int(z[f"{group}.stride"])
    rows = a.reshape(-1, a.shape[-1])[::stride]
    ref = z[f"{group}.rows"].astype(np.float64)
    err = np.abs(rows - ref)
    tol = atol + rtol * np.abs(ref)
    worst = float(err.max()) if err.size else 0.0
    assert (err <= tol).all(), f"{what or group}: max-abs err {worst:.3e} > tol {atol:.1e} (+{rtol:.1e} rel)"
    # full-tensor guards (catch errors in rows that were not sampled)
    n = a.size
    assert abs(a.sum() - float(z[f"{group}.sum"])) <= (atol + rtol) * n * 0.05 + 1e-6 * abs(float(z[f"{group}.abssum"])) + 4 * atol * np.sqrt(n), \
        f"{what or group}: checksum mismatch"
    assert abs(np.abs(a).sum() - float(z[f"{group}.abssum"])) <= (atol + rtol * 1.0) * n + 1e-6 * float(z[f"{group}.abssum"]), \
        f"{what or group}: abs-checksum mismatch"
    return worst
