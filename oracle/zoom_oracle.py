"""CPU restatement of `scipy.ndimage.zoom` exactly as the reference calls it (utils.py:69 `zoom(slice, (P/x, P/y), order=3)` and
utils.py:77 `zoom(out, (x/P, y/P), order=0)`): mode='constant', cval=0, prefilter=True, grid_mode=False.

TEST INFRASTRUCTURE ONLY (checker for the GPU resampling kernels; pinned against scipy itself in tests/test_oracle_golden.py —
scipy is the third-party dependency that holds this arithmetic, requirements.txt:10, unpinned; installed here: 1.18).
Algorithm (scipy/ndimage/src/ni_splines.c, ni_interpolation.c NI_ZoomShift):
  * output length = round(in * zoom); coordinate of output o along an axis: cc = o * ((in - 1) / (out - 1)) in float64;
  * cc < 0 or cc > in - 1 (as float64!) -> the output pixel is cval = 0.  With 512 -> 224 the last coordinate is
    223 * (511 / 223) = 511.00000000000006 > 511, so scipy's LAST ROW AND COLUMN ARE ZERO; the reference feeds that to the net;
  * order 3: separable cubic B-spline; coefficients from the recursive prefilter (pole sqrt(3) - 2, gain 6, mirror boundaries:
    for mode 'constant' scipy filters with mirror initialisation), float64; taps floor(cc) - 1 .. + 2, indices that leave the
    array are mirrored (i < 0 -> -i, i >= n -> 2n - 2 - i); weights of the cubic B-spline at t = cc - floor(cc); result cast
    to the input dtype;
  * order 0: index floor(cc + 0.5).
"""
from __future__ import annotations

import numpy as np

POLE = np.sqrt(3.0) - 2.0


def out_len(n_in: int, factor: float) -> int:
    return int(round(n_in * factor))


def spline_prefilter_axis(c: np.ndarray, axis: int) -> np.ndarray:
    """Cubic B-spline prefilter along `axis` with mirror boundaries (float64), all lines at once."""
    c = np.moveaxis(np.array(c, dtype=np.float64), axis, 0).copy()
    n = c.shape[0]
    if n < 2:
        return np.moveaxis(c, 0, axis)
    z = POLE
    c *= (1.0 - z) * (1.0 - 1.0 / z)
    z_n_1 = z ** (n - 1)
    c0 = c[0] + z_n_1 * c[n - 1]
    z_i = z
    for i in range(1, n - 1):
        c0 = c0 + z_i * (c[i] + z_n_1 * c[n - 1 - i])
        z_i *= z
    c[0] = c0 / (1.0 - z_n_1 * z_n_1)
    for i in range(1, n):
        c[i] += z * c[i - 1]
    c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1.0)
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])
    return np.moveaxis(c, 0, axis)


def _axis_taps(n_in: int, n_out: int):
    """(indices (n_out, 4), weights (n_out, 4), inside (n_out,)) of the cubic interpolation along one axis."""
    zf = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
    cc = np.arange(n_out, dtype=np.float64) * zf
    inside = (cc >= 0) & (cc <= n_in - 1)
    fl = np.floor(cc)
    t = cc - fl
    w = np.stack([(1 - t) ** 3 / 6, (3 * t ** 3 - 6 * t ** 2 + 4) / 6, (-3 * t ** 3 + 3 * t ** 2 + 3 * t + 1) / 6, t ** 3 / 6], 1)
    idx = fl.astype(np.int64)[:, None] - 1 + np.arange(4)[None, :]
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= n_in, 2 * n_in - 2 - idx, idx)
    idx = np.clip(idx, 0, n_in - 1)                    # only reachable for outside coordinates, which are zeroed anyway
    return idx, w, inside


def zoom_cubic(x: np.ndarray, out_hw) -> np.ndarray:
    """== scipy.ndimage.zoom(x, (oh / H, ow / W), order=3) for a 2-D (or batched (n, H, W)) float array."""
    x = np.asarray(x)
    batched = x.ndim == 3
    xb = x if batched else x[None]
    H, W = xb.shape[1:]
    oh, ow = out_hw
    c = spline_prefilter_axis(spline_prefilter_axis(xb.astype(np.float64), 1), 2)
    iy, wy, in_y = _axis_taps(H, oh)
    ix, wx, in_x = _axis_taps(W, ow)
    out = np.zeros((xb.shape[0], oh, ow), np.float64)
    for a in range(4):
        rows = c[:, iy[:, a], :]                                        # (n, oh, W)
        for b in range(4):
            out += wy[None, :, a, None] * wx[None, None, :, b] * rows[:, :, ix[:, b]]
    out *= (in_y[:, None] & in_x[None, :])[None]
    out = out.astype(x.dtype)
    return out if batched else out[0]


def zoom_nearest(x: np.ndarray, out_hw) -> np.ndarray:
    """== scipy.ndimage.zoom(x, (oh / H, ow / W), order=0) (label maps)."""
    x = np.asarray(x)
    batched = x.ndim == 3
    xb = x if batched else x[None]
    H, W = xb.shape[1:]
    oh, ow = out_hw

    def axis(n_in, n_out):
        zf = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
        cc = np.arange(n_out, dtype=np.float64) * zf
        inside = (cc >= 0) & (cc <= n_in - 1)
        return np.clip(np.floor(cc + 0.5).astype(np.int64), 0, n_in - 1), inside
    iy, in_y = axis(H, oh)
    ix, in_x = axis(W, ow)
    out = xb[:, iy][:, :, ix] * (in_y[:, None] & in_x[None, :])[None].astype(x.dtype)
    return out if batched else out[0]
