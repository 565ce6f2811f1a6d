import numpy as np
import pytest
import torch


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    t = t.cuda()
    return t.to(dtype) if dtype is not None else t


def host(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def assert_close(actual, desired, rtol, atol_frac=0.0, atol=0.0, what=""):
    """|a - d| <= atol + atol_frac * max|d| + rtol * |d|   elementwise."""
    a = np.asarray(actual, np.float64)
    d = np.asarray(desired, np.float64)
    assert a.shape == d.shape, (what, a.shape, d.shape)
    tol = atol + atol_frac * np.abs(d).max() + rtol * np.abs(d)
    err = np.abs(a - d)
    bad = err > tol
    if bad.any() or not np.isfinite(a).all():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError("%s: %d / %d out of tolerance; worst at %s: got %r want %r (tol %g); max|d|=%g" % (
            what, bad.sum(), bad.size, i, a[i], d[i], tol[i], np.abs(d).max()))
