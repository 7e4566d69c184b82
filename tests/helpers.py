"""Shared helpers of the parity tests."""
import numpy as np


def assert_close(actual, ref, rel=1e-4, what="", per_column=None):
    """|actual - ref| <= rel * (|ref| + rms) elementwise, NaN / inf patterns identical.

    'within 1e-4 relative' (BASELINE.json) read against a scale: entries of a mean-normalised feature column pass through
    zero, where a pure per-element relative error is meaningless.  For 2-D arrays the rms is taken PER COLUMN (default): the
    columns of a feature matrix differ by two orders of magnitude (c1 ~ 50, a delta-delta ~ 1), and one global rms would
    hold the small columns to the tolerance of the large ones."""
    actual = np.asarray(actual, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert actual.shape == ref.shape, f"{what}: shape {actual.shape} != {ref.shape}"
    fin = np.isfinite(ref)
    assert np.array_equal(np.isnan(actual), np.isnan(ref)), f"{what}: NaN pattern differs"
    assert np.array_equal(actual[~fin & ~np.isnan(ref)], ref[~fin & ~np.isnan(ref)]), f"{what}: inf pattern differs"
    if not fin.any():
        return
    if per_column is None:
        per_column = ref.ndim == 2
    if per_column and ref.ndim == 2:
        sq = np.where(fin, ref, 0.0) ** 2
        cnt = np.maximum(fin.sum(axis=0), 1)
        scale = np.broadcast_to(np.sqrt(sq.sum(axis=0) / cnt)[None, :], ref.shape)[fin]
    else:
        scale = np.sqrt(np.mean(ref[fin] ** 2))
    err = np.abs(actual[fin] - ref[fin])
    bound = rel * (np.abs(ref[fin]) + scale)
    worst = np.argmax(err - bound)
    assert np.all(err <= bound), (f"{what}: max excess at flat index {worst}: got {actual[fin][worst]!r}, "
                                  f"ref {ref[fin][worst]!r}, err {err[worst]:.3e} > {bound[worst]:.3e}")


def path_score(path, pi, A, b):
    """log score of a state path under (pi, A) with emissions b [N,T] (float64)."""
    path = np.asarray(path, dtype=np.int64)
    s = pi[path[0]] + b[path[0], 0]
    for t in range(1, len(path)):
        s += A[path[t - 1], path[t]] + b[path[t], t]
    return s


def assert_paths_equal_ties_aside(path, ref_path, pi, A, b, tol=1e-9, what=""):
    """Bit-exact paths, ties aside (BASELINE.json): a mismatch is accepted only if both paths score the same."""
    path, ref_path = np.asarray(path, dtype=np.int64), np.asarray(ref_path, dtype=np.int64)
    if np.array_equal(path, ref_path):
        return
    s1, s2 = path_score(path, pi, A, b), path_score(ref_path, pi, A, b)
    assert abs(s1 - s2) <= tol * max(1.0, abs(s2)), f"{what}: paths differ and scores differ: {s1} vs {s2}"
