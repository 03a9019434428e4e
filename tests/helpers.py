"""Shared comparison helpers for the parity tests."""
import numpy as np

GEOM_FIELDS = ("origin", "direction", "throughput", "normal", "distance", "bounces", "pixel_index")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_records_equal(got, want, fields=GEOM_FIELDS, what="records"):
    """Bit-for-bit equality of ray/shadow records on the given fields (floats compared as bit patterns)."""
    assert got.shape == want.shape, "%s: %s vs %s records" % (what, got.shape, want.shape)
    if got.shape[0] == 0:
        return
    for f in fields:
        x, y = got[f], want[f]
        if x.dtype.kind == "f":
            x, y = bits(x), bits(y)
        bad = np.flatnonzero((x != y).reshape(got.shape[0], -1).any(axis=1))
        assert bad.size == 0, "%s: field %s differs in %d of %d records, first at %s: %s vs %s" % (
            what, f, bad.size, got.shape[0], bad[:3], got[f][bad[:3]], want[f][bad[:3]])


def max_rel_err(got, want, floor=1e-30):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want) / np.maximum(np.abs(want), floor)
    err[got == want] = 0
    return float(err.max()) if err.size else 0.0


def assert_close_rel(got, want, tol, what):
    e = max_rel_err(got, want)
    assert e <= tol, "%s: max relative error %.3e > %.1e" % (what, e, tol)


def tile_means(acc):
    h, w = acc.shape[:2]
    th, tw = h // 8, w // 8
    return acc[: th * 8, : tw * 8].reshape(8, th, 8, tw, 4).astype(np.float64).mean(axis=(1, 3)).astype(np.float32)
