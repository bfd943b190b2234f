"""The invariant behind `key_less` (csrc/knn.cuh): a candidate key bits(d2) << 32 | index, d2 a finite non-negative float,
is also the bit pattern of a finite non-negative double, and such doubles order exactly like their bit patterns - so the
device compares keys with one double comparison.  Checked here on the host over the corner cases of both words."""
import numpy as np


def _keys():
    rng = np.random.default_rng(5)
    d = np.concatenate([
        np.array([0.0, np.finfo(np.float32).tiny, np.finfo(np.float32).max, 1.0, 1e-45, 1e-38, 3.0e38], dtype=np.float32),
        np.nextafter(np.float32(1.0), np.float32(2.0), dtype=np.float32).reshape(1),
        rng.uniform(0.0, 100.0, 400).astype(np.float32),
        (rng.uniform(0.0, 1.0, 100) ** 8 * 1e-30).astype(np.float32),      # tiny and denormal floats
        np.abs(rng.standard_normal(100).astype(np.float32)) * np.float32(1e30),
    ])
    d = np.minimum(d, np.finfo(np.float32).max)
    idx = np.concatenate([
        np.array([0, 1, 2, 0x7FFFFFFF, 0x7FFFFFFE, 65535, 65536], dtype=np.uint64),
        rng.integers(0, 2**31 - 1, len(d) - 7).astype(np.uint64),
    ])
    keys = (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | idx
    extra = np.array([0, (int(np.float32(np.finfo(np.float32).max).view(np.uint32)) << 32) | 0x7FFFFFFF], dtype=np.uint64)  # "nothing to the left", the empty entry
    return np.concatenate([keys, extra])


def test_candidate_keys_order_like_doubles():
    k = _keys()
    as_double = k.view(np.float64)
    assert np.all(np.isfinite(as_double)) and np.all(as_double >= 0.0)
    lt_int = k[:, None] < k[None, :]
    lt_dbl = as_double[:, None] < as_double[None, :]
    assert np.array_equal(lt_int, lt_dbl)
    # equal keys stay equal, distinct keys stay distinct (denormals are not flushed)
    assert np.array_equal(k[:, None] == k[None, :], as_double[:, None] == as_double[None, :])


def test_the_no_candidate_pattern_is_not_a_key():
    """~0 marks "no candidate" in the scans; as a double it is a NaN, which is why it is compared as an integer there."""
    assert np.isnan(np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64).view(np.float64)[0])
