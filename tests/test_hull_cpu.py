"""The hull code of the keyframe store (csrc/hull.hpp, host only) against Qhull itself (scipy): pcl::ConvexHull and
pcl::ConcaveHull, which OdomNode::getSubmapKeyframes uses on the keyframe positions (odom.cc:993-1065), call Qhull, and
neither PCL nor Qhull is under /root/reference.  No GPU needed: the functions are plain C++ exported by the library."""
import numpy as np
import pytest

from oracle_backend import pcl_dimension, qhull_concave, qhull_convex


@pytest.fixture(scope="module")
def ng(ddlo_lib):
    from dynamic_direct_lidar_odometry_b200 import nano_gicp

    return nano_gicp


def planar(rng, n, tilt=0.0):
    p = np.zeros((n, 3))
    p[:, :2] = rng.normal(0, 10, (n, 2))
    p[:, 2] = 1.5 + rng.normal(0, 0.01, n) + tilt * p[:, 0]
    return p


def test_convex_and_concave_2d_match_qhull(ng):
    rng = np.random.default_rng(0)
    for trial in range(150):
        p = planar(rng, int(rng.integers(5, 120)), tilt=float(rng.uniform(0, 0.3)))
        assert pcl_dimension(p) == 2
        assert ng.hull_convex(p) == qhull_convex(p), trial
        alpha = float(rng.uniform(2, 15))
        assert ng.hull_concave(p, alpha) == qhull_concave(p, alpha), (trial, alpha)


def test_projection_plane_follows_the_normal(ng):
    """a trajectory in a vertical plane (normal along x or y): PCL projects onto yz / xz instead of xy"""
    rng = np.random.default_rng(1)
    for axis in (0, 1):
        p = np.zeros((40, 3))
        cols = [c for c in range(3) if c != axis]
        p[:, cols] = rng.normal(0, 5, (40, 2))
        p[:, axis] = rng.normal(0, 1e-3, 40)
        assert pcl_dimension(p) == 2
        assert ng.hull_convex(p) == qhull_convex(p)
        assert ng.hull_concave(p, 3.0) == qhull_concave(p, 3.0)


def test_convex_3d_matches_qhull_and_concave_3d_is_reported(ng):
    rng = np.random.default_rng(2)
    for trial in range(60):
        p = rng.normal(0, 5, (int(rng.integers(6, 150)), 3))
        assert pcl_dimension(p) == 3
        assert ng.hull_convex(p) == qhull_convex(p), trial
        assert ng.hull_concave(p, 3.0) is None  # not implemented for 3-D position sets (hull.hpp)


def test_known_answers_and_degenerate_inputs(ng):
    sq = np.array([[0, 0, 0], [4, 0, 0], [4, 4, 0], [0, 4, 0], [2, 2, 0], [1, 3, 0]], dtype=np.float64)
    assert ng.hull_convex(sq) == [0, 1, 2, 3]
    assert ng.hull_concave(sq, 100.0) == [0, 1, 2, 3]      # a huge alpha gives the convex hull
    assert ng.hull_concave(sq, 1e-3) == []                 # no triangle is that small
    # a point in the middle of an edge is not an extreme point
    e = np.array([[0, 0, 0], [2, 0, 0], [4, 0, 0], [4, 4, 0], [0, 4, 0]], dtype=np.float64)
    assert ng.hull_convex(e) == [0, 2, 3, 4]
    line = np.array([[i, 2 * i, 0] for i in range(6)], dtype=np.float64)
    assert ng.hull_convex(line) == [0, 5]
    assert ng.hull_concave(line, 5.0) == []
    assert ng.hull_convex(np.zeros((2, 3))) == []
    dup = np.array([[0, 0, 0], [0, 0, 0], [3, 0, 0], [0, 3, 0], [3, 3, 0]], dtype=np.float64)
    assert len(ng.hull_convex(dup)) == 4
    # an L-shaped corridor
    xs = [(x, 0.0) for x in np.arange(0, 10.5, 1.0)] + [(10.0, y) for y in np.arange(1, 10.5, 1.0)]
    xs += [(x, 1.0) for x in np.arange(0, 9.5, 1.0)] + [(9.0, y) for y in np.arange(2, 10.5, 1.0)]
    L = np.array([[x, y, 0.0] for x, y in xs])
    inner = int(np.flatnonzero((L[:, 0] == 9.0) & (L[:, 1] == 1.0))[0])
    # (the exact lattice is degenerate - four co-circular points per cell, any diagonal is a Delaunay triangulation -
    # so the comparison with Qhull is made on a slightly jittered copy)
    L[:, :2] += np.random.default_rng(5).normal(0, 0.02, (len(L), 2))
    assert inner not in ng.hull_convex(L)
    assert ng.hull_concave(L, 1.0) == qhull_concave(L, 1.0)
    assert len(ng.hull_concave(L, 1.0)) > 2 * len(ng.hull_convex(L))  # the alpha shape follows the corridor, the convex hull does not
