"""The oracle's restatement of the two PCL filters in front of the registration path (SURVEY.md §8f row 2:
pcl::VoxelGrid / pcl::CropBox as OdomNode::preprocessPoints uses them, odom.cc:442-478) against independent
numpy statements of the published algorithms.  PCL itself is not installable here."""
import numpy as np
import pytest

from dynamic_direct_lidar_odometry_b200 import synth


def _numpy_voxel(points, leaf):
    """independent float64 statement: group by floor(p / leaf), mean per group, groups ordered by x-fastest index"""
    p = points[:, :3]
    ok = np.isfinite(p).all(axis=1)
    p = p[ok]
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(p * inv).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    idx = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    order = np.argsort(idx, kind="stable")
    idx_s = idx[order]
    start = np.flatnonzero(np.concatenate([[True], idx_s[1:] != idx_s[:-1]]))
    sums = np.add.reduceat(p[order].astype(np.float64), start, axis=0)
    cnt = np.diff(np.concatenate([start, [len(idx_s)]]))
    return sums / cnt[:, None], cnt


@pytest.mark.parametrize("leaf", [0.25, 0.5, 1.0])
def test_voxel_filter_matches_numpy(oracle, leaf):
    scan = synth.scan(2, 32, 512)
    got = oracle.voxel_filter(scan, leaf)
    want, cnt = _numpy_voxel(scan, leaf)
    assert got.shape == (len(want), 4) and np.all(got[:, 3] == 1.0)
    # float sums of up to a few hundred points: a few ulp of the coordinate magnitude
    assert np.abs(got[:, :3] - want).max() < 1e-4
    assert len(got) < len(scan)


def test_voxel_filter_edge_cases(oracle):
    assert oracle.voxel_filter(np.zeros((0, 4), np.float32), 0.5).shape == (0, 4)
    p = np.array([[0.1, 0.1, 0.1, 1], [0.2, 0.2, 0.2, 1], [np.nan, 0, 0, 1], [5.0, 5.0, 5.0, 1], [np.inf, 1, 1, 1]], dtype=np.float32)
    out = oracle.voxel_filter(p, 1.0)
    assert out.shape == (2, 4)
    assert np.array_equal(out[0, :3], ((p[0, :3] + p[1, :3]) / np.float32(2)).astype(np.float32))
    assert np.array_equal(out[1, :3], p[3, :3])
    # idempotence: centroids of a filtered cloud stay in their own voxels
    scan = synth.scan(0, 16, 256)
    once = oracle.voxel_filter(scan, 0.5)
    assert len(oracle.voxel_filter(once, 0.5)) == len(once)
    with pytest.raises(OverflowError):
        oracle.voxel_filter(np.array([[0, 0, 0, 1], [1e4, 1e4, 1e4, 1]], np.float32), 1e-3)


def test_crop_box(oracle):
    scan = synth.scan(1, 16, 256)
    lo, hi = np.array([-2.0, -2.0, -2.0], np.float32), np.array([2.0, 2.0, 2.0], np.float32)
    inside = np.all((scan[:, :3] >= lo) & (scan[:, :3] <= hi), axis=1)
    assert np.array_equal(oracle.crop_box(scan, lo, hi)[:, :3], scan[inside, :3])
    assert np.array_equal(oracle.crop_box(scan, lo, hi, negative=True)[:, :3], scan[~inside, :3])
    org = oracle.crop_box(scan, lo, hi, negative=True, keep_organized=True)
    assert len(org) == len(scan) and np.isnan(org[inside, :3]).all() and np.array_equal(org[~inside, :3], scan[~inside, :3])


def test_residual_image_oracle(oracle):
    """odom.cc:804-827 restated: last point per cell wins, empty cells zero, out-of-view points skipped."""
    scan = synth.scan(4, 32, 512)
    res = np.linalg.norm(scan[:, :3], axis=1).astype(np.float64) * 1e-3
    img = oracle.residual_image(scan, res, 64, 48, -np.pi / 3, np.pi / 3)
    assert img.shape == (48, 64, 4)
    # independent numpy statement
    x, y, z = (scan[:, i].astype(np.float64) for i in range(3))
    th, ph = np.arctan2(x, z), np.arctan2(y, np.sqrt(x * x + z * z))
    u = np.trunc((th + np.pi / 3) / (2 * np.pi / 3) * 64).astype(int)
    v = np.trunc((ph + np.pi / 3) / (2 * np.pi / 3) * 48).astype(int)
    want = np.zeros((48, 64, 4), np.float32)
    for i in np.flatnonzero((u >= 0) & (u < 64) & (v >= 0) & (v < 48)):
        want[v[i], u[i]] = (scan[i, 0], scan[i, 1], scan[i, 2], np.float32(res[i]))
    assert np.array_equal(img, want)
    assert (img[..., 3] > 0).any() and (img[..., 3] == 0).any()


def test_extract_stride_known_answer(oracle):
    """the strided mask of odom.cc:124-130 on a 4 x 6 scan with strides (2, 3): rows 0 and 2, columns 0 and 3"""
    pts = np.ones((24, 4), np.float32)
    pts[:, 0] = np.arange(24)
    out = oracle.extract_stride(pts, 6, 4, 2, 3)
    kept = np.flatnonzero(~np.isnan(out[:, 0]))
    assert kept.tolist() == [0, 3, 12, 15]
    assert np.array_equal(out[kept, 0], pts[kept, 0]) and (out[:, 3] == 1.0).all()
