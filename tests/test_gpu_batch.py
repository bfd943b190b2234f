"""The C++ batch driver (ddlo_batch_*, BASELINE config C5) and handle sharing between runtimes.  B200 box."""
import ctypes as C

import numpy as np
import pytest

from dynamic_direct_lidar_odometry_b200 import binding as B
from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng
from dynamic_direct_lidar_odometry_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scans():
    w = synth.make_world()
    return [synth.scan(f, 16, 256, w) for f in range(7)]


def single_align(device, blocks, src, tgt, guess=None, shared=None):
    rt = ng.Runtime(device)
    rt.set_align_blocks(blocks)
    g = ng.NanoGICP(rt)
    g.setInputSource(ng.PointCloud(rt, src))
    if shared is None:
        g.setInputTarget(ng.PointCloud(rt, tgt))
    else:
        g.setInputTarget(shared[0])
        g.setTargetCovariances(shared[1])
    r = g.align(guess)
    del g
    rt.close()
    return r


@pytest.mark.parametrize("mode", ["waves", "lanes"])
def test_batch_s2s_bit_identical_to_single_align(scans, mode):
    """every unit of a batch returns the pose, Hessian and iteration counts a single ddlo_gicp_align gives (waves: an
    ordinary engine; lanes: one with the same align-block limit) - bit for bit, whatever lane or wave it ran in"""
    b = ng.Batch(0, lanes=3, host_threads=2, mode=mode, wave_units=4)
    ids = [b.stage(s) for s in scans]
    scans_by_id = dict(zip(ids, scans))
    rng = np.random.default_rng(3)
    units = []
    for u in range(11):
        s = u % 6
        guess = np.eye(4, dtype=np.float32)
        guess[:3, 3] = rng.normal(0, 0.02, 3)
        units.append((ids[s + 1], ids[s], guess))
    res = b.run(units)
    res2 = b.run(units)  # lanes are reused: nothing of the first run may leak into the second
    for (s, t, guess), r, r2 in zip(units, res, res2):
        ref = single_align(0, b.align_blocks if mode == "lanes" else 0, scans_by_id[s], scans_by_id[t], guess)
        assert r.converged and ref.converged
        assert (r.iterations, r.n_linearize, r.n_compute_error) == (ref.iterations, ref.n_linearize, ref.n_compute_error)
        assert np.array_equal(r.T, ref.T) and np.array_equal(r.hessian, ref.hessian) and r.final_error == ref.final_error
        assert r.covs_computed
        assert np.array_equal(r.T, r2.T) and np.array_equal(r.hessian, r2.hessian)
    assert b.launch_count() > 0
    if mode == "waves":
        rounds, polls = b.stats()
        assert rounds >= 2 * 3 * 4 and polls >= 6  # 11 units in waves of 4, twice
    b.close()


@pytest.mark.parametrize("mode", ["waves", "lanes"])
def test_batch_shared_target_matches_single_engine(scans, mode):
    """scan-to-map units (target = -1) against one submap whose index and covariances are shared by all lanes"""
    submap = np.concatenate([synth.transform(scans[f], synth.pose(f)) for f in (0, 2, 4)])
    b = ng.Batch(0, lanes=4, host_threads=1, mode=mode, wave_units=5)
    sub_id = b.stage(submap)
    ids = [b.stage(scans[f]) for f in (1, 3, 5)]
    b.set_shared_target(sub_id)
    guesses = [synth.pose(f).astype(np.float32) for f in (1, 3, 5)]
    units = [(ids[i % 3], -1, guesses[i % 3]) for i in range(9)]
    res = b.run(units)
    # the same on one ordinary engine
    rt = ng.Runtime(0)
    tgt = ng.PointCloud(rt, submap).share()
    cov = ng.Covariances.compute(tgt, 20).share(tgt)
    for i, r in enumerate(res):
        ref = single_align(0, b.align_blocks if mode == "lanes" else 0, scans[(1, 3, 5)[i % 3]], None, guesses[i % 3], shared=(tgt, cov))
        assert r.converged == ref.converged and r.iterations == ref.iterations
        assert np.array_equal(r.T, ref.T) and np.array_equal(r.hessian, ref.hessian)
        assert r.covs_computed  # source covariances are computed inside the unit
    # mixing unit kinds on the same lanes: S2S after S2M and back
    mixed = [(ids[0], -1, guesses[0]), (ids[1], ids[0], None), (ids[2], -1, guesses[2]), (ids[0], ids[2], None), (ids[1], -1, guesses[1])]
    rm = b.run(mixed)
    assert np.array_equal(rm[0].T, res[0].T) and np.array_equal(rm[2].T, res[2].T) and np.array_equal(rm[4].T, res[1].T)
    del tgt, cov
    rt.close()
    b.close()


@pytest.mark.parametrize("mode", ["waves", "lanes"])
def test_batch_errors(scans, mode):
    b = ng.Batch(0, lanes=2, mode=mode)
    i0 = b.stage(scans[0])
    with pytest.raises(ng.DdloError) as e:
        b.run([(i0, -1, None)])  # no shared target
    assert e.value.code == -5
    with pytest.raises(ng.DdloError) as e:
        b.run([(7, i0, None)])
    assert e.value.code == -1
    r = b.run([(i0, i0, None)])  # still usable afterwards
    assert r[0].converged
    assert b.run([]) == []
    b.close()
    with pytest.raises(ng.DdloError):
        ng.Batch(0, lanes=0)


def test_foreign_handles_need_share(rt, scans):
    """a cloud of another runtime is refused until it has been shared; shared handles are read-only inputs"""
    rt2 = ng.Runtime(0)
    c = ng.PointCloud(rt2, scans[0])
    g = ng.NanoGICP(rt)
    with pytest.raises(ng.DdloError) as e:
        g.setInputTarget(c)
    assert e.value.code == -1
    c.share()
    g.setInputTarget(c)
    g.setInputSource(ng.PointCloud(rt, scans[1]))
    r = g.align()
    g2 = ng.NanoGICP(rt2)
    g2.setInputTarget(c)
    g2.setInputSource(ng.PointCloud(rt2, scans[1]))
    r2 = g2.align()
    assert r.converged and np.array_equal(r.T, r2.T)
    del g, g2, c
    rt2.close()


def test_new_input_invalidates_stored_correspondences(rt, scans):
    """after setInputTarget / setInputSource with another cloud the correspondences of the previous align
    must not be readable (they index the old clouds); the reference throws from at() in that situation"""
    g = ng.NanoGICP(rt)
    src = ng.PointCloud(rt, scans[1])
    g.setInputSource(src)
    g.setInputTarget(ng.PointCloud(rt, scans[0]))
    g.align()
    assert len(g.getResiduals()) == len(scans[1])
    g.setInputTarget(ng.PointCloud(rt, scans[0][:50]))  # a much smaller target: stale indices would point behind it
    for call in (lambda: g.getResiduals(), lambda: g.getResidualVectors(np.eye(4)), lambda: g.compute_error(np.eye(4)), lambda: g.correspondences()):
        with pytest.raises(ng.DdloError) as e:
            call()
        assert e.value.code == -5
    g.align()
    g.getResiduals()
    other = ng.PointCloud(rt, scans[1].copy())  # same size, another handle
    g.registerInputSource(other)
    with pytest.raises(ng.DdloError):
        g.getResiduals()
    g.clearTarget()
    with pytest.raises(ng.DdloError):
        g.getResiduals()


def test_cpp_batch_program_shards_over_devices(scans, tmp_path):
    """tests/cpp/batch_protocol.cpp: the batched path driven from C++ only (one batch + one host thread per visible
    device, contiguous shards); it checks every unit against a single engine itself, and its poses must equal the ones
    the Python mirror of the same ABI gets"""
    import subprocess
    from pathlib import Path

    exe = Path(__file__).resolve().parent / "cpp" / "_build" / "batch_protocol"
    if not exe.exists():
        import __graft_entry__ as ge

        ge.build_cpp_tests()
    path = tmp_path / "scans.bin"
    with open(path, "wb") as fh:
        fh.write(np.int32(len(scans)).tobytes())
        for s in scans:
            fh.write(np.int32(len(s)).tobytes())
            fh.write(np.ascontiguousarray(s, dtype=np.float32).tobytes())
    units, lanes = 13, 3
    for mode in ("lanes", "waves"):
        out = subprocess.run([str(exe), str(path), str(units), str(lanes), "2", "8", mode, "5"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr
    got = {}
    for line in out.stdout.splitlines():
        f = line.split()
        if f[0] == "unit":
            got[int(f[1])] = (int(f[3]), int(f[4]), np.array(f[5:], dtype=np.float32).reshape(4, 4))
        elif f[0] == "summary":
            assert int(f[-1]) == 0
    assert len(got) == units
    b = ng.Batch(0, lanes=lanes, mode="waves")
    ids = [b.stage(s) for s in scans]
    n_pairs = len(scans) - 1
    res = b.run([(ids[u % n_pairs + 1], ids[u % n_pairs], None) for u in range(units)])
    for u, r in enumerate(res):
        assert got[u][0] == int(r.converged) and got[u][1] == r.iterations and np.array_equal(got[u][2], r.T)
    b.close()


def test_batch_sources_from_host_memory(scans):
    """ddlo_batch_submit_host: units whose source scan is uploaded from (pinned) host memory inside the submission give
    the same results as units over staged clouds"""
    b = ng.Batch(0, lanes=2, wave_units=3)
    ids = [b.stage(s) for s in scans]
    units = [(ids[u % 6 + 1], ids[u % 6], None) for u in range(8)]
    staged = b.run(units)
    host = []
    for u in range(8):
        a = ng.pinned_array(scans[u % 6 + 1].shape, np.float32)
        a[:] = scans[u % 6 + 1]
        host.append(a)
    b.submit_host([(-1, ids[u % 6], None) for u in range(8)], host)  # the staged source id is not needed
    got = b.wait()
    for s, g in zip(staged, got):
        assert g.converged and np.array_equal(s.T, g.T) and np.array_equal(s.hessian, g.hessian)
    b.close()


def test_batch_parameter_corner_cases(scans):
    """engine parameters reach every lane and wave slot: Gauss-Newton, an iteration limit of 0 (the guess comes back),
    a single LM trial; every result equals the ordinary engine's"""
    b = ng.Batch(0, lanes=2, wave_units=3)
    ids = [b.stage(s) for s in scans]
    rng = np.random.default_rng(9)
    guess = np.eye(4, dtype=np.float32)
    guess[:3, 3] = rng.normal(0, 0.03, 3)
    units = [(ids[u % 5 + 1], ids[u % 5], guess) for u in range(5)]
    for params in (dict(optimizer=ng.OPT_GAUSS_NEWTON), dict(max_iterations=0), dict(lm_max_iterations=1, max_iterations=3),
                   dict(max_correspondence_distance=0.5, k_correspondences=10)):
        b.set_params(**params)
        got = b.run(units)
        rt = ng.Runtime(0)
        for (s, t, g), r in zip(units, got):
            e = ng.NanoGICP(rt)
            for k, v in params.items():
                setattr(e._p, k, v)
            e._push()
            e.setInputSource(ng.PointCloud(rt, scans[ids.index(s)]))
            e.setInputTarget(ng.PointCloud(rt, scans[ids.index(t)]))
            ref = e.align(g)
            assert (r.flags, r.iterations, r.n_linearize, r.n_compute_error) == (ref.flags, ref.iterations, ref.n_linearize, ref.n_compute_error), params
            assert np.array_equal(r.T, ref.T) and np.array_equal(r.hessian, ref.hessian), params
            del e
        rt.close()
    b.close()


def test_batch_unit_that_cannot_run_is_reported(scans):
    """a unit whose source has fewer points than k fails alone (TOO_FEW); the other units of the submission complete"""
    b = ng.Batch(0, lanes=2, wave_units=4)
    ids = [b.stage(s) for s in scans[:3]]
    tiny = b.stage(scans[0][:7])
    with pytest.raises(ng.DdloError) as e:
        b.run([(ids[1], ids[0], None), (tiny, ids[0], None), (ids[2], ids[1], None)])
    assert e.value.code == -4
    ok = b.run([(ids[1], ids[0], None), (ids[2], ids[1], None)])
    assert all(r.converged for r in ok)
    b.close()
