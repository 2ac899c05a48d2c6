"""BASELINE.json config 2 at its full size (1 000 points x 10 000 particles x 100 frames of 4288 x 2848) through the
public API: properties that do not need the (far too slow) oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_2_full_size(cuda):
    import bench
    import glimpse_b200 as gb
    from glimpse_b200 import synthetic

    W = bench.WORKLOAD
    scene = bench.build_scene(W["n_points"], W["n_frames"])
    observers, models = synthetic.build(scene, gb)
    tracker = gb.Tracker(observers, seed=20260101)
    tracks = tracker.track(models, tile_size=scene.tile_size)
    P, T = W["n_points"], W["n_frames"]
    assert tracks.means.shape == (P, T, 6) and tracks.sigmas.shape == (P, T, 6)
    assert all(e is None for e in tracks.errors)
    assert np.isfinite(tracks.means).all() and np.isfinite(tracks.sigmas).all() and (tracks.sigmas[:, 1:, :2] > 0).all()
    # the filter locks on: every point recovers the synthetic velocity, and its position after 99 days
    v = tracks.vxyz[:, -1]
    assert np.median(np.abs(v[:, 0] - scene.truth_velocity[0])) < 0.01 and np.abs(v[:, 0] - scene.truth_velocity[0]).max() < 0.05
    travelled = tracks.means[:, -1, 0] - tracks.means[:, 0, 0]
    # (the frames shift uniformly in the image; with lens distortion the ground distance differs by up to ~2 % at the edges)
    assert np.abs(travelled / (scene.truth_velocity[0] * (T - 1)) - 1).max() < 0.04
    # the velocity is better known at the end than after the first update (initial sigma 0.2 m/d)
    assert np.median(tracks.sigmas[:, -1, 3]) < np.median(tracks.sigmas[:, 1, 3]) < 0.2
    # counter-based draws: the same seed reproduces the run bit for bit, whatever the batching of points
    os.environ["GB_STREAM_SLOTS"], os.environ["GB_STREAM_BATCH"] = "2", "333"
    try:
        again = gb.Tracker(observers, seed=20260101).track(models, tile_size=scene.tile_size)
    finally:
        del os.environ["GB_STREAM_SLOTS"], os.environ["GB_STREAM_BATCH"]
    np.testing.assert_array_equal(again.means, tracks.means)
    np.testing.assert_array_equal(again.sigmas, tracks.sigmas)


@pytest.mark.parametrize("config", [3, 4])
def test_configs_3_and_4_full_size(cuda, config):
    """BASELINE.json configs 3 (CylindricalMotion, uncertain elevation, 2 observers, 10 000 points x 10 000 particles x
    100 frames) and 4 (31 x 31 template, ~100 px search windows, 1 000 points x 100 000 particles x 50 frames) at full
    size: the same size-independent properties as config 2 (tools/full_size_configs.py prints the rates)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import full_size_configs as fs

    import glimpse_b200 as gb
    from glimpse_b200 import synthetic

    scene = fs.scene_for(config)
    observers, models = synthetic.build(scene, gb)
    tracker = gb.Tracker(observers, seed=20260100 + config)
    tracks = tracker.track(models, tile_size=scene.tile_size)
    P, T = len(scene.points), len(scene.datetimes)
    assert tracks.means.shape == (P, T, 6) and tracks.sigmas.shape == (P, T, 6)
    assert all(e is None for e in tracks.errors)
    assert np.isfinite(tracks.means).all() and np.isfinite(tracks.sigmas).all()
    dv = np.abs(tracks.vxyz[:, -1, 0] - scene.truth_velocity[0])
    assert np.median(dv) < 0.01 and dv.max() < 0.05
    travelled = tracks.means[:, -1, 0] - tracks.means[:, 0, 0]
    assert np.abs(travelled / (scene.truth_velocity[0] * (T - 1)) - 1).max() < 0.04
    assert np.median(tracks.sigmas[:, -1, 3]) < np.median(tracks.sigmas[:, 1, 3]) < 0.3
    win = tracker.last_run["window_width"]
    if config == 4:  # the large-tile stress really has large windows (and none exceeded the plan)
        assert np.median(win) >= 50 and win.max() >= 100
    tracker.clear_device_cache()
