"""Test helpers: scene -> oracle inputs, golden loading, reference-order draws."""
import os

import numpy as np

from oracle import tracker_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def oracle_inputs(scene, points=None):
    """(observers, models, taus, image_index) for ``oracle.tracker_oracle.track``."""
    datetimes = scene.datetimes
    observers = [orc.ObserverSpec(o.frames, np.asarray(o.cams), o.sigma) for o in scene.observers]
    image_index = np.full((len(datetimes), len(observers)), -1, dtype=int)
    for j, o in enumerate(scene.observers):
        lookup = {d: i for i, d in enumerate(o.datetimes)}
        for t, d in enumerate(datetimes):
            image_index[t, j] = lookup.get(d, -1)
    unit = scene.time_unit.total_seconds()
    taus = np.array([dt.total_seconds() / unit for dt in np.diff(datetimes)])
    params = dict(scene.motion)
    kind = params.pop("kind")
    names = {"cartesian": ("vxyz", "axyz"), "cylindrical": ("vrthz", "arthz"), "tangent_cartesian": ("vxy", "axy"),
             "tangent_cylindrical": ("vrth", "arth")}[kind]
    pad = lambda x: tuple(x) + (0.0,) * (3 - len(x))  # noqa: E731
    kw = dict(v=pad(params[names[0]]), v_sigma=pad(params[names[0] + "_sigma"]), a=pad(params[names[1]]),
              a_sigma=pad(params[names[1] + "_sigma"]), slope_sigma=params.get("slope_sigma", 0.0))

    def surface(value):
        if isinstance(value, dict):
            return orc.Surface(value["array"], tuple(value["x"]), tuple(value["y"]))
        return orc.Surface(value)

    sel = range(len(scene.points)) if points is None else points
    models = [
        orc.MotionSpec(xy=scene.points[i], n=scene.n_particles, kind=kind, dem=surface(params["dem"]),
                       dem_sigma=surface(params["dem_sigma"]), xy_sigma=params["xy_sigma"], **kw)
        for i in sel
    ]
    return observers, models, taus, image_index


def reference_draws(seed, n_points, n_particles, n_steps_per_point, tangent=False):
    """Draws in the reference's order: per point randn(n,2), randn(n), randn(n,3), then per
    later frame randn(n,3) and one random(); the tangent models draw randn(n,2), randn(n), randn(n,2), then per
    frame randn(n,2), randn(n) and one random().  Returns (init (P, n, 6), step (P, S, n, 3), uniforms (P, S))."""
    np.random.seed(seed)
    init = np.zeros((n_points, n_particles, 6))
    step = np.empty((n_points, n_steps_per_point, n_particles, 3))
    unif = np.empty((n_points, n_steps_per_point))
    for p in range(n_points):
        init[p, :, 0:2] = np.random.randn(n_particles, 2)
        init[p, :, 2] = np.random.randn(n_particles)
        if tangent:
            init[p, :, 3:5] = np.random.randn(n_particles, 2)
        else:
            init[p, :, 3:6] = np.random.randn(n_particles, 3)
        for s in range(n_steps_per_point):
            if tangent:
                step[p, s, :, 0:2] = np.random.randn(n_particles, 2)
                step[p, s, :, 2] = np.random.randn(n_particles)
            else:
                step[p, s] = np.random.randn(n_particles, 3)
            unif[p, s] = np.random.random()
    return init, step, unif


def shape_scene(name):
    import scenes
    from glimpse_b200 import synthetic

    case = scenes.shape_cases()[name]
    scene = synthetic.nadir_scene(**case["scene_kwargs"])
    if case.get("post"):
        scene = case["post"](scene)
    return scene, case


_SHAPE_CACHE = {}


def shape_golden(name):
    """Full intermediates of a BASELINE-shape case in the key layout of the ``track_*`` goldens, regenerated with the
    oracle and pinned to the compact fixture the unmodified reference produced (``tests/golden/shape_*.npz``): ancestor
    indices, uniform draws, checksums of the evolved particles and weights, means and sigmas must be bit-identical, so
    the regenerated arrays ARE the reference's."""
    import warnings

    if name in _SHAPE_CACHE:
        return _SHAPE_CACHE[name]
    scene, case = shape_scene(name)
    small = load_golden(name)
    crc = np.array([int(np.asarray(f, dtype=np.int64).sum()) for o in scene.observers for f in o.frames])
    np.testing.assert_array_equal(crc, small["frame_crc"])
    obs, models, taus, idx = oracle_inputs(scene)
    np.testing.assert_array_equal(idx, small["images"])
    np.random.seed(int(small["seed"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = orc.track(obs, models, taus, idx, tile_size=scene.tile_size, return_particles=True,
                        return_covariances="covariances" in small, trace=True)
    steps = [s for s in res.trace if "indices" in s]
    assert len(steps) == int(small["n_steps"])
    np.testing.assert_array_equal(np.stack([s["indices"] for s in steps]), small["indices"])
    np.testing.assert_array_equal(np.array([s["u"] for s in steps]), small["u"])
    np.testing.assert_array_equal(np.stack([s["evolved"].sum(axis=0) for s in steps]), small["evolved_sum"])
    np.testing.assert_array_equal(np.array([s["weights"].sum() for s in steps]), small["weights_sum"])
    np.testing.assert_array_equal(np.array([s["weights"].max() for s in steps]), small["weights_max"])
    np.testing.assert_array_equal(res.means, small["means"])
    np.testing.assert_array_equal(res.sigmas, small["covariances"] if "covariances" in small else small["sigmas"])
    np.testing.assert_array_equal(np.nansum(res.particles[:, -1], axis=1), small["final_particles_sum"])
    per = len(steps) // len(models)
    for o, rec in steps[per - 1]["obs"].items():
        np.testing.assert_array_equal(rec["sse"], small[f"last0.obs.{o}.sse"])
        np.testing.assert_array_equal(rec["box"], small[f"last0.obs.{o}.box"])
    g = {k: small[k] for k in small.files}
    g["particles"], g["weights"] = res.particles, res.weights
    for i, s in enumerate(steps):
        for key in ("evolved", "weights", "u", "indices"):
            g[f"step{i}.{key}"] = np.asarray(s[key])
        for o, rec in s.get("obs", {}).items():
            for key, val in rec.items():
                g[f"step{i}.obs.{o}.{key}"] = np.asarray(val)
    assert len(res.templates) == int(small["n_templates"])
    for i, t in enumerate(res.templates):
        np.testing.assert_array_equal(t["box"], small[f"template{i}.box"])
        for key in ("obs", "box", "tile", "values", "quantiles"):
            g[f"template{i}.{key}"] = np.asarray(t[key])
    _SHAPE_CACHE[name] = g
    return g
