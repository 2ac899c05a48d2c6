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
