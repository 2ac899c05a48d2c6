"""Generate golden vectors by running the UNMODIFIED reference (ezwelty/glimpse) in the build container.

Usage (build container only; the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  The reference is imported through ``oracle/ref_shim.py`` (stubs for
the six third-party modules that are absent here and never executed on this path).  Scenes come
from ``glimpse_b200.synthetic`` (seeded), so the fixtures hold only the seed/config plus the
reference's outputs and intermediates, captured by wrapping:

* ``Tracker.extract_tile``                      -> template / search tiles (tracker.py:494-534)
* ``Observer.sample_tile``                      -> uv, SSE surface, SSE box, sampled SSE (observer.py:178-214)
* ``Tracker.resample_particles`` + ``np.searchsorted`` + ``np.random.random`` -> evolved particles,
  weights, the uniform draw and the resampled indices (tracker.py:151-223)
"""
import datetime
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_shim import import_reference  # noqa: E402

glimpse = import_reference()
from glimpse_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_reference(scene, seed, points=None, return_covariances=False, observer_mask=None, viewshed=None,
                  datetimes=None, capture=True, resample_method="systematic", highpass=None, interpolation=None):
    observers, models = synthetic.build(scene, glimpse, points=points)
    extra = {} if highpass is None else {"highpass": highpass}
    if interpolation is not None:
        extra["interpolation"] = interpolation
    tracker = glimpse.Tracker(observers, viewshed=viewshed, resample_method=resample_method, **extra)
    steps = []  # one dict per resample call (point-major, time-minor)
    templates = []
    current = {"obs": {}}

    orig_extract = tracker.extract_tile

    def extract_tile(obs, img, box, histogram=None, return_histogram=False):
        out = orig_extract(obs=obs, img=img, box=box, histogram=histogram, return_histogram=return_histogram)
        if return_histogram:
            tile, hist = out
            templates.append({"obs": obs, "img": img, "box": np.asarray(box), "tile": tile.copy(),
                              "values": hist[0].copy(), "quantiles": hist[1].copy()})
        else:
            current["obs"].setdefault(obs, {}).update({"box": np.asarray(box).copy(), "search": out.copy()})
        return out

    tracker.extract_tile = extract_tile

    orig_sample = glimpse.Observer.sample_tile

    def sample_tile(self, uv, tile, box, grid=False, **kwargs):
        out = orig_sample(self, uv, tile=tile, box=box, grid=grid, **kwargs)
        obs = tracker.observers.index(self)
        current["obs"].setdefault(obs, {}).update(
            {"uv": np.array(uv), "sse": np.array(tile), "sse_box": np.array(box), "sampled": np.array(out)})
        return out

    glimpse.Observer.sample_tile = sample_tile

    orig_resample = tracker.resample_particles
    orig_random = np.random.random
    orig_search = np.searchsorted

    def resample_particles(method=None):
        rec = {"evolved": tracker.particles.copy(), "weights": tracker.weights.copy(), "obs": current["obs"]}

        def random(*a):
            rec["u"] = orig_random(*a)
            return rec["u"]

        def searchsorted(a, v, *args, **kw):
            out = orig_search(a, v, *args, **kw)
            rec["indices"] = np.array(out)
            return out

        orig_choice = np.random.choice

        def choice(a, size=None, replace=True, p=None):
            state = np.random.get_state()
            out = orig_choice(a, size=size, replace=replace, p=p)
            replay = np.random.RandomState()
            replay.set_state(state)
            rec["u"] = replay.random_sample(size)  # what RandomState.choice drew (mtrand.pyx)
            rec["indices"] = np.array(out)
            return out

        orig_hstack = np.hstack

        def hstack(tup, *a, **kw):  # residual resampling: the ancestors are hstack((repeated, searched))
            out = orig_hstack(tup, *a, **kw)
            rec["indices"] = np.array(out)
            return out

        np.random.random, np.searchsorted, np.random.choice = random, searchsorted, choice
        if (method or tracker.resample_method) == "residual":
            np.hstack = hstack
        try:
            orig_resample(method)
        finally:
            np.random.random, np.searchsorted, np.random.choice = orig_random, orig_search, orig_choice
            np.hstack = orig_hstack
        steps.append(rec)
        current["obs"] = {}

    tracker.resample_particles = resample_particles
    np.random.seed(seed)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            kwargs = dict(tile_size=scene.tile_size, return_particles=True, return_covariances=return_covariances)
            if observer_mask is not None:
                kwargs["observer_mask"] = observer_mask
            if datetimes is not None:
                kwargs["datetimes"] = datetimes
            tracks = tracker.track(models, **kwargs)
    finally:
        glimpse.Observer.sample_tile = orig_sample
    return tracks, steps, templates, tracker


def pack(prefix, d, out):
    for key, val in d.items():
        if isinstance(val, dict):
            pack(f"{prefix}{key}.", val, out)
        else:
            out[f"{prefix}{key}"] = np.asarray(val)


def save_track_case(name, scene_kwargs, seed, builder=synthetic.nadir_scene, post=None, **run_kwargs):
    scene = builder(**scene_kwargs)
    if post is not None:
        scene = post(scene)
    tracks, steps, templates, tracker = run_reference(scene, seed, **run_kwargs)
    out = {
        "seed": seed,
        "means": tracks.means,
        "particles": tracks.particles,
        "weights": tracks.weights,
        "n_steps": len(steps),
        "n_templates": len(templates),
        "error_types": np.array([type(e).__name__ if e is not None else "" for e in np.atleast_1d(tracks.errors)]),
        "images": np.array([[(-1 if v is None else v) for v in row] for row in tracks.images]),
        "frame_crc": np.array([int(np.asarray(f, dtype=np.int64).sum()) for o in scene.observers for f in o.frames]),
    }
    if tracks.sigmas is not None:
        out["sigmas"] = tracks.sigmas
    if tracks.covariances is not None:
        out["covariances"] = tracks.covariances
    for i, s in enumerate(steps):
        pack(f"step{i}.", s, out)
    for i, t in enumerate(templates):
        pack(f"template{i}.", t, out)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "steps", len(steps), "templates", len(templates), "errors", out["error_types"],
          os.path.getsize(path) // 1024, "KiB")
    return scene, tracks


def save_shape_case(name, scene_kwargs, seed, builder=synthetic.nadir_scene, post=None, **run_kwargs):
    """Compact fixture of a run at one of BASELINE.json's shapes: the reference's means / sigmas, every update's ancestor
    indices (int32) and uniform draw, checksums of its evolved particles and weights (enough to pin the oracle bit for bit;
    the tests regenerate the full intermediates with the oracle), and the SSE surface of point 0's last update."""
    scene = builder(**scene_kwargs)
    if post is not None:
        scene = post(scene)
    tracks, steps, templates, tracker = run_reference(scene, seed, **run_kwargs)
    out = {
        "seed": seed, "means": tracks.means, "n_steps": len(steps), "n_templates": len(templates),
        "error_types": np.array([type(e).__name__ if e is not None else "" for e in np.atleast_1d(tracks.errors)]),
        "images": np.array([[(-1 if v is None else v) for v in row] for row in tracks.images]),
        "frame_crc": np.array([int(np.asarray(f, dtype=np.int64).sum()) for o in scene.observers for f in o.frames]),
        "final_particles_sum": np.nansum(tracks.particles[:, -1], axis=1),
        "indices": np.stack([s["indices"] for s in steps]).astype(np.int32),
        "u": np.array([s["u"] for s in steps], dtype=float),
        "evolved_sum": np.stack([s["evolved"].sum(axis=0) for s in steps]),
        "weights_sum": np.array([s["weights"].sum() for s in steps]),
        "weights_max": np.array([s["weights"].max() for s in steps]),
    }
    if tracks.sigmas is not None:
        out["sigmas"] = tracks.sigmas
    if tracks.covariances is not None:
        out["covariances"] = tracks.covariances
    P = len(scene.points)
    per = len(steps) // P
    last = steps[per - 1]
    for o, rec in last["obs"].items():
        for key in ("box", "sse", "sse_box"):
            out[f"last0.obs.{o}.{key}"] = np.asarray(rec[key])
    for i, t in enumerate(templates):
        out[f"template{i}.obs"], out[f"template{i}.box"] = np.asarray(t["obs"]), np.asarray(t["box"])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "steps", len(steps), "errors", out["error_types"], os.path.getsize(path) // 1024, "KiB")


# ---- scene definitions shared with tests/scenes.py (the tests rebuild the same scenes) ----
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402


def camera_vectors():
    """Projection / inverse-projection known answers from the reference Camera."""
    rng = np.random.RandomState(7)
    out = {}
    configs = scenes.camera_configs()
    for name, kw in configs.items():
        corr = kw.pop("correction", False)
        cam = glimpse.Camera(correction=corr, **kw)
        # world points in front of the camera, spread over and a bit beyond the frame
        uv = rng.rand(64, 2) * cam.imgsz * 1.2 - cam.imgsz * 0.1
        depth = 200 + rng.rand(64) * 3000
        if any(cam.k) or any(cam.p):
            uv = rng.rand(64, 2) * cam.imgsz
        xyz = cam.uv_to_xyz(uv, directions=False, depth=depth)
        xyz[-4:] = cam.xyz - (xyz[-4:] - cam.xyz)  # behind the camera -> NaN
        out[f"{name}.vector"] = cam._vector.copy()
        out[f"{name}.corr"] = np.array([cam.correction["radius"], cam.correction["refraction"]]) if corr else np.zeros(0)
        out[f"{name}.xyz"] = xyz
        out[f"{name}.uv"] = cam.xyz_to_uv(xyz)
        out[f"{name}.uv_in"] = uv
        out[f"{name}.dirs"] = cam.uv_to_xyz(uv)
        out[f"{name}.R"] = cam.R
    np.savez_compressed(os.path.join(OUT, "camera.npz"), **out)
    print("camera", len(configs), "configs")


def project_image_vectors():
    """``Image.project`` outputs of the reference for the cases of ``scenes.project_image_cases``."""
    out = {}
    for name, (dtype, bands, method, src, dst) in scenes.project_image_cases().items():
        def cam_of(v):
            return glimpse.Camera(imgsz=tuple(int(x) for x in v[6:8]), f=tuple(v[8:10]), c=tuple(v[10:12]), k=tuple(v[12:18]),
                                  p=tuple(v[18:20]), xyz=tuple(v[0:3]), viewdir=tuple(v[3:6]))
        img = glimpse.Image(name, cam=cam_of(src), datetime=synthetic.T0)
        img.array = scenes.project_image_frame(name)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[name] = img.project(cam_of(dst), method=method)
    path = os.path.join(OUT, "project_image.npz")
    np.savez_compressed(path, **out)
    print("project_image", len(out), "cases", os.path.getsize(path) // 1024, "KiB")


def viewshed_vectors():
    """``Raster.viewshed`` outputs of the reference for the cases of ``scenes.viewshed_cases`` (bit-packed)."""
    out = {}
    for name, case in scenes.viewshed_cases().items():
        raster = glimpse.Raster(case["z"].copy(), x=case["xlim"], y=case["ylim"])
        corr = case["correction"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vis = raster.viewshed(case["origin"], correction=dict(radius=corr[0], refraction=corr[1]) if corr else False)
        out[name] = np.packbits(vis)
        print("  ", name, vis.shape, "visible %.3f" % vis.mean())
    case = scenes.viewshed_large_case()
    raster = glimpse.Raster(case["z"].copy(), x=case["xlim"], y=case["ylim"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vis = raster.viewshed(case["origin"], correction=dict(radius=case["correction"][0], refraction=case["correction"][1]))
    out["large_2000"] = np.packbits(vis)
    print("  ", "large_2000", vis.shape, "visible %.4f" % vis.mean())
    path = os.path.join(OUT, "viewshed.npz")
    np.savez_compressed(path, **out)
    print("viewshed", len(out), "cases", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    only = sys.argv[1:]  # optional: names of the track cases to (re)generate
    if not only:
        camera_vectors()
    if not only or "viewshed" in only:
        viewshed_vectors()
    if not only or "project_image" in only:
        project_image_vectors()
    for name, case in scenes.track_cases().items():
        if not only or name in only:
            save_track_case(name, **case)
    for name, case in scenes.shape_cases().items():
        if not only or name in only:
            save_shape_case(name, **case)
