"""The production kernels against the reference's own outputs (tests/golden/track_*.npz) and the oracle.

Teacher-forced: every update of every golden case is replayed through the production step kernel
(``gb_track_step``) with the reference's evolved particles (and, for resampling, its weights and
uniform draw) forced in, and each intermediate is compared at its own tolerance.
Free-running: ``Tracker.track(rng='numpy')`` with the reference's draw sequence end to end.
"""
import ctypes as C
import warnings

import numpy as np
import pytest

import helpers
import scenes
from glimpse_b200 import synthetic
from oracle import tracker_oracle as orc

pytestmark = pytest.mark.gpu

# tolerances (fp64 unless noted)
UV_TOL_PX = 1e-9          # projection
TILE_TOL = 1e-12          # template tile / CDF values (z-scores of order 1)
SEARCH_TOL = 2e-7         # float32 search tile vs the reference's float64 tile cast to float32 (1 ulp of ~2)
SSE_REL_CV2 = 5e-5        # vs OpenCV's float32 DFT-based matchTemplate (its own noise)
SSE_REL_EXACT = 2e-6      # vs exact float64 SSD of the same tiles (fp32 accumulation of <= 961 terms)
SAMPLED_ABS = 2e-5        # spline-sampled SSE vs FITPACK on OpenCV's surface
WEIGHT_REL = 3e-4         # exp(-ll): |d ll| <= SAMPLED_ABS / (2 sigma^2)
MOMENT_REL = 1e-9


def case_scene(name):
    case = scenes.track_cases()[name]
    scene = synthetic.nadir_scene(**case["scene_kwargs"])
    if case.get("post"):
        scene = case["post"](scene)
    return scene, case


CASES = list(scenes.track_cases())
# the tangent motion models run in the default streaming organisation only
# ... and so does stratified resampling (rank 2)
CASE_MODES = [(n, m) for n in CASES for m in ("stream", "fused") if not (m == "fused" and ("tangent" in n or "stratified" in n or "choice" in n))]


def make_session(scene, case, golden, return_particles=False, cluster=0, tile_bytes=None, mode="stream"):
    import glimpse_b200 as gb
    from glimpse_b200.session import Session, reference_order_draws

    observers, models = synthetic.build(scene, gb)
    method = case.get("resample_method", "systematic")
    tracker = gb.Tracker(observers, rng="numpy", cluster=cluster, mode=mode, resample_method=method,
                         highpass=case.get("highpass", {"size": (5, 5)}), interpolation=case.get("interpolation", {"kx": 3, "ky": 3}))
    datetimes = tracker.datetimes
    matching = tracker.match_datetimes(datetimes)
    image_index = np.array([[-1 if v is None else int(v) for v in row] for row in matching], dtype=np.int32)
    np.testing.assert_array_equal(image_index, golden["images"])
    unit = scene.time_unit.total_seconds()
    taus = np.array([dt.total_seconds() / unit for dt in np.diff(datetimes)])
    P = len(models)
    mask = np.ones((P, len(observers)), dtype=bool)
    from glimpse_b200.session import point_span

    first, last = point_span(image_index, mask)
    np.random.seed(int(golden["seed"]))
    tangent = np.full(P, scene.motion["kind"].startswith("tangent"))
    draws = reference_order_draws(P, scene.n_particles, last - first, tangent, method in ("stratified", "choice"))
    session = Session(tracker, models, image_index, taus, scene.tile_size, mask,
                      return_covariances=bool(case.get("return_covariances", False)),
                      return_particles=return_particles, draws=draws)
    if tile_bytes is not None:  # shrink the on-chip tile capacity: every search window overflows to the slabs
        session.plan.tile_bytes = tile_bytes
        session.desc.plan.tile_bytes = tile_bytes
    return session, draws


def golden_steps(g, P, T_steps):
    """step records are point-major, time-minor: index = p * steps_per_point + s."""
    n = int(g["n_steps"])
    per = n // P
    return per


@pytest.mark.parametrize("name,mode", CASE_MODES)
def test_templates_match_reference(cuda, name, mode):
    scene, case = case_scene(name)
    g = helpers.load_golden(name)
    session, _ = make_session(scene, case, g, mode=mode)
    # run the whole track so that staggered templates (observer starting later) are built too
    session.run()
    cuda.cuda.synchronize()
    P, O = session.P, session.O
    n_t = int(g["n_templates"])
    assert n_t == P * O
    k = 0
    for p in range(P):
        tmpl = session.templates(p)
        for _ in range(O):
            o = int(g[f"template{k}.obs"])
            mine = tmpl[o]
            np.testing.assert_array_equal(mine["box"], g[f"template{k}.box"])
            assert np.max(np.abs(mine["tile"] - g[f"template{k}.tile"])) <= TILE_TOL
            vals, qs = mine["histogram"]
            assert len(vals) == len(g[f"template{k}.values"])
            assert np.max(np.abs(vals - g[f"template{k}.values"])) <= TILE_TOL
            np.testing.assert_array_equal(qs, g[f"template{k}.quantiles"])
            k += 1


@pytest.mark.parametrize("name,mode,cluster,tile_bytes",
                         [(n, m, 0, None) for n, m in CASE_MODES]
                         + [("track_c1", "fused", 2, None), ("track_cyl2", "fused", 4, None), ("track_jitter", "fused", 8, None),
                            ("track_c1", "fused", 4, 2048),
                            # k_s2_surface's other work-space organisations (negative = no interleaved path): every window on
                            # planes; a budget that sends windows to the staged, planar or global path depending on their size
                            ("track_c1", "stream", 0, -112640), ("track_cyl2", "stream", 0, -112640),
                            ("track_c1", "stream", 0, -18000), ("track_cyl2", "stream", 0, -30000), ("track_jitter", "stream", 0, -16000),
                            # other median sizes on the planar / staged / global organisations
                            ("track_hp37", "stream", 0, -112640), ("track_hp4", "stream", 0, -30000), ("track_hp37", "stream", 0, -16000)])
def test_step_teacher_forced(cuda, name, mode, cluster, tile_bytes):
    from glimpse_b200 import _lib

    torch = cuda
    scene, case = case_scene(name)
    g = helpers.load_golden(name)
    session, draws = make_session(scene, case, g, cluster=cluster, tile_bytes=tile_bytes, mode=mode)
    assert cluster == 0 or session.plan.cluster == cluster
    P, N, T, O = session.P, session.N, session.T, session.O
    per = int(g["n_steps"]) // P
    dev = session.device
    # build all templates exactly as a real run would (needs the filter to advance): run it once
    session.run()
    torch.cuda.synchronize()
    assert (session.buf["status"].cpu().numpy() == 0).all()
    cap = 200 * 200
    worst = dict(uv=0.0, search=0.0, sse_cv2=0.0, sse_exact=0.0, sampled=0.0, weights=0.0, mean=0.0, sigma=0.0)

    def upd(key, values):
        v = float(np.max(values))
        assert np.isfinite(v), f"{key}: non-finite difference (missing or NaN output)"
        worst[key] = max(worst[key], v)

    for s in range(per):
        t = int(session.first[0]) + 1 + s
        evolved = np.stack([g[f"step{p * per + s}.evolved"] for p in range(P)])  # (P, N, 6)
        weights_ref = np.stack([g[f"step{p * per + s}.weights"] for p in range(P)])
        u_ref = np.stack([np.asarray(g[f"step{p * per + s}.u"], dtype=float) for p in range(P)])  # (P,) or (P, N)
        session.buf["uniforms"][:, t - 1 - int(session.first[0])] = torch.as_tensor(u_ref).to(dev)
        f_ev = torch.as_tensor(np.ascontiguousarray(evolved.transpose(0, 2, 1))).to(dev)  # (P, 6, N)
        dump = {
            "uv": torch.full((P, O, N, 2), float("nan"), dtype=torch.float64, device=dev),
            "box": torch.full((P, O, 4), -1, dtype=torch.int32, device=dev),
            "search": torch.zeros((P, O, cap), dtype=torch.float32, device=dev),
            "sse": torch.zeros((P, O, cap), dtype=torch.float32, device=dev),
            "sampled": torch.zeros((P, O, N), dtype=torch.float64, device=dev),
            "weights": torch.zeros((P, N), dtype=torch.float64, device=dev),
            "indices": torch.full((P, N), -1, dtype=torch.int32, device=dev),
        }
        io = _lib.gb_stage_io()
        io.force_evolved = f_ev.data_ptr()
        io.dump_uv, io.dump_box = dump["uv"].data_ptr(), dump["box"].data_ptr()
        io.dump_search, io.dump_sse = dump["search"].data_ptr(), dump["sse"].data_ptr()
        io.dump_sampled, io.dump_weights = dump["sampled"].data_ptr(), dump["weights"].data_ptr()
        io.dump_indices = dump["indices"].data_ptr()
        io.dump_cap = cap
        session.buf["status"].zero_()
        session.step(t, io)
        torch.cuda.synchronize()
        assert (session.buf["status"].cpu().numpy() == 0).all()
        got = {k: v.cpu().numpy() for k, v in dump.items()}
        for p in range(P):
            rec = f"step{p * per + s}."
            for o in range(O):
                key = f"{rec}obs.{o}."
                if key + "uv" not in g:
                    continue
                uv_ref = g[key + "uv"]
                upd("uv", np.abs(got["uv"][p, o] - uv_ref))
                np.testing.assert_array_equal(got["box"][p, o], g[key + "box"])
                search_ref = g[key + "search"]
                sv, su = search_ref.shape
                mine = got["search"][p, o, : sv * su].reshape(sv, su)
                upd("search", np.abs(mine - search_ref.astype(np.float32)))
                sse_ref = g[key + "sse"]
                mv, mu = sse_ref.shape
                sse = got["sse"][p, o, : mv * mu].reshape(mv, mu)
                upd("sse_cv2", np.abs(sse - sse_ref) / np.maximum(sse_ref, 1e-3))
                tmpl = g[f"template{[int(g[f'template{k}.obs']) for k in range(p * O, (p + 1) * O)].index(o) + p * O}.tile"]
                exact = orc.ssd_surface(search_ref, tmpl, exact=True)
                upd("sse_exact", np.abs(sse - exact) / np.maximum(exact, 1e-3))
                upd("sampled", np.abs(got["sampled"][p, o] - g[key + "sampled"]))
            w_ref = weights_ref[p]
            upd("weights", np.abs(got["weights"][p] - w_ref) / w_ref)
        # resampling and moments with the reference's weights forced in: indices must be bit-exact
        f_w = torch.as_tensor(weights_ref).to(dev)
        io2 = _lib.gb_stage_io()
        io2.force_evolved, io2.force_weights = f_ev.data_ptr(), f_w.data_ptr()
        io2.dump_indices = dump["indices"].data_ptr()
        session.step(t, io2)
        torch.cuda.synchronize()
        idx = dump["indices"].cpu().numpy()
        means = session.buf["means"].cpu().numpy()
        sig = session.buf["sig"].cpu().numpy()
        for p in range(P):
            np.testing.assert_array_equal(idx[p], g[f"step{p * per + s}.indices"])
            m_ref = g["means"][p, t]
            upd("mean", np.abs(means[p, t] - m_ref) / np.maximum(np.abs(m_ref), 1e-3))
            if "covariances" in g:
                c_ref = g["covariances"][p, t].ravel()
                scale = np.sqrt(np.outer(np.diag(g["covariances"][p, t]), np.diag(g["covariances"][p, t]))).ravel()
                upd("sigma", np.abs(sig[p, t] - c_ref) / np.maximum(scale, 1e-12))
            else:
                s_ref = g["sigmas"][p, t]
                ok = s_ref > 0
                upd("sigma", np.abs(sig[p, t][ok] - s_ref[ok]) / s_ref[ok])
    print(name, {k: float(v) for k, v in worst.items()})
    assert worst["uv"] <= UV_TOL_PX
    assert worst["search"] <= SEARCH_TOL
    assert worst["sse_cv2"] <= SSE_REL_CV2
    assert worst["sse_exact"] <= SSE_REL_EXACT
    assert worst["sampled"] <= SAMPLED_ABS
    assert worst["weights"] <= WEIGHT_REL
    assert worst["mean"] <= MOMENT_REL
    assert worst["sigma"] <= 1e-7


@pytest.mark.parametrize("name,mode", CASE_MODES)
def test_track_free_running_matches_reference(cuda, name, mode):
    """Whole Tracker.track with the reference's draw sequence (rng='numpy')."""
    import glimpse_b200 as gb

    scene, case = case_scene(name)
    g = helpers.load_golden(name)
    observers, models = synthetic.build(scene, gb)
    tracker = gb.Tracker(observers, rng="numpy", mode=mode, resample_method=case.get("resample_method", "systematic"),
                         highpass=case.get("highpass", {"size": (5, 5)}), interpolation=case.get("interpolation", {"kx": 3, "ky": 3}))
    np.random.seed(int(g["seed"]))
    cov = bool(case.get("return_covariances", False))
    tracks = tracker.track(models, tile_size=scene.tile_size, return_particles=True, return_covariances=cov)
    assert all(e is None for e in tracks.errors)
    # ancestors: fraction of particles that differ from the reference's after each resampling
    same = np.isclose(tracks.particles, g["particles"], rtol=0, atol=1e-9).all(axis=3)
    frac = 1.0 - same.mean()
    sig_ref = g["sigmas"] if not cov else np.sqrt(np.einsum("ptii->pti", g["covariances"]))
    d = np.abs(tracks.means - g["means"]) / np.maximum(sig_ref, 1e-12)
    d = d[..., [0, 1, 3, 4]] if name != "track_cyl2" else d
    print(name, "particle mismatch fraction", frac, "max |dmean|/sigma", np.nanmax(d))
    assert frac < 0.02
    assert np.nanmax(d) < 0.05
    if not cov:
        ok = g["sigmas"] > 0
        assert np.nanmax(np.abs(tracks.sigmas[ok] - g["sigmas"][ok]) / g["sigmas"][ok]) < 0.05


@pytest.mark.parametrize("mode,kind,method", [("stream", "cartesian", "systematic"), ("fused", "cartesian", "systematic"),
                                              ("stream", "cartesian", "stratified"), ("stream", "cartesian", "choice"),
                                              ("stream", "tangent_cartesian", "systematic")])
def test_philox_run_recovers_velocity(cuda, mode, kind, method):
    """Device RNG: the filter recovers the synthetic ground-truth velocity (0.4 m/d along +x), for both resamplers and
    for a tangent motion model on a gridded DEM."""
    import glimpse_b200 as gb

    scene = synthetic.nadir_scene(seed=9, n_points=12, n_particles=2000, n_frames=10, imgsz=(600, 400), kind=kind)
    if kind.startswith("tangent"):
        rng = np.random.RandomState(2)
        scene.motion.update(dem=dict(array=0.3 * rng.rand(10, 14), x=(-80.0, 80.0), y=(60.0, -60.0)), dem_sigma=0.2)
    observers, models = synthetic.build(scene, gb)
    make = lambda seed: gb.Tracker(observers, seed=seed, mode=mode, resample_method=method)  # noqa: E731
    tracks = make(1234).track(models, tile_size=scene.tile_size)
    assert all(e is None for e in tracks.errors)
    v = tracks.vxyz[:, -1]
    assert np.all(np.abs(v[:, 0] - scene.truth_velocity[0]) < 0.03), v
    assert np.all(np.abs(v[:, 1] - scene.truth_velocity[1]) < 0.03), v
    # determinism: same seed, same answer; other seed, other answer
    again = make(1234).track(models, tile_size=scene.tile_size)
    np.testing.assert_array_equal(again.means, tracks.means)
    other = make(99).track(models, tile_size=scene.tile_size)
    assert not np.array_equal(other.means, tracks.means)
