"""Golden vectors for Tracks.from_multiple / Tracks.average from the UNMODIFIED reference (build container only).

    python tests/golden/make_tracks_golden.py   ->  tests/golden/tracks_merge.npz
"""
import os, sys, datetime
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference
glimpse = import_reference()
rng = np.random.RandomState(3)
P, T = 4, 7
dts = [datetime.datetime(2020, 1, 1) + datetime.timedelta(days=i) for i in range(T)]
day = datetime.timedelta(days=1)
def run():
    m = rng.randn(P, T, 6); s = 0.1 + rng.rand(P, T, 6)
    return m, s
(m1, s1), (m2, s2) = run(), run()
m1[0, 2] = np.nan; s1[0, 2] = np.nan           # one run misses a time
m1[1, 4] = np.nan; s1[1, 4] = np.nan; m2[1, 4] = np.nan; s2[1, 4] = np.nan   # both miss it
out = {"m1": m1, "s1": s1, "m2": m2, "s2": s2}
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for ign in (False, True):
        a = glimpse.Tracks(datetimes=dts, time_unit=day, means=m1, sigmas=s1)
        b = glimpse.Tracks(datetimes=dts, time_unit=day, means=m2, sigmas=s2)
        merged = glimpse.Tracks.from_multiple([a, b], ignore_nan=ign)
        out[f"merged_means_{int(ign)}"], out[f"merged_sigmas_{int(ign)}"] = merged.means, merged.sigmas
        am, asg = a.average(ignore_nan=ign)
        out[f"avg_means_{int(ign)}"], out[f"avg_sigmas_{int(ign)}"] = am, asg
np.savez_compressed(os.path.join(ROOT, "tests/golden/tracks_merge.npz"), **out)
print("ok", {k: v.shape for k, v in out.items()})
