"""GPU: both inter-search kernels (K2) give the reference's results.

The launcher picks k_inter_search_wide (16x16 targets per CTA, mptc_inter_wide.cu) when two of its CTAs
fit an SM and the threshold fits its int16 table, else k_inter_search_tiled (8x4 targets, mptc_inter.cu);
MPTC_K2 = wide | tiled forces one (read once per process, hence the subprocesses).  Every committed
sequence fixture of the unmodified reference is run through each, plus noise content (tiles with more
than 128 distinct words: the wide kernel's hand-over to the 8x4 tile search) and a 15-frame chunk of the
benchmarked 1080p sequence, where the two must agree with each other and with the full-GOP fixture."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from golden_util import check_sequence_against_golden, compare_with_full_fixture, full_fixture_name, load, sequence_cases, sha
from mptc_b200 import capi
from mptc_b200.synth import make_frame, make_sequence
ctx = capi.Context(0)
res = {}
def encode(frames, sa, thr, gop):
    n, h, w = frames.shape[:3]
    ctx.seq_reserve(w, h, n)
    ctx.seq_upload(np.ascontiguousarray(frames))
    ctx.seq_encode(0, n, sa, thr, gop)
    return ctx.seq_download(0, n)
for name in sequence_cases():
    g = load(name)
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    out = encode(make_sequence(w, h, n, seed=seed), sa, thr, gop)
    check_sequence_against_golden(g, [{"initial": out["initial"][i], "blocks": out["blocks"][i], "motion": out["motion"][i],
                                       "unique": out["unique"][i, :int(out["n_unique"][i])]} for i in range(n)])
    res[name] = sha(out["blocks"]) + sha(out["motion"])
rng = np.random.default_rng(7)
noise = rng.integers(0, 256, (3, 256, 512, 3), dtype=np.uint8)
for sa, thr in ((16, 50), (16, 0), (8, 40000), (5, 50), (20, 10), (16, 200), (7, 5000)):
    out = encode(noise, sa, thr, 3)
    res["noise_sa%%d_thr%%d" %% (sa, thr)] = sha(out["blocks"]) + sha(out["motion"])
frames = np.stack([make_frame(1920, 1080, f) for f in range(15)])
out = encode(frames, 16, 50, 15)
g = load(full_fixture_name(1920, 1080, 16, 50, 15))
n, bad = compare_with_full_fixture(g, out["blocks"], out["motion"], out["unique"], out["n_unique"], initial=out["initial"])
assert n == 15 and not bad, bad
res["full_1080p_gop0"] = sha(out["blocks"]) + sha(out["motion"])
res["work"] = ctx.last_work_count()
print("RESULT " + json.dumps(res))
"""


def run_variant(k2):
    env = dict(os.environ)
    env.pop("MPTC_K2", None)
    if k2:
        env["MPTC_K2"] = k2
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_inter_kernels_agree_with_the_reference_and_each_other():
    tiled, wide, default = run_variant("tiled"), run_variant("wide"), run_variant(None)
    wt, ww, wd = tiled.pop("work"), wide.pop("work"), default.pop("work")
    assert tiled == wide == default
    # the variants really are different kernels: one work record per CTA tile
    assert ww["inter_tiles"] < wt["inter_tiles"] and wd["inter_tiles"] == ww["inter_tiles"]
    assert ww["inter_evals"] == wt["inter_evals"]          # the same (word, target) evaluations
