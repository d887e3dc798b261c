import glob
import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sequence_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "seq_*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def check_sequence_against_golden(g, frames_results):
    """frames_results: list of dict(initial, blocks, motion, unique[, planes]) per frame."""
    hashes = g["hashes"]
    for i, r in enumerate(frames_results):
        assert sha(r["initial"]) == hashes[i][0], f"initial blocks differ in frame {i}"
        assert sha(r["motion"]) == hashes[i][2], f"motion differs in frame {i}"
        assert sha(r["blocks"]) == hashes[i][1], f"final blocks differ in frame {i}"
        assert sha(r["unique"]) == hashes[i][3], f"unique palette differs in frame {i}"
        if hashes.shape[1] > 4 and "planes" in r:
            assert sha(r["planes"]) == hashes[i][4], f"endpoint planes differ in frame {i}"
        if f"final_{i}" in g:
            assert np.array_equal(r["blocks"], g[f"final_{i}"])
            assert np.array_equal(r["motion"], g[f"motion_{i}"])


# ---- full-size fixtures of the benchmarked configurations (tests/golden/gen_golden_full.py) ----------
def full_fixture_name(w, h, sa, thr, gop, seed=1234):
    """Name of the committed full-GOP reference fixture that covers this configuration, or None."""
    for p in sorted(glob.glob(os.path.join(GOLDEN, "full_*.npz"))):
        g = np.load(p)
        gw, gh, _n, gseed, gsa, gthr, ggop = [int(x) for x in g["params"]]
        if (gw, gh, gseed, gsa, gthr, ggop) == (w, h, seed, sa, thr, gop):
            return os.path.basename(p)[:-4]
    return None


def compare_with_full_fixture(g, blocks, motion, unique, n_unique, initial=None, first_frame=0):
    """Compares encoder outputs (arrays indexed [frame][...], frame i = sequence frame first_frame + i)
    with the reference's per-frame SHA-256 hashes.  Returns (frames compared, list of mismatch
    descriptions); a mismatching frame is located down to its first differing block row through
    the fixture's per-row CRC-32."""
    import zlib
    hashes, n_fix = g["hashes"], int(g["params"][2])
    bh = int(g["params"][1]) // 4
    bad = []
    n = 0
    for i in range(len(blocks)):
        f = first_frame + i
        if f >= n_fix:
            break
        n += 1
        nu = int(n_unique[i])
        got = {"final blocks": (1, blocks[i]), "motion": (2, motion[i]), "unique palette": (3, unique[i][:nu])}
        if initial is not None:
            got["initial blocks"] = (0, initial[i])
        for what, (col, arr) in got.items():
            if sha(arr) == hashes[f][col]:
                continue
            where = ""
            rows = {"final blocks": "rows_final", "motion": "rows_motion"}.get(what)
            if rows is not None:
                a = np.ascontiguousarray(arr).reshape(bh, -1)
                crc = np.array([zlib.crc32(r.tobytes()) for r in a], dtype=np.uint32)
                diff = np.nonzero(crc != g[rows][f])[0]
                where = f" (first differing block row {int(diff[0])}, {diff.size} rows differ)" if diff.size else ""
            elif what == "unique palette":
                where = f" ({nu} unique words, reference {int(g['n_unique'][f])})"
            bad.append(f"frame {f}: {what} differ from the reference{where}")
    return n, bad
