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
