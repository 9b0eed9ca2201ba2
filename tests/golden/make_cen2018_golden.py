"""Regenerates tests/golden/cen2018_golden.npz: a small seeded polar frame with the per-pixel response and the
keypoints of the CPU oracle for cen2018features (oracle/frontend_ref.c loops with glibc, the Gaussian filter2D by
the real cv2 4.13).  The reference ships no fixtures for this path; the vectors pin OUR restatement
against regressions and let the GPU box check the CUDA path against values produced in the build container.
Run from the repo root:  python tests/golden/make_cen2018_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import frontend_ref as F, pyoracle as O  # noqa: E402


def main():
    O.build()
    rng = np.random.default_rng(2018)
    out = {}
    for tag, (rows, cols, zq, sg, mr) in {"a": (24, 1000, 3.0, 17, 58), "b": (10, 333, 2.0, 7, 12)}.items():
        fft = (rng.integers(0, 24, (rows, cols)) / 255.0).astype(np.float32)
        hot = rng.random((rows, cols)) < 0.025
        fft[hot] = (rng.integers(110, 256, int(hot.sum())) / 255.0).astype(np.float32)
        fft[:, :42] = 0
        st = F.cen2018_stages(fft, zq, sg, mr)
        out[f"{tag}_fft"] = fft
        out[f"{tag}_par"] = np.array([zq, sg, mr], dtype=np.float64)
        out[f"{tag}_y"] = st["y"]
        out[f"{tag}_sigma"] = st["sigma_q"]
        out[f"{tag}_targets"] = st["targets"]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cen2018_golden.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
