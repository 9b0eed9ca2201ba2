"""Regenerates tests/golden/orora_golden.npz: small seeded inputs with the outputs of the CPU oracle
(oracle/solver_ref.c; oracle/frontend_ref.py driving the real cv2 4.13).  The reference ships no
fixtures for this path, so these pin OUR restatement against regressions and let the
GPU box check the CUDA path against vectors produced in the build container.
Run from the repo root:  python tests/golden/make_orora_golden.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import frontend_ref as F, pyoracle as O  # noqa: E402

synth = importlib.import_module("navtech-radar-slam_b200.synth")


def main():
    O.build()
    out = {}
    for tag, (N, rho, seed) in {"a": (150, 0.6, 31), "b": (420, 0.45, 32)}.items():
        prev, curr, inl, T = synth.correspondences(N, rho, seed)
        r = O.orora_solve(curr, prev)
        out[f"sol_{tag}_prev"], out[f"sol_{tag}_curr"] = prev, curr
        out[f"sol_{tag}_T"], out[f"sol_{tag}_clique"] = r["T"], r["clique"]
        out[f"sol_{tag}_final"] = r["final_inliers"]
        out[f"sol_{tag}_meta"] = np.array([r["info"].n_edges, r["info"].M, r["info"].gnc_iters, r["info"].n_rot_inliers])
    rng = np.random.default_rng(33)
    fft = (rng.integers(0, 256, (48, 640)) / 255.0).astype(np.float32)
    fft[:, :42] = 0
    for a in range(0, 48, 5):
        c = int(rng.integers(80, 600))
        fft[a:a + 2, c:c + 4] = 0.95
    out["cen_fft_u8"] = np.rint(fft * 255).astype(np.uint8)
    out["cen_targets"] = F.cen2019features(fft, 400, 58)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "orora_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
