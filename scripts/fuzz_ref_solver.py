"""Randomised comparison of oracle/solver_ref.c with the reference's compiled solver (oracle/_ref/libref_solver.so):
python scripts/fuzz_ref_solver.py SEED N < /dev/null -- prints every mismatch and a count to stderr.  Measured in round 2:
9900 cases (N 3..260, outlier ratios 0..85 %, duplicates, near-collinear clouds, far points, heavy noise, both parameter
sets): 0 mismatches in edge count, clique, rotation-inlier count, final inliers; poses within 1e-9."""
import importlib, numpy as np, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import refsolver as RS, pyoracle as O
synth = importlib.import_module("navtech-radar-slam_b200.synth")
rng = np.random.default_rng(int(sys.argv[1]))
bad = 0
for it in range(int(sys.argv[2])):
    N = int(rng.integers(3, 260))
    rho = float(rng.uniform(0.15, 1.0))
    prev, curr, inl, T = synth.correspondences(N, rho, int(rng.integers(0, 1 << 30)))
    mode = it % 5
    if mode == 1:   # duplicates
        k = rng.integers(0, N, N // 4); prev[:, :len(k)] = prev[:, k]; curr[:, :len(k)] = curr[:, k]
    if mode == 2:   # near-collinear
        prev[1] *= 1e-3; curr[1] *= 1e-3
    if mode == 3:   # far points
        prev *= 3.0; curr *= 3.0
    if mode == 4:   # heavy noise
        curr = curr + rng.normal(0, 0.4, curr.shape)
    P = O.OroraParams(float(rng.choice([0.5, 0.75, 1.0])), 1.0, int(rng.choice([50, 100])), float(rng.choice([1.39, 1.4])),
                      float(rng.choice([1e-6, 1e-4])), float(rng.choice([0.1, 0.3536])), np.deg2rad(9.0))
    o = O.orora_solve(curr, prev, P)
    if o["info"].M <= 1:
        continue
    r = RS.solve(curr, prev, P)
    ok = (r["n_edges"] == o["info"].n_edges and np.array_equal(r["clique"], o["clique"]) and r["n_rot_inliers"] == o["info"].n_rot_inliers
          and np.array_equal(r["final_inliers"], o["final_inliers"]) and np.abs(r["T"] - o["T"]).max() < 1e-9 and r["valid"] == o["valid"])
    if not ok:
        bad += 1
        print("MISMATCH it", it, "N", N, "mode", mode, "edges", r["n_edges"], o["info"].n_edges, "clique", len(r["clique"]), len(o["clique"]),
              np.array_equal(r["clique"], o["clique"]), "rot", r["n_rot_inliers"], o["info"].n_rot_inliers, "final", len(r["final_inliers"]),
              len(o["final_inliers"]), "dT", np.abs(r["T"] - o["T"]).max(), file=sys.stderr)
print("done, mismatches:", bad, file=sys.stderr)
