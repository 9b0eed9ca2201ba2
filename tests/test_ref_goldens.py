"""The committed golden vectors (tests/golden/*.npz) are what the GPU suite checks the CUDA path against on the GPU box
(test_sc_gpu / test_solver_gpu / test_frontend_gpu ::*golden*).  They were written by the oracle; here every expected
value in them is REPRODUCED BY THE REFERENCE'S OWN COMPILED CODE (oracle/_ref/*.so, see tests/test_ref_*.py), so a GPU
result that equals a golden vector equals what the reference computes.  Runs on CPU; skipped where the libraries are absent."""
import importlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RSC = importlib.import_module("oracle.refsc")
RSO = importlib.import_module("oracle.refsolver")
RFE = importlib.import_module("oracle.refbridge")


@pytest.mark.skipif(not RSC.available(), reason="oracle/_ref/libref_sc.so not built")
def test_sc_golden_is_what_the_reference_computes():
    g = np.load(os.path.join(GOLD, "sc_golden.npz"))
    m = RSC.SCManager()
    for tag in ("bin", "gen"):
        db, q = g[f"{tag}_db"], g[f"{tag}_query"]
        best = (1e7, 0, 0)
        for i in range(95):
            d, s = m.distanceBtnScanContext(q, db[i])
            assert np.float64(d).view(np.uint64) == g[f"{tag}_dist"][i].view(np.uint64) and s == g[f"{tag}_shift"][i], (tag, i)
            if d < best[0]:
                best = (d, i, s)
        assert tuple(best) == tuple(g[f"{tag}_best"])
        assert np.array_equal(m.keys(db[0])[1], g[f"{tag}_sectorkey0"])
        k = RSC.SCManager()
        for i in range(len(db)):
            k.saveScancontextAndKeys(db[i])
        assert np.array_equal(np.stack([k.get(i)[1] for i in range(len(db))]), g[f"{tag}_ringkeys_f32"])
        k.close()
    assert np.array_equal(m.makeScancontext(g["cloud"]), g["cloud_desc"])
    m.close()


@pytest.mark.skipif(not RSO.available(), reason="oracle/_ref/libref_solver.so not built")
def test_solver_golden_is_what_the_reference_computes():
    g = np.load(os.path.join(GOLD, "orora_golden.npz"))
    for tag in ("a", "b"):
        r = RSO.solve(g[f"sol_{tag}_curr"], g[f"sol_{tag}_prev"])
        assert np.array_equal(r["clique"], g[f"sol_{tag}_clique"]) and np.array_equal(r["final_inliers"], g[f"sol_{tag}_final"])
        n_edges, M, _, n_rot = (int(v) for v in g[f"sol_{tag}_meta"])
        assert (r["n_edges"], len(r["clique"]), r["n_rot_inliers"]) == (n_edges, M, n_rot)
        assert np.abs(r["T"] - g[f"sol_{tag}_T"]).max() < 1e-12


@pytest.mark.skipif(not RFE.available(), reason="oracle/_ref/libref_frontend.so not built")
def test_frontend_goldens_are_what_the_reference_computes():
    g = np.load(os.path.join(GOLD, "cen2018_golden.npz"))
    for tag in ("a", "b"):
        zq, sg, mr = g[f"{tag}_par"]
        assert np.array_equal(RFE.cen2018features(g[f"{tag}_fft"], float(zq), int(sg), int(mr)), g[f"{tag}_targets"])
    g = np.load(os.path.join(GOLD, "orora_golden.npz"))
    fft = (g["cen_fft_u8"].astype(np.float64) / 255.0).astype(np.float32)
    # cen2019 golden (written with the oracle's stable tie order): on this frame the reference's std::sort order gives the same keypoints
    assert np.array_equal(RFE.cen2019features(fft, 400, 58), g["cen_targets"])
