import importlib, time, numpy as np, ctypes as C, sys
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("navtech-radar-slam_b200")
synth = importlib.import_module("navtech-radar-slam_b200.synth")
ctx = pkg.Context(0)
orr = pkg.orora
world = synth.radar_world(seed=1234)
poses = [(2.5 * k, 0.1 * k, np.deg2rad(1.0) * k) for k in range(8)]
loader = orr.RadarLoader("mulran", "cen2019", 0.6, True, False, ctx=ctx)
feats = [loader.extractFeatureFromRows(synth.radar_frame(world, poses[k], seed=1234 + k), k) for k in range(8)]
p = orr.default_params()
vel = 0.0
for k in range(1, 8):
    pm, cm = loader.matchFeatures(feats[k-1], feats[k])
    for v in (0.0, vel):
        t0 = time.perf_counter()
        r = orr.solve(ctx, pm.feat_in_cart_, cm.feat_in_cart_, p, v)
        dt = time.perf_counter() - t0
        i = r["info"]
        print(k, "vel", round(v,2), "N", pm.feat_in_cart_.shape[1], "edges", i.n_edges, "M", i.M, "nodes", i.clique_nodes, "exact", i.clique_exact, "iters", i.gnc_iters, f"{dt*1e3:.2f} ms", "T", np.round(r["T"][:2,3],3))
    vel = r["T"][0,3] / 0.25
