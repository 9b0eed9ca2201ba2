"""python scripts/run_odometry.py <dir with polar_oxford_form PNGs> [--dataset mulran|oxford] [--extractor cen2018|cen2019]
[--loops] [--out DIR]   or   python scripts/run_odometry.py --seq-dir SEQ ... (the node's own layout: SEQ/polar_oxford_form,
SEQ/gt/radar_odometry.csv, SEQ/outputs): the ROS-free counterpart of `roslaunch orora run_orora.launch` (+ the ScanContext
loop detector).  The C++ form of the same driver is navtech-radar-slam_b200/cpp/orora_odom."""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("datadir", nargs="?", default=None)
    ap.add_argument("--seq-dir", default=None)
    ap.add_argument("--dataset", default="mulran")
    ap.add_argument("--extractor", default="cen2018")
    ap.add_argument("--loops", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    stream = importlib.import_module("navtech-radar-slam_b200.stream")
    if (a.datadir is None) == (a.seq_dir is None):
        ap.error("give either a frame directory or --seq-dir")
    r = stream.run_odometry(a.datadir, a.dataset, a.extractor, loop_detection=a.loops, out_dir=a.out, seq_dir=a.seq_dir)
    n = len(r["seconds"])
    print(f"{n} pairs, {n / max(r['seconds'].sum(), 1e-9):.1f} pairs/s incl. PNG decode; final pose x={r['trajectory'][-1][0, 3]:.3f} "
          f"y={r['trajectory'][-1][1, 3]:.3f}; {len(r['loops'])} loop closures")
