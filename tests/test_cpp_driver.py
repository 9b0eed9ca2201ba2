"""The C++ ROS-free node (navtech-radar-slam_b200/cpp/orora_odom.cpp): its host-side pieces on CPU -- the PNG decoder
against the real cv2.imread(IMREAD_GRAYSCALE) the reference's load_radar calls (radar_utils.cpp:107), the
SE2PoseManager writer against the Python twin byte for byte (SE2posemanager.hpp:151-259) -- and, on the GPU, the whole
driver against stream.run_odometry on the same directory of PNGs."""
import importlib
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "navtech-radar-slam_b200", "cpp", "orora_odom")


def need_exe():
    if not os.path.exists(EXE):
        pytest.skip("orora_odom not built (python __graft_entry__.py)")


def decode(path, tmp_path):
    out = str(tmp_path / "decoded.raw")
    r = subprocess.run([EXE, "--decode-only", str(path), out], capture_output=True, text=True)
    if r.returncode != 0:
        return None, r.stderr
    raw = open(out, "rb").read()
    rows, cols = struct.unpack("<ii", raw[:8])
    return np.frombuffer(raw, dtype=np.uint8, offset=8).reshape(rows, cols), ""


def png_bytes(img, depth=8, ctype=0, filters=None, interlace=0, idat_split=1 << 30):
    """A PNG written by hand: img is rows x cols (x channels) of samples; filters[y] picks the row filter."""
    h, w = img.shape[:2]

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xFFFFFFFF)

    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    if depth == 16:
        rows = [np.asarray(img[y], dtype=">u2").tobytes() for y in range(h)]
    elif depth == 8:
        rows = [np.asarray(img[y], dtype=np.uint8).tobytes() for y in range(h)]
    else:                                   # packed sub-byte grey samples, leftmost in the high bits
        rows = []
        for y in range(h):
            bits = np.zeros(((w * depth + 7) // 8) * 8, dtype=np.uint8)
            for b in range(depth):
                bits[b:w * depth:depth] = (np.asarray(img[y]) >> (depth - 1 - b)) & 1
            rows.append(np.packbits(bits).tobytes())
    bpp = max(1, depth * ch // 8)
    stream = bytearray()
    prev = bytes(len(rows[0]))
    for y, row in enumerate(rows):
        f = 0 if filters is None else filters[y % len(filters)]
        out = bytearray(len(row))
        for x in range(len(row)):
            a = row[x - bpp] if x >= bpp else 0
            b = prev[x]
            c = prev[x - bpp] if x >= bpp else 0
            if f == 0:
                pred = 0
            elif f == 1:
                pred = a
            elif f == 2:
                pred = b
            elif f == 3:
                pred = (a + b) >> 1
            else:
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            out[x] = (row[x] - pred) & 0xFF
        stream.append(f)
        stream += out
        prev = row
    z = zlib.compress(bytes(stream), 6)
    idat = b"".join(chunk(b"IDAT", z[i:i + idat_split]) for i in range(0, len(z), idat_split))
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
            + chunk(b"tEXt", b"Comment\x00hand-made") + idat + chunk(b"IEND", b""))


def test_png_decoder_matches_cv2_on_cv2_written_frames(tmp_path, synth):
    """A synthetic Navtech frame (400 x 3371, stamps + encoder + power bytes) written by cv2.imwrite at several
    compression levels / strategies (libpng picks different row filters for each)."""
    import cv2
    need_exe()
    world = synth.radar_world(seed=3)
    raw = synth.radar_frame(world, (1.0, 0.5, 0.02), seed=11, dataset="mulran")
    for k, flags in enumerate(([], [cv2.IMWRITE_PNG_COMPRESSION, 0], [cv2.IMWRITE_PNG_COMPRESSION, 9],
                               [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FILTERED],
                               [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_RLE])):
        p = tmp_path / f"{1560000000000000000 + k}.png"
        assert cv2.imwrite(str(p), raw, flags)
        got, err = decode(p, tmp_path)
        assert got is not None, err
        ref = cv2.imread(str(p), cv2.IMREAD_GRAYSCALE)
        assert got.shape == ref.shape == raw.shape and np.array_equal(got, ref) and np.array_equal(got, raw)


def test_png_decoder_row_filters_depths_and_refusals(tmp_path):
    """Hand-made PNGs: every row filter (incl. mixed per row, IDAT split into pieces, an ancillary chunk), 16-bit and
    sub-byte grey, grey+alpha -- all equal to cv2.imread(IMREAD_GRAYSCALE); colour, palette, interlaced, truncated and
    corrupted files are refused with an error (never decoded to something else)."""
    import cv2
    need_exe()
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    img[5] = 0
    img[:, 7] = 255
    cases = {f"f{f}": png_bytes(img, filters=[f]) for f in range(5)}
    cases["mixed"] = png_bytes(img, filters=[4, 3, 2, 1, 0, 3, 4], idat_split=97)
    img16 = rng.integers(0, 65536, (19, 31)).astype(np.uint16)
    cases["g16"] = png_bytes(img16, depth=16, filters=[4, 1, 3])
    ga = rng.integers(0, 256, (17, 23, 2), dtype=np.uint8)
    cases["ga8"] = png_bytes(ga, ctype=4, filters=[3, 4])
    for d in (1, 2, 4):
        cases[f"g{d}"] = png_bytes(rng.integers(0, 1 << d, (13, 29)).astype(np.uint8), depth=d, filters=[0, 2, 1])
    for name, data in cases.items():
        p = tmp_path / f"{name}.png"
        p.write_bytes(data)
        ref = cv2.imread(str(p), cv2.IMREAD_GRAYSCALE)
        assert ref is not None, name
        got, err = decode(p, tmp_path)
        assert got is not None, (name, err)
        assert got.shape == ref.shape and np.array_equal(got, ref), name
    rgb = rng.integers(0, 256, (9, 11, 3), dtype=np.uint8)
    bad = {"rgb": png_bytes(rgb, ctype=2), "interlaced": png_bytes(img, interlace=1),
           "truncated": cases["f1"][:len(cases["f1"]) // 2], "notpng": b"P5 2 2 255 abcd"}
    corrupt = bytearray(cases["f2"])
    corrupt[len(corrupt) // 2] ^= 0x40
    bad["corrupt"] = bytes(corrupt)
    for name, data in bad.items():
        p = tmp_path / f"bad_{name}.png"
        p.write_bytes(data)
        got, err = decode(p, tmp_path)
        assert got is None and "orora_odom:" in err, name


def se2(x, y, yaw):
    T = np.eye(4)
    c, s = np.cos(yaw), np.sin(yaw)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1], T[0, 3], T[1, 3] = c, -s, s, c, x, y
    return T


@pytest.mark.parametrize("dataset", ["mulran", "oxford"])
def test_pose_files_equal_python_twin(tmp_path, dataset):
    """The five evaluation files, C++ writer vs posemanager.py on the same 300 (ground truth, estimate) pairs, including
    zero motions, exact identities, half turns, values that print in exponent form and negative zeros: byte for byte."""
    need_exe()
    pmod = importlib.import_module("navtech-radar-slam_b200.posemanager")
    rng = np.random.default_rng(9)
    den = 1e9 if dataset == "mulran" else 1e6
    pairs = []
    for k in range(300):
        g = (float(rng.normal(2.5, 1.0)), float(rng.normal(0, 0.2)), float(rng.normal(0, 0.03)))
        T = se2(g[0] + rng.normal(0, 0.05), g[1] + rng.normal(0, 0.05), (-g[2] if dataset == "mulran" else g[2]) + rng.normal(0, 0.004))
        if k % 17 == 0:
            g, T = (0.0, 0.0, 0.0), np.eye(4)
        if k % 29 == 1:
            g = (1e-5, -2e-7, 3e-6)
        if k % 41 == 2:
            T = se2(-3.0, 1e-9, np.pi - 1e-3)
        if k % 53 == 3:
            g = (0.002, 0.0, -0.5)                         # "stopped" by translation only if the rotation is small or negative
        pairs.append((g, T, 1560000000000000000 + 250000000 * k, 1560000000250000000 + 250000000 * k, 700 + 13 * k))
    blob = struct.pack("<i", len(pairs))
    for g, T, t1, t2, us in pairs:
        blob += struct.pack("<3f", *g) + np.asfortranarray(T).tobytes(order="F") + struct.pack("<qqdi", t1, t2, den, us)
    (tmp_path / "in.bin").write_bytes(blob)
    cdir, pdir = tmp_path / "c", tmp_path / "p"
    cdir.mkdir()
    r = subprocess.run([EXE, "--posefile-selftest", str(tmp_path / "in.bin"), str(cdir), dataset, "cen2018"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pm = pmod.SE2PoseManager(str(pdir), dataset, "ORORA", "cen2018")
    curr_gt, curr = np.eye(4), np.eye(4)
    for g, T, t1, t2, us in pairs:
        Tgt = pmod.xyyaw2se2(g, dataset)
        curr_gt, curr = pmod.mul4(curr_gt, Tgt), pmod.mul4(curr, T)
        pm.writePoseError(Tgt, T)
        pm.writePose(Tgt, T, t1, t2, den)
        pm.writeTraj(curr, pm.est_traj_path_)
        pm.writeTraj(curr_gt, pm.gt_traj_path_)
        pm.writeTimeTaken(us, Tgt)
    for name in ("rel_error_odom", "eval_odom", "time_taken", "est_traj", "gt_traj"):
        a = (cdir / f"{dataset}_ORORA_cen2018_{name}.txt").read_text().splitlines()
        b = (pdir / f"{dataset}_ORORA_cen2018_{name}.txt").read_text().splitlines()
        assert len(a) == len(b) == len(pairs), name
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y]
        assert not diff, (name, diff[:3])


def test_ground_truth_reader(tmp_path):
    """loadGTPoses / get_groundtruth_odometry behind getXYYawPose (SE2posemanager.hpp:17-36,137-166,
    radar_utils.cpp:483-501): C++ and Python agree, MulRan's y column is the yaw column (the reference's own indexing),
    Oxford rows are found by their two stamps, a missing stamp is an error in both."""
    need_exe()
    pmod = importlib.import_module("navtech-radar-slam_b200.posemanager")
    rng = np.random.default_rng(21)
    for dataset in ("mulran", "oxford"):
        seq = tmp_path / dataset
        (seq / "gt").mkdir(parents=True)
        stamps = [1560000000000000000 + 250000000 * k for k in range(40)]
        vals = rng.normal(0, 1, (40, 6))
        with open(seq / "gt" / "radar_odometry.csv", "w") as f:
            if dataset == "mulran":
                for t, v in zip(stamps, vals):
                    f.write(f"{t},{v[0]:.9g},{v[1]:.9g},{v[5]:.9g}\n")
            else:
                f.write("source_timestamp,destination_timestamp,x,y,z,roll,pitch,yaw,source_radar_timestamp,destination_radar_timestamp\n")
                for k, v in enumerate(vals[:-1]):
                    f.write(f"{k},{k + 1}," + ",".join(f"{x:.9g}" for x in v) + f",{stamps[k + 1]},{stamps[k]}\n")
        pm = pmod.SE2PoseManager.from_seq_dir(str(seq), dataset, "ORORA", "cen2018")
        assert os.path.isdir(seq / "outputs") and os.path.exists(pm.output_pose_path_)
        for k in (0, 7, 38):
            t1, t2 = stamps[k], stamps[k + 1]
            want = pm.getXYYawPose(t1, t2) if dataset == "oxford" else pm.getXYYawPose(t1)
            f32 = lambda x: float(np.float32(float(f"{x:.9g}")))
            if dataset == "mulran":
                assert want == [f32(vals[k][0]), f32(vals[k][5]), f32(vals[k][5])]          # y := column 3 (:31)
            else:
                assert want == [f32(vals[k][0]), f32(vals[k][1]), f32(vals[k][5])]
            r = subprocess.run([EXE, "--gt-lookup", str(seq), dataset, str(t1), str(t2)], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            assert [float(x) for x in r.stdout.split()] == [float(f"{v:.9g}") for v in want]
        r = subprocess.run([EXE, "--gt-lookup", str(seq), dataset, "12345", "67890"], capture_output=True, text=True)
        assert r.returncode == 1
        with pytest.raises(ValueError):
            pm.getXYYawPose(12345, 67890) if dataset == "oxford" else pm.getXYYawPose(12345)


def test_adapters_with_cv_and_eigen_types(tmp_path):
    """RadarFeatureT<Eigen::MatrixXd, cv::Mat, cv::KeyPoint> -- the reference's exact RadarFeature members -- RadarLoaderT,
    ORORAT and SCManagerT compiled and linked against OpenCV- / Eigen- / PCL-shaped types (tests/cpp/cv_types_check.cpp).
    Neither library exists in this image: the interface stand-ins of oracle/refshim supply the shapes (cv::Mat::create /
    data / rows / cols, cv::KeyPoint(x, y, size), MatrixXd::resize / data / operator(), PointCloud::points).  Without a GPU
    the program must stop at nrs_ctx_create (77); with one it extracts, matches and solves through these types (0;
    profiles/tests/r02_cpp_host_check.log holds a run on a B200)."""
    shim = os.path.join(ROOT, "oracle", "refshim")
    lib = os.path.join(ROOT, "navtech-radar-slam_b200")
    if not os.path.exists(os.path.join(lib, "libnrs_b200.so")):
        pytest.skip("libnrs_b200.so not built")
    exe = str(tmp_path / "cv_types_check")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + shim, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(lib, "cpp"),
                        os.path.join(ROOT, "tests", "cpp", "cv_types_check.cpp"), "-L" + lib, "-lnrs_b200", "-Wl,-rpath," + lib, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode in (0, 77), (r.returncode, r.stdout, r.stderr)
    assert "adapters" in r.stdout


def test_driver_without_gpu_or_input(tmp_path):
    """No CPU fallback in the C++ host either: without a device the driver says so and exits 77; a bad directory or
    flag is an error (1), not a crash."""
    need_exe()
    import torch
    r = subprocess.run([EXE, str(tmp_path / "missing"), "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown argument" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([EXE, str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 77 and "no GPU" in r.stdout


@pytest.mark.gpu
def test_cpp_node_equals_python_driver(nrs, ctx, synth, tmp_path):
    """orora_odom (C++: png_gray -> RadarLoaderT -> ORORAT -> SE2PoseManager, SCManagerT with --loops) and
    stream.run_odometry (Python mirrors over the same C ABI) on one directory of PNGs: every evaluation file identical
    (time_taken's wall-clock column aside), the loop list identical."""
    import cv2
    need_exe()
    stream = importlib.import_module("navtech-radar-slam_b200.stream")
    world = synth.radar_world(seed=4321)
    d = tmp_path / "polar_oxford_form"
    d.mkdir()
    for k in range(8):
        raw = synth.radar_frame(world, (2.0 * k, 0.05 * k, np.deg2rad(0.8) * k), seed=300 + k, dataset="mulran")
        assert cv2.imwrite(str(d / f"{1560000000000000000 + k * 250000000}.png"), raw)
    gt = [(2.0, 0.05, -np.deg2rad(0.8))] * 6
    (tmp_path / "gt.txt").write_text("".join(f"{x:.9g},{y:.9g},{w:.9g}\n" for x, y, w in gt))
    for extractor in ("cen2018", "cen2019"):
        cdir, pdir = tmp_path / f"c_{extractor}", tmp_path / f"p_{extractor}"
        r = subprocess.run([EXE, str(d), "--dataset", "mulran", "--extractor", extractor, "--out", str(cdir), "--gt",
                            str(tmp_path / "gt.txt"), "--loops", "--sc-dist-thres", "0.6"], capture_output=True, text=True)
        assert r.returncode == 0, (r.stdout, r.stderr)
        assert "6 pairs (0 skipped)" in r.stdout, r.stdout
        gt32 = [tuple(float(np.float32(float(f"{v:.9g}"))) for v in g) for g in gt]
        p = stream.run_odometry(str(d), "mulran", extractor, loop_detection=True, sc_dist_thres=0.6, ctx=ctx, out_dir=str(pdir), gt_rel=gt32)
        assert p["relative"].shape == (6, 4, 4)
        for name in ("rel_error_odom", "eval_odom", "est_traj", "gt_traj"):
            a = (cdir / f"mulran_ORORA_{extractor}_{name}.txt").read_text()
            b = (pdir / f"mulran_ORORA_{extractor}_{name}.txt").read_text()
            assert a == b and a.count("\n") == 6, (extractor, name)
        ta = (cdir / f"mulran_ORORA_{extractor}_time_taken.txt").read_text().splitlines()
        tb = (pdir / f"mulran_ORORA_{extractor}_time_taken.txt").read_text().splitlines()
        assert [x.split(",")[1] for x in ta] == [x.split(",")[1] for x in tb] and len(ta) == 6
        assert (cdir / "loops.txt").read_text() == (pdir / "loops.txt").read_text()
    # the node's own layout (seq_dir with polar_oxford_form/, gt/radar_odometry.csv, outputs/): two copies of the
    # sequence directory sharing the frames, one per driver
    seqs = []
    for who in ("c", "p"):
        seq = tmp_path / f"seq_{who}"
        (seq / "gt").mkdir(parents=True)
        os.symlink(d, seq / "polar_oxford_form")
        with open(seq / "gt" / "radar_odometry.csv", "w") as f:
            for k in range(8):
                f.write(f"{1560000000000000000 + k * 250000000},{2.0 + 0.01 * k:.9g},0.05,{-0.0139 - 0.001 * k:.9g}\n")
        seqs.append(seq)
    r = subprocess.run([EXE, "--seq-dir", str(seqs[0]), "--dataset", "mulran", "--extractor", "cen2018"], capture_output=True, text=True)
    assert r.returncode == 0 and "6 pairs (0 skipped)" in r.stdout, (r.stdout, r.stderr)
    stream.run_odometry(None, "mulran", "cen2018", ctx=ctx, seq_dir=str(seqs[1]))
    for name in ("rel_error_odom", "eval_odom", "est_traj", "gt_traj"):
        a = (seqs[0] / "outputs" / f"mulran_ORORA_cen2018_{name}.txt").read_text()
        b = (seqs[1] / "outputs" / f"mulran_ORORA_cen2018_{name}.txt").read_text()
        assert a == b and a.count("\n") == 6, name
    ev = np.loadtxt(seqs[0] / "outputs" / "mulran_ORORA_cen2018_eval_odom.txt", delimiter=",")
    # ground-truth columns: x of pair i = the row of file i's stamp; y = yaw = -(column 3) through XYYaw2SE2's -yaw
    assert np.allclose(ev[:, 0], [2.0 + 0.01 * k for k in range(1, 7)], atol=1e-5)
    assert np.allclose(ev[:, 1], [-0.0139 - 0.001 * k for k in range(1, 7)], atol=1e-6)
    assert np.allclose(ev[:, 2], [0.0139 + 0.001 * k for k in range(1, 7)], atol=1e-6)
