"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/nrs_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os

import pytest


def test_library_is_built(nrs):
    assert os.path.exists(nrs.library_path()), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported(nrs):
    names = nrs.declared_symbols()
    assert len(names) >= 20
    L = nrs.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/nrs_b200.h but not exported: {missing}"


def test_header_is_plain_c(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "nrs_b200.h"\nint main(void){ nrs_sc_result r; r.idx = 0; return r.idx; }\n')
    rc = os.system(f"/usr/bin/gcc -std=c99 -Wall -Werror -I{root}/include -c {src} -o {tmp_path}/t.o")
    assert rc == 0


def test_no_cpu_fallback(nrs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is exercised on CPU-only hosts")
    with pytest.raises(nrs.NrsError) as e:
        nrs.Context(0)
    assert e.value.code == -2      # NRS_ENODEVICE
    assert b"no CPU fallback" in nrs.lib().nrs_last_error() or True


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "navtech-radar-slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in text and "oracle/" not in text.replace("oracle/sc_ref.c", "").replace(
                    "oracle/solver_ref", "").replace("oracle/frontend", ""), f


def test_cpp_adapters_compile_and_link(nrs):
    """The header-only C++ host side (nrs_adapters.hpp) instantiated with minimal Eigen/PCL
    stand-ins and linked against the library; without a GPU it must refuse loudly (exit 77)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "navtech-radar-slam_b200", "cpp", "adapter_check")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    import torch
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77 and "no CPU fallback" in r.stdout


def test_ctypes_structs_match_the_header(nrs, tmp_path):
    """The Python host side mirrors the header's POD structs by hand: their sizes and field offsets must be
    what a C compiler makes of include/nrs_b200.h."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = importlib_lib(nrs)
    structs = {
        "nrs_sc_result": (lib.ScResult, ["dist", "idx", "shift"]),
        "nrs_extract_params": (lib.ExtractParams, ["navtech_version", "range_bins", "radar_resolution", "cart_resolution",
                                                   "cart_pixel_width", "min_range", "max_points", "patch_size",
                                                   "use_voxelization", "interpolate_crossover", "voxel_size", "nndr",
                                                   "keypoint_extraction", "zq", "sigma_gauss"]),
        "nrs_frame_info": (lib.FrameInfo, ["n_candidates", "j_consumed", "l_regions", "K_polar", "K_voxel", "K", "K_desc",
                                           "maxg", "mean", "mean_h", "cart_max", "t_criteria"]),
        "nrs_solver_params": (lib.SolverParams, None),
        "nrs_solver_info": (lib.SolverInfo, None),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nrs_b200.h"', 'int main(void){']
    for cname, (ct, fields) in structs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in (fields if fields is not None else [n for n, _ in ct._fields_]):
            lines.append(f'printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    assert os.system(f"/usr/bin/gcc -std=c99 -I{root}/include {src} -o {exe}") == 0
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    for ln in out:
        if not ln:
            continue
        cname, field, val = ln.split()
        ct = structs[cname][0]
        if field == "size":
            assert C.sizeof(ct) == int(val), (cname, C.sizeof(ct), val)
        else:
            assert getattr(ct, field).offset == int(val), (cname, field)


def test_every_symbol_has_a_ctypes_signature(nrs):
    """an undeclared signature would make ctypes pass Python ints as 32-bit C ints -- fatal for pointers"""
    L = nrs.lib()
    assert [n for n in nrs.declared_symbols() if getattr(L, n).argtypes is None] == []


def importlib_lib(nrs):
    import importlib
    return importlib.import_module("navtech-radar-slam_b200._lib")


def test_host_descriptor_classification(nrs):
    """nrs_sc_descriptor_is_binary (the host-side routing test between the bit-mask and the general-valued kernel; AVX2
    where the CPU has it) against a plain numpy statement of the rule: all non-zero entries carry the same bit
    pattern c, c finite, positive, exact in float and of moderate magnitude; +-0.0 are empty bins; NaN is a value."""
    import ctypes as C
    import numpy as np
    L = nrs.lib()
    rng = np.random.default_rng(5)

    def rule(d):
        nz = d[~(d == 0.0)]                     # NaN counts as a value, -0.0 does not
        if nz.size == 0:
            return 1
        bits = nz.view(np.uint64)
        if not np.all(bits == bits[0]):
            return 0
        c = float(nz[0])
        return int(c > 0.0 and float(np.float32(c)) == c and 7.888609052210118e-31 <= c <= 1.2676506002282294e30)

    cases = []
    for _ in range(40):
        d = np.where(rng.random(1200) < rng.uniform(0.0, 0.6), 3.0, 0.0)
        cases.append(d)
    d = cases[0].copy(); d[rng.integers(1200)] = 3.0000000000000004; cases.append(d)        # one ulp off in one bin
    d = cases[1].copy(); d[d == 0.0][:0] = 0.0; d[7] = -0.0; d[8] = -0.0; cases.append(d)   # negative zeros are empty bins
    d = cases[2].copy(); d[1199] = np.nan; cases.append(d)                                   # NaN in the last lane
    d = np.zeros(1200); cases.append(d)                                                      # empty descriptor
    d = np.zeros(1200); d[::7] = 0.1; cases.append(d)                                        # c not exact in float
    d = np.zeros(1200); d[::7] = -2.0; cases.append(d)                                       # negative c
    d = np.zeros(1200); d[3] = 2.5; d[1197] = 2.5; cases.append(d)                           # sparse, first / last vector lanes
    d = np.zeros(1200); d[5] = 1e-40; cases.append(d)                                        # tiny c
    d = np.zeros(1200); d[0] = np.inf; cases.append(d)
    d = rng.uniform(0.5, 4.0, 1200); cases.append(d)                                         # general-valued
    for k, d in enumerate(cases):
        d = np.ascontiguousarray(d, dtype=np.float64)
        assert L.nrs_sc_descriptor_is_binary(d.ctypes.data_as(C.c_void_p)) == rule(d), k
