"""navtech-radar-slam_b200 -- B200-native (sm_100a) hot paths of gisbi-kim/navtech-radar-slam.

The product is the C-ABI shared library ``libnrs_b200.so`` (hand-written CUDA, declared in
``include/nrs_b200.h``).  This Python package is the thin host-side mirror of the reference's
operator API used by the tests and the benchmark:

* ``scancontext.SCManager``   <-> pgo/SC-A-LOAM/include/scancontext/Scancontext.h:57-122
* ``scancontext.ShardedSCDatabase`` -- the keyframe database sharded over ranks (torch.distributed)
* ``orora.ORORA`` / ``orora.RadarFeature`` / ``orora.setParams`` <-> ORORA/include/ORORA.hpp, radar_loader.hpp
* ``loopverify.doICPVirtualRelative`` <-> pgo/SC-A-LOAM/src/laserPosegraphOptimization.cpp:355-406 (ICP loop verification)
* ``posemanager.SE2PoseManager`` -- the node's evaluation text files (ORORA/include/SE2posemanager.hpp:151-245)

The directory name carries a hyphen (it is the reference's name); import it with
``importlib.import_module("navtech-radar-slam_b200")`` or through ``nrs_b200_import()`` in
``__graft_entry__``.

There is no CPU fallback: importing works anywhere, every compute call raises ``NrsError``
(NRS_ENODEVICE) without a B200.
"""
from ._lib import NrsError, Context, lib, library_path, declared_symbols  # noqa: F401
from . import scancontext  # noqa: F401
from . import orora  # noqa: F401
from . import loopverify  # noqa: F401
from . import posemanager  # noqa: F401
