"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs) may
import this package.  Nothing under navtech-radar-slam_b200/ imports it; the product path has
no CPU fallback.

Layout: sc_ref.c / solver_ref.c / frontend_ref.{c,py} / pyoracle.py = the CPU restatement (the checker and the timed
CPU baseline); stdsort.cpp = this toolchain's std::sort for cen2019's tie order; refshim/ + ref_*_wrap.cpp + ref*.py =
the REFERENCE'S OWN SOURCES compiled in place against stand-in headers into oracle/_ref/ (Makefile target `ref`), which
the restatement is pinned against in tests/test_ref_*.py.
"""
