"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs) may
import this package.  Nothing under navtech-radar-slam_b200/ imports it; the product path has
no CPU fallback.
"""
