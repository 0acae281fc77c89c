"""CPU oracle — TEST INFRASTRUCTURE (see kernels_oracle.hpp).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package."""
