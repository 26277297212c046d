"""TEST INFRASTRUCTURE ONLY: CPU checkers for the CUDA RMSD path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (loos_b200/) must not.
"""
