"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatements of the reference's hot-path algorithms (NumPy / torch-CPU),
the canonical synthetic environments, stand-ins for absent third-party
packages (oracle/shims), and the loader that imports the *unmodified*
reference from /root/reference when it is present (build container only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import anything from here; the product package
``torchrl_b200`` never does (tests/test_boundary.py enforces it).
"""
