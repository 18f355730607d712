"""oracle -- TEST INFRASTRUCTURE ONLY (CPU checker for the x266 hot path).

Nothing in x266_b200/ imports this package.  Allowed users: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline leg and bench.py --impl reference.
"""
from .loader import Oracle, Ref, RefConv, build, have_ref, have_ref_conv  # noqa: F401
