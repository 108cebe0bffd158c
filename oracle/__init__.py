"""CPU oracle for the LEOD hot path.  TEST INFRASTRUCTURE ONLY.

A plain, un-optimised restatement (torch fp32 on the CPU for the floating-point network, numpy for
the integer / index work) of the reference algorithms on the hot path named in BASELINE.json.  It
exists so that the CUDA path in `leod_b200/` can be checked on machines where `/root/reference` is
absent (the GPU box).  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs of
`bench.py` may import it; nothing under `leod_b200/` does, and the product path never falls back
to it.

Parity status: PINNED.  Every function here is checked in `tests/test_oracle_golden.py` against
fixtures under `tests/golden/` that were produced by importing and running the reference itself
(`tests/golden/make_golden.py`, run in the build container where `/root/reference` is mounted).
The reference ships no tests / golden vectors of its own, so generated fixtures
are the only possible anchor.  NMS lives in torchvision (pinned 0.15.2 by the reference's
environment.yml, not vendored); its fixtures come from the installed torchvision op.

Each function cites the reference file:line it restates (paths relative to /root/reference).
"""
from .config import ModelCfg  # noqa: F401
