#!/usr/bin/env python
"""Prints "SKIP CNT" for ncu so that the captured window is exactly the one timed step of
`bench.py --steps 1 --warmup 3 --no-graph` (launch count from a dry plan; B-independent)."""
import os
import sys
import warnings

warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psld_b200 import NCSNpp, cifar10_config  # noqa: E402
from psld_b200.program import build_plan  # noqa: E402

net = NCSNpp(cifar10_config()).eval()
net.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
L = build_plan(net, 2, 1, True, dry=True).launches
pre = 3                      # prior_kernel + dtype cast + copy into the network input
warm = 1 + 3 * (L + 1)       # first half-step + 3 warm-up steps (program + fused update)
print(pre + warm, L + 1)
