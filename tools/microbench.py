#!/usr/bin/env python
"""Prints SM cycles per warp instruction for the shared-memory operations the fused kernel uses."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quack_b200 import capi

print(capi.microbench())
