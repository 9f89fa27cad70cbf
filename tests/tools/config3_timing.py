"""BASELINE configs[3] round timing on one GPU at a given chain count (bench.config3_collapsed): python tests/tools/config3_timing.py [chains]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
import grample_b200 as gb
from grample_b200 import distributed as gbd

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
print(json.dumps(bench.config3_collapsed(gb, gbd, torch, None, 0, 0, 1, total_chains=chains), indent=1), file=sys.stderr)
