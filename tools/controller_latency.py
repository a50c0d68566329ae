"""Wall time of the test-time controller (GCBF.apply) per call: library path (gcbf_apply) vs Python sequencing.  `python tools/controller_latency.py`"""
import os
import json
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import bench
dev=torch.device("cuda",0)
for c in ("C1","C2","C3"):
    print(json.dumps(bench.controller_leg(c, dev)))
from gcbf_b200 import ops
ops.NATIVE=False
for c in ("C1","C3"):
    r=bench.controller_leg(c, dev); r["path"]="python sequencing"; print(json.dumps(r))
