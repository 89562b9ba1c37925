"""C4 expression flow on the resident matrix, alone (development aid: run under the ncu launch list
to see which kernels the flow's wall time is made of)."""
import json, sys
sys.path.insert(0, ".")
import torch
import bench
r = bench.run_resident_flow(10000, 20000)
r.pop("regions", None)
print(json.dumps(r))
r = bench.run_resident_flow(10000, 20000)
r.pop("regions", None)
print(json.dumps(r))
