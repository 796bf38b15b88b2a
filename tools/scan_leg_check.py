import json, sys, torch
sys.path.insert(0, "/root/repo")
import bench
from bayesian_dlms_b200 import Engine
eng = Engine(0)
print(json.dumps(bench.scan_leg(eng, torch.device("cuda", 0), with_cpu=False)))
