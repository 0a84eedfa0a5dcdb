"""ncu target: the single-launch fixed-affinity loop (spn_iterate_fused.cu), 512 tiles of 128x128, T = 6."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["JSPSR_SPN_ITER_FUSED"] = "1"
import torch
from jspsr_b200 import functional as F
import bench
init, weight, offset, gout, w, b = bench.make_inputs(torch, 512, torch.device("cuda", 0), torch.float32, 4322)
aff = weight * 0.1
for _ in range(3):
    F.spn_iterate(init, aff, offset, 6)
torch.cuda.synchronize(); print("done")
