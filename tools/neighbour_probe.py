"""The SURVEY 8f rank 3/4 kernels alone (loss, RMSE/MAE sums, crop, merge) at bench.py's sizes: `variants.neighbours`
without the rest of the bench (dev tool; DESIGN section 4d numbers)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    peak, _ = bench.peak_hbm()
    for _ in range(2):
        print(json.dumps(bench.neighbour_rows(torch, dev, timed, peak, 4096), indent=1))
