"""Dev tool: backward-with-grad_init timing, both staged-halo sizes."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from tools.quick_bench import run
t = os.environ.get("TAG", "")
run(2048, 128, 128, torch.float32, gi=True, tag=t)
run(2048, 128, 128, torch.bfloat16, gi=True, tag=t)
run(1, 8192, 8192, torch.float32, gi=True, tag=t)
run(2048, 128, 128, torch.float32, gi=True, sigma=4.0, tag=t)
run(70, 128, 128, torch.float32, gi=True, tag=t)
''' % ROOT
for halo in ("narrow", "wide"):
    env = dict(os.environ, JSPSR_SPN_HALO=halo, TAG=halo)
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
