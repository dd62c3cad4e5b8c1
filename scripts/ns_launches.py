"""One NS step (512^3, 3/2 padding, RK4) between profiler start/stop -- used with
ncu --profile-from-start off --metrics gpu__time_duration.sum to list the step's kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shenfun_b200.demos.navier_stokes import FusedNS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ns = FusedNS((n, n, n), nu=0.000625, dt=0.001, padding_factor=1.5).initialize()
ns.step()
torch.cuda.synchronize(); torch.cuda.profiler.start()
ns.step()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
