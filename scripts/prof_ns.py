"""One RK stage of the Navier-Stokes step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shenfun_b200.demos.navier_stokes import FusedNS
n = int(os.environ.get('NS_N', 512))
ns = FusedNS((n, n, n), nu=0.000625, dt=0.001, padding_factor=1.5).initialize()
ns.step()
torch.cuda.synchronize(); torch.cuda.profiler.start()
ns.Tp.backward_cross_hat(ns._uw_hat, ns.Up)
ns.TVp.forward(ns.Up, ns.dU)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
