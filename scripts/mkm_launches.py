"""One channel-flow (MKM) step between profiler start/stop, for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shenfun_b200 as sf
from shenfun_b200.demos.channelflow import MKM
N = tuple(int(x) for x in sys.argv[1].split('x')) if len(sys.argv) > 1 else (512, 256, 512)
c = MKM(N=N, Re=180., dt=0.0005, conv=0, padding_factor=(1.5, 1.5, 1.5), timestepper='IMEXRK222', comm=sf.comm)
t, ts = c.initialize()
c.assemble()
t, ts = c.solve(t=t, tstep=ts, end_time=c.dt)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
t, ts = c.solve(t=t, tstep=ts, end_time=t+2*c.dt)
e1.record(); torch.cuda.synchronize()
print('ms/step (eager, 2 steps):', e0.elapsed_time(e1)/2, 'mem GB', torch.cuda.max_memory_allocated()/1e9)
torch.cuda.profiler.start()
t, ts = c.solve(t=t, tstep=ts, end_time=t+c.dt)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
