import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shenfun_b200 as sf
from shenfun_b200 import ops
from shenfun_b200.demos.channelflow import MKM
comm = sf.init_from_env('nccl')
N = tuple(int(x) for x in sys.argv[1].split('x'))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
conv = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
c = MKM(N=N, Re=180., dt=0.0005, conv=conv, padding_factor=(1.5, 1.5, 1.5), timestepper='IMEXRK222', comm=comm)
t, ts = c.initialize()
c.assemble()
torch.cuda.synchronize()
t_setup = time.time()-t0
t, ts = c.solve(t=t, tstep=ts, end_time=c.dt)      # warm-up step
torch.cuda.synchronize(); comm.barrier()
l0 = ops.LAUNCHES
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
t, ts = c.solve(t=t, tstep=ts, end_time=t+steps*c.dt)
e1.record(); torch.cuda.synchronize()
ms = comm.allreduce(e0.elapsed_time(e1)/steps, op='max')
e, div = c.energy_and_divergence()
if comm.Get_rank() == 0:
    print(json.dumps({'workload': f'MKM channel flow {N}, conv={conv}, 3/2 padding, IMEXRK222', 'n_gpus': comm.Get_size(),
                      'ms_per_step': ms, 'steps_per_s': 1e3/ms, 'setup_s': t_setup, 'energy': e, 'div': div,
                      'launches_per_step': (ops.LAUNCHES-l0)//steps,
                      'mem_GB': torch.cuda.max_memory_allocated()/1e9}))
