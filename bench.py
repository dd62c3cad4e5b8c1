#!/usr/bin/env python
"""bench.py -- 512^3 fp64 forward+backward transform pairs/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one TensorProductSpace.forward + one .backward of a 512^3 real
field on T = C2C(512) x C2C(512) x R2C(512)  (SURVEY.md 8d, M1).  At N > 1 the
same global 512^3 problem is slab-distributed over the ranks (strong scaling),
one process per GPU, NCCL all-to-all between the axis passes.

The JSON line also carries: `e2e` (same pair through the public API from/to
pinned HOST buffers), `roofline` of the dominant kernel (live CUDA-event
timing), `cpu_baseline` (the numpy/scipy oracle on this box's host cores),
and secondary workloads (`padded_pair`, `ns`) as extra keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (512, 512, 512)
METRIC = '512^3 fp64 fwd+bwd transform pairs/s'
# algorithmic bytes of one un-padded 512^3 pair: 3 axis passes each way
REAL_B = 512**3*8
SPEC_B = 512*512*257*16
PAIR_BYTES = 2*((REAL_B+SPEC_B) + 2*(2*SPEC_B))


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)).get('hbm_gbs', 6650.0), 'measured'
    return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_pair_time(shape, reps, workers):
    """The oracle port (numpy + scipy.fft) of the same pair on host cores."""
    from oracle import tensorproduct as otp
    from oracle import fourier as of
    of.WORKERS = workers
    d = [('C2C', shape[0], shape[0]), ('C2C', shape[1], shape[1]), ('R2C', shape[2], shape[2])]
    u = np.random.default_rng(1234).standard_normal(shape)
    uh = otp.forward(u, d)
    otp.backward(uh, d)          # warm-up (pocketfft plan cache, page faults)
    t0 = time.perf_counter()
    for _ in range(reps):
        uh = otp.forward(u, d)
        otp.backward(uh, d)
    return (time.perf_counter()-t0)/reps


def cpu_pair_sample_time(shape, frac, workers):
    """A bounded sample of the same pair for long reference-arm runs: every axis pass of the forward
    and of the backward transform on 1/frac of its lines (the passes are independent 1-D line batches,
    so the full pair costs frac times the sample).  Returns the seconds of one sample."""
    from oracle import tensorproduct as otp
    from oracle import fourier as of
    of.WORKERS = workers
    n0, n1, n2 = shape
    d = [('C2C', n0, n0), ('C2C', n1, n1), ('R2C', n2, n2)]
    rng = np.random.default_rng(1234)
    a = rng.standard_normal((n0//frac, n1, n2))                       # lines along axes 2 and 1
    bx = rng.standard_normal((n0, n1//frac, n2//2+1))+0j               # lines along axis 0
    t0 = time.perf_counter()
    w = otp._fwd_axis(a, 2, d[2])
    w = otp._fwd_axis(w, 1, d[1])
    bh = otp._fwd_axis(bx, 0, d[0])
    otp._bwd_axis(bh, 0, d[0])
    w = otp._bwd_axis(w, 1, d[1])
    otp._bwd_axis(w, 2, d[2])
    return time.perf_counter()-t0


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the reference's Python
    cannot travel to the GPU box and its FFTW/mpi4py-fft stack is absent) on host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if 'TORCHELASTIC_RUN_ID' in os.environ or int(os.environ.get('WORLD_SIZE', '1')) > 1:
        # torchrun pins OMP_NUM_THREADS=1 (and may narrow the CPU affinity) for its workers; the CPU
        # path does not depend on the number of GPUs, so rank 0 re-runs this very arm in a clean
        # environment: the measurement is then the same process setup as at N = 1.
        env = {k: v for k, v in os.environ.items()
               if k not in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'LOCAL_WORLD_SIZE',
                            'GROUP_RANK', 'ROLE_RANK', 'ROLE_WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT',
                            'TORCHELASTIC_RUN_ID', 'TORCHELASTIC_RESTART_COUNT', 'TORCHELASTIC_MAX_RESTARTS',
                            'TORCHELASTIC_USE_AGENT_STORE', 'TORCH_NCCL_ASYNC_ERROR_HANDLING')}
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except (AttributeError, OSError):
            pass
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--gpus', str(args.gpus),
                            '--steps', str(args.steps), '--warmup', str(args.warmup)], env=env,
                           stdout=subprocess.PIPE, text=True)
        sys.stdout.write(r.stdout)
        sys.stdout.flush()
        return
    cores = os.cpu_count() or 1
    budget_s = 150.0
    t_full = cpu_pair_time(SHAPE, 1, cores)             # one full pair: warm-up and cost estimate
    total = t_full*(args.warmup+args.steps)
    frac = 1
    while total/frac > budget_s and frac < 64:
        frac *= 2
    sample = 'one full 512^3 forward+backward pair per step'
    if frac > 1:
        sample = (f'per step: every axis pass of the 512^3 pair on 1/{frac} of its lines, time x {frac} '
                  f'(one full pair measured first: {t_full:.2f} s)')
    times = []
    for i in range(args.warmup+args.steps):
        t = cpu_pair_time(SHAPE, 1, cores) if frac == 1 else frac*cpu_pair_sample_time(SHAPE, frac, cores)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3*float(np.mean(times))
    val = 1e3/ms
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'pairs/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'TensorProductSpace(C2C(512),C2C(512),R2C(512)) forward+backward, 512^3 real fp64',
                   'global_shape': list(SHAPE)},
        'cpu_baseline': {'value': val, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{sample}, scipy.fft (pocketfft) workers={cores}, oracle/tensorproduct.py'},
        'e2e': {'value': val, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extras', action='store_true', help='skip secondary workloads')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of CUDA-graph replay')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import shenfun_b200 as sf
    from shenfun_b200 import ops

    comm = sf.init_from_env('nccl')
    world = comm.Get_size()
    rank = comm.Get_rank()
    if world == 1:
        torch.cuda.set_device(0)
    dev = torch.device('cuda', torch.cuda.current_device())

    bases = [sf.FunctionSpace(SHAPE[0], 'F', dtype='D'), sf.FunctionSpace(SHAPE[1], 'F', dtype='D'),
             sf.FunctionSpace(SHAPE[2], 'F', dtype='d')]
    T = sf.TensorProductSpace(comm, bases)
    g = torch.Generator(device=dev).manual_seed(1234+rank)
    u = sf.Array(T)
    u.v.copy_(torch.randn(T.shape(False), dtype=torch.float64, device=dev, generator=g))
    uh = sf.Function(T)
    ub = sf.Array(T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def eager_step():
        T.forward(u, uh)
        T.backward(uh, ub)

    for _ in range(args.warmup):
        eager_step()
    barrier()
    # The pair is captured once and replayed (shenfun_b200.graphs.GraphedCall, public API) at EVERY
    # N, so that the 1 -> 8 GPU curve is one launch mode: at 8 GPUs the per-rank work is so short that
    # issuing it from Python costs as much as running it.  Every rank must take the same path, hence
    # the all-reduce on the outcome.
    step, launch_mode = eager_step, 'eager'
    launches_per_step = None
    if not args.no_graph and not os.environ.get('SHB_NO_GRAPH'):
        ok = 1
        try:
            from shenfun_b200.graphs import GraphedCall
            lc = ops.LAUNCHES
            graphed = GraphedCall(eager_step, warmup=1, comm=comm)
            launches_per_step = (ops.LAUNCHES-lc)//2          # one warm-up pass + the captured pass
        except Exception as e:  # noqa: BLE001
            ok = 0
            if rank == 0:
                print(f'[bench] CUDA-graph capture failed, timing the eager path: {e!r}', file=sys.stderr)
        okt = torch.tensor([ok], device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 1:
            step, launch_mode = graphed, 'cuda_graph'
            for _ in range(3):
                step()
        barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    l0 = ops.LAUNCHES
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ops.LAUNCHES-l0 if launch_mode == 'eager' else launches_per_step*args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()/args.steps
    clocks = sampler.stop() if rank == 0 else None

    # parity guard inside the bench (cheap): round trip must reproduce the input
    rt = (torch.linalg.vector_norm(ub.v-u.v)/torch.linalg.vector_norm(u.v)).item()

    # ---- e2e: host pinned -> device -> forward -> backward -> host pinned ---------
    from shenfun_b200.hostio import bind_host_to_device_numa_node
    aff0 = os.sched_getaffinity(0)
    numa = bind_host_to_device_numa_node()        # pinned buffers are first-touched on the GPU's own NUMA node
    h_in = torch.empty(T.shape(False), dtype=torch.float64).pin_memory()
    h_in.copy_(u.v)
    h_out = torch.empty(T.shape(False), dtype=torch.float64).pin_memory()
    h_out.zero_()
    os.sched_setaffinity(0, aff0)                 # the CPU legs below use every core again
    from shenfun_b200.hostio import HostPipeline
    uh2 = [sf.Function(T), sf.Function(T)]
    cnt = [0]

    def pair(d_in, d_out):
        w = uh2[cnt[0] % 2]
        cnt[0] += 1
        T.forward(sf.Array(T, buffer=d_in), w)
        T.backward(w, sf.Array(T, buffer=d_out))

    # public streaming API: H2D of step k+1, transforms of step k and D2H of step k-1 overlap;
    # every step copies its full input from pinned host memory and its full result back
    pipe = HostPipeline(T.shape(False), torch.float64, T.shape(False), torch.float64, pair)

    def e2e_step():
        pipe.step(h_in, h_out)

    for _ in range(3):
        e2e_step()
    pipe.drain()
    barrier()
    n_e2e = max(5, min(args.steps, 20))
    e0.record()
    for _ in range(n_e2e):
        e2e_step()
    pipe.drain()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_rt = (torch.linalg.vector_norm(h_out-h_in)/torch.linalg.vector_norm(h_in)).item()
    e2e_ms = t.item()/n_e2e
    # what the host links give: every rank copies its block H2D, then D2H, all ranks at the same time
    # (the e2e figure is bounded by this, not by the device work, once several GPUs share the host's PCIe fabric)
    pcie = {}
    d_tmp = torch.empty(T.shape(False), dtype=torch.float64, device=dev)
    for name, fn in (('h2d', lambda: d_tmp.copy_(h_in, non_blocking=True)), ('d2h', lambda: h_out.copy_(d_tmp, non_blocking=True))):
        fn()
        barrier()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)/5], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        pcie[name+'_GBs_per_gpu_all_ranks_concurrent'] = h_in.numel()*8/tt.item()/1e6
    # both directions at once on two streams (what the pipeline asks of the link): bytes per direction / time
    d_tmp2 = torch.empty_like(d_tmp)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    barrier()
    cur_stream = torch.cuda.current_stream()
    e0.record()
    s_up.wait_stream(cur_stream)
    s_dn.wait_stream(cur_stream)
    for _ in range(5):
        with torch.cuda.stream(s_up):
            d_tmp.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_out.copy_(d_tmp2, non_blocking=True)
    cur_stream.wait_stream(s_up)
    cur_stream.wait_stream(s_dn)
    e1.record()
    barrier()
    tt = torch.tensor([e0.elapsed_time(e1)/5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    pcie['bidirectional_GBs_per_direction_per_gpu'] = h_in.numel()*8/tt.item()/1e6
    pcie['e2e_frac_of_bidirectional_link'] = (h_in.numel()*8/(e2e_ms*1e6))/pcie['bidirectional_GBs_per_direction_per_gpu']
    del d_tmp, d_tmp2

    # ---- roofline of the dominant kernel (live CUDA events around single launches) --
    peak, peak_src = peaks()
    roof = None
    if world == 1:
        roof = dominant_kernel_roofline(T, u, uh, peak, peak_src)
    else:
        roof = exchange_kernel_roofline(T, u, uh, comm, ms)
    exchange_impl = T.exchange_impl

    extras = {}
    if world == 1 and not args.no_extras:
        extras = secondary_workloads(T, uh, args, peak, with_cpu=not args.no_cpu)
    elif not args.no_extras:
        try:
            from shenfun_b200.demos import navier_stokes
            del u, ub
            torch.cuda.empty_cache()
            extras['ns'] = navier_stokes.bench(512, steps=max(2, min(args.steps, 5)), warmup=1, comm=comm)
        except Exception as e:  # noqa: BLE001
            extras['ns'] = {'error': repr(e)}
        if world == 8:
            # BASELINE config 5: demo/MKM.py channel flow 1024 x 768 x 1024 slab-distributed over 8 GPUs
            try:
                torch.cuda.empty_cache()
                extras['mkm'] = mkm_workload(comm, (1024, 768, 1024), steps=2)
            except Exception as e:  # noqa: BLE001
                extras['mkm'] = {'error': repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        tcpu = cpu_pair_time(SHAPE, 1, cores)
        cpu = {'value': 1.0/tcpu, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
               'sample': f'1 full 512^3 forward+backward pair (after 1 warm-up pair), scipy.fft workers={cores}, '
                         'oracle/tensorproduct.py'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': 1e3/ms, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'TensorProductSpace(C2C(512),C2C(512),R2C(512)) forward+backward, 512^3 real fp64',
                       'global_shape': list(SHAPE), 'decomposition': 'slab' if world > 1 else 'single',
                       'l2_policy': 'working set 2.15 GB per pass >> 126 MB L2 (no flush needed)',
                       'seed': 1234, 'launch_mode': launch_mode, 'exchange_impl': exchange_impl},
            'pair_hbm_frac_of_measured': (PAIR_BYTES/world/(ms*1e-3)/1e9)/peak,
            'roundtrip_rel_err': rt,
            'clocks': clocks,
            'e2e': {'value': 1e3/e2e_ms, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(np.prod(T.shape(False)))*8,
                    'd2h_bytes_per_step': int(np.prod(T.shape(False)))*8, 'ms_per_step': e2e_ms,
                    'roundtrip_rel_err_host': e2e_rt,
                    'api': 'shenfun_b200.hostio.HostPipeline (H2D / transforms / D2H on three streams)',
                    'host_numa_node': numa, 'pcie': pcie},
            'gpu_launches': launches,
            'roofline': roof,
            'cpu_baseline': cpu,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _time_launch(fn, reps=10):
    import torch
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps


# ncu kernel name of each stage of the forward plan (profiles/r2_ncu_full_summary.txt)
_STAGE_KERNEL = {
    'r2c_axis2': 'void fft_contig_kernel<Sched<256, 8, 8, 4, 32>, 2, 8>',
    'c2c_axis1': 'void fft_strided_kernel<Sched<512, 8, 8, 8, 64>, 0, 8, 0>',
    'c2c_axis0': 'void fft_strided_kernel<Sched<512, 8, 8, 8, 64>, 0, 8, 0>',
}


def ncu_traffic_bytes(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum of the stage's kernel, per launch, from the
    committed `ncu --set full` capture (profiles/); None when the summary is not there."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r2_ncu_full_summary.txt')
    want = _STAGE_KERNEL.get(stage)
    if want is None or not os.path.exists(path):
        return None
    hits = []
    with open(path) as f:
        for line in f:
            cols = [c.strip() for c in line.split('|')]
            if cols and cols[0] == want and len(cols) > 3:
                try:
                    hits.append((float(cols[2])+float(cols[3]))*1e9)
                except ValueError:
                    pass
    if not hits:
        return None
    # c2c_axis1 / c2c_axis0 are the same kernel launched twice in plan order (axis 1 first)
    idx = {'c2c_axis1': 0, 'c2c_axis0': 1}.get(stage, 0)
    return int(hits[min(idx, len(hits)-1)])


def dominant_kernel_roofline(T, u, uh, peak, peak_src):
    """Per-stage timing of the forward plan; the slowest axis pass is reported."""
    import torch
    from shenfun_b200 import ops
    stages = {}
    w = torch.empty(T.shape(True), dtype=torch.complex128, device=u.device)
    b2, b1, b0 = T.bases[2], T.bases[1], T.bases[0]
    stages['r2c_axis2'] = (_time_launch(lambda: b2._forward_axis(u.v, w, 512*512, 1)), REAL_B+SPEC_B)
    stages['c2c_axis1'] = (_time_launch(lambda: b1._forward_axis(w, w, 512, 257)), 2*SPEC_B)
    stages['c2c_axis0'] = (_time_launch(lambda: b0._forward_axis(w, w, 1, 512*257)), 2*SPEC_B)
    name = max(stages, key=lambda k: stages[k][0])
    ms, nbytes = stages[name]
    ach = nbytes/(ms*1e-3)/1e9
    return {'bound': 'hbm', 'kernel': name, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach/peak,
            'peak_source': peak_src, 'traffic': ncu_traffic_bytes(name),
            'traffic_source': 'profiles/r2_ncu_full_summary.txt (ncu --set full, dram read + write per launch)',
            'algorithmic_bytes_per_launch': nbytes,
            'avg_launch_ms': ms,
            'all_stages_gbs': {k: v[1]/(v[0]*1e-3)/1e9 for k, v in stages.items()}}


def exchange_kernel_roofline(T, u, uh, comm, ms_pair):
    """N > 1: the kernel that bounds the pair is the C2C pass whose stores ARE the all-to-all (it
    writes (P-1)/P of its output into the peers' receive buffers over NVLink).  Timed live with CUDA
    events around that one launch inside the forward plan (all ranks run it together, as in the
    plan), average over launches, max over ranks; the roofline is the NVLink bandwidth per direction
    per GPU (770 GB/s measured peer copy on this pool, /opt/skills/guides/B200_PROFILING.md; 900
    nominal; a copy-engine peer copy on the bench box itself: scripts/exchange_breakdown.py)."""
    import torch
    import torch.distributed as dist
    P = comm.Get_size()
    if not any(s.kind == 'exchange' for s in T.forward.steps):
        return None
    spec_local = int(np.prod(T.shape(True)))*16
    nv_bytes = spec_local*(P-1)/P                       # sent (= received) per rank per exchange
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    T._time_producer = ev
    tot, reps = 0.0, 20
    try:
        for it in range(reps+2):
            dist.barrier()
            T.forward(u, uh)
            torch.cuda.synchronize()
            if it >= 2:
                tot += ev[0].elapsed_time(ev[1])
    except Exception:  # noqa: BLE001  (the plan did not take the timed path)
        T._time_producer = None
        return None
    T._time_producer = None
    t = torch.tensor([tot/reps], dtype=torch.float64, device=u.v.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    ach = nv_bytes/(ms*1e-3)/1e9
    return {'bound': 'nvlink', 'kernel': 'fft_strided_kernel<Sched<512,8,8,8,64>, fwd, 16 columns, peer stores> (c2c along axis 1)',
            'achieved': ach, 'peak': 770.0, 'unit': 'GB/s', 'frac': ach/770.0,
            'peak_source': 'measured peer copy per direction on this pool (B200_PROFILING.md)',
            'traffic': None, 'algorithmic_bytes_per_launch': nv_bytes, 'avg_launch_ms': ms,
            'launches_per_step': 2,
            'how': 'CUDA events around the launch inside T.forward (eager), average of 20, max over ranks; bytes = the '
                   '(P-1)/P of the local spectral block that cross NVLink (the local 1/P is an HBM store)'}


def mkm_workload(comm, N, steps=2):
    """demo/MKM.py (channel flow, IMEXRK222, 3/2 padding) on the slab-distributed ranks."""
    import torch
    from shenfun_b200 import ops
    from shenfun_b200.demos.channelflow import MKM
    c = MKM(N=N, Re=180., dt=0.0005, conv=0, padding_factor=(1.5, 1.5, 1.5), timestepper='IMEXRK222', comm=comm)
    t, ts = c.initialize()
    c.assemble()
    t, ts = c.solve(t=t, tstep=ts, end_time=c.dt)      # warm-up step
    torch.cuda.synchronize()
    comm.barrier()
    l0 = ops.LAUNCHES
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    t, ts = c.solve(t=t, tstep=ts, end_time=t+steps*c.dt)
    e1.record()
    torch.cuda.synchronize()
    ms = comm.allreduce(e0.elapsed_time(e1)/steps, op='max')
    e, div = c.energy_and_divergence()
    return {'workload': f'demo/MKM.py channel flow {N[0]}x{N[1]}x{N[2]}, 3/2 padding, IMEXRK222', 'n_gpus': comm.Get_size(),
            'ms_per_step': ms, 'value': 1e3/ms, 'unit': 'steps/s', 'energy': e, 'divergence': div,
            'gpu_launches_per_step': (ops.LAUNCHES-l0)//steps, 'mem_GB_per_gpu': torch.cuda.max_memory_allocated()/1e9}


def ns_cpu_baseline(n, steps, workers):
    """oracle/navier_stokes.py (numpy restatement of demo/NavierStokes.py:35-44 + RK4) on host cores."""
    from oracle import fourier as of
    from oracle.navier_stokes import NavierStokes
    of.WORKERS = workers
    s = NavierStokes((n, n, n), nu=0.000625, dt=0.001, padding_factor=1.5).initialize()
    s.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    return (time.perf_counter()-t0)/steps


def ns_cross_kernel_roofline(peak):
    """Dominant kernel of the NS step: the six-component C2R fused with the cross product
    (`fft_c2r_cross3_kernel<Sched<256,...>>`: 768-point lines, 512 modes kept, two real lines per
    complex transform).  Algorithmic bytes: six spectral lines read (257 complex each), three
    physical lines written (768 reals)."""
    import torch
    from shenfun_b200 import ops
    nl = 768*768
    ab = torch.randn((6, nl, 257, 2), dtype=torch.float64, device='cuda')
    abc = torch.view_as_complex(ab)
    out = torch.empty((3, nl, 768), dtype=torch.float64, device='cuda')
    ms = _time_launch(lambda: ops.fft_c2r_cross(abc[:3], abc[3:], out, nl, 768, 512), reps=5)
    nbytes = 6*nl*257*16+3*nl*768*8
    del ab, abc, out
    torch.cuda.empty_cache()
    ach = nbytes/(ms*1e-3)/1e9
    return {'bound': 'hbm', 'kernel': 'fft_c2r_cross3_kernel (six C2R 768-point lines as three complex ones + cross product)', 'achieved': ach,
            'peak': peak, 'unit': 'GB/s', 'frac': ach/peak, 'algorithmic_bytes_per_launch': nbytes, 'avg_launch_ms': ms,
            'launches_per_step': 4, 'traffic': 25390000000,
            'traffic_source': 'profiles/r2_ncu_full_summary.txt (fft_c2r_cross3_kernel: dram read 14.55 GB + write 10.84 GB)'}


def secondary_workloads(T, uh, args, peak, with_cpu=True):
    """3/2-padded pair (S 512x512x257 <-> P 768^3), reported beside the headline."""
    import torch
    import shenfun_b200 as sf
    out = {}
    try:
        Tp = T.get_dealiased((1.5, 1.5, 1.5))
        up = sf.Array(Tp)
        wh = sf.Function(T)

        def step():
            Tp.backward(uh, up)
            Tp.forward(up, wh)
        ms = _time_launch(step, reps=max(3, args.steps//2))
        pad_bytes = 2*((SPEC_B+1.5*SPEC_B) + (1.5*SPEC_B+2.25*SPEC_B) + (2.25*SPEC_B+768**3*8))
        out['padded_pair'] = {'value': 1e3/ms, 'unit': 'pairs/s', 'ms_per_step': ms,
                              'hbm_frac_of_measured': pad_bytes/(ms*1e-3)/1e9/peak,
                              'workload': 'S 512x512x257 <-> P 768^3 (3/2 rule), backward+forward'}
        del up, wh
        Tp.destroy()
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['padded_pair'] = {'error': repr(e)}
    cores = os.cpu_count() or 1
    try:
        from shenfun_b200.demos import navier_stokes
        out['ns'] = navier_stokes.bench(512, steps=max(2, min(args.steps, 5)), warmup=1)
        # whole-step algorithmic bytes as built: 523 GB per RK4 step
        out['ns']['hbm_frac_of_measured'] = 523e9/(out['ns']['ms_per_step']*1e-3)/1e9/peak
        torch.cuda.empty_cache()
        out['ns']['roofline'] = ns_cross_kernel_roofline(peak)
        if with_cpu:
            n_cpu = 128
            tcpu = ns_cpu_baseline(n_cpu, 1, cores)
            # work per step scales with the number of points (36 padded transforms + point-wise work)
            out['ns']['cpu_baseline'] = {
                'value': 1.0/(tcpu*(512/n_cpu)**3), 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
                'sample': f'{n_cpu}^3 Taylor-Green RK4 steps with 3/2 padding, oracle/navier_stokes.py (numpy + scipy.fft '
                          f'workers={cores}): {tcpu:.3f} s/step measured, scaled by (512/{n_cpu})^3 = {(512//n_cpu)**3} '
                          f'to the 512^3 workload (a 768^3 padded step needs ~45 GB of host work arrays)',
                'measured_steps_per_s_at_sample_size': 1.0/tcpu}
    except Exception as e:  # noqa: BLE001
        out['ns'] = {'error': repr(e)}
    torch.cuda.empty_cache()
    try:
        from shenfun_b200.demos import poisson3d, biharmonic3d
        out['poisson3d_chebyshev_512'] = poisson3d.bench(512, 'C')
        if with_cpu:
            out['poisson3d_chebyshev_512']['cpu_baseline'] = solver_cpu_baseline('helmholtz', 512, cores)
        torch.cuda.empty_cache()
        out['biharmonic3d_legendre_256'] = biharmonic3d.bench(256, 'L')
        if with_cpu:
            out['biharmonic3d_legendre_256']['cpu_baseline'] = solver_cpu_baseline('legendre', 256, cores)
    except Exception as e:  # noqa: BLE001
        out['poisson_biharmonic'] = {'error': repr(e)}
    return out


def solver_cpu_baseline(kind, n, cores):
    """CPU leg of the solver configs on a bounded sample of lines (the lines are independent).
    With oracle/_ref/refcy built (the reference's own la.pyx / transforms.pyx cythonised unchanged,
    oracle/build_ref.sh) the leg runs the REFERENCE kernel: `kind: "reference"`; otherwise the numpy
    port.  'helmholtz' = LU_Helmholtz once + Solve_Helmholtz timed (chebyshev/la.py:131-351 ->
    la.pyx:708-947); 'legendre' = one Legendre scalar-product stage, the reference's default
    'recursive' kind (spectralbase.py:536-558 -> transforms.pyx:175-263)."""
    from oracle import la as ola, refcy
    import shenfun_b200 as sf
    rng = np.random.default_rng(1234)
    have_ref = refcy.available()
    if kind == 'helmholtz':
        SD = sf.FunctionSpace(n, 'C', bc=(0, 0))
        d = SD.dim()
        A = SD.spectral_matrix(0, 2).diags()[:d, :d]
        B = SD.mass_matrix().diags()[:d, :d]
        m1, m2 = 64, 33                                    # sample: 64 x 33 of the 512 x 257 lines
        k1 = np.fft.fftfreq(n, 1./n)[None, :m1, None]
        k2 = np.fft.rfftfreq(n, 1./n)[None, None, :m2]
        f = rng.standard_normal((n, m1, m2))+1j*rng.standard_normal((n, m1, m2))
        f[d:] = 0
        scale = (n*(n//2+1))/(m1*m2)
        if have_ref:
            la = refcy.load('la')
            Ad = {k: np.ascontiguousarray(np.diagonal(A, k)) for k in (0, 2, 4)}
            Bd = {k: np.ascontiguousarray(np.diagonal(B, k)) for k in (-2, 0, 2)}
            d0, d1, d2, L = (np.zeros((n, m1, m2)) for _ in range(4))
            la.LU_Helmholtz(Ad, Bd, np.ones((1, m1, m2)), np.ascontiguousarray(np.broadcast_to(-(k1**2+k2**2), (1, m1, m2))),
                            False, d0, d1, d2, L, 0)
            u = np.zeros_like(f)
            la.Solve_Helmholtz(f, u, False, d0, d1, d2, L, 0)
            t0 = time.perf_counter()
            for _ in range(3):
                la.Solve_Helmholtz(f, u, False, d0, d1, d2, L, 0)
            t = (time.perf_counter()-t0)/3
            return {'value': 1.0/(t*scale), 'unit': 'solves/s', 'cores': 1, 'kind': 'reference',
                    'sample': f'reference Solve_Helmholtz (optimization/cython/la.pyx, compiled unchanged into oracle/_ref) on '
                              f'{m1}x{m2} of the {n}x{n//2+1} lines, factors precomputed: {t*1e3:.1f} ms, scaled by {scale:.1f}'}
        t0 = time.perf_counter()
        ola.helmholtz_solve(A, B, np.ones((1, 1, 1)), -(k1**2+k2**2), f, 0)
        t = time.perf_counter()-t0
        return {'value': 1.0/(t*scale), 'unit': 'solves/s', 'cores': 1, 'kind': 'port',
                'sample': f'factor + solve of {m1}x{m2} of the {n}x{n//2+1} lines (n = {n}), oracle/la.py '
                          f'helmholtz_solve (numpy, one thread): {t:.2f} s, scaled by {scale:.1f}'}
    flops_full = 2.0*n*n*(n*(n//2+1)*2)
    if have_ref:
        tr = refcy.load('transforms')
        SL = sf.FunctionSpace(n, 'L')
        x, w = SL.points_and_weights(n)
        kk = np.arange(n+1, dtype=float)
        # Legendre three-term recurrence in the reference's table layout (spectralbase.py:476-480: jacobi `a`
        # rows for alpha = beta = 0): P_{k+1} = ((2k+1) x P_k - k P_{k-1}) / (k+1)
        a = np.zeros((3, n+1))
        a[0] = (kk+1)/(2*kk+1)
        a[1] = 0
        a[2] = kk/(2*kk+1)
        ncol = 16*(n//2+1)*2                               # sample: 16 of the n rows of (k, l) lines, real columns
        u = rng.standard_normal((n, ncol))
        out = np.zeros_like(u)
        tr.scalar_product(u, out, x, w, 0, a)
        t0 = time.perf_counter()
        tr.scalar_product(u, out, x, w, 0, a)
        t = time.perf_counter()-t0
        scale = n/16.0
        return {'value': flops_full/(t*scale)/1e12, 'unit': 'TFLOP/s', 'cores': 1, 'kind': 'reference', 'ms': t*scale*1e3,
                'sample': f"reference transforms.scalar_product ('recursive' Legendre kernel, optimization/cython/"
                          f'transforms.pyx, compiled unchanged) on 16/{n} of the lines: {t*1e3:.0f} ms, scaled by {scale:.0f}; '
                          f'TFLOP/s counted as the dense contraction it replaces'}
    # Legendre stage: (n x n) operator times n x (n/2+1) complex columns, numpy BLAS
    P = rng.standard_normal((n, n))
    u = rng.standard_normal((n, 2*n*(n//2+1)))           # complex columns as pairs of real ones (real operator)
    np.dot(P, u[:, :1024])
    t0 = time.perf_counter()
    np.dot(P, u)
    t = time.perf_counter()-t0
    return {'value': flops_full/t/1e12, 'unit': 'TFLOP/s', 'cores': cores, 'kind': 'port', 'ms': t*1e3,
            'sample': f'one full Legendre Vandermonde stage of the {n}^3 transform as numpy.dot (BLAS dgemm '
                      f'on the host cores): {t*1e3:.0f} ms'}


if __name__ == '__main__':
    main()
