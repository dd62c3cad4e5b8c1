"""Checkpointing of device-resident spectral arrays.

Host-side mirror of `shenfun.io.Checkpoint` (shenfun/io/__init__.py:41-125): same constructor,
`update(t, tstep)`, `write(step, d)`, `read(u, name, step=0)`, `check_if_kill()` and the same
data dictionary `{'0': {'U': [U_hat]}, '1': {...}}`.  The reference stores one parallel HDF5 file
through h5py / MPI-IO, neither of which this image has; the container here is a directory
`<filename>.chk/` with one `.npy` file per dataset (`<name>/<step>.npy`, the GLOBAL array) and
`attrs.json` (`t`, `tstep`).  Every rank writes its `local_slice()` of the global array through a
memory map, so a checkpoint written by P ranks reads back on any other number of ranks of a
shared file system.  Arrays cross PCIe once per write (device -> pinned host -> file).
"""
import json
import os
import shutil
import sys

import numpy as np
import torch

from .comm import comm

__all__ = ('Checkpoint',)


def _host(u):
    t = u.v if hasattr(u, 'v') else u
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class Checkpoint:
    def __init__(self, filename, checkevery=10, data=None):
        self.filename = filename
        self.data = {} if data is None else data
        self.checkevery = checkevery
        self.attrs = None

    # -- container ---------------------------------------------------------------------------
    @property
    def path(self):
        return self.filename+'.chk'

    def _dataset(self, name, step):
        return os.path.join(self.path, name, f'{int(step)}.npy')

    def open(self, mode='r+'):
        if mode == 'w':
            if comm.Get_rank() == 0:
                shutil.rmtree(self.path, ignore_errors=True)
                os.makedirs(self.path)
            comm.barrier()
            self.attrs = {'tstep': 0, 't': 0.0}
        else:
            with open(os.path.join(self.path, 'attrs.json')) as f:
                self.attrs = json.load(f)

    def close(self):
        if self.attrs is not None and comm.Get_rank() == 0 and os.path.isdir(self.path):
            tmp = os.path.join(self.path, 'attrs.json.tmp')
            with open(tmp, 'w') as f:
                json.dump(self.attrs, f)
            os.replace(tmp, os.path.join(self.path, 'attrs.json'))
        comm.barrier()

    # -- reference API -----------------------------------------------------------------------
    def update(self, t, tstep):
        """io/__init__.py:76-95: every `checkevery` steps (or when a file `killshenfun` appears)
        keep the previous checkpoint as `<filename>.old.chk` and write the current data."""
        if self.attrs is None:
            self.open(mode='w')
            self.close()
        kill = self.check_if_kill()
        if tstep % self.checkevery == 0 or kill:
            if comm.Get_rank() == 0:
                old = self.filename+'.old.chk'
                shutil.rmtree(old, ignore_errors=True)
                shutil.copytree(self.path, old)
            comm.barrier()
            self.open()
            for key, val in self.data.items():
                self.write(int(key), val)
            self.attrs['tstep'] = int(tstep)
            self.attrs['t'] = float(t)
            self.close()
            if kill:
                sys.exit(1)

    def write(self, step, d):
        """io/__init__.py:97-103: `d = {name: [array, ...]}`; each rank stores its local block."""
        for name, val in d.items():
            if comm.Get_rank() == 0:
                os.makedirs(os.path.join(self.path, name), exist_ok=True)
            comm.barrier()
            for u in val:
                h = _host(u)
                gshape = tuple(u.global_shape) if hasattr(u, 'global_shape') else h.shape
                fn = self._dataset(name, step)
                if comm.Get_rank() == 0:
                    ok = False
                    if os.path.exists(fn):
                        m = np.load(fn, mmap_mode='r')
                        ok = m.shape == gshape and m.dtype == h.dtype
                        del m
                    if not ok:
                        m = np.lib.format.open_memmap(fn, mode='w+', dtype=h.dtype, shape=gshape)
                        del m
                comm.barrier()
                m = np.load(fn, mmap_mode='r+')
                s = tuple(u.local_slice()) if hasattr(u, 'local_slice') else Ellipsis
                m[s] = h
                m.flush()
                del m
        comm.barrier()

    def read(self, u, name, **kw):
        """io/__init__.py:105-111: fill `u` with its local block of dataset `name`/`step`."""
        step = kw.get('step', 0)
        self.open()
        m = np.load(self._dataset(name, step), mmap_mode='r')
        s = tuple(u.local_slice()) if hasattr(u, 'local_slice') else Ellipsis
        blk = np.array(m[s])
        del m
        if isinstance(u, torch.Tensor):
            t = u.v if hasattr(u, 'v') else u
            t.copy_(torch.from_numpy(blk).to(t.dtype))
        else:
            u[...] = blk
        return u

    @staticmethod
    def check_if_kill():
        """io/__init__.py:113-125: a file named `killshenfun` in the working directory asks every
        rank to checkpoint and stop."""
        found = 1 if 'killshenfun' in os.listdir(os.getcwd()) else 0
        if comm.allreduce(found) > 0:
            if comm.Get_rank() == 0:
                os.remove('killshenfun')
                print('killshenfun Found! Stopping simulations cleanly by checkpointing...')
            comm.barrier()
            return True
        return False
