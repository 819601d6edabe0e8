#!/usr/bin/env python
"""Run a Python entry point of the repository (bench.py, __graft_entry__.smoke, an example) with
the EMULATED library in place of the GPU -- test infrastructure for the build container, which
has no GPU: the host layer's plumbing (option handling, buffer shapes, the JSON line of bench.py,
the phase marks ...) is exercised end to end against the same C ABI, compiled for the CPU
(tools/emu/README.md).  Nothing here is a product path and nothing under icqsol_b200/ knows
about it: the script patches, in this process only,
  * icqsol_b200._lib.load           -> the emulation build (device memory = host memory),
  * torch.cuda.*                    -> "one device is present" stand-ins,
  * tensor factories / Tensor.to    -> device 'cuda' means host memory.
usage: tools/emu/run_emulated.py <script.py | module:function> [args ...]"""
import contextlib
import ctypes
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def install():
    import torch
    import build as emu_build
    from icqsol_b200 import _lib

    os.environ.setdefault('ICQB_EMU_GUARD', '1')
    os.environ.setdefault('ICQB_EMU_SMS', '2')
    lib = ctypes.CDLL(emu_build.build())
    for name, restype, argtypes in _lib.SIGNATURES:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib.load = lambda: lib

    class FakeStream:
        cuda_stream = 0

        def synchronize(self):
            pass

    class FakeEvent:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t)

    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: 1
    torch.cuda.set_device = lambda d: None
    torch.cuda.current_device = lambda: 0
    torch.cuda.synchronize = lambda d=None: None
    torch.cuda.current_stream = lambda d=None: FakeStream()
    torch.cuda.Stream = lambda device=None: FakeStream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.Event = FakeEvent
    torch.cuda.mem_get_info = lambda d=None: (180 << 30, 180 << 30)
    torch.cuda.empty_cache = lambda: None

    def host(device):
        if device is None:
            return None
        return torch.device('cpu') if torch.device(device).type == 'cuda' else device

    def wrap(fn):
        def inner(*args, **kwargs):
            if 'device' in kwargs:
                kwargs['device'] = host(kwargs['device'])
            return fn(*args, **kwargs)
        return inner
    for name in ('empty', 'zeros', 'ones', 'full', 'tensor', 'empty_like', 'zeros_like'):
        setattr(torch, name, wrap(getattr(torch, name)))
    real_to = torch.Tensor.to

    def to(self, *args, **kwargs):
        args = tuple(host(a) if isinstance(a, (str, torch.device)) else a for a in args)
        if 'device' in kwargs:
            kwargs['device'] = host(kwargs['device'])
        out = real_to(self, *args, **kwargs)
        return out.clone() if out.data_ptr() == self.data_ptr() else out    # a copy, as H2D is
    torch.Tensor.to = to
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()    # H2D copy
    # a single-rank process group for scripts that ask for NCCL (tests/dist_worker.py)
    import torch.distributed as dist
    real_init = dist.init_process_group

    def init_process_group(backend=None, **kwargs):
        kwargs.pop('device_id', None)
        return real_init('gloo', **kwargs)
    dist.init_process_group = init_process_group
    real_cpu = torch.Tensor.cpu
    torch.Tensor.cpu = lambda self, *a, **k: real_cpu(self, *a, **k).clone()   # a copy, as D2H is
    return lib


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        return 2
    install()
    target = sys.argv[1]
    sys.argv = sys.argv[1:]
    if ':' in target and not target.endswith('.py'):
        modname, fn = target.split(':')
        mod = __import__(modname, fromlist=[fn])
        getattr(mod, fn)()
    else:
        runpy.run_path(target, run_name='__main__')
    return 0


if __name__ == '__main__':
    sys.exit(main())
