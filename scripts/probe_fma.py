import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tfcausalimpact_b200 import _native as nt
lib = nt.lib()
out = torch.zeros(1, device='cuda')
for name in ('ci_bench_fma', 'ci_bench_fma2'):
    f = getattr(lib, name)
    for blocks in (148 * 8, 148 * 2):
        fl = ctypes.c_double(0)
        for _ in range(2):
            nt.check(f(8192, blocks, nt.ptr(out), nt.stream_ptr(), ctypes.byref(fl)))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nt.check(f(8192, blocks, nt.ptr(out), nt.stream_ptr(), ctypes.byref(fl)))
        e1.record(); torch.cuda.synchronize()
        print(name, 'blocks', blocks, f'{fl.value / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s')
