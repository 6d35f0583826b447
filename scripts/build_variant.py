"""Developer tool: build a variant of libci_b200.so with extra -D macros into scripts/_variants/<name>.so
(git-ignored; travels to the GPU box) for same-box A/B timing with scripts/ab_eval.py."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tfcausalimpact_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, 'scripts', '_variants')
os.makedirs(os.path.join(out, name), exist_ok=True)
objs, procs = [], []
for src in sorted(glob.glob(os.path.join(B.CSRC, '*.cu'))):
    obj = os.path.join(out, name, os.path.basename(src)[:-3] + '.o')
    procs.append(subprocess.Popen([B.NVCC] + B.FLAGS + defs + ['-c', src, '-o', obj]))
    objs.append(obj)
assert all(p.wait() == 0 for p in procs)
subprocess.check_call([B.NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o',
                       os.path.join(out, name + '.so')] + objs + ['-lcudart'])
for o in objs:
    os.remove(o)
print(os.path.join(out, name + '.so'))
