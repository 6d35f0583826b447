"""Developer tool: evaluation timing at few lanes per series for every scripts/_variants/*.so."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'scripts'))
from tfcausalimpact_b200 import _native as nt
nt._LIB_PATH = sys.argv[1]
import dev_eval_bench as d
print(os.path.basename(sys.argv[1]))
for (N, G, k) in eval(os.environ.get('AB_CASES', '[(10000, 1, 50), (10000, 2, 50), (80000, 1, 50), (10000, 1, 10)]')):
    d.run(N, G, 365, 30, k, 7, 10)
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, 'scripts', '_variants', '*.so'))):
    subprocess.run([sys.executable, '-c', CHILD, lib], timeout=300)
