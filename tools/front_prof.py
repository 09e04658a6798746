"""Per-phase cycle breakdown of the policy front end (policy_features_umma2_kernel): builds a
profiling variant of the library (-DNAVGYM_PF_PROF: the first thread of every worker group sums
clock64() deltas per phase) and prints the mean cycles per pedestrian and phase."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, 'tools', '_prof', 'libnavgym_b200_pfprof.so')
os.makedirs(os.path.dirname(so), exist_ok=True)
os.environ['NAVGYM_LIB'] = so
from nav_gym_b200 import _lib  # noqa: E402
if not os.path.exists(so) or '--rebuild' in sys.argv:
    subprocess.check_call([_lib._nvcc()] + _lib.NVCC_FLAGS + ['-DNAVGYM_PF_PROF', '-o', so, _lib.SRC])
import torch  # noqa: E402
from nav_gym_b200.pedestrians import HumanPolicy, NativePolicy  # noqa: E402
n = 40960
torch.manual_seed(0)
pol = HumanPolicy().cuda().eval()
scan = (7.0 * torch.rand(n, 512, device='cuda')).contiguous()
goal, speed = torch.randn(n, 2, device='cuda'), torch.rand(n, 2, device='cuda')
nat = NativePolicy(pol, n, 'cuda:0')
out = torch.empty(n, 2, device='cuda')
lib = _lib.load()
buf = (C.c_ulonglong * 16)()
for _ in range(3):
    nat.mean(scan, goal, speed, out=out)
torch.cuda.synchronize()
lib.navgym_pf_prof_read(buf, 1)
iters = 5
for _ in range(iters):
    nat.mean(scan, goal, speed, out=out)
torch.cuda.synchronize()
lib.navgym_pf_prof_read(buf, 0)
names = ["wait conv1 MMAs", "P2", "-", "-", "P1a (next scan)", "wait conv2 MMAs", "P3a", "P3b + loop", "P1b"]
tot = 0
for i, nm in enumerate(names):
    c = buf[i] / (iters * n)
    tot += c
    print('%-24s %8.0f cycles per pedestrian' % (nm, c))
print('%-24s %8.0f  (x %d pedestrians per group and SM = %.0f us at 1.93 GHz)' % ('sum', tot, n // 148 // 4, tot * (n / 148 / 4) / 1.93e3))
