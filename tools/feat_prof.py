import sys, ctypes as C
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from nav_gym_b200 import _lib
from nav_gym_b200.pedestrians import HumanPolicy
lib = _lib.load()
N = 40960
pol = HumanPolicy().cuda()
x = torch.rand(N, 512, device='cuda') * 6
feat = torch.empty(N, 4096, device='cuda')
w = (pol.act_fea_cv1.weight.sum(1).contiguous(), pol.act_fea_cv1.bias, pol.act_fea_cv2.weight.contiguous(), pol.act_fea_cv2.bias)
vp = lambda t: C.c_void_p(t.data_ptr())
def run(): lib.navgym_policy_features(vp(x), N, *[vp(t) for t in w], vp(feat), None)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print('features kernel %.3f ms  -> %.1f TFLOP/s fp32 (conv2 only), writes %.0f GB/s' % (ms, N * 393216 * 2 / ms / 1e9, N * 16384 / ms / 1e6))
