import os, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from nav_gym_b200.pedestrians import HumanPolicy
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = 40960
pol = HumanPolicy().cuda().eval()
x = torch.rand(N, 1, 512, device='cuda') - 0.5
g = torch.rand(N, 2, device='cuda'); s = torch.rand(N, 2, device='cuda')
x3 = x.expand(-1, 3, -1).contiguous()
with torch.no_grad():
    print('mean fp32 3ch  %.3f ms' % t(lambda: pol.mean(x3, g, s)))
    w = pol.act_fea_cv1.weight.sum(1, keepdim=True)
    def folded(dt=None):
        h = F.relu(F.conv1d(x, w, pol.act_fea_cv1.bias, stride=2, padding=1))
        h = F.relu(pol.act_fea_cv2(h)); h = F.relu(pol.act_fc1(h.reshape(N, -1)))
        h = F.relu(pol.act_fc2(torch.cat((h, g, s), -1)))
        return torch.cat((torch.sigmoid(pol.actor1(h)), torch.tanh(pol.actor2(h))), -1)
    print('mean fp32 folded %.3f ms' % t(folded))
    print('  conv1 %.3f' % t(lambda: F.conv1d(x, w, pol.act_fea_cv1.bias, stride=2, padding=1)))
    h1 = F.relu(F.conv1d(x, w, pol.act_fea_cv1.bias, stride=2, padding=1))
    print('  conv2 %.3f' % t(lambda: pol.act_fea_cv2(h1)))
    h2 = F.relu(pol.act_fea_cv2(h1)).reshape(N, -1)
    print('  fc1 %.3f' % t(lambda: pol.act_fc1(h2)))
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    print('mean tf32 folded %.3f ms' % t(folded))
    print('  fc1 tf32 %.3f' % t(lambda: pol.act_fc1(h2)))
    with torch.autocast('cuda', dtype=torch.bfloat16):
        print('mean bf16 autocast folded %.3f ms' % t(folded))
        print('  conv2 bf16 %.3f' % t(lambda: pol.act_fea_cv2(h1)))
    xc = x.to(memory_format=torch.channels_last) if False else x
