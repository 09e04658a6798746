"""Time the native pedestrian policy (navgym_policy_mean: front end -> tcgen05 act_fc1 -> act_fc2 +
heads) at a crowd-sized batch, next to torch's float32 / TF32 forward of the same network.
  python tools/policy_prof.py [n] [iters]          CUDA-event times
  ncu --metrics gpu__time_duration.sum ... python tools/policy_prof.py 40960 3   per-kernel split
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nav_gym_b200.pedestrians import HumanPolicy, NativePolicy, preprocess_scan  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    torch.manual_seed(0)
    pol = HumanPolicy().cuda().eval()
    scan = (7.0 * torch.rand(n, 512, device='cuda')).contiguous()
    goal = torch.randn(n, 2, device='cuda')
    speed = torch.rand(n, 2, device='cuda')
    nat = NativePolicy(pol, n, 'cuda:0')
    out = torch.empty(n, 2, device='cuda')
    res = {'n': n, 'native_f16x3_ms': timed(lambda: nat.mean(scan, goal, speed, out=out), iters)}
    x3 = preprocess_scan(scan)[:, None, :].expand(-1, 3, -1).contiguous()
    with torch.no_grad():
        for name, flag in (('torch_fp32_ms', False), ('torch_tf32_ms', True)):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = flag
            res[name] = timed(lambda: pol.mean(x3, goal, speed), max(2, iters // 4))
            if not flag:
                want = pol.mean(x3, goal, speed)
    res['max_abs_err_vs_torch_fp32'] = float((out - want).abs().max())
    flop = 2.0 * n * (4096 * 256 * 3)
    res['fc1_f16_flop'] = flop
    print(json.dumps(res))


if __name__ == '__main__':
    main()
