import torch, time
for mb in (8.5, 64, 256):
    n = int(mb * 1e6 / 4)
    d = torch.empty(n, dtype=torch.float32, device='cuda'); h = torch.empty(n, dtype=torch.float32).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('D2H %.1f MB: %.3f ms  %.1f GB/s' % (mb, ms, mb / ms))
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('H2D %.1f MB: %.3f ms  %.1f GB/s' % (mb, ms, mb / ms))
import subprocess
print(subprocess.run(['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max', '--format=csv'], capture_output=True, text=True).stdout)
