"""D2H bandwidth of this box for the host pipeline's copy pattern: chunks of an [4096, 519] float32
observation block copied to pinned memory on several streams, with and without host syncs."""
import time, torch
B, W = 4096, 519
src = torch.rand(B, W, device='cuda')
dst = torch.empty(B, W).pin_memory()
def run(chunks, iters=300, sync_each=False):
    streams = [torch.cuda.Stream() for _ in range(chunks)]
    bounds = [(B * c // chunks, B * (c + 1) // chunks) for c in range(chunks)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(iters):
        for s, (a, b) in zip(streams, bounds):
            if sync_each: s.synchronize()
            with torch.cuda.stream(s):
                dst[a:b].copy_(src[a:b], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return B * W * 4 * iters / dt / 1e9, dt / iters * 1e6
for chunks in (1, 2, 4, 8):
    for sync_each in (False, True):
        gbs, us = run(chunks, sync_each=sync_each)
        print('chunks %d  sync-before-reuse %-5s  %.1f GB/s  %.0f us per 8.5 MB' % (chunks, sync_each, gbs, us))
