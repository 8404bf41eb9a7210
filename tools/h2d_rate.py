import torch, time
for mb, parts in ((85, 1), (85, 25), (78, 1), (23, 1)):
    h = torch.empty(mb * 1000 * 1000, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device='cuda')
    n = h.numel() // parts
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(parts): d[i*n:(i+1)*n].copy_(h[i*n:(i+1)*n], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('H2D %d MB in %d parts: %.2f ms  %.1f GB/s' % (mb, parts, dt*1e3, mb/1e3/dt))
    t0 = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('D2H %d MB: %.2f ms  %.1f GB/s' % (mb, dt*1e3, mb/1e3/dt))
