import torch, time
x = torch.empty(328*1024*1024//8, dtype=torch.float64, device="cuda")
h = torch.empty_like(x, device="cpu").pin_memory()
for n in (1,):
    torch.cuda.synchronize()
    for rep in range(3):
        t0=time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
        print("D2H pinned 328MB: %.2f ms  %.1f GB/s" % (dt*1e3, x.numel()*8/dt/1e9))
    for rep in range(2):
        t0=time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
        print("H2D pinned 328MB: %.2f ms  %.1f GB/s" % (dt*1e3, x.numel()*8/dt/1e9))
import numpy as np
dst = np.empty(x.numel()*4)   # fresh pageable
src = h.numpy()
t0=time.perf_counter(); dst[:x.numel()] = src; dt=time.perf_counter()-t0
print("host memcpy 1 thread into fresh pages: %.2f ms %.1f GB/s" % (dt*1e3, x.numel()*8/dt/1e9))
t0=time.perf_counter(); dst[:x.numel()] = src; dt=time.perf_counter()-t0
print("host memcpy 1 thread warm: %.2f ms %.1f GB/s" % (dt*1e3, x.numel()*8/dt/1e9))
