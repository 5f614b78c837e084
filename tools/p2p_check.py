import time, torch
n = torch.cuda.device_count()
print("devices", n, "can access 0->1", torch.cuda.can_device_access_peer(0, 1) if n > 1 else None)
if n > 1:
    a = torch.empty(1 << 27, dtype=torch.int64, device="cuda:0")   # 1 GiB
    b = torch.empty(1 << 27, dtype=torch.int64, device="cuda:1")
    for it in range(3):
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter(); b.copy_(a); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        print("torch d2d 1 GiB: %.2f ms -> %.1f GB/s" % ((time.perf_counter() - t0) * 1e3, 1.074 / (time.perf_counter() - t0)))
