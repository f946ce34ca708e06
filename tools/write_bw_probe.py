import torch, time
x = torch.empty(1 << 30, dtype=torch.float64, device="cuda")   # 8 GiB
y = torch.empty(1 << 29, dtype=torch.float32, device="cuda")
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(n):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_()); print("memset 8GiB: %.3f ms  %.0f GB/s" % (ms, x.numel()*8/ms/1e6))
ms = t(lambda: x.fill_(1.5)); print("fill 8GiB:   %.3f ms  %.0f GB/s" % (ms, x.numel()*8/ms/1e6))
x2 = torch.empty_like(x[: 1 << 29]); 
ms = t(lambda: x2.copy_(x[: 1 << 29])); print("copy 4GiB->4GiB: %.3f ms  %.0f GB/s (r+w)" % (ms, 2*x2.numel()*8/ms/1e6))
ms = t(lambda: x.sum()); print("read 8GiB: %.3f ms  %.0f GB/s" % (ms, x.numel()*8/ms/1e6))
