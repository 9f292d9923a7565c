import torch
x = torch.empty(2500 << 20, dtype=torch.uint8, device="cuda")
y = torch.empty(2500 << 20, dtype=torch.uint8, device="cuda")
for name, fn, nbytes in (("fill", lambda: x.zero_(), x.numel()), ("copy", lambda: y.copy_(x), 2 * x.numel()), ("read-sum", lambda: x.view(torch.int32).sum(), x.numel())):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, f"{ms:.3f} ms", f"{nbytes / ms / 1e6:.0f} GB/s")
