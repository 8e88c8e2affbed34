import torch
n = 4 * 1024**3
x = torch.empty(n, dtype=torch.uint8, device="cuda")
y = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.fill_(1)); print("fill (write only) GB/s", n / ms / 1e6)
ms = t(lambda: y.copy_(x)); print("copy (read+write) GB/s", 2 * n / ms / 1e6)
ms = t(lambda: x.sum()); print("sum over uint8 (read only; may be compute bound)")
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print("sum fp32 (read only) GB/s", n / ms / 1e6)
