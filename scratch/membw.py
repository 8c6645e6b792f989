import torch
n = 1 << 28   # 1 GiB of float32
x = torch.empty(n, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.zero_()); print("memset  1 GiB: %.3f ms  %.2f TB/s (write only)" % (ms, 4 * n / ms / 1e9))
ms = t(lambda: y.copy_(x)); print("copy    1 GiB: %.3f ms  %.2f TB/s (read+write)" % (ms, 8 * n / ms / 1e9))
ms = t(lambda: x.sum()); print("sum     1 GiB: %.3f ms  %.2f TB/s (read only)" % (ms, 4 * n / ms / 1e9))
