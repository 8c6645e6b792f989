import torch, time
n = 239357952
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
for chunks in (1, 2, 8):
    with torch.cuda.stream(s):
        for _ in range(2): d.copy_(h, non_blocking=True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10):
            c = n // chunks
            for i in range(chunks): d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("pinned H2D %d chunk(s): %.2f ms  %.1f GB/s" % (chunks, ms, n / ms / 1e6))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
