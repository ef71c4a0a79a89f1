import time, torch
torch.cuda.init()
n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
dt = t(lambda: d_a.copy_(h_in, non_blocking=True)); print("H2D GB/s", n / dt / 1e9)
dt = t(lambda: h_out.copy_(d_b, non_blocking=True)); print("D2H GB/s", n / dt / 1e9)
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
dt = t(both); print("H2D+D2H concurrently: each GB/s", n / dt / 1e9)
def both_ratio():
    with torch.cuda.stream(s1): d_a[:n//2].copy_(h_in[:n//2], non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
dt = t(both_ratio); print("D2H full + H2D half concurrently: D2H GB/s", n / dt / 1e9)
import subprocess
print(subprocess.run(["nvidia-smi","--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max","--format=csv"],capture_output=True,text=True).stdout)
