import torch, time
dev = torch.device("cuda:0")
n = 1 << 28   # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device=dev); d_out = torch.ones(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down):
    torch.cuda.synchronize(); t = time.perf_counter()
    if up:
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    if down:
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for _ in range(2): run(True, True)
for name, a, b in (("H2D", 1, 0), ("D2H", 0, 1), ("both", 1, 1)):
    t = min(run(a, b) for _ in range(5)); gb = (a + b) * n * 4 / 1e9
    print(f"{name}: {t*1e3:.1f} ms, {gb/t:.1f} GB/s total")
