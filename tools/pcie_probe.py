import torch, time
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
print("H2D GB/s", 1.0737 / t(h2d)); print("D2H GB/s", 1.0737 / t(d2h)); print("both: each GB/s", 1.0737 / t(both))
def chunks():
    for k in range(8):
        sl = slice(k * n // 8, (k + 1) * n // 8)
        with torch.cuda.stream(s1): d1[sl].copy_(h1[sl], non_blocking=True)
        with torch.cuda.stream(s2): h2[sl].copy_(d2[sl], non_blocking=True)
print("both chunked: each GB/s", 1.0737 / t(chunks))
