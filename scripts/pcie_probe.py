"""How fast are pinned H2D / D2H copies on this box, alone and concurrently? (e2e pipeline budget)"""
import time, torch
n = 160 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print("H2D %.2f ms (%.1f GB/s)  D2H %.2f ms (%.1f GB/s)  both %.2f ms" % (a, n / a / 1e6, b, n / b / 1e6, c))
for chunk in (1, 4, 16, 64):
    m = chunk * 1024 * 1024
    def chunks():
        for o in range(0, n, m):
            with torch.cuda.stream(s1): d_in[o:o + m].copy_(h_in[o:o + m], non_blocking=True)
            with torch.cuda.stream(s2): h_out[o:o + m].copy_(d_out[o:o + m], non_blocking=True)
    print("chunk %d MiB both directions: %.2f ms" % (chunk, t(chunks)))
