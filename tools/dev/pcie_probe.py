import torch, time
n_in, n_out = 262144000, 385351680
hi = torch.empty(n_in // 4, dtype=torch.float32).pin_memory()
ho = torch.empty(n_out // 4, dtype=torch.float32).pin_memory()
di = torch.empty(n_in // 4, dtype=torch.float32, device="cuda")
do = torch.empty(n_out // 4, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for _ in range(2): run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D 262MB {a:.2f} ms ({n_in/a/1e6:.1f} GB/s)  D2H 385MB {b:.2f} ms ({n_out/b/1e6:.1f} GB/s)  both {c:.2f} ms")
