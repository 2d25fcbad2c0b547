"""PCIe probe: pinned H2D / D2H bandwidth alone and concurrently (what bounds bench.py's e2e leg)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (4, 32, 256):
    n = mb * (1 << 20) // 4
    reps = max(4, 2048 // mb)
    h_in = torch.empty(n).pin_memory(); h_out = torch.empty(n).pin_memory()
    d_in = torch.empty(n, device=dev); d_out = torch.empty(n, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); return mb * reps / 1024 / (time.perf_counter() - t)
    run(True, True)
    print(f"{mb:4d} MB copies: H2D {run(True, False):5.1f} GB/s  D2H {run(False, True):5.1f} GB/s  both {run(True, True):5.1f} GB/s each way")
