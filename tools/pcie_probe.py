"""Development aid: raw PCIe copy bandwidth on this box (H2D, D2H, both at once), pinned host memory."""
import time
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, chunk, reps=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for a in range(0, n, chunk):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a[a:a + chunk].copy_(h_in[a:a + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[a:a + chunk].copy_(d_b[a:a + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9


for chunk in (1 << 30, 1 << 27, 1 << 24):
    run(True, True, chunk, 1)
    print("chunk %4d MiB  H2D %.1f GB/s  D2H %.1f GB/s  both %.1f GB/s per direction" % (
        chunk >> 20, run(True, False, chunk), run(False, True, chunk), run(True, True, chunk)), flush=True)
