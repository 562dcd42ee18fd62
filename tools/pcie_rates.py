"""PCIe floor for the end-to-end figure: 512 MiB pinned host <-> device, each direction alone and both at once."""
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda:0")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    s1.synchronize()
    s2.synchronize()


t_in = timed(lambda: d_a.copy_(h_in, non_blocking=True))
t_out = timed(lambda: h_out.copy_(d_b, non_blocking=True))
t_both = timed(both)
gb = n / 1e9
print(f"H2D {gb / t_in * 1e3:.1f} GB/s ({t_in:.2f} ms), D2H {gb / t_out * 1e3:.1f} GB/s ({t_out:.2f} ms), both at once {t_both:.2f} ms "
      f"({gb / t_both * 1e3:.1f} GB/s each way)")
