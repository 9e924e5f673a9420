"""Measure host<->device copy bandwidth on this box (pinned memory, 1-D and pitched 2-D,
one direction and both at once) -- sizing information for the host-buffer path."""
import time
import torch

dev = torch.device('cuda', 0)
N = 256 << 20
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device=dev)
d_b = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = timeit(lambda: d_a.copy_(h_in, non_blocking=True))
print(f'H2D 256 MiB pinned 1-D : {N / t / 1e9:6.1f} GB/s')
t = timeit(lambda: h_out.copy_(d_b, non_blocking=True))
print(f'D2H 256 MiB pinned 1-D : {N / t / 1e9:6.1f} GB/s')


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


t = timeit(both)
print(f'H2D + D2H concurrently : {2 * N / t / 1e9:6.1f} GB/s total')
# pitched: 38 rows of 16 KiB chunk out of 128 KiB rows (what a chunked [row][ld] copy looks like)
rows, ld, n = 2048, 128 << 10, 32 << 10
hv = h_in[:rows * ld].view(rows, ld)
dv = d_a[:rows * n].view(rows, n)
t = timeit(lambda: dv.copy_(hv[:, :n], non_blocking=True))
print(f'H2D pitched 2048 x 32 KiB of 128 KiB rows : {rows * n / t / 1e9:6.1f} GB/s')
for sz in (1 << 20, 4 << 20, 16 << 20, 64 << 20):
    t = timeit(lambda: d_a[:sz].copy_(h_in[:sz], non_blocking=True), reps=20)
    t2 = timeit(lambda: h_out[:sz].copy_(d_b[:sz], non_blocking=True), reps=20)
    print(f'{sz >> 20:3d} MiB: H2D {sz / t / 1e9:6.1f} GB/s  D2H {sz / t2 / 1e9:6.1f} GB/s')
