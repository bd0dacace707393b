"""Development aid: pinned host -> device copy bandwidth on this box (one stream, two streams, small/large chunks)."""
import time
import torch

dev = torch.device("cuda:0")
GB = 1 << 30
host = torch.empty(8 * GB, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dst = torch.empty(8 * GB, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()


def run(chunk, nstreams):
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i, off in enumerate(range(0, host.numel(), chunk)):
        with torch.cuda.stream(streams[i % nstreams]):
            dst[off:off + chunk].copy_(host[off:off + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"chunk {chunk / 2**20:7.0f} MiB x {nstreams} stream(s): {host.numel() / dt / 1e9:6.1f} GB/s", flush=True)


for chunk in (64 << 20, 503316480, 2 * GB):
    for ns in (1, 2):
        run(chunk, ns)
# with a bandwidth-heavy kernel running on the GPU at the same time
a = torch.empty(4 * GB // 8, dtype=torch.float64, device=dev)
b = torch.empty_like(a)
side = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(side):
    for _ in range(60):
        b.copy_(a)
for off in range(0, host.numel(), 503316480):
    dst[off:off + 503316480].copy_(host[off:off + 503316480], non_blocking=True)
torch.cuda.current_stream().synchronize()
dt = time.perf_counter() - t0
print(f"503 MiB chunks while a D2D copy loop runs: {host.numel() / dt / 1e9:6.1f} GB/s", flush=True)
torch.cuda.synchronize()
