"""Pinned-memory D2H / H2D bandwidth of the box (context for the PCIe-bound e2e number)."""
import time, torch
dev = torch.device("cuda", 0)
a = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
for name, (dst, src) in (("D2H", (h, a)), ("H2D", (a, h))):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print(f"{name}: {8 * (256 << 20) / (time.perf_counter() - t0) / 1e9:.1f} GB/s pinned", flush=True)
