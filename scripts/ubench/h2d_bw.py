import torch, time
n = 1 << 30
for streams in (1, 2, 4):
    hs = [torch.empty(n // streams, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ds = [torch.empty(n // streams, dtype=torch.uint8, device="cuda") for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go():
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
    go(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(streams, "streams:", round(n / dt / 1e9, 1), "GB/s")
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
