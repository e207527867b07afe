"""How fast is THIS box right now?  cuBLAS bf16 8192^3 for ~1.5 s and a 2 GiB device copy; prints TF/s, GB/s and clocks.
(gpurun boxes differ by up to 30 % under the shared power cap: compare kernels only within one call.)"""
import subprocess, torch
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn_like(a)
for _ in range(5): a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
n = 0
import time
t0 = time.time()
while time.time() - t0 < 1.5:
    for _ in range(10): a @ b
    n += 10
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,power.limit,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
tf = n * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
x = torch.empty(1 << 30, device="cuda", dtype=torch.bfloat16); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize(); e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
gb = 10 * 2 * x.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9
print("box speed: cuBLAS bf16 %.0f TF/s sustained, copy %.0f GB/s | after load: %s" % (tf, gb, q))
# write-only and read-only HBM bandwidth (is a write-heavy kernel bound below the copy figure?)
z = torch.empty(1 << 31, device="cuda", dtype=torch.uint8)
for _ in range(3): z.zero_()
torch.cuda.synchronize(); e0.record()
for _ in range(10): z.zero_()
e1.record(); torch.cuda.synchronize()
wr = 10 * z.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
zf = z.view(torch.float32)
for _ in range(3): zf.sum()
torch.cuda.synchronize(); e0.record()
for _ in range(10): zf.sum()
e1.record(); torch.cuda.synchronize()
rd = 10 * z.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
print("box speed: write-only (memset 2 GiB) %.0f GB/s, read-only (sum 2 GiB) %.0f GB/s" % (wr, rd))
