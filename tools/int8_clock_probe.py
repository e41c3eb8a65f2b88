"""SM clock / power while the int8 contraction kernel runs back to back (is the tensor pipe power-limited?)."""
import os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import _native

M, N, K = 4096, 16384, 400
g = torch.Generator().manual_seed(0)
A = torch.randn(M, K, generator=g, dtype=torch.double).cuda()
B = torch.randn(N, K, generator=g, dtype=torch.double).cuda()
for _ in range(3):
    _native.int8_matmul(A, B)
samples = []
stop = False
def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.05)
th = threading.Thread(target=sampler); th.start()
t0 = time.time(); n = 0
while time.time() - t0 < 4.0:
    _native.int8_matmul(A, B); n += 1
stop = True; th.join()
print("calls", n, "wall per call ms", 4000.0 / n, "(includes slicing both operands + allocation)")
print("\n".join(samples[:: max(1, len(samples) // 12)]))
