"""Where the host-API tick spends its time beyond the kernels (run on the GPU box)."""
import time, statistics
import numpy as np, torch
from project_morpheus_b200 import weights
from project_morpheus_b200.engine import SnacEngine
from oracle import speechpipe_ref as sp

S = 1024
eng = SnacEngine(weights.random_state_dict(0, "w1"), device=0, precision="fp16")
tok = np.stack([sp.synth_codes(i, 4) for i in range(S)]).astype(np.int32)
keys = np.arange(S, dtype=np.uint64)
tok_dev = torch.from_numpy(tok).cuda()
pcm_dev = torch.empty((S, 2048), dtype=torch.int16, device="cuda"); st_dev = torch.empty((S,), dtype=torch.int32, device="cuda")
pin = torch.empty((S, 2048), dtype=torch.int16, pin_memory=True)
def med(f, n=30):
    xs = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); xs.append(time.perf_counter() - t0)
    return 1e3 * statistics.median(xs)
for _ in range(3):
    eng.decode_windows(tok, noise="philox", seed=1, keys=keys); eng.decode_windows_device(tok_dev, noise="philox", seed=1, keys=keys, pcm=pcm_dev, status=st_dev)
print("device tick wall ms", med(lambda: eng.decode_windows_device(tok_dev, noise="philox", seed=1, keys=keys, pcm=pcm_dev, status=st_dev)))
print("host tick wall ms  ", med(lambda: eng.decode_windows(tok, noise="philox", seed=1, keys=keys)))
print("D2H 4.2 MB ms      ", med(lambda: pin.copy_(pcm_dev, non_blocking=True)))
print("H2D 112 KB ms      ", med(lambda: tok_dev.copy_(torch.from_numpy(tok).pin_memory(), non_blocking=True)))
print("empty sync ms      ", med(lambda: None))
# CPU enqueue time of the device tick (call returns before the GPU finishes) and GPU time by events
enq, gpu = [], []
for _ in range(30):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    eng.decode_windows_device(tok_dev, noise="philox", seed=1, keys=keys, pcm=pcm_dev, status=st_dev)
    b.record(); enq.append(time.perf_counter() - t0); torch.cuda.synchronize(); gpu.append(a.elapsed_time(b))
print("device tick: CPU enqueue ms", 1e3 * statistics.median(enq), "GPU event ms", statistics.median(gpu))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
gpu = []
for _ in range(30):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_(); a.record()
    eng.decode_windows_device(tok_dev, noise="philox", seed=1, keys=keys, pcm=pcm_dev, status=st_dev)
    b.record(); torch.cuda.synchronize(); gpu.append(a.elapsed_time(b))
print("device tick behind an L2 flush: GPU event ms", statistics.median(gpu))
