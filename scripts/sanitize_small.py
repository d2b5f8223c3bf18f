"""Small decode through every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck)."""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from project_morpheus_b200 import weights
from project_morpheus_b200.engine import SnacEngine
sd = weights.random_state_dict(0, "w1")
tok = np.stack([np.random.Generator(np.random.PCG64(1234 + i)).integers(1, 4096, 28) for i in range(3)]).astype(np.int32)
for kw in (dict(precision="fp16"), dict(precision="fp16", persistent_ru=True), dict(precision="fp16", fuse_ru=False, fuse_convt_noise=False), dict(precision="fp32")):
    eng = SnacEngine(sd, device=0, **kw)
    pcm, st = eng.decode_windows(tok, noise="philox", seed=1, keys=[1, 2, 3])
    pcm, st = eng.decode_windows(tok, noise="philox", seed=1, keys=[1, 2, 3])   # graph capture
    pcm, st = eng.decode_windows(tok, noise="philox", seed=1, keys=[1, 2, 3])   # graph replay
    c = [torch.randint(0, 4096, (1, 20 * k), dtype=torch.int32) for k in (1, 2, 4)]
    wav = eng.decode_codes(c, noise="off")                                       # tiled long path
    torch.cuda.synchronize()
    print(kw, int(st.sum()), float(np.abs(pcm).max()), tuple(wav.shape))
    eng.close()
print("sanitize_small ok")
