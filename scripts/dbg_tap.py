import numpy as np, torch, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
torch.set_grad_enabled(False)
from helpers import windows_tokens
from oracle import snac_ref, speechpipe_ref as sp
from project_morpheus_b200 import weights
from project_morpheus_b200.engine import SnacEngine
from test_gpu_parity import _oracle_taps
sd = weights.random_state_dict(0, "w1")
oracle = snac_ref.SNAC.from_state_dict(sd).eval()
n, F = 3, 4
tok = windows_tokens(n, F, base_stream=100)
noise = snac_ref.make_noise(n, F, seed=5)
lv = [sp.split_levels(row.tolist()) for row in tok]
codes = [torch.from_numpy(np.stack([l[k] for l in lv]).astype(np.int64)) for k in range(3)]
want = _oracle_taps(oracle, codes, noise)
eng = SnacEngine(sd, device=0, precision="fp32", trim=False)
for rep in range(2):
  for stage in sorted(want)[:14]:
    eng.set_tap(stage, 3 * 8192 * 1024)
    eng.decode_windows_device(torch.from_numpy(tok).cuda(), noise=snac_ref.pack_noise(noise))
    got, lo = eng.get_tap(); got = got.cpu().numpy()
    ref = want[stage][:, lo: lo + got.shape[1], :]
    d = np.abs(got - ref); i = np.unravel_index(d.argmax(), d.shape)
    print(rep, stage, got.shape, "maxerr", d.max(), "at", i, "ref", ref[i], "got", got[i], "scale", np.abs(ref).max())
    if stage == 3:
        x2 = want[2][:, lo: lo + got.shape[1], :]
        print("   input x at that point:", x2[i])
