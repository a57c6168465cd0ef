"""Run the same synthetic island twice from its seed and compare every field bit for bit (longer than the test in tests/test_gpu_fullsize.py)."""
import sys
sys.path.insert(0, '.')
import torch
from sapiogenesis_b200.synth import build_world
n, ticks = int(sys.argv[1]), int(sys.argv[2])
snaps = []
for _ in range(2):
    w, eng = build_world(n, seed=7)
    eng.exclusive = True
    eng.tick(ticks); eng.synchronize()
    snaps.append(({k: v.clone() for k, v in w.t.items() if k != "org_rect"}, w.stats()))
    del w, eng
    torch.cuda.empty_cache()
bad = [k for k in snaps[0][0] if not (torch.equal(snaps[0][0][k], snaps[1][0][k]) or torch.equal(torch.nan_to_num(snaps[0][0][k].double()), torch.nan_to_num(snaps[1][0][k].double())))]
print("stats", snaps[0][1])
print("fields that differ:", bad if bad else "none", "| stats equal:", snaps[0][1] == snaps[1][1])
