"""Soak run: a synthetic island for many ticks, invariants checked along the way (no NaN, no overflow, gas conservation)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from sapiogenesis_b200.synth import build_world
n, ticks, every = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w, eng = build_world(n, seed=5)
eng.exclusive = True
gas0 = float(w.t["g_co2"][0] + w.t["g_o2"][0])
t0 = time.perf_counter()
for k in range(0, ticks, every):
    eng.tick(every); eng.synchronize()
    st = w.stats()
    live = w.t["c_alive"] == 1
    ok = bool(torch.isfinite(w.t["c_pos"][live]).all() and torch.isfinite(w.t["c_vel"][live]).all() and torch.isfinite(w.t["d_vel"][w.t["d_state"] > 0]).all()
              and torch.isfinite(w.t["org_brain"][w.t["org_state"] == 1]).all())
    gas = float(w.t["g_co2"][0] + w.t["g_o2"][0])
    vmax = float(w.t["c_vel"][live].abs().max()) if live.any() else 0.0
    sp = w.t["c_vel"].norm(dim=1) * live
    imax = int(sp.argmax()); omax = imax // 64
    fast = f"fast(>1500) {int((sp > 1500).sum())} fastest: org {omax} age {w.tick - int(w.t['org_birth_tick'][omax])} cells {int(w.t['org_ncells'][omax])} fails {st['birth_fail_place']}/{st['birth_fail_caps']}"
    print(f"tick {w.tick}: {fast} pop {int((w.t['org_state'] == 1).sum())} cells {int(live.sum())} dead {int((w.t['d_state'] > 0).sum())} births {st['births']} deaths {st['deaths']} "
          f"overflow {st['overflow']} finite {ok} gas drift {abs(gas - gas0) / gas0:.2e} vmax {vmax:.0f} xc {st['xcontacts']} coupled {st['coupled']} levels {st['levels']} "
          f"{(time.perf_counter() - t0) / (k + every) * 1e3:.2f} ms/tick", flush=True)
    assert ok
