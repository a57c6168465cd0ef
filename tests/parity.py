"""Field-by-field comparison of a GPU world against the oracle's world after the same stage(s) from identical state.

Bit-exact (north star): every integer / index / flag field -- living-cell sets, contact-pair sets (cross list, adjacency,
per-organism intra lists), decisions (requests, deaths, births, slots, uids, genome tables), memory counters.
fp32 fields: |gpu - oracle| <= tol * max(rms(field), floor), tol = 1e-5 unless stated per field below."""
from __future__ import annotations

import numpy as np

from sapiogenesis_b200.world import FIELDS, HIST, ICAP, MAXC, MEMROWS, NPAIR, STM

TOL = 1e-5
SKIP = {"org_gas_a", "org_gas_b", "c_dmg", "d_eaten", "org_rect", "g_stats", "org_out_raw"}
# fields whose meaningful extent depends on counters / rings: handled explicitly
SPECIAL = {"org_brain", "org_adam_m", "org_adam_v", "org_hist_in", "org_hist_out", "org_hist_dopa", "org_stm_in", "org_stm_out",
           "org_stm_rew", "org_mem_in", "org_mem_out", "org_mem_rew", "org_mem_seq", "org_incell", "org_ldim", "ic_code", "ic_jn", "ic_jt", "j_jn",
           "x_key", "x_jn", "x_jt", "xa_off", "xa_other", "c_spos", "c_svel", "c_svbias", "c_spin_jn", "c_wbias", "d_wbias"}
# dopamine / pain = (percent now - percent at the last snapshot) / dt with percents O(100) held in fp32 and dt >= 0.09 s: the
# subtraction carries 100 * 6e-8 / 0.09 ~ 7e-5 of absolute noise whatever the size of the result, so these signals are compared
# on the scale of the percent-per-second they are made of (floor 10), like every other fp32 field at 1e-5.
DOPA_FLOOR = 10.0
FLOORS = {"c_wbias": 1.0, "d_wbias": 1.0, "c_vbias": 1e-2, "d_vbias": 1e-2, "org_energy_diff": DOPA_FLOOR, "org_dopamine": DOPA_FLOOR,
          "org_pain": DOPA_FLOOR,
          # the refund is what is left of (energy + share) - storage along the add_energy carry chain: energies O(1e2..1e3) in fp32,
          # so its ABSOLUTE error is a few ulps of those (~1e-4) however small the refund itself: entries below 1 are held to 2e-4 absolute
          "org_gas_refund": 1.0, "org_move": 1.0}
# The rotate actuator adds max_rot * command to omega every tick (sprites.py:1547, not scaled by dt) against 1.8 %/tick of
# friction: cells of the test worlds spin at hundreds of rad/s, and ten Gauss-Seidel sweeps of friction impulses on top of
# that accumulate a few fp32 ulps of omega itself. 3e-5 of the field's rms for angular velocities, 1e-5 everywhere else.
TOLS = {"c_angvel": 3e-5, "d_angvel": 3e-5,
        # org_move is the clamped 3-decimal-rounded network output (neural.py:292, 320-349): at a rounding tie the fp32 forward's 1e-7
        # flips the last digit, like org_hist_out below -- 1.5e-3 absolute (floor 1), seen once in ~4000 thinks
        "org_move": 1.5e-3,
        # co2 refund = overflow of add_energy = (energy + share) - storage with energy ~ storage ~ 1e3 in fp32: the difference
        # carries the operands' 1e-7, i.e. ~1e-4 of its own size; its effect on the global co2 (~3e4, checked at 1e-5) is 1e-9
        "org_gas_refund": 2e-4}


def field_err(got, ref, floor=1e-3, elementwise=False):
    """max |got - ref| / scale, scale = max(rms(ref), floor); elementwise: scale_i = max(|ref_i|, rms(ref), floor), i.e.
    `tol` is a relative tolerance with an absolute floor of tol * rms (numpy.allclose semantics)."""
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    scale = max(float(np.sqrt(np.mean(ref * ref))), floor)
    if elementwise:
        return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), scale)))
    return float(np.max(np.abs(got - ref)) / scale)


def elem_rel_err(got, ref):
    """max over entries of |got - ref| / |ref|, entries with |ref| <= 1e-3 * rms(ref) left out (their relative error is noise)."""
    got = np.asarray(got, dtype=np.float64).reshape(-1); ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    if ref.size == 0:
        return 0.0
    rms = float(np.sqrt(np.mean(ref * ref)))
    m = np.abs(ref) > 1e-3 * rms
    return float(np.max(np.abs(got[m] - ref[m]) / np.abs(ref[m]))) if m.any() else 0.0


# cached impulses span orders of magnitude inside one list (a settled pile: one resting contact carries the weight, the rest
# are ~0): each entry is held to 1e-5 of its own size, with an absolute floor of 1e-5 of max(the list's rms, the momentum
# scale). The momentum scale m/2 * rms(v) is the size of the ten increments -(bounce + v_rel.n) * nMass an accumulated
# impulse is the (mostly cancelling) sum of: an fp32 solver cannot resolve an accumulator finer than an ulp of its terms.
ELEMENTWISE = {"x_jn", "x_jt", "ic_jn", "ic_jt", "j_jn", "c_spin_jn", "org_gas_refund"}


def velocity_rms(world) -> float:
    live = world.np("c_alive") == 1; dead = world.np("d_state") > 0
    v = np.concatenate([world.np("c_vel")[live].reshape(-1), world.np("d_vel")[dead].reshape(-1)]).astype(np.float64)
    return float(np.sqrt(np.mean(v * v))) if v.size else 0.0


def momentum_scale(world) -> float:
    live = world.np("c_alive") == 1; dead = world.np("d_state") > 0
    m = np.concatenate([world.np("c_mass")[live], world.np("d_mass")[dead]]).astype(np.float64)
    v = np.concatenate([world.np("c_vel")[live].reshape(-1), world.np("d_vel")[dead].reshape(-1)]).astype(np.float64)
    return 0.5 * float(np.median(m)) * float(np.sqrt(np.mean(v * v))) if m.size else 0.0


# outputs of the iterative solver: in a dense pile their fp32 error grows with the length of the contact chains; a test may
# state a looser tolerance for them (phys_tol) while every other field keeps its own
SOLVER_FIELDS = {"c_vel", "c_angvel", "c_vbias", "c_svel", "c_spin_jn", "c_wbias", "d_vel", "d_angvel", "d_vbias", "d_wbias",
                 "x_jn", "x_jt", "ic_jn", "ic_jt", "j_jn"}


class Report:
    def __init__(self, tag, phys_tol=None, extra_tols=None):
        self.tag, self.errs, self.bad, self.phys_tol, self.extra = tag, {}, [], phys_tol, dict(extra_tols or {})
        self.rel = {}          # per-element relative error max |got - ref| / |ref| over entries with |ref| > 1e-3 * rms: reported, not asserted

    def exact(self, name, got, ref):
        got, ref = np.asarray(got), np.asarray(ref)
        if got.shape != ref.shape or not np.array_equal(got, ref):
            n = int(np.sum(got != ref)) if got.shape == ref.shape else -1
            where = np.argwhere(got != ref)[:5].tolist() if got.shape == ref.shape else []
            self.bad.append(f"{name}: {n} of {ref.size} entries differ, first at {where}")

    def close(self, name, got, ref, tol=None, floor=None):
        e = field_err(got, ref, FLOORS.get(name, 1e-3) if floor is None else floor, elementwise=name in ELEMENTWISE)
        self.errs[name] = max(e, self.errs.get(name, 0.0))
        self.rel[name] = max(elem_rel_err(got, ref), self.rel.get(name, 0.0))
        lim = TOLS.get(name, TOL) if tol is None else tol
        if self.phys_tol is not None and name in SOLVER_FIELDS:
            lim = max(lim, self.phys_tol)
        if name in self.extra:
            lim = max(lim, self.extra[name])
        if not e <= lim:
            self.bad.append(f"{name}: {e:.3e}")

    def finish(self, verbose=True):
        if verbose:
            worst = sorted(self.errs.items(), key=lambda kv: -kv[1])[:8]
            print(f"{self.tag}: worst " + " ".join(f"{k}={v:.1e}" for k, v in worst))
            wr = sorted(self.rel.items(), key=lambda kv: -kv[1])[:8]
            print(f"{self.tag}: per-element relative (reported) " + " ".join(f"{k}={v:.1e}" for k, v in wr))
        assert not self.bad, f"{self.tag}: " + "; ".join(self.bad[:12])


class _Lazy(dict):
    """Field name -> host array, fetched from the world on first use (a 100k-organism world is tens of GB: one field at a time)."""
    def __init__(self, world):
        super().__init__(); self.world = world

    def __missing__(self, k):
        v = self.world.np(k); self[k] = v
        return v

    def drop(self, *names):
        for n in names:
            self.pop(n, None)


_PAIR_I, _PAIR_J = np.triu_indices(MAXC, 1)            # lexicographic (i, j), i < j == pair_index order


def compare_worlds(wg, w, tag, brain_tol=TOL, check_brain=True, skip=(), verbose=True, phys_tol=None, detail_orgs=None, before=None, adam_tol=1e-4,
                   extra_tols=None):
    """wg: GPU world after the stage(s); w: oracle world after the same stage(s).
    detail_orgs: organism slots whose brains / rings / memories are compared entry by entry (default: all). With `before` (the common
    start state) every OTHER living organism's brain, optimiser state and memories must be bit-identical on both sides to what they
    were before -- nothing may touch an organism that neither thought nor trained nor was born."""
    rep = Report(tag, phys_tol, extra_tols)
    p = w.p
    G = _Lazy(wg)
    O = _Lazy(w)
    rep.exact("org_state", G["org_state"], O["org_state"])
    orgs = np.nonzero(O["org_state"] == 1)[0]
    # organisms whose network output sits on a 3-decimal rounding tie this tick (see below): their movement command differs by 1e-3 on
    # the two sides, the push actuator turns that into a different velocity of their own cells -- left out of the fp32 comparisons
    tie_orgs = np.zeros(p.n_org, dtype=bool)
    if detail_orgs is not None and len(detail_orgs):
        dsel = np.intersect1d(orgs, np.asarray(detail_orgs, dtype=np.int64))
        mv = np.abs(G["org_move"][dsel].astype(np.float64) - O["org_move"][dsel]).reshape(len(dsel), -1).max(axis=1) if len(dsel) else np.zeros(0)
        tie_orgs[dsel[mv > 5e-4]] = True
    tie_dead = np.zeros(p.n_dead, dtype=bool)
    if tie_orgs.any():
        # ... and so does everything the contact solver couples to them in this tick: the connected component of the cross-contact graph
        nx_ = int(O["g_x_n"][0]); key = O["x_key"][:nx_].astype(np.uint64)
        NCb = p.n_org * MAXC
        ka = (key >> np.uint64(32)).astype(np.int64); kb = (key & np.uint64(0xFFFFFFFF)).astype(np.int64)
        node = lambda b: np.where(b < NCb, b // MAXC, np.where(b < NCb + p.n_dead, p.n_org + (b - NCb), -1))
        na_, nb_ = node(ka), node(kb)
        ok = nb_ >= 0
        tainted = np.concatenate([tie_orgs, tie_dead])
        for _ in range(64):
            grow = np.zeros_like(tainted)
            grow[nb_[ok][tainted[na_[ok]]]] = True; grow[na_[ok][tainted[nb_[ok]]]] = True
            if not (grow & ~tainted).any():
                break
            tainted |= grow
        tie_orgs, tie_dead = tainted[: p.n_org].copy(), tainted[p.n_org:].copy()
    tie_rows = np.repeat(tie_orgs, MAXC)
    alive_org = O["org_state"] == 1
    org_rows = ((np.arange(MAXC)[None, :] < O["org_ncells"].astype(np.int64)[:, None]) & alive_org[:, None]).reshape(-1)
    rep.exact("c_alive", G["c_alive"][org_rows], O["c_alive"][org_rows])
    live = org_rows & (O["c_alive"] == 1)
    rep.exact("d_state", G["d_state"] > 0, O["d_state"] > 0)
    dead = O["d_state"] > 0
    sens = live & np.isin(O["c_type"], (2, 3))
    for f in FIELDS:
        name = f.name
        if name in SKIP or name in SPECIAL or name in skip or name in ("org_state", "c_alive", "d_state"):
            continue
        g, o = G[name], O[name]
        if f.scope == "ORG":
            sel = orgs[~tie_orgs[orgs]] if f.ctype in ("float", "double") else orgs
            g, o = g[sel], o[sel]
        elif f.scope == "CELL":
            m = org_rows if name in ("c_type", "c_parent", "c_mirror", "c_gorder", "c_id", "c_size", "c_mass", "c_elast", "c_fric", "c_rel",
                                     "c_grow", "c_maxhealth", "c_use_idle", "c_use_act", "c_storage", "c_damage") else live
            if f.ctype in ("float", "double"):
                m = m & ~tie_rows
            g, o = g[m], o[m]
        elif f.scope == "DEAD":
            md = dead & ~tie_dead if f.ctype in ("float", "double") else dead
            g, o = g[md], o[md]
        elif f.scope == "GLOB":
            pass
        else:
            continue
        if f.ctype in ("float", "double"):
            if name in ("c_angvel", "d_angvel"):
                # omega * radius is a rim speed: the error of an angular velocity is measured on the scale of the linear
                # velocities over a typical cell radius (10 px)
                vel = O["c_vel"][live] if name == "c_angvel" else O["d_vel"][dead]
                rep.close(name, g, o, tol=TOLS[name], floor=max(1e-3, (float(np.sqrt(np.mean(vel.astype(np.float64) ** 2))) if vel.size else 0.0) / 10.0))
            else:
                rep.close(name, g, o)
        else:
            rep.exact(name, g, o)
        if G[name].nbytes > (1 << 26) and name not in ("c_vel", "d_vel", "c_type", "c_mass", "d_mass"):
            G.drop(name); O.drop(name)
    if verbose:                                             # where the largest velocity error sits: organism, size, contacts, its own velocity scale
        gv, ov = G["c_vel"].astype(np.float64), O["c_vel"].astype(np.float64)
        err = np.abs(gv - ov).max(axis=1) * (live & ~tie_rows)
        r = int(np.argmax(err)); oo = r // MAXC
        rows = slice(oo * MAXC, oo * MAXC + int(O["org_ncells"][oo]))
        lv = O["c_alive"][rows] == 1
        deg = O["xa_off"][oo * MAXC + 1: oo * MAXC + MAXC + 1] - O["xa_off"][oo * MAXC: oo * MAXC + MAXC]
        vmag = np.sqrt((ov[rows][lv] ** 2).sum(axis=1))
        bmag = np.sqrt((before.np("c_vel").astype(np.float64)[rows][lv] ** 2).sum(axis=1)) if before is not None else vmag * 0
        print(f"{tag}: largest c_vel error {err[r]:.3e} px/s at row {r} (organism {oo}, cell {r % MAXC}): rows {int(O['org_ncells'][oo])}, living {int(lv.sum())}, "
              f"intra contacts {int(O['org_ic_n'][oo])}, cross partners {int(deg.sum())}, |v| of its cells max {vmag.max():.2f} (before the tick {bmag.max():.2f}), "
              f"gpu {gv[r]} oracle {ov[r]}")
    # integrated positions are one fp32 fma of identical inputs: bit-exact (an ulp of position is 0.002 .. 0.016 px in a large world, and
    # the pin joints turn it into 0.01 .. 0.08 px/s of bias velocity)
    rep.exact("c_pos[exact]", G["c_pos"][live], O["c_pos"][live]); rep.exact("d_pos[exact]", G["d_pos"][dead], O["d_pos"][dead])
    # sensors
    sens_x = sens                                           # exact fields keep every sensor
    sens = sens & ~tie_rows
    for name in ("c_spos",):
        rep.exact(name, G[name][sens_x], O[name][sens_x])
    rep.close("c_svel", G["c_svel"][sens], O["c_svel"][sens])
    rep.close("c_spin_jn", G["c_spin_jn"][sens], O["c_spin_jn"][sens], floor=max(1e-3, velocity_rms(w)))   # sensor body: mass 1
    rep.close("c_wbias", G["c_wbias"][live], O["c_wbias"][live]); rep.close("d_wbias", G["d_wbias"][dead], O["d_wbias"][dead])
    # contacts
    nx = int(O["g_x_n"][0])
    rep.exact("x_key", G["x_key"][:nx], O["x_key"][:nx])
    # friction impulses are bounded by mu * jn: their error is measured on the scale of the normal impulses
    rms = lambda a: float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2))) if len(a) else 0.0
    pf = max(1e-3, momentum_scale(w))
    xm = np.ones(nx, dtype=bool)
    if tie_orgs.any():                                      # contacts of a tie organism's component (see tie_orgs)
        kk = O["x_key"][:nx].astype(np.uint64); NCb = p.n_org * MAXC
        for b in ((kk >> np.uint64(32)).astype(np.int64), (kk & np.uint64(0xFFFFFFFF)).astype(np.int64)):
            io = b < NCb; idd = (b >= NCb) & (b < NCb + p.n_dead)
            xm[io] &= ~tie_orgs[b[io] // MAXC]; xm[idd] &= ~tie_dead[b[idd] - NCb]
    rep.close("x_jn", G["x_jn"][:nx][xm], O["x_jn"][:nx][xm], floor=pf); rep.close("x_jt", G["x_jt"][:nx][xm], O["x_jt"][:nx][xm], floor=max(pf, rms(O["x_jn"][:nx])))
    rep.exact("xa_off", G["xa_off"], O["xa_off"])
    na = int(O["xa_off"][-1])
    rep.exact("xa_other", G["xa_other"].reshape(-1)[:na], O["xa_other"].reshape(-1)[:na])
    icm = ((np.arange(ICAP)[None, :] < O["org_ic_n"].astype(np.int64)[:, None]) & alive_org[:, None]).reshape(-1)
    icm_f = icm & np.repeat(~tie_orgs, ICAP)                # impulses of a tie organism's own contacts and pins: see tie_orgs
    al2 = (live & ~tie_rows).reshape(p.n_org, MAXC)
    jm = np.empty((p.n_org, NPAIR), dtype=bool)
    for s0 in range(0, p.n_org, 8192):                   # pin (i, j) exists iff both cells live
        jm[s0:s0 + 8192] = al2[s0:s0 + 8192][:, _PAIR_I] & al2[s0:s0 + 8192][:, _PAIR_J]
    jm = jm.reshape(-1)
    rep.exact("ic_code", G["ic_code"][icm], O["ic_code"][icm])
    rep.close("ic_jn", G["ic_jn"][icm_f], O["ic_jn"][icm_f], floor=pf); rep.close("ic_jt", G["ic_jt"][icm_f], O["ic_jt"][icm_f], floor=max(pf, rms(O["ic_jn"][icm_f])))
    rep.close("j_jn", G["j_jn"][jm], O["j_jn"][jm], floor=pf)
    G.drop("j_jn", "ic_jn", "ic_jt", "ic_code"); O.drop("j_jn", "ic_jn", "ic_jt", "ic_code")
    del jm, icm, icm_f
    # brains and memories
    IN = p.in_cap
    detail = orgs if detail_orgs is None else np.intersect1d(orgs, np.asarray(detail_orgs, dtype=np.int64))
    if detail_orgs is not None:
        # everybody else neither thought nor trained nor was born this tick: brain, optimiser state, rings and memories must be
        # the same bytes on both sides (and, with `before`, the bytes they were)
        rest = np.setdiff1d(orgs, detail)
        for name in ("org_brain", "org_adam_m", "org_adam_v", "org_mem_in", "org_mem_out", "org_mem_rew", "org_stm_in", "org_stm_out",
                     "org_stm_rew", "org_hist_in", "org_hist_out", "org_hist_dopa", "org_ldim", "org_incell"):
            g, o = G[name], O[name]
            b = before.np(name) if before is not None else None
            ndiff = 0
            for s0 in range(0, len(rest), 2048):
                r = rest[s0:s0 + 2048]
                gv = g[r].view(np.uint8)
                ndiff += int(np.sum(gv != o[r].view(np.uint8)))
                if b is not None:
                    ndiff += int(np.sum(gv != b[r].view(np.uint8)))
            if ndiff:
                rep.bad.append(f"{name}[untouched organisms]: {ndiff} bytes differ")
            if name not in ("org_ldim", "org_incell") and len(detail) == 0:
                G.drop(name); O.drop(name)
    ties = 0
    for o in detail:
        nl = int(O["org_nlayers"][o]); ld = O["org_ldim"][o]
        rep.exact(f"org_ldim[{o}]", G["org_ldim"][o][: nl + 1], ld[: nl + 1])
        nin = int(O["org_nincell"][o])
        rep.exact(f"org_incell[{o}]", G["org_incell"][o][:nin], O["org_incell"][o][:nin])
        if G["org_nlayers"][o] != nl:
            continue
        I = int(ld[0])
        H = int(O["org_rnn_h"][o])                           # RNN brain: H more input columns and the hidden head after the output layer
        P = sum(int(ld[l + 1]) * (int(ld[l]) + (H if l == 0 else 0)) + int(ld[l + 1]) for l in range(nl)) + (H * int(ld[nl - 1]) + H if H else 0)
        # Rounding cliff: the network output is rounded to 3 decimals (neural.py:292); at a tie the fp32 forward's 1e-7
        # flips the last digit, and the flipped row is a training TARGET of the same tick (|output - target| <= 5e-4 becomes 1.5e-3):
        # the gradient of that organism is a different one on the two sides. Such organisms are counted, held to the 1.5e-3 of the
        # rounded outputs below, and left out of the brain / optimiser comparison; at most 2 (or 0.2 %) per comparison.
        hn_ = min(int(O["org_hist_n"][o]), HIST); sn_ = min(int(O["org_stm_n"][o]), STM); mn_ = int(O["org_mem_n"][o])
        tie = any(float(np.max(np.abs(G[k][o][: c * 4].astype(np.float64) - O[k][o][: c * 4]), initial=0.0)) > 5e-4
                  for k, c in (("org_hist_out", hn_), ("org_stm_out", sn_), ("org_mem_out", mn_)))
        ties += tie
        if check_brain and not tie:
            gb, ob = G["org_brain"][o][:P].astype(np.float64), O["org_brain"][o][:P].astype(np.float64)
            scale = np.maximum(np.abs(ob), np.sqrt(np.mean(ob ** 2)) if P else 1.0)
            rel = np.abs(gb - ob) / scale
            # Trained weights. Adam moves each weight by lr * m_hat / (sqrt(v_hat) + 1e-8): on the first steps that is
            # lr * g / (|g| + 1e-8), ill-conditioned in the gradient's own fp32 noise wherever |g| is within a few orders of
            # 1e-8 (common right after birth, when the batch holds one or two samples). The well-conditioned quantities are the
            # gradient moments themselves: they must agree to 1e-5 of their scale. For the weights: >= 90 % of an organism's
            # parameters within 1e-5 of the weight scale, none further than 1 % of one Adam step (lr = 0.02).
            if P and (np.mean(rel <= brain_tol) < 0.90 or np.max(np.abs(gb - ob)) > 1e-2 * 0.02):
                rep.bad.append(f"org_brain[{o}]: {np.mean(rel <= brain_tol):.4f} within {brain_tol}, max abs {np.max(np.abs(gb - ob)):.2e}")
            if int(O["org_adam_t"][o]) > 0:
                h0, hl = int(ld[1]), int(ld[nl - 1])
                npar = h0 * I + h0 + 4 * hl + 4
                # the gradient is built on (output - target) where the target is an earlier, 3-decimal-rounded output of the
                # same network: |output - target| <= 5e-4, so the fp32 forward's 1e-7 is 1e-3 of it. The moments are therefore
                # compared on the scale of the outputs (O(1)), not of their own (tiny) magnitude.
                rep.close("org_adam_m", G["org_adam_m"][o][:npar], O["org_adam_m"][o][:npar], tol=adam_tol, floor=1e-3)
                rep.close("org_adam_v", G["org_adam_v"][o][:npar], O["org_adam_v"][o][:npar], tol=adam_tol, floor=1e-6)
            rep.errs["org_brain"] = max(rep.errs.get("org_brain", 0.0), float(np.max(rel)) if P else 0.0)
        hn = min(int(O["org_hist_n"][o]), HIST)
        rep.exact(f"org_hist_in[{o}]", G["org_hist_in"][o].reshape(HIST, IN)[:hn, :I], O["org_hist_in"][o].reshape(HIST, IN)[:hn, :I])
        rep.close("org_hist_out", G["org_hist_out"][o][: hn * 4], O["org_hist_out"][o][: hn * 4], tol=1.5e-3, floor=1.0)
        rep.close("org_hist_dopa", G["org_hist_dopa"][o][:hn], O["org_hist_dopa"][o][:hn], floor=DOPA_FLOOR)
        sn = min(int(O["org_stm_n"][o]), STM)
        rep.exact(f"org_stm_in[{o}]", G["org_stm_in"][o].reshape(STM, IN)[:sn, :I], O["org_stm_in"][o].reshape(STM, IN)[:sn, :I])
        rep.close("org_stm_out", G["org_stm_out"][o][: sn * 4], O["org_stm_out"][o][: sn * 4], tol=1.5e-3, floor=1.0)
        rep.close("org_stm_rew", G["org_stm_rew"][o][:sn], O["org_stm_rew"][o][:sn], floor=DOPA_FLOOR)
        M = int(O["org_mem_n"][o])
        if int(G["org_mem_n"][o]) == M and M:
            rep.exact(f"org_mem_in[{o}]", G["org_mem_in"][o].reshape(MEMROWS, IN)[:M, :I], O["org_mem_in"][o].reshape(MEMROWS, IN)[:M, :I])
            rep.close("org_mem_out", G["org_mem_out"][o][: M * 4], O["org_mem_out"][o][: M * 4], tol=1.5e-3, floor=1.0)
            rep.close("org_mem_rew", G["org_mem_rew"][o][:M], O["org_mem_rew"][o][:M], floor=DOPA_FLOOR)
            rep.exact(f"org_mem_seq[{o}]", G["org_mem_seq"][o][:M], O["org_mem_seq"][o][:M])
    if ties:
        print(f"{tag}: {ties} of {len(detail)} organisms at a 3-decimal rounding tie of their outputs (brain / optimiser comparison skipped for them)")
        if ties > max(2, 0.002 * len(detail)):
            rep.bad.append(f"rounding ties: {ties} of {len(detail)} organisms")
    rep.ties = ties
    rep.finish(verbose)
    return rep
