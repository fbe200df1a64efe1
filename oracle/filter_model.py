"""TEST INFRASTRUCTURE — numpy model of the certified fast policy evaluation of the B200 descent (csrc/descend_fx.cu).

Two evaluations of the regularised policy at one node (policy + newton_search + the sampling loop of descend_kernel,
boardlaw/mcts/cpp/cpu.cpp:38-102,138-175), vectorised over envs:

  * ``exact_eval``: the reference's arithmetic restated operation by operation in numpy float32 (sequential sums in the
    order a = 0..A-1, IEEE single division), i.e. what the reference computes bit for bit;
  * ``fast_eval``: the closed form the CUDA kernel evaluates.  Every child-less action has q = 0, so its terms share the
    divisor alpha:  S(alpha) = lambda*P/alpha + sum_children t_c/(alpha - q_c)  with P = sum of pi over child-less actions;
    O(children) work per Newton iteration instead of O(A).  The result is NOT bit-identical to the reference's sequential
    sums, so each decision the reference derives from them (stop / continue of the Newton loop, the sampled action) is
    CERTIFIED with a running bound on |reference - ours|, and an evaluation whose bound does not separate the decision is
    flagged for the exact path.  Certified decisions equal the reference's; flagged ones are recomputed exactly.

Error model (u = 2^-24, all terms are >= 0):
  reference, same alpha:  |S_ref - S| <= (A+2) u S   (A-term sequential sum of correctly rounded quotients, Higham gamma_n),
                          |g_ref - g| <= (A+3) u |g|
  ours, same alpha:       |S_ours - S| <= 32 u S, |g_ours - g| <= 32 u |g|  (a handful of roundings + approximate reciprocals)
  alpha:  both sequences are perturbed Newton iterations alpha' = N(alpha) + rho, N(x) = x - F/F', F = S - 1, so
          e' <= L e + R,  L >= sup N' = sup F F''/F'^2 over the segment between the two iterates,
          R = (cS_ref + cS_ours) S/|g| + (cG_ref + cG_ours + 4u) |F/g| + 2u alpha
  decisions at alpha:  |x_ref - x_ours| <= cS S + |g| e  for x = S - 1 and every prefix sum of the sampling loop.
Functions of alpha are evaluated at our iterate only; the guard e <= 2^-8 (alpha - q_max) makes them vary by < 2% across the
segment, which the inflation factors (1.05, 1.2) absorb.
"""
import numpy as np

f32 = np.float32
U = f32(2.0 ** -24)
OURS = f32(32.0)                 # our own roundings, in units of u
GUARD = f32(2.0 ** -8)
MAX_FAST_ITERS = 24
TINY = f32(7.888609052210118e-31)   # 2^-100: rows with a smaller nonzero lambda*pi take the exact path (denormal quotients)


class Stats:
    def __init__(self):
        self.evals = self.flag_stop = self.flag_sample = self.flag_guard = self.flag_tiny = 0
        self.bad_stop = self.bad_action = 0
        self.max_ne_ratio = self.max_cum_ratio = self.max_alpha_ratio = 0.
        self.iters = 0
        self.children = 0
        self.delta_sum = 0.
        self.resolved = self.resolve_passes = self.bad_resolved = 0
        self.resolve_hist = [0] * 10
        self.resolve_budget = 1 << 30

    def report(self):
        n = max(self.evals, 1)
        flagged = self.flag_stop + self.flag_sample + self.flag_guard + self.flag_tiny
        print(f'evaluations {self.evals}, newton iterations/eval {self.iters / n:.2f}, children/eval {self.children / n:.2f}')
        print(f'flagged {flagged} ({100 * flagged / n:.3f} %): stop {self.flag_stop}, sample {self.flag_sample}, guard {self.flag_guard}, tiny {self.flag_tiny}')
        print(f'certified but different: stop {self.bad_stop}, action {self.bad_action}  (must be 0)')
        print(f'largest observed deviation / bound: S-1 {self.max_ne_ratio:.3f}, prefix sums {self.max_cum_ratio:.3f}, alpha {self.max_alpha_ratio:.3f}')
        print(f'mean sampling bound delta {self.delta_sum / n:.3e}')
        print(f'flagged evaluations resolved by the kernel\'s control flow (eval_one): {self.resolved}, exact passes {self.resolve_passes} '
              f'({self.resolve_passes / max(self.resolved, 1):.2f} per flagged evaluation, histogram {self.resolve_hist}), wrong {self.bad_resolved} (must be 0)')


def exact_eval(pi, q, lam, r, trace=12):
    """pi, q (n,A) f32 (q = 0 where there is no child), lam, r (n,) f32.  Returns action (n,), iters (n,), the first `trace`
    iterates (alpha_k, ne_k) as (n,trace) arrays (nan beyond the last), and the final prefix sums (n,A)."""
    n, A = pi.shape
    top = (lam[:, None] * pi).astype(f32)
    alpha = np.zeros(n, f32)
    for a in range(A):
        alpha = np.maximum(alpha, q[:, a] + np.maximum(top[:, a], f32(1e-4)))
    error = np.full(n, np.inf, f32)
    active = np.ones(n, bool)
    iters = np.zeros(n, np.int64)
    tr_alpha = np.full((n, trace), np.nan, f32)
    tr_ne = np.full((n, trace), np.nan, f32)
    with np.errstate(all='ignore'):
        for it in range(100):
            if not active.any():
                break
            S = np.zeros(n, f32)
            g = np.zeros(n, f32)
            for a in range(A):
                bot = alpha - q[:, a]
                S = S + top[:, a] / bot
                g = g + (-top[:, a]) / (bot * bot)
            ne = S - f32(1.)
            if it < trace:
                tr_alpha[active, it] = alpha[active]
                tr_ne[active, it] = ne[active]
            iters[active] += 1
            stop = (ne < f32(1e-3)) | (error == ne)
            upd = active & ~stop
            alpha = np.where(upd, alpha - ne / g, alpha)
            error = np.where(upd, ne, error)
            active = upd
        # sampling loop
        total = np.zeros(n, f32)
        cum = np.zeros((n, A), f32)
        action = np.full(n, -1, np.int64)
        valid = np.full(n, -1, np.int64)
        done = np.zeros(n, bool)
        for a in range(A):
            p = top[:, a] / (alpha - q[:, a])
            total = total + p
            cum[:, a] = total
            hit = ~done & (p > 0) & (total >= r)
            action = np.where(hit, a, action)
            done |= hit
            valid = np.where(~done & (p > 0), a, valid)
        action = np.where(action >= 0, action, valid)
    return action, iters, tr_alpha, tr_ne, cum, alpha


def fast_eval(pi, q, child, lam, r, lanes=8):
    """The certified closed-form evaluation.  child (n,A) bool.  Returns (action, iters, flag, detail) where flag != 0 means
    'not certified' (1 stop test, 2 sample, 3 guard, 4 tiny) and detail carries the iterates / bounds for the study."""
    n, A = pi.shape
    top = (lam[:, None] * pi).astype(f32)
    f64 = np.float64
    with np.errstate(all='ignore'):
        alpha = np.max(q + np.maximum(top, f32(1e-4)), axis=1).astype(f32)                   # the exact seed (a max of exact values)
        nz = top > 0
        first_nz = np.where(nz.any(1), nz.argmax(1), -1)
        last_nz = np.where(nz.any(1), A - 1 - nz[:, ::-1].argmax(1), -1)
        tiny = np.where(nz, top, np.inf).min(1) < TINY
        # once per evaluation, in double: mass and "how many later additions see this term" weight of the child-less actions
        wgt = np.maximum(last_nz[:, None] + 1 - np.arange(A)[None], 0).astype(f64)            # >= number of rounded additions at or after a
        pi_cl = np.where(child, 0, pi).astype(f64)
        P_cl = pi_cl.sum(1).astype(f32)
        W_cl = (pi_cl * wgt).sum(1).astype(f32)
        nc = child.sum(1)
        c_ours = (f32(8) + np.ceil(nc / lanes).astype(f32) + f32(np.log2(lanes))) * U         # our own roundings (kernel: see descend_fx.cu)
        qmax = np.where(child, q, 0).max(1)
        tc = np.where(child, top, 0).astype(f32)
        wc = np.where(child, wgt, 0).astype(f32)
        e = np.zeros(n, f32)
        flag = np.where(tiny, 4, 0)
        active = flag == 0
        iters = np.zeros(n, np.int64)
        ne_prev = np.full(n, np.inf, f32)
        D_prev = np.zeros(n, f32)
        trace = 12
        tr_alpha = np.full((n, trace), np.nan, f32); tr_ne = np.full((n, trace), np.nan, f32)
        tr_D = np.full((n, trace), np.nan, f32); tr_e = np.full((n, trace), np.nan, f32)
        S = np.zeros(n, f32); G = np.ones(n, f32); ES = np.zeros(n, f32)
        for it in range(MAX_FAST_ITERS + 1):
            if not active.any():
                break
            d = (alpha[:, None] - q).astype(f32)
            rc = (f32(1) / d).astype(f32)
            s_c = np.where(child, tc * rc, 0).astype(f32)
            g_c = (s_c * rc).astype(f32)
            h_c = (g_c * rc).astype(f32)
            ra = (f32(1) / alpha).astype(f32)
            k = (lam * ra).astype(f32)
            kg = (k * ra).astype(f32)
            S_n = (k * P_cl + s_c.sum(1, dtype=f32)).astype(f32)
            G_n = (kg * P_cl + g_c.sum(1, dtype=f32)).astype(f32)            # |g|
            H_n = f32(2) * (kg * ra * P_cl + h_c.sum(1, dtype=f32))
            # the reference's own rounding: u * (sum over its rounded additions of the partial sum) + term roundings
            ES_n = (U * f32(1.01) * (k * W_cl + (wc * s_c).sum(1, dtype=f32)) + (f32(2) * U + c_ours) * S_n).astype(f32)
            EG_n = (U * f32(1.01) * (kg * W_cl + (wc * g_c).sum(1, dtype=f32)) + (f32(4) * U + c_ours) * G_n).astype(f32)
            S = np.where(active, S_n, S); G = np.where(active, G_n, G); ES = np.where(active, ES_n, ES)
            ne = (S_n - f32(1)).astype(f32)
            Dk = (ES_n + G_n * e * f32(1.05) + f32(2) * U * np.abs(ne)).astype(f32)
            if it < trace:
                tr_alpha[active, it] = alpha[active]; tr_ne[active, it] = ne[active]; tr_D[active, it] = Dk[active]; tr_e[active, it] = e[active]
            guard_ok = (e <= GUARD * (alpha - qmax)) & np.isfinite(S_n) & np.isfinite(G_n) & (G_n > 0) & (it < MAX_FAST_ITERS)
            stop_sure = ne < f32(1e-3) - Dk
            cont_sure = (ne > f32(1e-3) + Dk) & (np.abs(ne - ne_prev) > Dk + D_prev)
            iters[active] += 1
            newflag = np.where(~guard_ok, 3, np.where(stop_sure | cont_sure, 0, 1))
            flag = np.where(active & (newflag != 0), newflag, flag)
            cont = active & (newflag == 0) & cont_sure
            # error recurrence for the next iterate: e' <= L e + R bounds the distance of the two UNROUNDED updates; both are then
            # rounded to fp32, and when no rounding boundary lies within that distance of ours the reference's alpha is our float
            # exactly (e' = 0) — which matters because |g| e is the sensitivity of every later decision to alpha
            L = f32(1.2) * (np.maximum(ne, 0) + f32(2) * G_n * e) * H_n / (G_n * G_n)
            R = f32(1.05) * (Dk / G_n + np.abs(ne) / G_n * (EG_n / G_n + f32(4) * U))
            eps = (L * e + R).astype(f32)
            step = (ne / G_n).astype(f32)
            a_new = (alpha + step).astype(f32)
            err = (step - (a_new - alpha)).astype(f32)                   # FastTwoSum residual (|alpha| >= |step|): alpha + step = a_new + err exactly
            ulp = (np.abs(a_new).view(np.uint32) & np.uint32(0x7f800000)).view(f32) * f32(2.0 ** -23)
            pow2 = (np.abs(a_new).view(np.uint32) & np.uint32(0x007fffff)) == 0
            exact = (eps < f32(0.5) * ulp - np.abs(err)) & ~pow2 & (np.abs(step) <= np.abs(alpha))
            e_new = np.where(exact, f32(0), eps + ulp).astype(f32)        # two roundings of at most ulp/2 each
            e = np.where(cont, e_new, e).astype(f32)
            alpha = np.where(cont, a_new, alpha).astype(f32)
            ne_prev = np.where(cont, ne, ne_prev); D_prev = np.where(cont, Dk, D_prev)
            active = cont
        # sampling at the final alpha: prefix sums in double
        d = (alpha[:, None] - q).astype(f32)
        ra = (f32(1) / alpha).astype(f32)
        k = (lam * ra).astype(f32)
        p = np.where(child, (tc * (f32(1) / d).astype(f32)).astype(f32).astype(f64), k.astype(f64)[:, None] * pi.astype(f64))
        cum = np.cumsum(p, axis=1)
        delta = (ES + G * e * f32(1.05)).astype(f32)
        below = cum < r[:, None]
        l = below.sum(1)
        gap = np.abs(cum - r[:, None]).min(1)
        action = np.where(l < A, l, last_nz)
        action = np.where(r <= 0, first_nz, action)
        unsure = (gap <= delta) & (r > 0)
        flag = np.where((flag == 0) & unsure, 2, flag)
    return action, iters, flag, dict(alpha=tr_alpha, ne=tr_ne, D=tr_D, e=tr_e, cum=cum.astype(f32), delta=delta, alpha_final=alpha)


def replay_descend(logits, q, n, c_puct, seats, terminal, children, rands, stats):
    """Replays mctscuda.descend level by level for all envs with both evaluations; updates `stats`; returns the exact
    path's (parents, actions) for checking against the oracle."""
    B, T, A = logits.shape
    exp_lut = _exp_lut()
    cur = np.zeros(B, np.int64)
    parents = np.zeros(B, np.int16)
    actions = np.full(B, -1, np.int16)
    live = ~terminal[np.arange(B), 0]
    while live.any():
        idx = np.nonzero(live)[0]
        t = cur[idx]
        row_logits = logits[idx, t]                                          # (n,A) half
        pi = exp_lut[row_logits.view(np.uint16)]
        ch = children[idx, t].astype(np.int64)                               # (n,A)
        child = ch >= 0
        seat = seats[idx, t].astype(np.int64)
        chq = q[idx[:, None], np.maximum(ch, 0), seat[:, None]].astype(f32)
        qa = np.where(child, chq, f32(0)).astype(f32)
        N = np.where(child, n[idx[:, None], np.maximum(ch, 0)].astype(np.int64), 1).sum(1)
        lam = (c_puct[idx].astype(f32) * N.astype(f32)).astype(f32) / (N + A).astype(f32)
        r = rands[idx, t].astype(f32)
        act, iters, tra, trn, cum, alpha = exact_eval(pi, qa, lam, r)
        fact, fiters, flag, det = fast_eval(pi, qa, child, lam, r)
        # ---- statistics
        stats.evals += len(idx)
        stats.iters += int(iters.sum())
        stats.children += int(child.sum())
        stats.flag_stop += int((flag == 1).sum()); stats.flag_sample += int((flag == 2).sum())
        stats.flag_guard += int((flag == 3).sum()); stats.flag_tiny += int((flag == 4).sum())
        for i in np.nonzero(flag != 0)[0][:stats.resolve_budget]:
            a1, it1, xp = eval_one(pi[i], qa[i], child[i], lam[i], r[i])
            stats.resolved += 1
            stats.resolve_passes += xp
            stats.resolve_hist[min(xp, 9)] += 1
            stats.bad_resolved += int(a1 != act[i]) + int(it1 != iters[i])
        for i in np.nonzero(flag == 0)[0][:2]:                               # and a couple of unflagged ones: same answer, no exact pass
            a1, it1, xp = eval_one(pi[i], qa[i], child[i], lam[i], r[i])
            stats.bad_resolved += int(a1 != act[i]) + int(it1 != iters[i])
        ok = flag == 0
        stats.bad_stop += int((ok & (fiters != iters)).sum())
        stats.bad_action += int((ok & (fact != act)).sum())
        stats.delta_sum += float(det['delta'][ok].sum())
        with np.errstate(all='ignore'):
            same = ok & (fiters == iters)
            rat = np.abs(det['ne'] - trn) / det['D']
            rat = np.where(np.isfinite(rat), rat, 0)[same]
            if rat.size:
                stats.max_ne_ratio = max(stats.max_ne_ratio, float(rat.max()))
            ra = np.abs(det['alpha'] - tra) / det['e']
            ra = np.where(np.isfinite(ra), ra, 0)[same]
            if ra.size:
                stats.max_alpha_ratio = max(stats.max_alpha_ratio, float(ra.max()))
            rc = (np.abs(det['cum'] - cum).max(1) / det['delta'])[same]
            if rc.size:
                stats.max_cum_ratio = max(stats.max_cum_ratio, float(rc.max()))
        # ---- advance along the exact path
        parents[idx] = t
        actions[idx] = act
        nxt = np.where(act >= 0, ch[np.arange(len(idx)), np.maximum(act, 0)], -1)
        cur[idx] = nxt
        live[idx] = (nxt >= 0) & (act >= 0)
        alive = np.nonzero(live)[0]
        live[alive] = ~terminal[alive, cur[alive]]
    return parents, actions


_lut = None


def _exp_lut():
    global _lut
    if _lut is None:
        import oracle
        _lut = oracle.exp_table()
    return _lut


# ------------------------------------------------------------------------------------------------------------------------
# the kernel's complete control flow for ONE evaluation (scalar): fast Newton with a safe point, exact passes on demand
# ------------------------------------------------------------------------------------------------------------------------
def _exact_pass(top, q, alpha):
    """One pass of the reference's loops at alpha: (S, g, prefix sums, terms) in fp32, sequential order."""
    A = len(top)
    S = f32(0); g = f32(0)
    cum = np.zeros(A, f32); terms = np.zeros(A, f32)
    with np.errstate(all='ignore'):
        for a in range(A):
            bot = f32(alpha - q[a])
            p = f32(top[a] / bot)
            S = f32(S + p)
            g = f32(g + f32(f32(-top[a]) / f32(bot * bot)))
            cum[a] = S; terms[a] = p
    return S, g, cum, terms


def _exact_sample(terms, cum, r):
    action, valid = -1, -1
    for a in range(len(terms)):
        if terms[a] > 0:
            if cum[a] >= r:
                return a
            valid = a
    return valid


def _all_exact(top, q, alpha0, r):
    """The reference's loops verbatim (newton_search + sampling), one exact pass per Newton pass."""
    alpha, it, ne_prev, xp = alpha0, 0, f32(np.inf), 0
    with np.errstate(all='ignore'):
        while True:
            xS, xg, cum, terms = _exact_pass(top, q, alpha)
            xp += 1; it += 1
            ne = f32(xS - f32(1))
            if it > 100 or ne < f32(1e-3) or ne_prev == ne:
                return _exact_sample(terms, cum, r), min(it, 100), xp
            alpha = f32(alpha - f32(ne / xg)); ne_prev = ne


def eval_one(pi, q, child, lam, r):
    """Returns (action, newton passes as the reference counts them, exact passes used).  pi, q (A,) f32, child (A,) bool."""
    A = len(pi)
    top = (lam * pi).astype(f32)
    nz = top > 0
    if not nz.any():
        return -1, 0, 0
    first_nz, last_nz = int(nz.argmax()), int(A - 1 - nz[::-1].argmax())
    alpha0 = f32(np.max(q + np.maximum(top, f32(1e-4))))
    wgt = np.maximum(last_nz + 1 - np.arange(A), 0).astype(np.float64)
    pi_cl = np.where(child, 0, pi).astype(np.float64)
    # (the kernel: M = lambda*P_all - sum t_c in double; same quantity up to u)
    M = f32(lam * pi_cl.sum()); MW = f32(lam * (pi_cl * wgt).sum())
    cidx = np.nonzero(child)[0]
    nc = len(cidx)
    cours = f32(nc + 13) * U
    qmax = f32(q[cidx].max()) if nc else f32(0)
    if np.where(nz, top, np.inf).min() < TINY:
        return _all_exact(top, q, alpha0, r)
    xpasses = 0
    alpha, e, it = alpha0, f32(0), 0
    ne_prev, D_prev = f32(np.inf), f32(0)
    safe = (alpha, it, ne_prev, D_prev)
    with np.errstate(all='ignore'):
        while True:
            # ---- fast Newton until a decision or a doubt
            stopped = False
            while True:
                if e == 0:
                    safe = (alpha, it, ne_prev, D_prev)
                ra = f32(f32(1) / alpha)
                S = f32(M * ra); G = f32(S * ra); H = f32(G * ra)
                ES = f32(MW * ra); EG = f32(ES * ra)
                for a in cidx:
                    rc = f32(f32(1) / f32(alpha - q[a]))
                    s = f32(top[a] * rc); g = f32(s * rc)
                    S = f32(S + s); G = f32(G + g); H = f32(H + f32(g * rc))
                    ES = f32(ES + f32(wgt[a]) * s); EG = f32(EG + f32(wgt[a]) * g)
                H = f32(2) * H
                ESn = f32(U * f32(1.01) * ES + (f32(2) * U + cours) * S)
                EGn = f32(U * f32(1.01) * EG + (f32(4) * U + cours) * G)
                ne = f32(S - f32(1))
                Dk = f32(ESn + G * e * f32(1.05) + f32(2) * U * abs(ne))
                it += 1
                guard_ok = (e <= GUARD * (alpha - qmax)) and np.isfinite(S) and np.isfinite(G) and G > 0 and it <= MAX_FAST_ITERS
                if not guard_ok:
                    a1, i1, xp = _all_exact(top, q, alpha0, r)
                    return a1, i1, xp + xpasses
                if ne < f32(1e-3) - Dk:
                    stopped = True; break
                if not ((ne > f32(1e-3) + Dk) and (abs(ne - ne_prev) > Dk + D_prev)):
                    break
                rG = f32(1) / G
                L = f32(1.2) * (max(ne, f32(0)) + f32(2) * G * e) * H * rG * rG
                R = f32(1.05) * (Dk * rG + abs(ne) * rG * (EGn * rG + f32(4) * U))
                eps = f32(L * e + R)
                step = f32(ne / G)
                a_new = f32(alpha + step)
                err = f32(step - f32(a_new - alpha))
                ab = np.float32(abs(a_new)).view(np.uint32)
                ulp = (ab & np.uint32(0x7f800000)).view(f32) * f32(2.0 ** -23)
                exact = (eps < f32(0.5) * ulp - abs(err)) and (ab & np.uint32(0x007fffff)) != 0 and abs(step) <= abs(alpha)
                e = f32(0) if exact else f32(eps + ulp)
                alpha = a_new; ne_prev = ne; D_prev = Dk
            if stopped:
                # ---- certified sampling
                if r <= 0:
                    return first_nz, it, xpasses
                ra = f32(f32(1) / alpha); k = f32(lam * ra)
                p = np.where(child, 0, k.astype(np.float64) * pi.astype(np.float64))
                for a in cidx:
                    p[a] = f32(top[a] * f32(f32(1) / f32(alpha - q[a])))
                cum = np.cumsum(p)
                delta = f32(ESn + G * e * f32(1.05))
                lo_n = int((cum < float(r) - float(delta)).sum()); hi_n = int((cum <= float(r) + float(delta)).sum())
                if lo_n == hi_n:
                    return (lo_n if lo_n < A else last_nz), it, xpasses
            # ---- a doubt: one exact pass at the safe point (the last iterate known to be the reference's float)
            alpha, it, ne_prev, D_prev = safe
            xS, xg, cum, terms = _exact_pass(top, q, alpha)
            xpasses += 1
            it += 1
            ne = f32(xS - f32(1))
            if ne < f32(1e-3):
                return _exact_sample(terms, cum, r), it, xpasses
            # `error == new_error`: error is the reference's previous S-1, known exactly when D_prev == 0 and to D_prev otherwise
            if (D_prev == 0 and ne_prev == ne):
                return _exact_sample(terms, cum, r), it, xpasses
            if D_prev != 0 and not (abs(ne - ne_prev) > D_prev):
                a1, i1, xp = _all_exact(top, q, alpha0, r)
                return a1, i1, xp + xpasses
            alpha = f32(alpha - f32(ne / xg))
            e = f32(0); ne_prev = ne; D_prev = f32(0)
