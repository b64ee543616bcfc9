"""Pure-Python models of the device permutation kernels (ls_spa_b200/csrc/perms.cu).

They restate, line for line, the integer logic of the CUDA kernels so that the
algorithms can be checked against numpy/scipy on a machine without a GPU.  They are
test infrastructure, not a product path.
"""

import numpy as np

MASK64 = (1 << 64) - 1
MASK128 = (1 << 128) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


def pcg_output(state):
    hi, lo = state >> 64, state & MASK64
    rot = hi >> 58
    x = hi ^ lo
    return ((x >> rot) | (x << ((64 - rot) & 63))) & MASK64


def pcg_advance(state, inc, delta):
    acc_mult, acc_plus, cur_mult, cur_plus = 1, 0, PCG_MULT, inc
    while delta > 0:
        if delta & 1:
            acc_mult = (acc_mult * cur_mult) & MASK128
            acc_plus = (acc_plus * cur_mult + cur_plus) & MASK128
        cur_plus = ((cur_mult + 1) * cur_plus) & MASK128
        cur_mult = (cur_mult * cur_mult) & MASK128
        delta >>= 1
    return (acc_mult * state + acc_plus) & MASK128


def raw_budget(p, count):
    e = sum(((0xFFFFFFFF >> (32 - i.bit_length())) + 1.0) / (i + 1.0) for i in range(1, p))
    n = int(e * count * 1.15 + 65536.0)
    return (n + 1) & ~1


def pcg64_raw(gen_state, nout):
    """pcg64_raw_kernel: raw 32-bit draw stream (buffered uinteger first, then lo/hi halves)."""
    s0 = (gen_state[0] << 64) | gen_state[1]
    inc = (gen_state[2] << 64) | gen_state[3]
    raw = []
    if gen_state[4]:
        raw.append(gen_state[5] & 0xFFFFFFFF)
    s = s0
    for _ in range(nout):
        s = (s * PCG_MULT + inc) & MASK128
        v = pcg_output(s)
        raw.append(v & 0xFFFFFFFF)
        raw.append(v >> 32)
    return raw


def pcg64_scan(p, count, raw, ndraws, gen_state):
    """pcg64_scan_kernel: 32 lanes, speculative ballot iteration. Returns (accepted, flag)."""
    steps = p - 1
    total = count * steps
    accepted = [None] * total
    t = pos = consumed = 0
    tmod = 0
    max_iters = 0
    while t < total and pos < ndraws:
        d = [raw[pos + l] if pos + l < ndraws else 0 for l in range(32)]
        inr = [pos + l < ndraws for l in range(32)]
        accmask = 0xFFFFFFFF
        for it in range(33):
            acc, tl, val = [False] * 32, [0] * 32, [0] * 32
            for l in range(32):
                prior = bin(accmask & ((1 << l) - 1)).count("1")
                tl[l] = t + prior
                m = (tmod + prior) % steps
                i = steps - m
                msk = 0xFFFFFFFF >> (32 - i.bit_length())
                val[l] = d[l] & msk
                acc[l] = inr[l] and tl[l] < total and val[l] <= i
            nm = sum(1 << l for l in range(32) if acc[l])
            if nm == accmask:
                break
            accmask = nm
        max_iters = max(max_iters, it + 1)
        for l in range(32):
            if acc[l]:
                accepted[tl[l]] = val[l]
        nacc = bin(accmask).count("1")
        if t + nacc >= total:
            last = max(l for l in range(32) if acc[l] and tl[l] + 1 == total)
            consumed = pos + last + 1
        else:
            consumed = pos + min(32, ndraws - pos)
        t += nacc
        tmod = (tmod + nacc) % steps
        pos += 32
    flag = int(t < total)
    s0 = (gen_state[0] << 64) | gen_state[1]
    inc = (gen_state[2] << 64) | gen_state[3]
    has = bool(gen_state[4])
    from_outputs = consumed - (1 if (has and consumed > 0) else 0)
    if consumed > 0:
        nout = (from_outputs + 1) // 2
        s = pcg_advance(s0, inc, nout)
        gen_state[0], gen_state[1] = s >> 64, s & MASK64
        if from_outputs & 1:
            gen_state[4], gen_state[5] = 1, pcg_output(s) >> 32
        else:
            gen_state[4], gen_state[5] = 0, 0
    return accepted, flag, max_iters


def _fy_accept(d, steps, m):
    i = steps - m
    val = d & (0xFFFFFFFF >> (32 - i.bit_length()))
    return val <= i, val


def pcg64_scan_fsm(p, count, raw, ndraws, gen_state, chunk=256, group=64):
    """pcg64_fsm / group / prefix / emit / finish kernels: the acceptance automaton evaluated by
    chunks from every start state, chunk maps composed per group, short serial walk over the
    groups, replay of every chunk from its true start state.  Returns (accepted, flag)."""
    steps = p - 1
    total = count * steps
    nchunks = -(-ndraws // chunk)
    ngroups = -(-nchunks // group)
    acc_tab = np.zeros((nchunks, steps), dtype=np.int64)
    for c in range(nchunks):                       # pcg64_fsm_kernel: block = chunk, thread = start state
        n0 = c * chunk
        nvalid = min(chunk, ndraws - n0)
        for m0 in range(steps):
            m = m0
            a = 0
            for n in range(nvalid):
                ok, _ = _fy_accept(raw[n0 + n], steps, m)
                if ok:
                    a += 1
                    m = 0 if m + 1 == steps else m + 1
            acc_tab[c, m0] = a
    grp_tab = np.zeros((ngroups, steps), dtype=np.int64)
    for g in range(ngroups):                       # pcg64_group_kernel
        for m0 in range(steps):
            m, t = m0, 0
            for c in range(g * group, min((g + 1) * group, nchunks)):
                a = int(acc_tab[c, m])
                t += a
                m = (m + a) % steps
            grp_tab[g, m0] = t
    Tg = [0] * (ngroups + 1)                       # pcg64_prefix_kernel
    t = 0
    for g in range(ngroups):
        Tg[g] = t
        t += int(grp_tab[g, t % steps])
    Tg[ngroups] = t
    consumed = -1
    accepted = [None] * total
    for g in range(ngroups):                       # pcg64_emit_kernel: block = group, thread = chunk
        tc = []
        t = Tg[g]
        for k in range(group):
            tc.append(t)
            if g * group + k < nchunks:
                t += int(acc_tab[g * group + k, t % steps])
        for k in range(group):
            c = g * group + k
            if c >= nchunks:
                continue
            t = tc[k]
            if t >= total:
                continue
            m = t % steps
            n0 = c * chunk
            for n in range(min(chunk, ndraws - n0)):
                ok, val = _fy_accept(raw[n0 + n], steps, m)
                if ok:
                    accepted[t] = val
                    t += 1
                    m = 0 if m + 1 == steps else m + 1
                    if t == total:
                        consumed = n0 + n + 1
                        break
    flag = 0                                       # pcg64_finish_kernel
    if consumed < 0:
        flag, consumed = 1, ndraws
    s0 = (gen_state[0] << 64) | gen_state[1]
    inc = (gen_state[2] << 64) | gen_state[3]
    has = bool(gen_state[4])
    from_outputs = consumed - (1 if (has and consumed > 0) else 0)
    if consumed > 0:
        nout = (from_outputs + 1) // 2
        s = pcg_advance(s0, inc, nout)
        gen_state[0], gen_state[1] = s >> 64, s & MASK64
        if from_outputs & 1:
            gen_state[4], gen_state[5] = 1, pcg_output(s) >> 32
        else:
            gen_state[4], gen_state[5] = 0, 0
    return accepted, flag


def pcg64_shuffle(p, count, accepted):
    out = np.empty((count, p), dtype=np.int64)
    for n in range(count):
        a = list(range(p))
        for s, i in enumerate(range(p - 1, 0, -1)):
            j = accepted[n * (p - 1) + s]
            a[i], a[j] = a[j], a[i]
        out[n] = a
    return out


def pcg64_perms(p, count, gen_state, fsm=False, budget=None):
    """Whole lsspa_perms_pcg64 call; gen_state (list of 6 ints) is advanced in place.  fsm selects
    the parallel automaton kernels (the default of the library), else the single-warp walk."""
    if p > 1:
        budget = raw_budget(p, count) if budget is None else budget
        raw = pcg64_raw(gen_state, budget // 2)
        if fsm:
            accepted, flag = pcg64_scan_fsm(p, count, raw, budget, gen_state)
            iters = 0
        else:
            accepted, flag, iters = pcg64_scan(p, count, raw, budget, gen_state)
        assert flag == 0
    else:
        accepted, iters = [], 0
    return pcg64_shuffle(p, count, accepted), iters


def state_from_numpy(seed):
    st = np.random.default_rng(seed).bit_generator.state
    s, inc = st["state"]["state"], st["state"]["inc"]
    return [s >> 64, s & MASK64, inc >> 64, inc & MASK64, st["has_uint32"], st["uinteger"]]


def exact_perm(p, rank):
    """perms_exact_kernel"""
    m = min(p, 20)
    base = p - m
    pool = list(range(m))
    dig = [0] * m
    r = rank
    for i in range(1, m + 1):
        dig[m - i] = r % i
        r //= i
    row = list(range(base))
    for pos in range(m):
        row.append(base + pool.pop(dig[pos]))
    return row


def sobol_ints(sv, shift, bits, k):
    gray = k ^ (k >> 1)
    x = shift.copy()
    b = 0
    while gray and b < bits:
        if gray & 1:
            x ^= sv[:, b]
        gray >>= 1
        b += 1
    return x


def rank_scatter(keys):
    p = len(keys)
    row = [0] * p
    for j in range(p):
        r = sum(1 for l in range(p) if keys[l] < keys[j] or (keys[l] == keys[j] and l < j))
        row[r] = j
    return row


def sobol_argsort_perm(sv, shift, bits, k):
    return np.argsort(sobol_ints(sv, shift, bits, k), kind="stable")


def permutohedron_perm(p, sv, shift, bits, k):
    """permutohedron_kernel (suffix-sum form of the projection on U)."""
    dim = p - 1
    x = sobol_ints(sv, shift, bits, k).astype(np.float64) * 2.0 ** -bits
    npairs = (dim + 1) // 2
    w = np.zeros(p)
    for t in range(npairs):
        rad = np.sqrt(-2.0 * np.log(x[2 * t]))
        th = (2.0 * 3.141592653589793) * x[2 * t + 1]
        r0, r1 = 2 * t, 2 * t + 1
        w[r0] = rad * np.cos(th) / np.sqrt(float(r0 + 1) * float(r0 + 2))
        if r1 < dim:
            w[r1] = rad * np.sin(th) / np.sqrt(float(r1 + 1) * float(r1 + 2))
    proj = np.zeros(p)
    suffix = 0.0
    for j in range(dim - 1, -1, -1):
        suffix += w[j]
        proj[j] = suffix
    for j in range(1, p):
        proj[j] -= j * w[j - 1]
    return np.argsort(proj, kind="stable"), proj
