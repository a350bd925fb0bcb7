#!/usr/bin/env python
"""CPU model of the shared-memory wavefronts of the per-nonzero tile kernel (csrc/recompute_tiles.cu).

Rebuilds the kernel's plan (tiles of consecutive column nodes, per-tile element slots, pairs sorted into 32-lane tasks)
in numpy from an element->dof table and counts, for every task step, the wavefronts of the nine 64-bit geometry loads and
the nine 64-bit reference-tensor loads under the bank model of B300_MICROARCH.md (32 banks x 4 B; a 64-bit warp load is
served in two half-warp passes; inside a pass, lanes reading the same address are merged and the pass costs the largest
number of distinct addresses that fall in one bank pair).  Used to rank plan variants (pair order inside a task, slot
numbering, row strides) WITHOUT spending GPU time; the absolute numbers are checked against ncu
(profiles/round1_ncu_tiles_v2_c3_n110.txt: 4.25 wavefronts per geometry load, 2.0 per reference-tensor load).

    python tools/tile_bank_sim.py /path/to/dump_dir [variant ...]
"""
import sys

import numpy as np

ND, NB, GSZ, INREC = 10, 100, 9, 10
CAP_INC, CAP_PAIRS = 510, 1280


def build(edof):
    ne = edof.shape[0]
    node = edof // 3
    J = np.repeat(node[:, :, None], ND, 2).reshape(-1)  # (e, j, i): j slow
    I = np.repeat(node[:, None, :], ND, 1).reshape(-1)
    cid = np.arange(ne * NB)
    order = np.lexsort((cid, I, J))
    Js, Is = J[order], I[order]
    newp = np.r_[True, (Js[1:] != Js[:-1]) | (Is[1:] != Is[:-1])]
    cstart = np.r_[np.nonzero(newp)[0], len(order)]
    pJ = Js[newp]
    csrc = order
    cols, colstart = np.unique(pJ, return_index=True)
    colstart = np.r_[colstart, len(pJ)]
    inc = np.bincount(node.reshape(-1), minlength=cols.max() + 1)[cols]
    rstart = np.r_[0, np.cumsum(inc)]
    # incidences per node: elements
    nflat = node.reshape(-1)
    io = np.argsort(nflat, kind="stable")
    inc_el = (io // ND)
    inc_start = np.r_[0, np.cumsum(np.bincount(nflat))]
    return dict(ne=ne, cstart=cstart, csrc=csrc, pJ=pJ, cols=cols, colstart=colstart, rstart=rstart, inc_el=inc_el,
                inc_start=inc_start, node=node)


def tiles(st):
    out = []
    k, n = 0, len(st["cols"])
    rs, cs = st["rstart"], st["colstart"]
    while k < n:
        k1 = k + 1
        while k1 < n and rs[k1 + 1] - rs[k] <= CAP_INC and cs[k1 + 1] - cs[k] <= CAP_PAIRS:
            k1 += 1
        out.append((k, k1))
        k = k1
    return out


def sig_of(seq):
    s = 0
    for rr in seq:
        s = ((s ^ int(rr)) * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF
        s ^= s >> 15
    return s & 0xFFFFFF


def half_cost(addr):
    """addr: int array [16] of 8-byte word addresses (or -1 = inactive); wavefronts of one half-warp pass."""
    a = np.unique(addr[addr >= 0])
    if a.size == 0:
        return 0
    return int(np.bincount(a % 16).max())


def load_cost(rows, stride, k_count=9):
    """rows [32]: row index per lane; the lane reads words rows*stride + k, k < k_count.  Sum over k of both halves."""
    tot = 0
    for k in range(k_count):
        addr = rows * stride + k
        tot += half_cost(addr[:16]) + half_cost(addr[16:])
    return tot


def simulate(st, variant, max_tiles=60, seed=0):
    rng = np.random.default_rng(seed)
    tl = tiles(st)
    # interior tiles only (the boundary of a small mesh is not representative)
    mid = tl[len(tl) // 4: len(tl) // 4 + max_tiles]
    g_w = m_w = steps_tot = 0
    s_w = s_n = 0
    distinct_rows = []
    for (k0, k1) in mid:
        p0, p1 = st["colstart"][k0], st["colstart"][k1]
        # distinct elements of the tile, sorted -> slot
        els = np.unique(np.concatenate([st["inc_el"][st["inc_start"][c]:st["inc_start"][c + 1]] for c in st["cols"][k0:k1]]))
        slot_of = {int(e): s for s, e in enumerate(els)}
        if variant.get("slots") == "random":
            perm = rng.permutation(len(els))
            slot_of = {int(e): int(perm[s]) for s, e in enumerate(els)}
        pairs = []
        # image offsets (all entries kept): column node J holds npJ pairs -> its three columns are 3*npJ entries each
        off0 = {}
        run = 0
        for c in range(k0, k1):
            a, b = st["colstart"][c], st["colstart"][c + 1]
            for r, p in enumerate(range(a, b)):
                off0[p] = (run + 3 * r, 3 * (b - a))  # offset of the component-0 piece, distance between components
            run += 9 * (b - a)
        for p in range(p0, p1):
            c = st["csrc"][st["cstart"][p]:st["cstart"][p + 1]]
            el, rr = c // NB, c % NB
            pairs.append((len(c), el, rr, p))
        normal = [q for q in pairs if q[0] <= INREC]
        key = variant.get("sort", "sig")
        if key == "sig":
            normal.sort(key=lambda q: (-q[0], sig_of(q[2]), q[3]))
        elif key == "csc":
            normal.sort(key=lambda q: (-q[0], q[3]))
        elif key == "elem":
            normal.sort(key=lambda q: (-q[0], int(q[1][0]), q[3]))
        elif key == "elem_sig":
            normal.sort(key=lambda q: (-q[0], int(q[1][0]), sig_of(q[2]), q[3]))
        elif key == "sig_elem":
            normal.sort(key=lambda q: (-q[0], sig_of(q[2]), int(q[1][0]), q[3]))
        if variant.get("slots") == "color":
            slot_of = color_slots(normal, els)
        zslot = len(els)
        gstride = variant.get("gstride", 9)
        mstride = variant.get("mstride", 9)
        mrow = variant.get("mrow", lambda rr: rr)
        for t0 in range(0, len(normal), 32):
            task = normal[t0:t0 + 32]
            steps = task[0][0]
            for b in range(3):  # flush: nine 64-bit stores per lane
                for r in range(3):
                    addr = np.full(32, -1, np.int64)
                    for l, q in enumerate(task):
                        addr[l] = off0[q[3]][0] + b * off0[q[3]][1] + r
                    s_w += half_cost(addr[:16]) + half_cost(addr[16:])
                    s_n += 1
            for c in range(steps):
                rows = np.full(32, -1, np.int64)
                mr = np.full(32, -1, np.int64)
                for l, q in enumerate(task):
                    if c < q[0]:
                        rows[l] = slot_of[int(q[1][c])]
                        mr[l] = mrow(int(q[2][c]))
                    else:
                        rows[l] = zslot
                        mr[l] = mrow(0)
                # lanes beyond the task's pairs still execute with the zero descriptor
                rows[rows < 0] = zslot
                mr[mr < 0] = mrow(0)
                g_w += load_cost(rows, gstride)
                m_w += load_cost(mr, mstride)
                distinct_rows.append(len(np.unique(rows[:16])) + len(np.unique(rows[16:])))
                steps_tot += 1
    return g_w / (9 * steps_tot), m_w / (9 * steps_tot), steps_tot, float(np.mean(distinct_rows)) / 2, s_w / s_n, s_w, g_w + m_w


def color_slots(normal, els):
    """Greedy slot numbering: elements are visited in order of first use; each takes the free slot whose bank class
    (slot mod 16) collides least with the elements it is read together with (same half-warp, same step)."""
    co = {}  # element -> list of half-warp groups (sets of elements) it appears in
    groups = []
    for t0 in range(0, len(normal), 32):
        task = normal[t0:t0 + 32]
        for c in range(task[0][0]):
            for h in (0, 16):
                g = set(int(q[1][c]) for q in task[h:h + 16] if c < q[0])
                if len(g) > 1:
                    groups.append(g)
    for gi, g in enumerate(groups):
        for e in g:
            co.setdefault(e, []).append(gi)
    n = len(els)
    cls_of, slot_of = {}, {}
    free = {c: list(range(c, n, 16)) for c in range(16)}
    order = sorted((int(e) for e in els), key=lambda e: -len(co.get(e, [])))
    for e in order:
        cost = np.zeros(16)
        for gi in co.get(e, []):
            for e2 in groups[gi]:
                if e2 != e and e2 in cls_of:
                    cost[cls_of[e2]] += 1
        for c in range(16):
            if not free[c]:
                cost[c] = 1e9
        c = int(np.argmin(cost))
        cls_of[e] = c
        slot_of[e] = free[c].pop(0)
    return slot_of


def simulate_groups(st, KG=3, max_tiles=60, sort="sig", same_column=True):
    """Variant F: a lane owns up to KG pairs of ONE column node that have the same element list; the geometry row of a
    step is loaded once and reused for the KG pairs.  Returns wavefronts per pair-contribution (G, M), stores per pair."""
    tl = tiles(st)
    mid = tl[len(tl) // 4: len(tl) // 4 + max_tiles]
    g_w = m_w = contribs = lanesteps = 0
    s_w = npairs = nitems = 0
    for (k0, k1) in mid:
        p0, p1 = st["colstart"][k0], st["colstart"][k1]
        els = np.unique(np.concatenate([st["inc_el"][st["inc_start"][c]:st["inc_start"][c + 1]] for c in st["cols"][k0:k1]]))
        slot_of = {int(e): s for s, e in enumerate(els)}
        zslot = len(els)
        off0, run = {}, 0
        for c in range(k0, k1):
            a, b = st["colstart"][c], st["colstart"][c + 1]
            for r, p in enumerate(range(a, b)):
                off0[p] = (run + 3 * r, 3 * (b - a))
            run += 9 * (b - a)
        groups = {}
        for p in range(p0, p1):
            c = st["csrc"][st["cstart"][p]:st["cstart"][p + 1]]
            if len(c) > INREC:
                continue
            el, rr = c // NB, c % NB
            groups.setdefault((int(st["pJ"][p]) if same_column else -1, tuple(int(e) for e in el)), []).append((rr, p))
        items = []
        for (J, el), lst in groups.items():
            for a in range(0, len(lst), KG):
                ch = lst[a:a + KG]
                items.append((len(el), el, [q[0] for q in ch], [q[1] for q in ch]))
        if sort == "sig":
            items.sort(key=lambda it: (-it[0], -len(it[2]), sig_of(np.concatenate(it[2])), it[3][0]))
        else:
            items.sort(key=lambda it: (-it[0], -len(it[2]), it[1][0], it[3][0]))
        nitems += len(items)
        for t0 in range(0, len(items), 32):
            task = items[t0:t0 + 32]
            steps = task[0][0]
            kmax = max(len(it[2]) for it in task)
            for g in range(kmax):
                for b in range(3):
                    for r in range(3):
                        addr = np.full(32, -1, np.int64)
                        for l, it in enumerate(task):
                            if g < len(it[3]):
                                addr[l] = off0[it[3][g]][0] + b * off0[it[3][g]][1] + r
                        s_w += half_cost(addr[:16]) + half_cost(addr[16:])
            for c in range(steps):
                rows = np.full(32, zslot, np.int64)
                for l, it in enumerate(task):
                    if c < it[0]:
                        rows[l] = slot_of[it[1][c]]
                g_w += load_cost(rows, 9)
                for g in range(kmax):
                    mr = np.full(32, -1, np.int64)
                    for l, it in enumerate(task):
                        if g < len(it[2]):
                            mr[l] = int(it[2][g][c]) if c < it[0] else 0
                    m_w += load_cost(mr, 9)
                    contribs += sum(1 for it in task if g < len(it[2]) and c < it[0])
                    lanesteps += 32
        npairs += sum(len(it[2]) for it in items)
    return dict(G_per_contrib=g_w / contribs * 32, M_per_contrib=m_w / contribs * 32, store_per_pair=s_w / npairs * 32,
                pairs_per_item=npairs / nitems, lane_util=contribs / lanesteps, total=g_w + m_w + s_w, contribs=contribs)


VARIANTS = {
    "current(sig)": {"sort": "sig"},
    "csc-order": {"sort": "csc"},
    "elem-order": {"sort": "elem"},
    "elem+sig": {"sort": "elem_sig"},
    "sig+elem": {"sort": "sig_elem"},
    "sig,random-slots": {"sort": "sig", "slots": "random"},
    "sig,colored-slots": {"sort": "sig", "slots": "color"},
    "sig+elem,colored": {"sort": "sig_elem", "slots": "color"},
    "elem,colored": {"sort": "elem", "slots": "color"},
    "elem,M16rows": {"sort": "elem", "mrow": lambda rr: (rr // 10) * 16 + rr % 10},
    "elem,colored,M16rows": {"sort": "elem", "slots": "color", "mrow": lambda rr: (rr // 10) * 16 + rr % 10},
}


def main():
    d = sys.argv[1]
    edof = np.load(d + "/elem_dof.npy")
    st = build(edof)
    if len(sys.argv) > 2 and sys.argv[2] == "groups":
        for sc in (True, False):
            for KG in (1, 2, 3, 4, 6):
                r = simulate_groups(st, KG, sort="sig", same_column=sc)
                print("same_column", sc, "KG", KG, {k: round(v, 2) for k, v in r.items()})
        return
    names = sys.argv[2:] or list(VARIANTS)
    print("%-26s %8s %8s %8s %8s %8s %10s" % ("variant", "G wf/ld", "M wf/ld", "sum/step", "rows/half", "wf/store", "total wf"))
    for nm in names:
        g, m, steps, dr, sw, stot, ltot = simulate(st, VARIANTS[nm])
        print("%-26s %8.3f %8.3f %8.2f %8.2f %8.2f %10d   (%d steps)" % (nm, g, m, 9 * (g + m), dr, sw, stot + ltot, steps))


if __name__ == "__main__":
    main()
