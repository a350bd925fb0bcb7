"""CPU, world_size 2 and 3 under gloo: the halo protocol of getfem_b200/halo.py (ownership bounds, pair
announcements, merge order, slice exchange) with an ORACLE-backed term standing in for the device term.
The oracle (plain C restatement, oracle/asm_oracle.c) assembles each rank's element block; after the exchange
every rank's owned slab must equal the oracle's assembly of the whole mesh (pattern identical, values 1e-13)."""
import os
import sys
import tempfile

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleTerm:
    """Same halo_* surface as capi.DeviceTerm, numpy + oracle inside (test infrastructure)."""

    def __init__(self, g, e0, e1):
        self.g, self.e0, self.e1 = g, e0, e1
        self.Q = g["Q"]
        self.ndof = int(g["meta"]["ndof"])
        self.sources = []
        self.own = (0, self.ndof)

    def _assemble_local(self):
        from oracle import oracle
        g = self.g
        return oracle.assemble(g["pts"], g["conn"][self.e0:self.e1], g["elem_dof"][self.e0:self.e1], self.ndof, self.Q,
                               g["quad_w"], g["gt_grad"], g["phi"], g["gphi"], bool(g["gt_linear"]), g["family"],
                               g["fparams"], g["U"])

    def _pairs(self, jc, ir, lo, hi):
        """{(J, I): mask} of the columns [lo, hi) of a CSC matrix."""
        Q, out = self.Q, {}
        for c in range(lo, hi):
            J, b = c - c % Q, c % Q
            for r in ir[jc[c]:jc[c + 1]]:
                I, a = int(r) - int(r) % Q, int(r) % Q
                out[(J, I)] = out.get((J, I), 0) | (1 << (b * Q + a))
        return out

    def halo_begin(self, U=None):
        self.ljc, self.lir, self.lpr, self.lR = self._assemble_local()
        ed = self.g["elem_dof"][self.e0:self.e1]
        return int(ed.min()), int(ed.max()) + self.Q

    def halo_ghost_pairs(self, a, b):
        pr = self._pairs(self.ljc, self.lir, a, b)
        keys = sorted(pr)
        return (np.array([k[0] for k in keys], np.int32), np.array([k[1] for k in keys], np.int32),
                np.array([pr[k] for k in keys], np.uint16))

    def halo_add_source(self, src, J, I, mask, r_lo, r_hi):
        assert not self.sources or self.sources[-1]["rank"] < src
        self.sources.append({"rank": src, "J": J, "I": I, "mask": mask, "r": (r_lo, r_hi)})

    def halo_commit(self, lo, hi):
        import torch
        self.own = (lo, hi)
        Q = self.Q
        merged = self._pairs(self.ljc, self.lir, lo, hi)
        for s in self.sources:
            for J, I, m in zip(s["J"], s["I"], s["mask"]):
                assert lo <= J < hi
                merged[(int(J), int(I))] = merged.get((int(J), int(I)), 0) | int(m)
        # merged CSC of the owned slab
        cols = {c: [] for c in range(lo, hi)}
        for (J, I), m in sorted(merged.items()):
            for b in range(Q):
                for a in range(Q):
                    if m & (1 << (b * Q + a)):
                        cols[J + b].append(I + a)
        self.mjc = np.zeros(hi - lo + 1, np.int64)
        for c in range(lo, hi):
            self.mjc[c - lo + 1] = self.mjc[c - lo] + len(cols[c])
        self.mir = np.array([r for c in range(lo, hi) for r in cols[c]], np.int64)
        pos = {(c, r): self.mjc[c - lo] + k for c in range(lo, hi) for k, r in enumerate(cols[c])}
        self.own_map = np.array([pos[(c, int(r))] for c in range(lo, hi) for r in self.lir[self.ljc[c]:self.ljc[c + 1]]],
                                np.int64)
        for s in self.sources:  # the source's segment = its CSC of these columns: column by column, rows ascending
            scols = {}
            for J, I, m in zip(s["J"], s["I"], s["mask"]):
                for b in range(Q):
                    for a in range(Q):
                        if int(m) & (1 << (b * Q + a)):
                            scols.setdefault(int(J) + b, []).append(int(I) + a)
            s["map"] = np.array([pos[(c, r)] for c in sorted(scols) for r in scols[c]], np.int64)
            s["recv"] = torch.zeros(len(s["map"]), dtype=torch.float64)
            s["rrecv"] = torch.zeros(s["r"][1] - s["r"][0], dtype=torch.float64)

    def assemble_dev(self, U, order):
        _, _, self.lpr, self.lR = self._assemble_local()
        lo, hi = self.own
        self.mpr = np.zeros(len(self.mir))
        self.mpr[self.own_map] = self.lpr[self.ljc[lo]:self.ljc[hi]]
        self.R = self.lR.copy()

    def halo_send_buffers(self, a, b, r_lo):
        import torch
        return (torch.from_numpy(np.ascontiguousarray(self.lpr[self.ljc[a]:self.ljc[b]])),
                torch.from_numpy(np.ascontiguousarray(self.R[r_lo:b])))

    def halo_recv_buffers(self, src):
        s = [x for x in self.sources if x["rank"] == src][0]
        return s["recv"], s["rrecv"]

    def halo_accumulate(self, order):
        for s in self.sources:
            np.add.at(self.mpr, s["map"], s["recv"].numpy())
            self.R[s["r"][0]:s["r"][1]] += s["rrecv"].numpy()

    def ctx_synchronize(self):
        pass


def _worker(rank, world, store_path, name, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from getfem_b200 import halo
    from oracle import oracle
    store = dist.FileStore(store_path, world)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)
    try:
        g = load_golden(name)
        ne = g["conn"].shape[0]
        cuts = [round(r * ne / world) for r in range(world + 1)]
        term = OracleTerm(g, cuts[rank], cuts[rank + 1])
        plan = halo.setup_distributed(term)
        term.assemble_dev(None, 3)
        halo.exchange_distributed(term, plan, 3)
        ndof = int(g["meta"]["ndof"])
        jc, ir, pr, R = oracle.assemble(g["pts"], g["conn"], g["elem_dof"], ndof, g["Q"], g["quad_w"], g["gt_grad"],
                                        g["phi"], g["gphi"], bool(g["gt_linear"]), g["family"], g["fparams"], g["U"])
        lo, hi = plan.own
        assert plan.D[0] == 0 and plan.D[-1] == ndof
        assert np.array_equal(term.mjc, jc[lo:hi + 1] - jc[lo]), "owned column counts"
        assert np.array_equal(term.mir, ir[jc[lo]:jc[hi]]), "owned row indices"
        ref = pr[jc[lo]:jc[hi]]
        assert np.linalg.norm(term.mpr - ref) <= 1e-13 * max(np.linalg.norm(ref), 1e-300)
        assert np.linalg.norm(term.R[lo:hi] - R[lo:hi]) <= 1e-13 * max(np.linalg.norm(R[lo:hi]), 1e-300)
        open(result_path + ".%d" % rank, "w").write("ok sends=%d sources=%d own=%d-%d" % (len(plan.sends), len(plan.sources), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("c3_elast3d_p2_n2", 2), ("c3_elast3d_p2_n2", 3), ("c2_lap3d_p1_n3", 2),
                                        ("c4b_svk_q2_n2", 2), ("x_lap3d_p1_ragged", 3)])
def test_halo_protocol_gloo(name, world):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        store, res = os.path.join(d, "store"), os.path.join(d, "res")
        mp.spawn(_worker, args=(world, store, name, res), nprocs=world, join=True)
        outs = [open(res + ".%d" % r).read() for r in range(world)]
        assert all(o.startswith("ok") for o in outs)
        assert any("sends=1" in o or "sends=2" in o for o in outs), outs


def test_owner_bounds():
    from getfem_b200.halo import owner_bounds
    assert owner_bounds([10, 25, 25, 40], 50) == [0, 10, 25, 25, 50]
    assert owner_bounds([30, 20], 30) == [0, 30, 30]
