"""Multi-GPU halo exchange of the assembled tangent / residual (include/gfgpu.h "multi-GPU").

One process per GPU; every rank assembles the element block it was given (DeviceTerm.set_element_range).
Dofs are owned by the rank whose block touches them first, which with the reference's first-touch numbering
(getfem_mesh_fem.cc:320-446) gives contiguous ranges [D[r], D[r+1]); the columns a rank touches below its
range are ghosts.  Replaces the reference's MPI scheme -- every rank holds a full-size partial matrix / vector
and MPI_SUM_SPARSE_MATRIX / MPI_SUM_VECTOR reduce them (getfem_generic_assembly_workspace.cc:855-858,
getfem_models.cc:586,2572,2608) -- by column-owned slabs and ONE point-to-point exchange per assembly:

  setup    (once)  touched ranges all-gathered -> owner bounds; every rank announces the (J, I, keep mask) pairs of
                   its ghost columns to their owners (host mediated, pickled objects); owners merge them into
                   their pattern (drop rule = OR of the masks);
  exchange (every assembly)  ghost values = a contiguous slice of pr, residual = a slice of R, sent with
                   isend/irecv (NCCL over NVLink on device tensors; gloo on CPU tensors in the tests), then the
                   owner adds the received parts in ascending source rank: fixed order, no atomics.

The term object only has to provide the halo_* methods of capi.DeviceTerm, so the protocol is tested on CPU under
gloo with an oracle-backed term (tests/test_halo_gloo.py) and on one GPU with several terms (tests/test_gpu_halo.py).
"""
from dataclasses import dataclass, field


def owner_bounds(touched_hi, ndof):
    """D[r] = first dof owned by rank r: everything below max(hi of the lower ranks); the last rank keeps the rest."""
    D = [0]
    for h in touched_hi:
        D.append(max(D[-1], int(h)))
    D[-1] = max(D[-1], int(ndof))
    return D


@dataclass
class HaloPlan:
    rank: int
    world: int
    D: list
    touched: tuple
    sends: list = field(default_factory=list)    # (owner rank q, dof_lo, dof_hi, r_lo)
    sources: list = field(default_factory=list)  # source ranks, ascending

    @property
    def own(self):
        return self.D[self.rank], self.D[self.rank + 1]


def announcements(term, rank, D, touched):
    """{owner q: (J, I, mask, r_lo, r_hi)} for the ghost columns of `term`, and the matching send list."""
    lo, hi = touched
    out, sends = {}, []
    for q in range(rank):
        a, b = D[q], D[q + 1]
        if b <= a or lo >= b or hi <= a:
            continue
        J, I, m = term.halo_ghost_pairs(a, b)
        r_lo = max(lo, a)
        if len(J) == 0:
            continue
        out[q] = (J, I, m, r_lo, b)
        sends.append((q, a, b, r_lo))
    return out, sends


def merge(term, plan, incoming):
    """incoming: {source rank: (J, I, mask, r_lo, r_hi)} announced for the columns this rank owns."""
    for src in sorted(incoming):
        J, I, m, r_lo, r_hi = incoming[src]
        term.halo_add_source(src, J, I, m, r_lo, r_hi)
        plan.sources.append(src)
    term.halo_commit(*plan.own)


# ---------------------------------------------------------------- torch.distributed driver (one process per GPU)
def setup_distributed(term, U_dev_ptr=None, group=None):
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    touched = term.halo_begin(U_dev_ptr)
    his = [None] * world
    dist.all_gather_object(his, int(touched[1]), group=group)
    D = owner_bounds(his, term.ndof)
    plan = HaloPlan(rank, world, D, touched)
    out, plan.sends = announcements(term, rank, D, touched)
    everything = [None] * world
    dist.all_gather_object(everything, out, group=group)
    incoming = {src: everything[src][rank] for src in range(world) if rank in everything[src]}
    merge(term, plan, incoming)
    return plan


def make_communicator(ctx, group=None):
    """gfgpu communicator (NCCL, created inside the C ABI) over the ranks of a torch.distributed group: rank 0's
    ncclGetUniqueId travels through the group's object broadcast, the exchange itself never touches torch."""
    import torch.distributed as dist
    from . import capi
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return capi.Communicator(ctx, world, rank, box[0])


def register_sends(term, plan):
    """tells the library what this rank sends at every exchange (gfgpu_term_halo_add_send)"""
    for q, a, b, r_lo in plan.sends:
        term.halo_add_send(q, a, b, r_lo)


def exchange_nccl(term, comm, order_mask):
    """Call after term.assemble_dev(..., order_mask): gfgpu_term_halo_exchange, asynchronous on the context's stream."""
    term.halo_exchange(comm, order_mask)


def exchange_distributed(term, plan, order_mask, group=None):
    """Call after term.assemble_dev(..., order_mask), on the stream the term's context was created with."""
    import torch.distributed as dist
    from . import capi
    ops, keep = [], []
    for q, a, b, r_lo in plan.sends:
        pr, R = term.halo_send_buffers(a, b, r_lo)
        if (order_mask & capi.TANGENT) and pr.numel():
            ops.append(dist.P2POp(dist.isend, pr, q, group))
        if (order_mask & capi.RESIDUAL) and R.numel():
            ops.append(dist.P2POp(dist.isend, R, q, group))
        keep += [pr, R]
    for src in plan.sources:
        pr, R = term.halo_recv_buffers(src)
        if (order_mask & capi.TANGENT) and pr.numel():
            ops.append(dist.P2POp(dist.irecv, pr, src, group))
        if (order_mask & capi.RESIDUAL) and R.numel():
            ops.append(dist.P2POp(dist.irecv, R, src, group))
        keep += [pr, R]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    term.halo_accumulate(order_mask)


# ---------------------------------------------------------------- in-process driver (several terms, one process)
def setup_local(terms, U_dev_ptr=None):
    """terms[r] plays rank r (same mesh / fem, disjoint element blocks in rank order)."""
    world = len(terms)
    touched = [t.halo_begin(U_dev_ptr) for t in terms]
    D = owner_bounds([h for _, h in touched], terms[0].ndof)
    plans, outs = [], []
    for r, t in enumerate(terms):
        plan = HaloPlan(r, world, D, touched[r])
        out, plan.sends = announcements(t, r, D, touched[r])
        plans.append(plan)
        outs.append(out)
    for r, t in enumerate(terms):
        merge(t, plans[r], {src: outs[src][r] for src in range(world) if r in outs[src]})
    return plans


def exchange_local(terms, plans, order_mask):
    """Copies the send slices into the owners' receive buffers (same device or host), then accumulates."""
    for t in terms:
        t.ctx_synchronize()
    cuda = False
    for r, t in enumerate(terms):
        for q, a, b, r_lo in plans[r].sends:
            pr, R = t.halo_send_buffers(a, b, r_lo)
            dpr, dR = terms[q].halo_recv_buffers(r)
            assert dpr.numel() == pr.numel() and dR.numel() == R.numel(), (dpr.numel(), pr.numel(), dR.numel(), R.numel())
            dpr.copy_(pr)
            dR.copy_(R)
            cuda = cuda or pr.is_cuda
    if cuda:
        import torch
        torch.cuda.synchronize()
    for t in terms:
        t.halo_accumulate(order_mask)


def selfcheck_distributed(ctx, dim, nsub, gt, k, Q, im, family, params, comm=None, group=None):
    """Parity of the REAL multi-GPU path at a size every rank can also assemble alone: element blocks + halo exchange
    against the single-GPU assembly of the same mesh.  Returns this rank's figures (pattern of the owned slab bit-exact,
    relative errors of its values and of the residual slice); bench.py gathers them on rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import capi, fem_tables
    from .workspace import mesh as Mesh, mesh_fem as MeshFem, regular_unit_mesh
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    m = Mesh()
    regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    mf = MeshFem(m, Q)
    mf.set_classical_finite_element(k)
    dmesh, dfem = m.device(ctx), mf.device(ctx)
    t = fem_tables.classical_tables(gt, dim, k, im)
    tab = capi.DeviceTables(ctx, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    ndof = dfem.ndof
    # finite strain laws: nodal noise well below the cell size, so that det(I + Grad_u) stays positive on every cell
    U = np.random.default_rng(7).uniform(-1, 1, ndof) * (0.05 / max(nsub) if family in ("svk", "nh_ciarlet", "nh_bonet") else 1.0)
    U_dev = torch.from_numpy(U).to("cuda:%d" % ctx.device)
    order = capi.TANGENT | capi.RESIDUAL
    full = capi.DeviceTerm(ctx, dmesh, dfem, tab, family, params, 1.0, 0)
    full.assemble_dev(U_dev.data_ptr(), order)
    jc, ir, pr = full.export_csc()
    R = full.export_residual()
    ne = m.nb_convex()
    term = capi.DeviceTerm(ctx, dmesh, dfem, tab, family, params, 1.0, 0)
    term.set_element_range(rank * ne // world, (rank + 1) * ne // world)
    plan = setup_distributed(term, U_dev.data_ptr(), group)
    if comm is not None:
        register_sends(term, plan)
    term.assemble_dev(U_dev.data_ptr(), order)
    if comm is not None:
        exchange_nccl(term, comm, order)
    else:
        exchange_distributed(term, plan, order, group)
    ctx.synchronize()
    lo, hi = term.owned_range()
    tjc, tir, tpr = term.export_csc()
    tR = term.export_residual()
    a, b, A, B = tjc[lo], tjc[hi], jc[lo], jc[hi]
    ok = bool(np.array_equal(tjc[lo:hi + 1] - a, jc[lo:hi + 1] - A) and b - a == B - A and np.array_equal(tir[a:b], ir[A:B]))
    out = {"rank": rank, "own": [int(lo), int(hi)], "pattern_ok": ok, "elements": int(ne), "ndof": int(ndof)}
    if ok and hi > lo:
        out["rel_K"] = float(np.linalg.norm(tpr[a:b] - pr[A:B]) / max(np.linalg.norm(pr[A:B]), 1e-300))
        out["rel_R"] = float(np.linalg.norm(tR[lo:hi] - R[lo:hi]) / max(np.linalg.norm(R[lo:hi]), 1e-300))
    return out
