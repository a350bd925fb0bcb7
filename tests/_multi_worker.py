"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): element blocks per rank,
halo exchange over NCCL -- through the library's own communicator (csrc/comm.cu) and through torch.distributed P2P --
against the single-GPU assembly of the same mesh computed by every rank.  Prints one JSON line per rank."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import getfem_b200 as gf
    from getfem_b200 import capi, fem_tables, halo
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dim, nsub, gt, k, Q, im, family, params = json.loads(sys.argv[1])
    stream = torch.cuda.Stream()
    ctx = capi.Context(local, stream.cuda_stream)
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    mf = gf.mesh_fem(m, Q)
    mf.set_classical_finite_element(k)
    dmesh, dfem = m.device(ctx), mf.device(ctx)
    t = fem_tables.classical_tables(gt, dim, k, im)
    tab = capi.DeviceTables(ctx, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    ndof = dfem.ndof
    U = np.random.default_rng(7).uniform(-1, 1, ndof) * (0.02 if family in ("svk", "nh_ciarlet") else 1.0)
    ORDER = capi.TANGENT | capi.RESIDUAL
    out = {"rank": rank, "world": world}
    with torch.cuda.stream(stream):
        U_dev = torch.from_numpy(U).cuda()
        full = capi.DeviceTerm(ctx, dmesh, dfem, tab, family, params, 1.0, 0)
        full.assemble_dev(U_dev.data_ptr(), ORDER)
        jc, ir, pr = full.export_csc()
        R = full.export_residual()
        ne = m.nb_convex()
        e0, e1 = rank * ne // world, (rank + 1) * ne // world
        for mode in ("library", "torch"):
            term = capi.DeviceTerm(ctx, dmesh, dfem, tab, family, params, 1.0, 0)
            term.set_element_range(e0, e1)
            plan = halo.setup_distributed(term, U_dev.data_ptr())
            comm = None
            if mode == "library":
                comm = halo.make_communicator(ctx)
                halo.register_sends(term, plan)
            for rep in range(2):
                term.assemble_dev(U_dev.data_ptr(), ORDER)
                if comm is not None:
                    halo.exchange_nccl(term, comm, ORDER)
                else:
                    halo.exchange_distributed(term, plan, ORDER)
            ctx.synchronize()
            lo, hi = term.owned_range()
            tjc, tir, tpr = term.export_csc()
            tR = term.export_residual()
            a, b, A, B = tjc[lo], tjc[hi], jc[lo], jc[hi]
            res = {"own": [int(lo), int(hi)], "sends": len(plan.sends), "sources": len(plan.sources),
                   "cols_ok": bool(np.array_equal(tjc[lo:hi + 1] - a, jc[lo:hi + 1] - A)),
                   "rows_ok": bool(b - a == B - A and np.array_equal(tir[a:b], ir[A:B]))}
            if res["rows_ok"] and hi > lo:
                res["rel_K"] = float(np.linalg.norm(tpr[a:b] - pr[A:B]) / max(np.linalg.norm(pr[A:B]), 1e-300))
                res["rel_R"] = float(np.linalg.norm(tR[lo:hi] - R[lo:hi]) / max(np.linalg.norm(R[lo:hi]), 1e-300))
                res["slab_checksum"] = float(np.sum(tpr[a:b]))
            out[mode] = res
            del term, comm
    print("MULTI " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
