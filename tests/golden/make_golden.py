"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/gf_ref_driver,
built by oracle/Makefile from /root/reference/src).  Run in the build container only:

    make -C oracle && python tests/golden/make_golden.py

Each fixture holds the inputs read through the reference's own accessors (mesh, dof table,
reference tables at the quadrature points, U) and the reference result of
ga_workspace::assembly(2) / assembly(1): tangent as CSC (K_jc, K_ir, K_pr) and residual R.
"""
import glob
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DRV = os.path.join(HERE, "..", "..", "oracle", "_ref", "gf_ref_driver")

# name -> driver arguments.  Sizes are kept tiny: the fixtures are committed.
CASES = {
    # BASELINE.json configs at CPU-checkable size
    "c1_lap2d_p1_n6": "dim=2 n=6 gt=pk k=1 q=1 im=2 family=laplace u=random",
    "c2_lap3d_p1_n3": "dim=3 n=3 gt=pk k=1 q=1 im=2 family=laplace u=random",
    "c3_elast3d_p2_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=elast u=random lambda=1 mu=1",
    "c4_nh_ciarlet_q2_n2": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=nh_ciarlet u=smooth lambda=1 mu=1 uamp=0.02",
    "c4b_svk_q2_n2": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=svk u=smooth lambda=1 mu=1 uamp=0.02",
    "c5_lap_q4_n1": "dim=3 n=1 gt=qk k=4 q=1 im=8 family=laplace u=random",
    # extra coverage of the same path
    "x_nh_bonet_p2tet_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=nh_bonet u=smooth lambda=1.3 mu=0.7 uamp=0.05",
    "x_nh_ciarlet_p1tet_n2": "dim=3 n=2 gt=pk k=1 q=3 im=2 family=nh_ciarlet u=smooth lambda=1.3 mu=0.7 uamp=0.05",
    "x_elast2d_p2_n3": "dim=2 n=3 gt=pk k=2 q=2 im=4 family=elast u=random lambda=2 mu=0.5",
    "x_elast3d_p1_n2": "dim=3 n=2 gt=pk k=1 q=3 im=2 family=elast u=random lambda=1.3 mu=0.7",
    "x_lap3d_q1_n3": "dim=3 n=3 gt=qk k=1 q=1 im=3 family=laplace u=random a=2.5",
    "x_lap3d_p2_n2": "dim=3 n=2 gt=pk k=2 q=1 im=4 family=laplace u=random",
    "x_lapvec3d_p1_n2": "dim=3 n=2 gt=pk k=1 q=3 im=2 family=laplace_vec u=random",
    "x_lap2d_q2_n3": "dim=2 n=3 gt=qk k=2 q=1 im=4 family=laplace u=random",
    "x_lap3d_p1_ragged": "dim=3 nx=1 ny=2 nz=3 gt=pk k=1 q=1 im=2 family=laplace u=random",
    "x_elast3d_q2_n1": "dim=3 n=1 gt=qk k=2 q=3 im=4 family=elast u=random lambda=1 mu=1",
    # volumic source term (order 1 only, empty tangent): the RHS of BASELINE config 1
    "c1b_source2d_p1_n6": "dim=2 n=6 gt=pk k=1 q=1 im=2 family=source u=random a=1.5",
    "x_source3d_p2vec_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=source u=random a=0.5",
    "x_source3d_q2_n2": "dim=3 n=2 gt=qk k=2 q=1 im=6 family=source u=random a=2",
    # mesh regions: boundary faces (unit normal, surface Jacobian, C&E.cc:8827-8848) with the
    # Neumann / normal source / boundary mass (Robin, Dirichlet penalisation) terms, and sub-regions of convexes
    "r_mass3d_p2vec_outer": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=mass region=outer u=random a=1.5",
    "r_source3d_p2vec_xmax": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=source region=xmax u=random a=0.5",
    "r_nsource3d_p2vec_outer": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=nsource region=outer u=random a=0.8",
    "r_nsource2d_p1_outer": "dim=2 n=4 gt=pk k=1 q=1 im=2 family=nsource region=outer u=random a=1.2",
    "r_mass2d_p2_outer": "dim=2 n=3 gt=pk k=2 q=1 im=4 family=mass region=outer u=random a=2",
    "r_mass3d_q2_zmin": "dim=3 n=2 gt=qk k=2 q=1 im=6 family=mass region=zmin u=random a=1.5",
    "r_nsource3d_q2vec_outer": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=nsource region=outer u=random a=0.7",
    "r_elast3d_p2_half": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=elast region=half u=random lambda=1 mu=1",
    "r_lap3d_p1_half": "dim=3 n=4 gt=pk k=1 q=1 im=2 family=laplace region=half u=random",
    # several expressions in ONE workspace (what a model's bricks add up to): stiffness + Robin boundary mass + Neumann
    # load + normal source; K = sum of the order-2 trees (union pattern), V = sum of the order-1 trees
    "m_elast_robin_neumann_p2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=elast u=random lambda=1.2 mu=0.8 "
                                "family2=mass region2=outer a2=3.5 family3=source region3=xmax a3=0.6 "
                                "family4=nsource region4=zmin a4=0.9",
    "m_poisson_rhs_robin_2d": "dim=2 n=6 gt=pk k=1 q=1 im=2 family=laplace u=random a=1.3 family2=source a2=2.0 "
                              "family3=mass region3=outer a3=5.0 family4=nsource region4=outer a4=0.7",
    "m_lap_q2_robin": "dim=3 n=2 gt=qk k=2 q=1 im=6 family=laplace u=random a=0.9 family2=mass region2=zmin a2=4.0 "
                      "family3=source a3=1.1",
    # fem-data coefficients (ws.add_fem_constant): heterogeneous material, distributed loads
    "f_lap3d_p2_coefp1": "dim=3 n=2 gt=pk k=2 q=1 im=4 family=laplace u=random a=1.3 coef=fem kd=1",
    "f_elast3d_p2_coefp1": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=elast u=random lambda=1.2 mu=0.8 coef=fem kd=1",
    "f_elast2d_p2_coefp2": "dim=2 n=3 gt=pk k=2 q=2 im=4 family=elast u=random lambda=2 mu=0.5 coef=fem kd=2",
    "f_source3d_p2vec_coefp2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=source u=random a=0.5 coef=fem kd=2",
    "f_source2d_p1_coefp1": "dim=2 n=6 gt=pk k=1 q=1 im=2 family=source u=random a=1.5 coef=fem kd=1",
    "f_mass3d_q2_coefq1": "dim=3 n=2 gt=qk k=2 q=1 im=6 family=mass u=random a=1.5 coef=fem kd=1",
    "f_lap3d_q2_coefq2": "dim=3 n=2 gt=qk k=2 q=1 im=6 family=laplace u=random a=0.7 coef=fem kd=2",
    "f_source3d_p2vec_xmax_coefp1": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=source region=xmax u=random a=0.5 coef=fem kd=1",
    "f_mass3d_p2_outer_coefp1": "dim=3 n=2 gt=pk k=2 q=1 im=4 family=mass region=outer u=random a=2.5 coef=fem kd=1",
    "r_nh_ciarlet_q2_half": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=nh_ciarlet region=half u=smooth lambda=1 mu=1 uamp=0.02",
    # DISTORTED meshes (noise=amp: every node moved by amp*h*uniform(-1,1)): every simplex has its own K / B / J, every GT_QK
    # cell is non-affine (geometry evaluated at every Gauss point, C&E.cc:8789-8866), boundary normals vary face by face
    "d_elast3d_p2_n3": "dim=3 n=3 gt=pk k=2 q=3 im=4 family=elast u=random lambda=1.3 mu=0.7 noise=0.2",
    "d_lap3d_p1_n3": "dim=3 n=3 gt=pk k=1 q=1 im=2 family=laplace u=random a=1.7 noise=0.2",
    "d_lap3d_p2_n2": "dim=3 n=2 gt=pk k=2 q=1 im=4 family=laplace u=random noise=0.2",
    "d_elast2d_p2_n4": "dim=2 n=4 gt=pk k=2 q=2 im=4 family=elast u=random lambda=2 mu=0.5 noise=0.2",
    "d_mass3d_p2vec_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=mass u=random a=1.5 noise=0.2",
    "d_lap3d_q2_n2": "dim=3 n=2 gt=qk k=2 q=1 im=6 family=laplace u=random a=0.9 noise=0.15",
    "d_lap_q4_n1": "dim=3 n=1 gt=qk k=4 q=1 im=8 family=laplace u=random noise=0.15",
    "d_lap_q4_n2": "dim=3 nx=2 ny=1 nz=1 gt=qk k=4 q=1 im=8 family=laplace u=random noise=0.15",
    "d_elast3d_q2_n2": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=elast u=random lambda=1 mu=1 noise=0.15",
    "d_nh_ciarlet_q2_n2": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=nh_ciarlet u=smooth lambda=1 mu=1 uamp=0.02 noise=0.15",
    "d_svk_q2_n2": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=svk u=smooth lambda=1 mu=1 uamp=0.02 noise=0.15",
    "d_nh_bonet_p2tet_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=nh_bonet u=smooth lambda=1.3 mu=0.7 uamp=0.05 noise=0.2",
    "d_source3d_p2vec_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=source u=random a=0.5 noise=0.2",
    "d_mass3d_p2vec_outer": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=mass region=outer u=random a=1.5 noise=0.2",
    "d_nsource3d_q2vec_outer": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=nsource region=outer u=random a=0.7 noise=0.15",
    "d_elast3d_p2_half": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=elast region=half u=random lambda=1 mu=1 noise=0.2",
    # mid-size fixtures of the two headline configurations (values checked beyond n = 2)
    "c3_elast3d_p2_n8": "dim=3 n=8 gt=pk k=2 q=3 im=4 family=elast u=random lambda=1 mu=1",
    "c4_nh_ciarlet_q2_n4": "dim=3 n=4 gt=qk k=2 q=3 im=6 family=nh_ciarlet u=smooth lambda=1 mu=1 uamp=0.02",
    # P4 simplices (the degree of the reference's published table, contrib/opt_assembly/opt_assembly.cc:704-712), distorted
    "p4_lap3d_n1": "dim=3 n=1 gt=pk k=4 q=1 im=8 family=laplace u=random a=1.3 noise=0.15",
    "p4_elast3d_n1": "dim=3 n=1 gt=pk k=4 q=3 im=8 family=elast u=random lambda=1.3 mu=0.7 noise=0.15",
    "p4_svk3d_n1": "dim=3 n=1 gt=pk k=4 q=3 im=8 family=svk u=smooth lambda=1 mu=1 uamp=0.03",
    "p4_elast2d_n3": "dim=2 n=3 gt=pk k=4 q=2 im=8 family=elast u=random lambda=1.3 mu=0.7 noise=0.15",
    "p4_mass2d_n2": "dim=2 n=2 gt=pk k=4 q=1 im=8 family=mass u=random a=0.8",
    # ORACLE-ONLY fixtures (prefix o_: not yet a device family; tests/conftest.py keeps them out of the GPU parametrisations).
    # Compressible Mooney-Rivlin (the law of the reference's tests/nonlinear_elastostatic.cc), C10 = lambda, C01 = mu, D1 = a
    "o_mooney_rivlin_q2_n2": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=mooney_rivlin u=smooth lambda=0.8 mu=0.3 a=2.0 uamp=0.03",
    "o_mooney_rivlin_p2tet_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=mooney_rivlin u=smooth lambda=1.1 mu=0.2 a=1.5 uamp=0.05",
    # Ciarlet-Geymonat (lambda, mu, a) and generalized Blatz-Ko (a, b, c, d, n: the law's default set), the other laws of
    # add_finite_strain_elasticity_brick (getfem_nonlinear_elasticity.cc:2276-2290)
    "o_ciarlet_geymonat_q2_n2": "dim=3 n=2 gt=qk k=2 q=3 im=6 family=ciarlet_geymonat u=smooth lambda=1.3 mu=0.7 a=0.25 uamp=0.03",
    "o_blatz_ko_p2tet_n2": "dim=3 n=2 gt=pk k=2 q=3 im=4 family=blatz_ko u=smooth params=1.0,1.0,1.5,-0.5,1.5 uamp=0.04",
}


def main():
    only = sys.argv[1:]
    for name, args in CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            out = subprocess.check_output([DRV] + args.split() + ["out=" + d]).decode()
            meta = json.loads(out.strip().splitlines()[-1])
            arrs = {os.path.basename(f)[:-4]: np.load(f) for f in glob.glob(d + "/*.npy")}
        meta["driver_args"] = args
        arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        for k in ("conn", "elem_dof", "K_jc", "K_ir"):
            arrs[k] = arrs[k].astype(np.int32)
        if "coef=fem" in args:  # keep the all-point tables: the data fem's basis table covers all points
            arrs["d_elem_dof"] = arrs["d_elem_dof"].astype(np.int32)
        elif "region" not in args:  # the all-point tables and face data only travel with the region fixtures
            for k in ("all_w", "all_x", "all_gt_grad", "all_phi", "all_gphi", "face_first", "face_nq", "ref_normals",
                      "gdata"):
                arrs.pop(k, None)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
        print(name, meta["fem"], meta["im"], "ne", meta["ne"], "ndof", meta["ndof"], "nnz", meta["nnz"],
              os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
