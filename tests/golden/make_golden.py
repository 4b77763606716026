"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

Run here (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
It imports /root/reference/src/python with a throw-away matplotlib stub (the reference
imports matplotlib at module level, diffusion.py:2, plot.py:2-6, solution.py:2-3; the image
has none), runs the reference functions on seeded inputs, cross-checks the oracle
(`oracle/v3s_oracle.py`) against them and writes small .npz fixtures next to this file.
Nothing under tests/ reads /root/reference at test time.
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/python"
sys.dont_write_bytecode = True


def install_matplotlib_stub():
    d = tempfile.mkdtemp(prefix="mplstub_")
    pkg = os.path.join(d, "matplotlib")
    os.makedirs(pkg)
    open(os.path.join(pkg, "__init__.py"), "w").write("")
    open(os.path.join(pkg, "pylab.py"), "w").write("import numpy as np\npinv = np.linalg.pinv\n")
    open(os.path.join(pkg, "pyplot.py"), "w").write(
        "class _M:\n"
        "  def __getattr__(self, n):\n    return _M()\n"
        "  def __call__(self, *a, **k):\n    return _M()\n"
        "  def __iter__(self):\n    return iter((_M(), _M()))\n"
        "def __getattr__(n):\n  return _M()\n")
    open(os.path.join(pkg, "axes.py"), "w").write("class Axes: pass\n")
    open(os.path.join(pkg, "figure.py"), "w").write("class Figure: pass\n")
    sys.path.insert(0, d)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    install_matplotlib_stub()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import v3s_oracle as orc
    from projection import Projector
    from rotation import Rotation
    from distance import DistanceMatrix
    from diffusion import DiffusionMapSolver
    from compute import ComputeUtils

    rep = {}

    # ---- (1) single images at n=15 and n=31 with proper rotations ------------------
    for n, nrot in ((7, 4), (15, 4), (31, 3)):
        R9, _ = orc.random_rotations(nrot, seed=100 + n)
        Exp = Projector.experiment_parameters(n)["Experiment"]
        Prot = quiet(Projector.synsthesize_3d, n)
        Q, mask = Projector.q_and_mask(Exp)
        k = Projector.fourier_scaled_axes_from_grid(Prot["Grid_3D"])[2]
        imgs, edrot, edf = [], [], []
        for c in range(nrot):
            Rm = R9[:, c].reshape(3, 3).T
            im, er, ef, _ = Projector.generate_single_image(Prot["ED"], Rm, k, Q, mask)
            imgs.append(im.copy()); edrot.append(er.copy()); edf.append(ef.copy())
        imgs = np.array(imgs)
        o_ed = orc.synthesize_3d(n)["ED"]
        assert np.array_equal(o_ed, Prot["ED"]), "oracle object differs"
        o_imgs = orc.generate_from_rotations(R9, n)
        err = np.max(np.abs(o_imgs - imgs)) / np.max(np.abs(imgs))
        rep[f"single_image_n{n}_rel_err"] = float(err)
        assert err < 1e-12, err
        np.savez_compressed(os.path.join(HERE, f"images_n{n}.npz"), R9=R9, ED=Prot["ED"],
                            Images=imgs, ED_rot0=edrot[0], ED_rot_f0=edf[0],
                            kx=k["x"], Qx=Q["x"], Qy=Q["y"], Qz=Q["z"], mask=mask)

    # ---- (2) rotations: Hopf grid + clip bug ------------------------------------------
    for nr in (2, 3, 8):
        R, Axis, Quat = quiet(Rotation.all_rot_variables, nr ** 3)
        oR, oAxis, oQuat = orc.all_rot_variables(nr ** 3)
        assert np.allclose(R, oR, rtol=0, atol=1e-15) and np.allclose(Axis, oAxis, rtol=0, atol=1e-15)
        assert np.allclose(Quat, oQuat, rtol=0, atol=1e-15)
    np.savez_compressed(os.path.join(HERE, "hopf_8.npz"), R=R, Axis=Axis, Quat=Quat)

    # ---- (3) default config: stages 1-3a are bit-reproducible ------------
    Images, R, Axis, Quat = quiet(Projector.generate, 8, 7)
    o_Images = orc.generate(8, 7)[0]
    err = np.max(np.abs(o_Images - Images)) / np.max(np.abs(Images))
    rep["default_images_rel_err"] = float(err)
    assert err < 1e-12
    B = quiet(DistanceMatrix.convert_to_round_distance, Images)
    S2, N = quiet(DistanceMatrix.knn_from_round_distance, B, 24)
    oB = orc.convert_to_round_distance(Images)
    assert np.array_equal(B, oB), "oracle round distance differs bitwise"
    oS2, oN = orc.knn_from_round_distance(oB, 24)
    assert np.array_equal(S2, oS2) and np.array_equal(N, oN)
    S2m = S2.copy()
    W = quiet(DiffusionMapSolver.s2_to_w_matrix, S2m, N, 0.7)
    oS2m = S2.copy()
    oW = orc.s2_to_w_matrix(oS2m, N, 0.7)
    assert np.array_equal(W, oW) and np.array_equal(S2m, oS2m)
    W2 = quiet(DiffusionMapSolver.anisotropic_norm, W)
    oW2 = orc.anisotropic_norm(oW)
    rep["default_W2_max_abs"] = float(np.max(np.abs(W2 - oW2)))
    assert np.allclose(W2, oW2, rtol=1e-14, atol=0)
    P_ep, D = quiet(DiffusionMapSolver.dm_cov, W2)
    oP, odi = orc.dm_cov(oW2, dense_D=False)
    assert np.allclose(P_ep, oP, rtol=1e-14, atol=0) and np.allclose(np.diag(D), odi, rtol=1e-15)
    lam_all = np.linalg.eigvalsh(P_ep)
    z = np.random.default_rng(7).standard_normal(512)
    np.savez_compressed(os.path.join(HERE, "default_cfg.npz"), Images=Images, R=R, Axis=Axis, Quat=Quat,
                        S2=S2, N=N, S2_mutated=S2m, W_rowsum=W.sum(1), W_colsum=W.sum(0),
                        W2_rowsum=W2.sum(1), P_ep_z=P_ep @ z, z=z, D_diag=np.diag(D).copy(),
                        P_ep_row0=P_ep[0].copy(), P_ep_row511=P_ep[511].copy(),
                        lam_top=np.sort(lam_all)[::-1][:200], B_row0=B[0].copy(), B_row100=B[100].copy())

    # ---- (4) connected case: proper rotations, n=15 (31x31), full downstream ----------
    Nc, kc, eps = 1000, 150, 0.7
    R9, Quat = orc.random_rotations(Nc, seed=1)
    Imgs = orc.generate_from_rotations(R9, 15).astype(np.float32).astype(np.float64)  # fp32-exact shared input
    S2, N = quiet(DistanceMatrix.knn, Imgs, kc)
    oS2, oN = orc.knn(Imgs, kc)
    assert np.array_equal(S2, oS2) and np.array_equal(N, oN)
    S2_in = S2.copy()
    W = quiet(DiffusionMapSolver.s2_to_w_matrix, S2.copy(), N, eps)
    P_ep, D = quiet(DiffusionMapSolver.dm_cov, quiet(DiffusionMapSolver.anisotropic_norm, W))
    # graph connectivity
    import scipy.sparse.csgraph as csg
    import scipy.sparse as sp
    ncomp = csg.connected_components(sp.csr_array(P_ep > 0))[0]
    rep["connected_case_components"] = int(ncomp)
    S2_a = S2_in.copy()
    Ps, Lambda = quiet(DiffusionMapSolver.diffusion_coordinate, S2_a, N, eps)
    oPs, oLambda, _, _, oP, odi = orc.diffusion_coordinate(S2_in.copy(), N, eps, return_all=True)
    rep["connected_lambda"] = [float(x) for x in Lambda]
    rep["connected_lambda_err_oracle_vs_ref"] = float(np.max(np.abs(Lambda - oLambda)))
    ang = orc.principal_angles(Ps, oPs)
    rep["connected_subspace_angle_oracle_vs_ref"] = float(np.max(ang))
    c, c0, Sigma_T = quiet(DiffusionMapSolver.dm_fit_all, Ps, R9, True)
    oc, _, oSig, oRecon_pre, _, _, _ = orc.dm_fit_all(oPs, R9, True, return_all=True)
    Recon_pre = Ps[:, 1:] @ c
    rep["connected_recon_pre_err"] = float(np.max(np.abs(Recon_pre - oRecon_pre)))
    rep["connected_sigma_T_ref"] = float(Sigma_T)
    rep["connected_sigma_T_oracle"] = float(oSig)
    lam_all = np.sort(np.linalg.eigvalsh(P_ep))[::-1]
    rep["connected_lambda_11"] = float(lam_all[10])
    z = np.random.default_rng(8).standard_normal(Nc)
    np.savez_compressed(os.path.join(HERE, "connected_n15.npz"), R9=R9, Images_f32=Imgs.astype(np.float32),
                        S2=S2_in, N=N.astype(np.int32), eps=eps, k=kc, Lambda=Lambda, Ps=Ps, c=c,
                        Recon_pre=Recon_pre, Sigma_T=Sigma_T, D_diag=np.diag(D).copy(), P_ep_z=P_ep @ z, z=z,
                        lam_top12=lam_all[:12])

    # ---- (5) lsq / polar / quaternion helpers on random input ---------------------------
    rng = np.random.default_rng(3)
    A = rng.standard_normal((50, 9)); b = rng.standard_normal((50, 9))
    M = ComputeUtils.lsq_linear(A, b)
    assert np.allclose(M, orc.lsq_linear(A, b), rtol=1e-13)
    Ms = rng.standard_normal((20, 3, 3))
    pol = np.array([ComputeUtils.polar_decompose(m) for m in Ms])
    opol = np.array([orc.polar_decompose(m) for m in Ms])
    assert np.allclose(pol, opol, atol=1e-14)
    Rq, _ = orc.random_rotations(20, seed=5)
    Rb = Rq.T.reshape(20, 3, 3)
    qs = np.hstack([Rotation.rot_mat_to_quat(r) for r in Rb])
    assert np.allclose(qs, orc.rot_mat_to_quat_batch(Rb), atol=1e-15)
    q2 = orc.unit_mag_pos(qs + 0.05 * rng.standard_normal(qs.shape))
    sig = Rotation.mean_geodesic_distance_degrees(qs, q2)
    assert abs(sig - orc.mean_geodesic_distance_degrees(qs, q2)) < 1e-12
    np.savez_compressed(os.path.join(HERE, "helpers.npz"), A=A, b=b, M=M, Ms=Ms, polar_compat=pol,
                        Rb=Rb, quats=qs, q2=q2, sigma=sig)

    import json
    json.dump(rep, open(os.path.join(HERE, "make_golden_report.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
