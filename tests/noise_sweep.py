"""Stand-alone GPU diagnostic (not a pytest file): how the recovered orientations degrade with the
photon budget of the Poisson-noise transform (BASELINE config 5's input model), at a given size."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
from v3s import _util, orientation_errors_deg
from v3s.pipeline import ShardPlan, run_pipeline
from v3s.projection import Projector
from v3s.rotation import Rotation

N, n, n_p, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 31, int(sys.argv[2]) if len(sys.argv) > 2 else 64, 150
ops = _util.get_ops()
R9, _ = Rotation.random_rotations(N, seed=0)
ed = Projector.object_on_device(n, ops)
rot9 = Projector.rot9_rowmajor(R9, ops)
rot_y = ops.to_device(np.ascontiguousarray(R9.T), torch.float64)
for photons in (None, 1e9, 1e8, 1e7, 1e6, 1e5):
    st = {}
    out = run_pipeline(ops, ShardPlan(N), ed, rot9, rot_y, n_p, k, 0.7, polar_mode=2, check=False, validate=False, stats=st,
                       photons_per_pattern=photons)
    lam = out["Lambda"].cpu().numpy()
    err = orientation_errors_deg(out["recon_pre"], R9)
    img = out["Images_local"]
    print("photons %-8s lambda[1:5] %s  converged %s matvecs %d  Sigma_T %.2f  aligned error median %.2f deg p90 %.2f  max count %.0f"
          % (photons, np.round(lam[1:5], 6), st.get("converged"), st.get("matvecs"), out["sigma_T"].item(), np.median(err),
             np.percentile(err, 90), img.max().item()), flush=True)
