import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "progressive-x_b200"))
import pyprogressivex
G = np.load(ROOT / "tests" / "golden" / "reference_scenes.npz")
scene, seed, sampler = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
corrs, ref = G[f"{scene}_corrs"], G[f"{scene}_labels"]
F, lab = pyprogressivex.findTwoViewMotions(corrs, 640, 480, 640, 480, threshold=0.75, conf=0.5, spatial_coherence_weight=float(sys.argv[4]) if len(sys.argv) > 4 else 0.5,
    neighborhood_ball_radius=50.0, maximum_tanimoto_similarity=0.4, max_iters=10000, minimum_point_number=7,
    maximum_model_number=4, sampler_id=sampler, scoring_exponent=1.0, seed=seed, do_logging=True)
print("M", F.shape[0] // 3, np.bincount(lab), np.bincount(ref))
