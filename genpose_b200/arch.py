"""Static description of the network on the hot path (default GenPose configuration).

Everything here restates constants of the reference so that host code, kernels and tests agree:
  * encoder levels   : networks/pts_encoder/pointnet2.py:57-66 (ClsMSG_CFG_Light, `--pointnet2_params light`)
                       + use_xyz channel bump networks/pts_encoder/pointnet2_utils/pointnet2/pointnet2_modules.py:89-90
  * score/energy net : networks/gf_algorithms/scorenet.py:104-170, energynet.py:52-120 (`Rx_Ry_and_T`, pose_dim 9)
  * VE SDE           : networks/gf_algorithms/sde.py:15-28, :90-97 (sigma_min 0.01, sigma_max 50, eps 1e-5)
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

NUM_POINTS = 1024
POSE_DIM = 9
PTS_FEAT_DIM = 1024
T_EMBED_DIM = 128
POSE_FEAT_DIM = 256
HEAD_HIDDEN = 256
HEADS = ("rot_x", "rot_y", "trans")
FUSED_IN = PTS_FEAT_DIM + T_EMBED_DIM + POSE_FEAT_DIM  # 1408, column order [pts | t | pose] (scorenet.py:204)

SIGMA_MIN = 0.01
SIGMA_MAX = 50.0
SAMPLING_EPS = 1e-5
SNR = 0.16  # posenet.py:94
BN_EPS = 1e-5


@dataclass(frozen=True)
class SALevel:
    n_in: int                       # points entering the level
    npoint: Optional[int]           # centroids kept (None = GroupAll)
    c_in: int                       # feature channels entering (without xyz)
    radii: Tuple[Optional[float], Optional[float]]
    nsamples: Tuple[Optional[int], Optional[int]]
    mlps: Tuple[Tuple[int, ...], Tuple[int, ...]]   # channel specs INCLUDING the 3+c_in input

    @property
    def c_out(self) -> int:
        return self.mlps[0][-1] + self.mlps[1][-1]


SA_LEVELS: List[SALevel] = [
    SALevel(1024, 512, 0, (0.02, 0.04), (16, 32), ((3, 16, 16, 32), (3, 32, 32, 64))),
    SALevel(512, 256, 96, (0.04, 0.08), (16, 32), ((99, 64, 64, 128), (99, 64, 96, 128))),
    SALevel(256, 128, 256, (0.08, 0.16), (16, 32), ((259, 128, 196, 256), (259, 128, 196, 256))),
    SALevel(128, None, 512, (None, None), (None, None), ((515, 256, 256, 512), (515, 256, 384, 512))),
]

# Algorithmic work — the numerators of bench.py's roofline.
ENCODER_FLOP_PER_OBJECT = 2.201e9
SCORE_FLOP_PER_EVAL_HOISTED = 2 * 266_752   # 0.5335 MFLOP
SCORE_FLOP_PER_OBJECT_ONCE = 2 * 3 * 1024 * 256


def encoder_macs_per_object() -> int:
    total = 0
    for lv in SA_LEVELS:
        for s in range(2):
            rows = (lv.npoint * lv.nsamples[s]) if lv.npoint is not None else lv.n_in
            spec = lv.mlps[s]
            total += rows * sum(spec[i] * spec[i + 1] for i in range(len(spec) - 1))
    return total
