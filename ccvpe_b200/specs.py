"""Per-class constants of the four reference models (reference models.py)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

#: channels of the aerial encoder block outputs used as decoder skips, coarse -> fine
#: (blocks 15, 10, 4, 2, 0; reference models.py:167-171)
SKIP_CHANNELS: Tuple[int, ...] = (320, 112, 40, 24, 16)
SKIP_BLOCKS: Tuple[int, ...] = (15, 10, 4, 2, 0)
ENCODER_CHANNELS = 1280


@dataclass(frozen=True)
class VariantSpec:
    name: str
    grd_feat_h: int                      # height of the ground feature volume (second 1x1 conv reduces it to 1)
    head_channels: Tuple[int, ...]       # c_l of the six ground descriptor heads
    sat_dim: int                         # aerial cell descriptor length D
    n_rolls: int                         # orientations of a full sweep
    roll_strides: Tuple[int, ...]        # channel shift per orientation step at levels 1..6
    centred: bool                        # Oxford: window centred in the channel axis (models.py:1094)
    loc_deconv_out: Tuple[int, ...]      # deconv6..deconv1 output channels
    loc_conv_out: Tuple[int, ...]        # conv6..conv2 output channels (= aerial map channels at levels 2..6)
    ori_deconv_out: Tuple[int, ...]
    ori_conv_out: Tuple[int, ...]

    def level_channels(self) -> Tuple[int, ...]:
        """C_l of the aerial map matched at levels 1..6."""
        return (self.sat_dim,) + tuple(self.loc_conv_out)

    def window_offset(self, C: int, L: int) -> int:
        return int(C / 2 - L / 2) if self.centred else 0

    def window_len(self, C: int, L: int) -> int:
        if self.centred:
            return int(C / 2 + L / 2) - int(C / 2 - L / 2)
        return L


_VIGOR_KW = dict(
    sat_dim=1280, n_rolls=20, roll_strides=(64, 32, 16, 8, 4, 2),
    loc_deconv_out=(1024, 320, 160, 80, 40, 16), loc_conv_out=(640, 320, 160, 80, 40),
    ori_deconv_out=(1024, 256, 128, 64, 32, 16), ori_conv_out=(640, 256, 128, 64, 32),
)

VIGOR = VariantSpec(name="vigor", grd_feat_h=10, head_channels=(64, 32, 16, 8, 4, 2), centred=False, **_VIGOR_KW)
OXFORD = VariantSpec(name="oxford", grd_feat_h=4, head_channels=(32, 16, 8, 4, 2, 1), centred=True, **_VIGOR_KW)
KITTI = VariantSpec(
    name="kitti", grd_feat_h=8, head_channels=(16, 8, 4, 2, 1, 1), sat_dim=2048, n_rolls=16,
    roll_strides=(128, 64, 32, 16, 8, 8), centred=False,
    loc_deconv_out=(1024, 256, 128, 64, 32, 16), loc_conv_out=(512, 256, 128, 128, 32),
    ori_deconv_out=(1024, 256, 128, 64, 32, 16), ori_conv_out=(512, 256, 128, 64, 32),
)


def loc_roll_indices(spec: VariantSpec, ori_noise: Optional[float]) -> List[int]:
    """Orientation indices swept by the localisation branch (models.py:191 / :489 / :793 / :1092)."""
    if ori_noise is None:
        return list(range(spec.n_rolls))
    k = int(ori_noise / 18)
    return list(range(-k, k + 1))
