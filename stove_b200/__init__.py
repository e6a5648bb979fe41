"""stove_b200 -- B200-native (sm_100a) implementation of STOVE's data-parallel hot path.

Drop-in for the reference's module API:
    from stove_b200.video_prediction.stove import Stove          # model/video_prediction/stove.py
    from stove_b200.video_prediction.config import StoveConfig   # model/video_prediction/config.py
    from stove_b200.spn.rat_torch import RatSpn                  # model/spn/rat_torch.py
"""
from .video_prediction.config import StoveConfig          # noqa: F401
from .video_prediction.stove import Stove                 # noqa: F401
from .video_prediction.supair import Supair               # noqa: F401
from .video_prediction.dynamics import Dynamics           # noqa: F401
from .spn.rat_torch import RatSpn, SpnArgs                # noqa: F401
from .spn.region_graph import RegionGraph                 # noqa: F401
