from .bundle_adjustment import TorchBA  # noqa: F401
from .global_positioning import TorchGP  # noqa: F401
from .track_filter import (FilterTracksByAngle, FilterTracksByReprojectionNormalized,  # noqa: F401
                           FilterTracksTriangulationAngle)
