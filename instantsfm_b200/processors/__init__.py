from .bundle_adjustment import TorchBA  # noqa: F401
from .global_positioning import TorchGP  # noqa: F401
from .image_undistortion import UndistortImages, undistort_process  # noqa: F401
from .track_filter import (FilterTracksByAngle, FilterTracksByReprojection,  # noqa: F401
                           FilterTracksByReprojectionNormalized, FilterTracksTriangulationAngle)
from .track_retriangulation import (RetriangulateTracks, complete_and_merge_tracks, complete_tracks,  # noqa: F401
                                    filter_points)
