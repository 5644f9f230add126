"""The reference's wrappers (tinycarlo/wrapper/__init__.py:1-3) for the single-env drop-in and the vectorised env.
NoiseObservationWrapper (unseeded global np.random in the reference) is not part of the parity scope, see DESIGN.md."""
from .reward import CTELinearRewardWrapper, CTESparseRewardWrapper, LanelineLinearRewardWrapper, LanelineSparseRewardWrapper  # noqa: F401
from .termination import CrashTerminationWrapper, CTETerminationWrapper, LanelineCrossingTerminationWrapper  # noqa: F401
from .utils import linear_reward, sparse_reward  # noqa: F401
