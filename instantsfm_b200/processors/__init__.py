from .bundle_adjustment import TorchBA  # noqa: F401
from .global_positioning import TorchGP  # noqa: F401
