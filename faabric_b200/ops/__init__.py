"""Device ops.  Collectives live on :class:`faabric_b200.parallel.Communicator`;
this package holds the snapshot / state kernels' Python wrappers."""
from . import snapshot  # noqa: F401
from . import state  # noqa: F401
