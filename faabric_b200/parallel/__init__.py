from .comm import (  # noqa: F401
    ALGOS,
    OPS,
    CommError,
    Communicator,
    LocalGroup,
    init_from_env,
    make_config,
)
