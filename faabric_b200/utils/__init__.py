from .clocks import ClockSampler  # noqa: F401

from .numa import bind_process_near_gpu, gpu_local_cpus  # noqa: E402,F401
