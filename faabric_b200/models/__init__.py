"""Workloads.  The reference has no ML models; its headline benchmark replays
ResNet-50 gradient tensor sizes through MPI_Allreduce, which is what
:class:`GradientSync` does on GPUs."""
from .grad_sync import GradientSync  # noqa: F401
from .resnet50_grads import resnet50_grad_sizes, small_sizes  # noqa: F401
