"""The gradient-synchronisation workload of the reference's own all-reduce
benchmark (tests/dist/mpi/benchmarks/mpi_bench.cpp:25-56 in the reference):
the 214 gradient tensors of a ResNet-50 (TF-slim style: every convolution
carries a kernel plus three per-channel vectors), all-reduced one tensor per
MPI_Allreduce call in reverse layer order, MPI_INT / MPI_SUM.

The list is *generated from the architecture* rather than transcribed."""

from __future__ import annotations


def resnet50_grad_sizes() -> list[int]:
    fwd: list[int] = []

    def conv(k, cin, cout):
        return k * k * cin * cout

    def vecs(c):
        return [c, c, c]

    # stem
    fwd += [conv(7, 3, 64)] + vecs(64)
    cin = 64
    for width, blocks in ((64, 3), (128, 4), (256, 6), (512, 3)):
        cout = width * 4
        for b in range(blocks):
            fwd += [conv(1, cin, width)] + vecs(width)
            fwd += [conv(3, width, width)] + vecs(width)
            fwd += [conv(1, width, cout)]
            if b == 0:
                # projection shortcut
                fwd += [conv(1, cin, cout)] + vecs(cout)
            fwd += vecs(cout)
            cin = cout
    fwd += [2048 * 1000, 1000]
    return list(reversed(fwd))


def small_sizes() -> list[int]:
    """The reference's "small" payload: 1000 messages of 8 ints."""
    return [8] * 1000


TOTAL_ELEMS = 25_583_592
N_TENSORS = 214
