#pragma once

namespace faabric::mpi {

// Migration point for long-running (MPI or plain) functions: call it at a
// point where no messages are in flight (typically right after a barrier).
//
// Asks the planner - through group idx 0 - whether the app should be
// re-distributed.  If this function must move, its memory is snapshotted and
// pushed to the destination, a MIGRATION request is dispatched there with
// `entrypointArg` as input data (the function resumes from it) and
// FunctionMigratedException unwinds this execution.  If the policy says the
// app must be FROZEN (spot eviction without spare capacity), the snapshot goes
// to the planner and FunctionFrozenException is thrown; the app thaws when
// capacity returns.  Functions that stay put line up with the new group.
//
// (The reference keeps this logic in its distributed tests,
// tests/dist/mpi/mpi_native.cpp:783-913; Faasm has its own copy.)
void mpiMigrationPoint(int entrypointArg);

}
