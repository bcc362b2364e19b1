// Forwarding header: the declarations live in faabric/mpi/mpi_runtime.h
#pragma once

#include <faabric/mpi/mpi_runtime.h>
