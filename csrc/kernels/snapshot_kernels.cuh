// Snapshot kernels: see snapshot_kernels.cu
#pragma once
#include "fb_prims.cuh"
#include "launch_api.h"
