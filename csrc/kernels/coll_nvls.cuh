// NVLS collectives: see coll_nvls.cu
#pragma once
#include "fb_prims.cuh"
#include "launch_api.h"
