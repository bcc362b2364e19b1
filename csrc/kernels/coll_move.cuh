// Data-movement collectives: see coll_move.cu
#pragma once
#include "fb_prims.cuh"
#include "launch_api.h"
