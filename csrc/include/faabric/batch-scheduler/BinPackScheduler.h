#pragma once
#include <faabric/batch-scheduler/BatchScheduler.h>
