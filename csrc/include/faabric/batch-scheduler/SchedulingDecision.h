// Forwarding header: the declarations live in faabric/batch-scheduler/batch_scheduler.h
#pragma once

#include <faabric/batch-scheduler/batch_scheduler.h>
