// Forwarding header: the declarations live in faabric/scheduler/scheduler_module.h
#pragma once

#include <faabric/scheduler/scheduler_module.h>
