// Forwarding header: the declarations live in faabric/planner/planner_module.h
#pragma once

#include <faabric/planner/planner_module.h>
