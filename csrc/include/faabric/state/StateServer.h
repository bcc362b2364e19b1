// Forwarding header: the declarations live in faabric/state/state_module.h
#pragma once

#include <faabric/state/state_module.h>
