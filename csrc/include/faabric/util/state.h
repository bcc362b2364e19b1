// Forwarding header: the declarations live in faabric/util/util_module.h
#pragma once

#include <faabric/util/util_module.h>
