// Forwarding header: the declarations live in faabric/transport/transport_module.h
#pragma once

#include <faabric/transport/transport_module.h>
