// Forwarding header: the declarations live in faabric/snapshot/snapshot_module.h
#pragma once

#include <faabric/snapshot/snapshot_module.h>
