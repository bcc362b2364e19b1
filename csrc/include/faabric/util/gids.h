#pragma once

namespace faabric::util {

// Globally unique-ish ids: a per-process random base mixed with host identity
// plus an atomic counter (reference: src/util/gids.cpp:16-35)
unsigned int generateGid();

}
