// Forwarding header: the declarations live in faabric/runner/runner_module.h
#pragma once

#include <faabric/runner/runner_module.h>
