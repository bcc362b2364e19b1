// Forwarding header: the declarations live in faabric/executor/executor_module.h
#pragma once

#include <faabric/executor/executor_module.h>
