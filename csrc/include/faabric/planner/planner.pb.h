#pragma once
#include <faabric/proto/faabric.pb.h>
