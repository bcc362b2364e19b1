#pragma once

#include <string>

#define LOCALHOST "127.0.0.1"

namespace faabric::util {

std::string getIPFromHostname(const std::string& hostname);

std::string getPrimaryIPForThisHost(const std::string& interface);

// "gpu3" style host alias used when GPUs are registered as planner hosts
std::string gpuHostName(int gpuIdx);

// -1 if `host` is not a gpu alias
int gpuIndexFromHostName(const std::string& host);

}
