#pragma once

#include <string>

namespace faabric::util {

std::string getEnvVar(const std::string& key, const std::string& deflt);

std::string setEnvVar(const std::string& varName, const std::string& value);

void unsetEnvVar(const std::string& varName);

// Hardware threads usable by this process (OVERRIDE_CPU_COUNT wins)
unsigned int getUsableCores();

// Number of visible CUDA devices (0 on a CPU-only machine)
int getUsableGpus();

}
