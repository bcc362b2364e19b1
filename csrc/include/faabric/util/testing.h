#pragma once

namespace faabric::util {

// Test mode relaxes some checks; mock mode makes every RPC client record its
// calls instead of opening sockets (reference: src/util/testing.cpp:6-26)
void setTestMode(bool val);

bool isTestMode();

void setMockMode(bool val);

bool isMockMode();

}
