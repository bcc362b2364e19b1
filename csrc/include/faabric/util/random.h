#pragma once

#include <string>

namespace faabric::util {

std::string randomString(int len);

std::string randomStringFromSet(int len, const std::string& charSet);

int randomInteger(int iStart, int iEnd);

}
