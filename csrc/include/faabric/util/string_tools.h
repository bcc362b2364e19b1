#pragma once

#include <string>
#include <vector>

namespace faabric::util {

bool isAllWhitespace(const std::string& input);

bool startsWith(const std::string& input, const std::string& subStr);

bool endsWith(const std::string& value, const std::string& ending);

bool contains(const std::string& input, const std::string& subStr);

std::string removeSubstr(const std::string& input, const std::string& toErase);

bool stringIsInt(const std::string& input);

std::vector<std::string> splitString(const std::string& input, char delim);

std::string trim(const std::string& input);

std::string toLower(const std::string& input);

}
