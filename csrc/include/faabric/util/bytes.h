#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

namespace faabric::util {

std::vector<uint8_t> stringToBytes(const std::string& str);

std::string bytesToString(const std::vector<uint8_t>& bytes);

std::string formatByteArrayToIntString(const std::vector<uint8_t>& bytes);

void trimTrailingZeros(std::vector<uint8_t>& vectorIn);

// Copy a string into a fixed byte buffer, failing if it does not fit
int safeCopyToBuffer(const std::vector<uint8_t>& dataIn,
                     uint8_t* buffer,
                     int bufferLen);

int safeCopyToBuffer(const uint8_t* dataIn,
                     int dataLen,
                     uint8_t* buffer,
                     int bufferLen);

std::string byteArrayToHexString(const uint8_t* data, int dataSize);

std::vector<uint8_t> hexStringToByteArray(const std::string& hexString);

template<typename T>
T unalignedRead(const uint8_t* bytes)
{
    T value;
    std::memcpy(&value, bytes, sizeof(T));
    return value;
}

template<typename T>
void unalignedWrite(const T& value, uint8_t* destination)
{
    std::memcpy(destination, &value, sizeof(T));
}

template<typename T>
std::vector<uint8_t> valueToBytes(T val)
{
    std::vector<uint8_t> out(sizeof(T));
    std::memcpy(out.data(), &val, sizeof(T));
    return out;
}

template<typename T>
size_t appendDataToBytes(std::vector<uint8_t>& bytes, const T& val)
{
    size_t before = bytes.size();
    bytes.resize(before + sizeof(T));
    std::memcpy(bytes.data() + before, &val, sizeof(T));
    return bytes.size();
}

template<typename T>
size_t readBytesOf(const std::vector<uint8_t>& container, size_t offset, T* out)
{
    if (offset + sizeof(T) > container.size()) {
        throw std::range_error("readBytesOf past end of buffer");
    }
    std::memcpy(out, container.data() + offset, sizeof(T));
    return offset + sizeof(T);
}

} // namespace faabric::util
