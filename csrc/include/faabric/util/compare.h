#pragma once

#include <cstddef>

namespace faabric::util {

// Element-wise equality of two arrays (reference: include/faabric/util/compare.h)
template<typename T>
bool compareArrays(const T* v1, const T* v2, size_t size)
{
    for (size_t i = 0; i < size; i++) {
        if (!(v1[i] == v2[i])) {
            return false;
        }
    }
    return true;
}

}
