// Smoke check that the library links and basic message plumbing works
// (reference: examples/check.cpp)
#include <faabric/util/func.h>
#include <faabric/util/logging.h>

#include <cstdio>

int main()
{
    faabric::util::initLogging();
    faabric::Message msg = faabric::util::messageFactory("foo", "bar");
    std::string msgString = faabric::util::funcToString(msg, true);
    printf("Message: %s\n", msgString.c_str());
    return msgString.empty() ? 1 : 0;
}
