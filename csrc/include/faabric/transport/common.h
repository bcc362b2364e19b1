#pragma once

#include <string>

// Well-known ports (reference: include/faabric/transport/common.h:9-29).  With
// one worker process per GPU on a box, FAABRIC_PORT_OFFSET shifts the whole
// block so workers do not collide.
#define DEFAULT_STATE_HOST "0.0.0.0"
#define STATE_ASYNC_PORT 8003
#define STATE_SYNC_PORT 8004
#define STATE_INPROC_LABEL "state"

#define DEFAULT_FUNCTION_CALL_HOST "0.0.0.0"
#define FUNCTION_CALL_ASYNC_PORT 8005
#define FUNCTION_CALL_SYNC_PORT 8006
#define FUNCTION_INPROC_LABEL "function"

#define DEFAULT_SNAPSHOT_HOST "0.0.0.0"
#define SNAPSHOT_ASYNC_PORT 8007
#define SNAPSHOT_SYNC_PORT 8008
#define SNAPSHOT_INPROC_LABEL "snapshot"

#define DEFAULT_POINT_TO_POINT_HOST "0.0.0.0"
#define POINT_TO_POINT_ASYNC_PORT 8009
#define POINT_TO_POINT_SYNC_PORT 8010
#define POINT_TO_POINT_INPROC_LABEL "ptp"

#define PLANNER_ASYNC_PORT 8011
#define PLANNER_SYNC_PORT 8012
#define PLANNER_INPROC_LABEL "planner"

#define MPI_BASE_PORT 8020

namespace faabric::transport {

// host may be "ip" or "ip:offset": several workers of one box (one per GPU)
// register with the planner as distinct hosts distinguished by a port offset
struct HostAddress
{
    std::string ip;
    int portOffset = 0;
};

HostAddress parseHostAddress(const std::string& host);

// Virtual host names (e.g. one per GPU of this box) served by another address
void registerHostAlias(const std::string& alias, const std::string& realAddress);

void clearHostAliases();

std::string resolveHostAlias(const std::string& host);

std::string makeHostAddress(const std::string& ip, int portOffset);

// Address other workers use to reach this worker
std::string getThisHostAddress();

}
