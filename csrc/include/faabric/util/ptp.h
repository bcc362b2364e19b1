#pragma once

#include <faabric/proto/faabric.pb.h>

namespace faabric::batch_scheduler {
class SchedulingDecision;
}

namespace faabric::util {

// Unlike the reference (src/util/ptp.cpp:4-19) the MPI port / mailbox slot of
// every mapping is carried across.
faabric::PointToPointMappings ptpMappingsFromSchedulingDecision(
  std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> decision);

}
