// Queue of function messages (reference: include/faabric/scheduler/InMemoryMessageQueue.h)
#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/queue.h>

#include <string>
#include <utility>

namespace faabric::scheduler {

typedef faabric::util::Queue<faabric::Message> InMemoryMessageQueue;

typedef std::pair<std::string, InMemoryMessageQueue*> InMemoryMessageQueuePair;

}
