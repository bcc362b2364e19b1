#pragma once

#include <cstdint>
#include <vector>

namespace faabric::mpi {

enum MpiMessageType : int32_t
{
    NORMAL = 0,
    BARRIER_JOIN = 1,
    BARRIER_DONE = 2,
    SCATTER = 3,
    GATHER = 4,
    ALLGATHER = 5,
    REDUCE = 6,
    SCAN = 7,
    ALLREDUCE = 8,
    ALLTOALL = 9,
    ALLTOALL_PACKED = 10,
    SENDRECV = 11,
    BROADCAST = 12,
    // Special message types for async messaging
    UNACKED_MPI_MESSAGE = 13,
    HANDSHAKE = 14,
};

// POD descriptor travelling through the per-pair queues / sockets (40 bytes,
// reference: include/faabric/mpi/MpiMessage.h:40-53).  `buffer` points at an
// eagerly copied payload: host memory, or device memory when bufferOnDevice.
struct MpiMessage
{
    int32_t id;
    int32_t worldId;
    int32_t sendRank;
    int32_t recvRank;
    int32_t typeSize;
    int32_t count;
    // For async messaging: id of the request this message satisfies
    int32_t requestId;
    MpiMessageType messageType;
    void* buffer;
};
static_assert(sizeof(MpiMessage) == 40, "MpiMessage must stay a 40-byte POD");

inline size_t payloadSize(const MpiMessage& msg)
{
    return (size_t)msg.typeSize * (size_t)msg.count;
}

inline size_t msgSize(const MpiMessage& msg)
{
    return sizeof(MpiMessage) + payloadSize(msg);
}

// Header followed by the payload bytes
void serializeMpiMsg(std::vector<uint8_t>& buffer, const MpiMessage& msg);

// Allocates msg.buffer with malloc (caller frees)
void parseMpiMsg(const std::vector<uint8_t>& bytes, MpiMessage* msg);

}
