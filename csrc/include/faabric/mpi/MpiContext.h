#pragma once

#include <faabric/mpi/MpiWorld.h>
#include <faabric/proto/faabric.pb.h>

namespace faabric::mpi {

// Per-thread (= per-rank) MPI state used by the C shim
class MpiContext
{
  public:
    MpiContext();

    int createWorld(faabric::Message& msg);

    void joinWorld(faabric::Message& msg);

    bool getIsMpi() const;

    int getRank() const;

    int getWorldId() const;

  private:
    bool isMpi;
    int rank;
    int worldId;
};

}
