// User-defined reduction operations and derived datatypes: process-wide
// registries behind MPI_Op_create / MPI_Type_contiguous.
#include <faabric/mpi/MpiWorld.h>

#include <map>
#include <shared_mutex>
#include <stdexcept>

namespace faabric::mpi {

// ---------------------------------------------------------------------------
// User-defined operations
// ---------------------------------------------------------------------------
namespace {
struct UserOp
{
    MPI_User_function* fn;
    bool commutes;
};
std::shared_mutex userOpsMx;
std::map<int, UserOp> userOps;
int nextUserOpId = FAABRIC_OP_USER_BASE;

}

bool isOrderedUserOp(const faabric_op_t* op)
{
    MPI_User_function* fn = nullptr;
    bool commutes = true;
    return isUserOp(op) && getUserOp(op->id, &fn, &commutes) && !commutes;
}

int registerUserOp(MPI_User_function* fn, bool commutes)
{
    if (fn == nullptr) {
        throw std::invalid_argument("Null user function for MPI_Op_create");
    }
    std::unique_lock<std::shared_mutex> lk(userOpsMx);
    int id = nextUserOpId++;
    userOps[id] = UserOp{ fn, commutes };
    return id;
}

bool unregisterUserOp(int opId)
{
    std::unique_lock<std::shared_mutex> lk(userOpsMx);
    return userOps.erase(opId) > 0;
}

bool getUserOp(int opId, MPI_User_function** fn, bool* commutes)
{
    std::shared_lock<std::shared_mutex> lk(userOpsMx);
    auto it = userOps.find(opId);
    if (it == userOps.end()) {
        return false;
    }
    *fn = it->second.fn;
    *commutes = it->second.commutes;
    return true;
}

// ---------------------------------------------------------------------------
// Derived (contiguous) datatypes
// ---------------------------------------------------------------------------
namespace {
std::shared_mutex derivedTypesMx;
std::map<int, std::pair<int, int>> derivedTypes;
int nextDerivedTypeId = FAABRIC_DERIVED_TYPE_BASE;
}

int registerContiguousType(int baseTypeId, int count)
{
    if (count <= 0) {
        throw std::invalid_argument("Contiguous datatype needs a positive count");
    }
    std::unique_lock<std::shared_mutex> lk(derivedTypesMx);
    int id = nextDerivedTypeId++;
    derivedTypes[id] = { baseTypeId, count };
    return id;
}

bool getContiguousType(int typeId, int* baseTypeId, int* count)
{
    std::shared_lock<std::shared_mutex> lk(derivedTypesMx);
    auto it = derivedTypes.find(typeId);
    if (it == derivedTypes.end()) {
        return false;
    }
    *baseTypeId = it->second.first;
    *count = it->second.second;
    return true;
}

bool unregisterContiguousType(int typeId)
{
    std::unique_lock<std::shared_mutex> lk(derivedTypesMx);
    return derivedTypes.erase(typeId) > 0;
}

}
