// MPI tests: every rank runs as a real function scheduled by the planner and
// talks through the C API (strategy: reference tests/test/mpi/*.cpp and the
// dist-test functions in tests/dist/mpi/examples).
#include "fixtures.h"

#include <faabric/mpi/MpiWorldRegistry.h>
#include <faabric/mpi/mpi.h>

#include <numeric>

using namespace tests;

namespace {
typedef std::function<int(int rank, int size)> RankBody;

#define CHECK_RANK(cond)                                                       \
    do {                                                                       \
        if (!(cond)) {                                                         \
            printf("         rank %d: check failed at line %d: %s\n", rank, __LINE__, #cond); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

// Runs `body` on `worldSize` ranks spread over `nHosts` virtual hosts
void runMpi(const std::string& name, int worldSize, int nHosts, const RankBody& body)
{
    // ceil(worldSize / nHosts) slots per host so the world spans all hosts
    int perHost = (worldSize + nHosts - 1) / nHosts;
    ClusterFixture f(nHosts == 1 ? worldSize : 0, nHosts == 1 ? 0 : nHosts, perHost);
    registerTestFunction("mpi", name, [&](auto*, int, int, auto) {
        MPI_Init(nullptr, nullptr);
        int rank = -1, size = -1;
        MPI_Comm_rank(MPI_COMM_WORLD, &rank);
        MPI_Comm_size(MPI_COMM_WORLD, &size);
        int rc = body(rank, size);
        MPI_Finalize();
        return rc;
    });
    auto req = faabric::util::batchExecFactory("mpi", name, 1);
    auto& msg = *req->mutable_messages(0);
    msg.set_ismpi(true);
    msg.set_mpiworldsize(worldSize);
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.nFunctions, 1);
    auto status = f.awaitBatch(req, 30000);
    REQUIRE_EQ(status->messageresults_size(), worldSize);
    std::set<int> ranks;
    std::set<std::string> hosts;
    for (auto& m : status->messageresults()) {
        if (m.returnvalue() != 0) {
            fbtest::fail(__FILE__, __LINE__, name + ": rank " + std::to_string(m.mpirank()) + " failed: " + m.outputdata());
        }
        ranks.insert(m.mpirank());
        hosts.insert(m.executedhost());
    }
    REQUIRE_EQ((int)ranks.size(), worldSize);
    REQUIRE_EQ((int)hosts.size(), nHosts);
    faabric::mpi::getMpiWorldRegistry().clear();
}

int bodyPointToPoint(int rank, int size)
{
    MPI_Status status{};
    // Ring with blocking send/recv (even/odd ordering)
    int right = (rank + 1) % size, left = (rank + size - 1) % size;
    int token = rank * 10, got = -1;
    if (rank % 2 == 0) {
        MPI_Send(&token, 1, MPI_INT, right, 0, MPI_COMM_WORLD);
        MPI_Recv(&got, 1, MPI_INT, left, 0, MPI_COMM_WORLD, &status);
    } else {
        MPI_Recv(&got, 1, MPI_INT, left, 0, MPI_COMM_WORLD, &status);
        MPI_Send(&token, 1, MPI_INT, right, 0, MPI_COMM_WORLD);
    }
    CHECK_RANK(got == left * 10);
    CHECK_RANK(status.MPI_SOURCE == left);
    int count = 0;
    MPI_Get_count(&status, MPI_INT, &count);
    CHECK_RANK(count == 1);

    // Sendrecv ring with bigger payloads
    std::vector<double> out(1000, rank + 0.5), in(1000, -1);
    MPI_Sendrecv(out.data(), 1000, MPI_DOUBLE, right, 0, in.data(), 1000, MPI_DOUBLE, left, 0, MPI_COMM_WORLD, &status);
    CHECK_RANK(in[0] == left + 0.5 && in[999] == left + 0.5);

    // Non-blocking all-pairs exchange
    std::vector<int> sendVals(size), recvVals(size, -1);
    std::vector<MPI_Request> reqs;
    for (int r = 0; r < size; r++) {
        if (r == rank) {
            continue;
        }
        sendVals[r] = rank * 100 + r;
        MPI_Request rq;
        MPI_Irecv(&recvVals[r], 1, MPI_INT, r, 0, MPI_COMM_WORLD, &rq);
        reqs.push_back(rq);
    }
    for (int r = 0; r < size; r++) {
        if (r == rank) {
            continue;
        }
        MPI_Request rq;
        MPI_Isend(&sendVals[r], 1, MPI_INT, r, 0, MPI_COMM_WORLD, &rq);
        reqs.push_back(rq);
    }
    MPI_Waitall((int)reqs.size(), reqs.data(), MPI_STATUSES_IGNORE);
    for (int r = 0; r < size; r++) {
        if (r != rank) {
            CHECK_RANK(recvVals[r] == r * 100 + rank);
        }
    }

    // Probe reports the pending message's size
    if (rank == 0) {
        std::vector<long> big(77, 5);
        MPI_Send(big.data(), 77, MPI_LONG, 1, 0, MPI_COMM_WORLD);
    } else if (rank == 1) {
        MPI_Probe(0, 0, MPI_COMM_WORLD, &status);
        MPI_Get_count(&status, MPI_LONG, &count);
        CHECK_RANK(count == 77);
        std::vector<long> big(77, 0);
        MPI_Recv(big.data(), 77, MPI_LONG, 0, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        CHECK_RANK(big[76] == 5);
    }
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}

int bodyCollectives(int rank, int size)
{
    // Broadcast from a non-zero root
    int root = size - 1;
    std::vector<int> bc(300, rank == root ? 7 : 0);
    MPI_Bcast(bc.data(), 300, MPI_INT, root, MPI_COMM_WORLD);
    CHECK_RANK(bc[0] == 7 && bc[299] == 7);

    // Scatter / gather
    const int per = 4;
    std::vector<int> all(per * size), mine(per, -1);
    if (rank == 1) {
        std::iota(all.begin(), all.end(), 0);
    }
    MPI_Scatter(all.data(), per, MPI_INT, mine.data(), per, MPI_INT, 1, MPI_COMM_WORLD);
    for (int i = 0; i < per; i++) {
        CHECK_RANK(mine[i] == rank * per + i);
        mine[i] *= 2;
    }
    std::vector<int> gathered(per * size, -1);
    MPI_Gather(mine.data(), per, MPI_INT, gathered.data(), per, MPI_INT, 0, MPI_COMM_WORLD);
    if (rank == 0) {
        for (int i = 0; i < per * size; i++) {
            CHECK_RANK(gathered[i] == 2 * i);
        }
    }

    // Allgather
    std::vector<int> ag(2 * size, -1);
    int pair[2] = { rank, rank * rank };
    MPI_Allgather(pair, 2, MPI_INT, ag.data(), 2, MPI_INT, MPI_COMM_WORLD);
    for (int r = 0; r < size; r++) {
        CHECK_RANK(ag[2 * r] == r && ag[2 * r + 1] == r * r);
    }

    // Reduce + Allreduce over several types and ops
    {
        std::vector<int> v(50, rank + 1), res(50, 0);
        MPI_Reduce(v.data(), res.data(), 50, MPI_INT, MPI_SUM, 2 % size, MPI_COMM_WORLD);
        if (rank == 2 % size) {
            CHECK_RANK(res[49] == size * (size + 1) / 2);
        }
        MPI_Allreduce(v.data(), res.data(), 50, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
        CHECK_RANK(res[0] == size);
        MPI_Allreduce(v.data(), res.data(), 50, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
        CHECK_RANK(res[0] == 1);
        // In place
        MPI_Allreduce(MPI_IN_PLACE, v.data(), 50, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
        CHECK_RANK(v[10] == size * (size + 1) / 2);
    }
    {
        std::vector<double> v(33, 0.5 * (rank + 1)), res(33, 0);
        MPI_Allreduce(v.data(), res.data(), 33, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
        CHECK_RANK(std::fabs(res[32] - 0.25 * size * (size + 1)) < 1e-9);
        std::vector<long long> lv(5, 2), lres(5, 0);
        MPI_Allreduce(lv.data(), lres.data(), 5, MPI_LONG_LONG, MPI_PROD, MPI_COMM_WORLD);
        CHECK_RANK(lres[0] == (1LL << size));
        std::vector<float> fv(9, (float)rank), fres(9, 0);
        MPI_Allreduce(fv.data(), fres.data(), 9, MPI_FLOAT, MPI_MAX, MPI_COMM_WORLD);
        CHECK_RANK(fres[8] == (float)(size - 1));
    }

    // Scan (inclusive prefix sum)
    int sv = rank + 1, sres = 0;
    MPI_Scan(&sv, &sres, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    CHECK_RANK(sres == (rank + 1) * (rank + 2) / 2);

    // Alltoall
    std::vector<int> a2aSend(2 * size), a2aRecv(2 * size, -1);
    for (int r = 0; r < size; r++) {
        a2aSend[2 * r] = rank * 1000 + r;
        a2aSend[2 * r + 1] = -(rank * 1000 + r);
    }
    MPI_Alltoall(a2aSend.data(), 2, MPI_INT, a2aRecv.data(), 2, MPI_INT, MPI_COMM_WORLD);
    for (int r = 0; r < size; r++) {
        CHECK_RANK(a2aRecv[2 * r] == r * 1000 + rank);
        CHECK_RANK(a2aRecv[2 * r + 1] == -(r * 1000 + rank));
    }

    // Reduce_scatter
    std::vector<int> rsIn(3 * size), rsOut(3, 0), counts(size, 3);
    for (int i = 0; i < 3 * size; i++) {
        rsIn[i] = i + rank;
    }
    MPI_Reduce_scatter(rsIn.data(), rsOut.data(), counts.data(), MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    for (int i = 0; i < 3; i++) {
        int idx = rank * 3 + i;
        CHECK_RANK(rsOut[i] == idx * size + size * (size - 1) / 2);
    }
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}

int bodyMisc(int rank, int size)
{
    int flag = 0;
    MPI_Initialized(&flag);
    CHECK_RANK(flag == 1);
    int tsize = 0;
    MPI_Type_size(MPI_DOUBLE, &tsize);
    CHECK_RANK(tsize == 8);
    MPI_Type_size(MPI_LONG_LONG, &tsize);
    CHECK_RANK(tsize == 8);
    MPI_Type_size(MPI_CHAR, &tsize);
    CHECK_RANK(tsize == 1);
    double t0 = MPI_Wtime();
    CHECK_RANK(t0 >= 0);
    char name[MPI_MAX_PROCESSOR_NAME];
    int len = 0;
    MPI_Get_processor_name(name, &len);
    CHECK_RANK(len > 0);

    // 2D periodic cartesian grid
    int side = 1;
    while ((side + 1) * (side + 1) <= size) {
        side++;
    }
    if (side * side == size) {
        int dims[2] = { side, side }, periods[2] = { 1, 1 };
        MPI_Comm cart;
        CHECK_RANK(MPI_Cart_create(MPI_COMM_WORLD, 2, dims, periods, 0, &cart) == MPI_SUCCESS);
        int coords[2], gotDims[2], gotPeriods[2];
        MPI_Cart_get(cart, 2, gotDims, gotPeriods, coords);
        CHECK_RANK(coords[0] == rank / side && coords[1] == rank % side);
        int back = -1;
        MPI_Cart_rank(cart, coords, &back);
        CHECK_RANK(back == rank);
        int src = -1, dst = -1;
        MPI_Cart_shift(cart, 1, 1, &src, &dst);
        CHECK_RANK(dst == (rank / side) * side + (rank % side + 1) % side);
        CHECK_RANK(src == (rank / side) * side + (rank % side + side - 1) % side);
        MPI_Cart_shift(cart, 0, 1, &src, &dst);
        CHECK_RANK(dst == ((rank / side + 1) % side) * side + rank % side);
    }

    // User-allocated memory and gatherv
    int* mem = nullptr;
    MPI_Alloc_mem(64 * sizeof(int), MPI_INFO_NULL, &mem);
    CHECK_RANK(mem != nullptr);
    mem[63] = rank;
    std::vector<int> recvCounts(size), displs(size);
    int total = 0;
    for (int r = 0; r < size; r++) {
        recvCounts[r] = r + 1;
        displs[r] = total;
        total += r + 1;
    }
    std::vector<int> mineV(rank + 1, rank), allV(total, -1);
    MPI_Allgatherv(mineV.data(), rank + 1, MPI_INT, allV.data(), recvCounts.data(), displs.data(), MPI_INT, MPI_COMM_WORLD);
    for (int r = 0; r < size; r++) {
        for (int i = 0; i < r + 1; i++) {
            CHECK_RANK(allV[displs[r] + i] == r);
        }
    }
    MPI_Free_mem(mem);
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}

// The BASELINE workloads: ping-pong and a burst of int32 all-reduces
int bodyPingPongAndAllreduce(int rank, int size)
{
    std::vector<uint8_t> buf(8192, (uint8_t)rank);
    for (int i = 0; i < 200; i++) {
        if (rank == 0) {
            MPI_Send(buf.data(), 8192, MPI_BYTE, 1, 0, MPI_COMM_WORLD);
            MPI_Recv(buf.data(), 8192, MPI_BYTE, 1, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        } else if (rank == 1) {
            MPI_Recv(buf.data(), 8192, MPI_BYTE, 0, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
            MPI_Send(buf.data(), 8192, MPI_BYTE, 0, 0, MPI_COMM_WORLD);
        }
    }
    std::vector<int> grad(100000, rank + 1), sum(100000, 0);
    for (int i = 0; i < 20; i++) {
        MPI_Allreduce(grad.data(), sum.data(), 100000, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
        CHECK_RANK(sum[0] == size * (size + 1) / 2 && sum[99999] == sum[0]);
    }
    return 0;
}
}

namespace {
// Large host all-reduces go through the shared-memory slice-parallel path
int bodyLargeHostAllreduce(int rank, int size)
{
    // count not divisible by the world size, several types / ops, in place
    const int n = 100003;
    std::vector<double> d(n), dOut(n, 0);
    for (int i = 0; i < n; i++) {
        d[i] = (double)(i % 97) + rank;
    }
    MPI_Allreduce(d.data(), dOut.data(), n, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    for (int i : { 0, 1, n / 2, n - 2, n - 1 }) {
        CHECK_RANK(dOut[i] == (double)(i % 97) * size + size * (size - 1) / 2.0);
    }
    std::vector<float> f(n, (float)rank);
    f[n - 1] = (float)(100 - rank);
    MPI_Allreduce(MPI_IN_PLACE, f.data(), n, MPI_FLOAT, MPI_MAX, MPI_COMM_WORLD);
    CHECK_RANK(f[0] == (float)(size - 1) && f[n - 1] == 100.0f);
    std::vector<long long> l(n, rank + 1), lOut(n, 0);
    MPI_Allreduce(l.data(), lOut.data(), n, MPI_LONG_LONG, MPI_MIN, MPI_COMM_WORLD);
    CHECK_RANK(lOut[12345] == 1);
    std::vector<int> v(40000, 1), vOut(40000, 0);
    for (int round = 0; round < 5; round++) {
        MPI_Allreduce(v.data(), vOut.data(), 40000, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
        CHECK_RANK(vOut[39999] == size);
        v.swap(vOut);
        std::fill(v.begin(), v.end(), 1);
    }
    // fewer elements than ranks per slice boundary cases
    std::vector<int> tiny(8200, rank), tinyOut(8200, -1);
    MPI_Allreduce(tiny.data(), tinyOut.data(), 8200, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    CHECK_RANK(tinyOut[8199] == size * (size - 1) / 2);
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}
}

TEST_CASE("mpi: large host all-reduces use the shared-memory path", "[mpi]")
{
    runMpi("large-host-allreduce", 5, 1, bodyLargeHostAllreduce);
    runMpi("large-host-allreduce-gpuhosts", 6, 3, bodyLargeHostAllreduce);
    // ...and agree with the reference algorithm
    setenv("FAABRIC_MPI_HOST_ALLREDUCE", "reference", 1);
    runMpi("large-host-allreduce-ref", 3, 1, bodyLargeHostAllreduce);
    unsetenv("FAABRIC_MPI_HOST_ALLREDUCE");
}

namespace {
// Large host broadcast / reduce / allgather / alltoall: direct copies between
// the user buffers when every rank lives in this process
int bodyLargeHostCollectives(int rank, int size)
{
    // (above the 256 KiB break-even used when ranks outnumber cores)
    const int n = 80001;
    // broadcast from a non-zero root
    const int root = size - 1;
    std::vector<int> b(n, rank == root ? 7 : -1);
    if (rank == root) {
        b[n - 1] = 99;
    }
    MPI_Bcast(b.data(), n, MPI_INT, root, MPI_COMM_WORLD);
    CHECK_RANK(b[0] == 7 && b[n / 2] == 7 && b[n - 1] == 99);

    // reduce to a non-zero root, out of place then in place
    std::vector<double> d(n), dOut(rank == 1 ? n : 0);
    for (int i = 0; i < n; i++) {
        d[i] = (double)(i % 13) + rank;
    }
    MPI_Reduce(d.data(), dOut.data(), n, MPI_DOUBLE, MPI_SUM, 1, MPI_COMM_WORLD);
    if (rank == 1) {
        for (int i : { 0, 5, n / 2, n - 1 }) {
            CHECK_RANK(dOut[i] == (double)(i % 13) * size + size * (size - 1) / 2.0);
        }
    }
    std::vector<int> m(n, rank);
    m[3] = 100 - rank;
    if (rank == 0) {
        MPI_Reduce(MPI_IN_PLACE, m.data(), n, MPI_INT, MPI_MAX, 0, MPI_COMM_WORLD);
        CHECK_RANK(m[0] == size - 1 && m[3] == 100 && m[n - 1] == size - 1);
    } else {
        MPI_Reduce(m.data(), nullptr, n, MPI_INT, MPI_MAX, 0, MPI_COMM_WORLD);
        CHECK_RANK(m[0] == rank && m[3] == 100 - rank);
    }

    // allgather, out of place and in place
    const int per = 20001;
    std::vector<int> mine(per, rank + 1), all((size_t)per * size, 0);
    MPI_Allgather(mine.data(), per, MPI_INT, all.data(), per, MPI_INT, MPI_COMM_WORLD);
    for (int r = 0; r < size; r++) {
        CHECK_RANK(all[(size_t)r * per] == r + 1 && all[(size_t)r * per + per - 1] == r + 1);
    }
    std::vector<int> inPlace((size_t)per * size, -1);
    std::fill(inPlace.begin() + (size_t)rank * per, inPlace.begin() + (size_t)(rank + 1) * per, 10 * rank);
    MPI_Allgather(MPI_IN_PLACE, 0, MPI_DATATYPE_NULL, inPlace.data(), per, MPI_INT, MPI_COMM_WORLD);
    for (int r = 0; r < size; r++) {
        CHECK_RANK(inPlace[(size_t)r * per + 17] == 10 * r);
    }

    // gather to / scatter from a non-zero root, plus the in-place forms
    const int gRoot = size / 2;
    std::vector<int> gathered(rank == gRoot ? (size_t)per * size : 0, -1);
    MPI_Gather(mine.data(), per, MPI_INT, gathered.data(), per, MPI_INT, gRoot, MPI_COMM_WORLD);
    if (rank == gRoot) {
        for (int r = 0; r < size; r++) {
            CHECK_RANK(gathered[(size_t)r * per] == r + 1 && gathered[(size_t)r * per + per - 1] == r + 1);
        }
        std::fill(gathered.begin(), gathered.end(), -1);
        std::fill(gathered.begin() + (size_t)rank * per, gathered.begin() + (size_t)(rank + 1) * per, rank + 1);
        MPI_Gather(MPI_IN_PLACE, 0, MPI_DATATYPE_NULL, gathered.data(), per, MPI_INT, gRoot, MPI_COMM_WORLD);
        for (int r = 0; r < size; r++) {
            CHECK_RANK(gathered[(size_t)r * per + 3] == r + 1);
        }
    } else {
        MPI_Gather(mine.data(), per, MPI_INT, nullptr, 0, MPI_DATATYPE_NULL, gRoot, MPI_COMM_WORLD);
    }
    std::vector<int> toScatter(rank == gRoot ? (size_t)per * size : 0), piece(per, -1);
    for (size_t i = 0; i < toScatter.size(); i++) {
        toScatter[i] = (int)(i / per) * 7;
    }
    MPI_Scatter(toScatter.data(), per, MPI_INT, piece.data(), per, MPI_INT, gRoot, MPI_COMM_WORLD);
    CHECK_RANK(piece[0] == rank * 7 && piece[per - 1] == rank * 7);

    // alltoall: chunk for rank r carries (me, r)
    std::vector<int> out((size_t)per * size), in((size_t)per * size, -1);
    for (int r = 0; r < size; r++) {
        std::fill(out.begin() + (size_t)r * per, out.begin() + (size_t)(r + 1) * per, rank * 100 + r);
    }
    for (int round = 0; round < 3; round++) {
        MPI_Alltoall(out.data(), per, MPI_INT, in.data(), per, MPI_INT, MPI_COMM_WORLD);
        for (int r = 0; r < size; r++) {
            CHECK_RANK(in[(size_t)r * per] == r * 100 + rank && in[(size_t)r * per + per - 1] == r * 100 + rank);
        }
        std::fill(in.begin(), in.end(), -1);
    }
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}
}

TEST_CASE("mpi: large host collectives copy between user buffers", "[mpi]")
{
    runMpi("large-host-coll", 5, 1, bodyLargeHostCollectives);
    runMpi("large-host-coll-gpuhosts", 6, 3, bodyLargeHostCollectives);
    setenv("FAABRIC_MPI_HOST_ALLREDUCE", "reference", 1);
    runMpi("large-host-coll-ref", 3, 1, bodyLargeHostCollectives);
    unsetenv("FAABRIC_MPI_HOST_ALLREDUCE");
}

namespace {
// ---- user-defined operations ----
// commutative: product modulo a prime
void mulModFn(void* in, void* inout, int* len, MPI_Datatype*)
{
    auto* a = (long*)in;
    auto* b = (long*)inout;
    for (int i = 0; i < *len; i++) {
        b[i] = (a[i] * b[i]) % 1000003L;
    }
}

// associative but NOT commutative: composition of affine maps x -> a*x + b,
// stored as consecutive (a, b) int pairs; inout = in o inout
void composeFn(void* in, void* inout, int* len, MPI_Datatype*)
{
    auto* f = (int*)in;
    auto* g = (int*)inout;
    for (int i = 0; i + 1 < *len; i += 2) {
        int a = f[i], b = f[i + 1], c = g[i], d = g[i + 1];
        g[i] = a * c;
        g[i + 1] = a * d + b;
    }
}

int bodyUserOps(int rank, int size)
{
    MPI_Op mulMod = nullptr, compose = nullptr;
    CHECK_RANK(MPI_Op_create(mulModFn, 1, &mulMod) == MPI_SUCCESS);
    CHECK_RANK(MPI_Op_create(composeFn, 0, &compose) == MPI_SUCCESS);
    CHECK_RANK(mulMod != nullptr && compose != nullptr && mulMod->id != compose->id);

    long expectedProd = 1;
    for (int r = 0; r < size; r++) {
        expectedProd = (expectedProd * (r + 2)) % 1000003L;
    }
    // rank-ordered composition f0 o f1 o ... with f_r = ((r % 3) + 1, r + 1)
    std::vector<std::pair<int, int>> prefix(size);
    int ea = 1, eb = 0;
    for (int r = 0; r < size; r++) {
        int c = (r % 3) + 1, d = r + 1;
        eb = ea * d + eb;
        ea = ea * c;
        prefix[r] = { ea, eb };
    }

    // small and large (>= 32 KiB) messages take different host paths
    for (int n : { 6, 10000 }) {
        std::vector<long> mine(n, rank + 2), out(n, 0);
        MPI_Reduce(mine.data(), out.data(), n, MPI_LONG, mulMod, size - 1, MPI_COMM_WORLD);
        if (rank == size - 1) {
            CHECK_RANK(out[0] == expectedProd && out[n - 1] == expectedProd);
        }
        MPI_Allreduce(mine.data(), out.data(), n, MPI_LONG, mulMod, MPI_COMM_WORLD);
        CHECK_RANK(out[0] == expectedProd && out[n / 2] == expectedProd);
        CHECK_RANK(mine[0] == rank + 2);
        MPI_Allreduce(MPI_IN_PLACE, mine.data(), n, MPI_LONG, mulMod, MPI_COMM_WORLD);
        CHECK_RANK(mine[n - 1] == expectedProd);

        std::vector<int> f(2 * n), g(2 * n, -1);
        for (int i = 0; i < n; i++) {
            f[2 * i] = (rank % 3) + 1;
            f[2 * i + 1] = rank + 1;
        }
        MPI_Reduce(f.data(), g.data(), 2 * n, MPI_INT, compose, 1 % size, MPI_COMM_WORLD);
        if (rank == 1 % size) {
            CHECK_RANK(g[0] == ea && g[1] == eb);
            CHECK_RANK(g[2 * n - 2] == ea && g[2 * n - 1] == eb);
        }
        std::fill(g.begin(), g.end(), -1);
        MPI_Allreduce(f.data(), g.data(), 2 * n, MPI_INT, compose, MPI_COMM_WORLD);
        CHECK_RANK(g[0] == ea && g[1] == eb && g[2 * n - 1] == eb);
        // in place at the root
        std::vector<int> h = f;
        if (rank == 0) {
            MPI_Reduce(MPI_IN_PLACE, h.data(), 2 * n, MPI_INT, compose, 0, MPI_COMM_WORLD);
            CHECK_RANK(h[0] == ea && h[1] == eb);
        } else {
            MPI_Reduce(h.data(), nullptr, 2 * n, MPI_INT, compose, 0, MPI_COMM_WORLD);
        }
        // inclusive prefix in rank order
        std::fill(g.begin(), g.end(), -1);
        MPI_Scan(f.data(), g.data(), 2 * n, MPI_INT, compose, MPI_COMM_WORLD);
        CHECK_RANK(g[0] == prefix[rank].first && g[1] == prefix[rank].second);
    }

    CHECK_RANK(MPI_Op_free(&mulMod) == MPI_SUCCESS && mulMod == MPI_OP_NULL);
    CHECK_RANK(MPI_Op_free(&compose) == MPI_SUCCESS);
    MPI_Op predefined = MPI_SUM;
    CHECK_RANK(MPI_Op_free(&predefined) == MPI_ERR_OP);
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}
}

TEST_CASE("mpi: user-defined operations, commutative and rank-ordered", "[mpi]")
{
    runMpi("user-ops", 5, 1, bodyUserOps);
    runMpi("user-ops-gpuhosts", 6, 3, bodyUserOps);
    setenv("FAABRIC_MPI_HOST_ALLREDUCE", "reference", 1);
    runMpi("user-ops-ref", 4, 1, bodyUserOps);
    unsetenv("FAABRIC_MPI_HOST_ALLREDUCE");
}

namespace {
// ---- one-sided communication ----
int bodyRma(int rank, int size)
{
    const int right = (rank + 1) % size, left = (rank + size - 1) % size;
    // Each rank exposes `size` slots of doubles plus a tail
    const int n = size + 4;
    std::vector<double> window(n, -1.0);
    MPI_Win win = nullptr;
    CHECK_RANK(MPI_Win_create(window.data(), n * sizeof(double), sizeof(double), MPI_INFO_NULL, MPI_COMM_WORLD, &win) == MPI_SUCCESS);
    void* base = nullptr;
    int flag = 0;
    MPI_Win_get_attr(win, MPI_WIN_BASE, &base, &flag);
    CHECK_RANK(flag == 1 && base == window.data());
    MPI_Aint winSize = 0;
    MPI_Win_get_attr(win, MPI_WIN_SIZE, &winSize, &flag);
    CHECK_RANK(winSize == (MPI_Aint)(n * sizeof(double)));

    MPI_Win_fence(0, win);
    // everybody deposits its signature in slot [rank] of EVERY window
    double mine = 100.0 + rank;
    for (int t = 0; t < size; t++) {
        MPI_Put(&mine, 1, MPI_DOUBLE, t, rank, 1, MPI_DOUBLE, win);
    }
    MPI_Win_fence(0, win);
    for (int r = 0; r < size; r++) {
        CHECK_RANK(window[r] == 100.0 + r);
    }
    CHECK_RANK(window[size] == -1.0);

    // get a strip from the right neighbour, put two elements into the left's tail
    window[size + 1] = 1000.0 + rank;
    MPI_Win_fence(0, win);
    std::vector<double> strip(3, 0.0);
    MPI_Get(strip.data(), 3, MPI_DOUBLE, right, size - 1, 3, MPI_DOUBLE, win);
    double pair[2] = { 7.0 + rank, 8.0 + rank };
    MPI_Put(pair, 2, MPI_DOUBLE, left, size + 2, 2, MPI_DOUBLE, win);
    MPI_Win_fence(0, win);
    CHECK_RANK(strip[0] == 100.0 + (size - 1) && strip[1] == -1.0 && strip[2] == 1000.0 + right);
    CHECK_RANK(window[size + 2] == 7.0 + right && window[size + 3] == 8.0 + right);

    // accesses outside the target's segment are refused
    bool threw = false;
    try {
        MPI_Put(pair, 2, MPI_DOUBLE, right, n - 1, 2, MPI_DOUBLE, win);
    } catch (const std::runtime_error&) {
        threw = true;
    }
    CHECK_RANK(threw);
    CHECK_RANK(MPI_Put(pair, 2, MPI_DOUBLE, right, 0, 1, MPI_DOUBLE, win) == MPI_ERR_ARG);
    CHECK_RANK(MPI_Win_free(&win) == MPI_SUCCESS && win == nullptr);

    // Shared windows: load/store straight into a peer's segment
    long* myShared = nullptr;
    MPI_Win shared = nullptr;
    CHECK_RANK(MPI_Win_allocate_shared(8 * sizeof(long), sizeof(long), MPI_INFO_NULL, MPI_COMM_WORLD, &myShared, &shared) == MPI_SUCCESS);
    CHECK_RANK(myShared != nullptr && ((uintptr_t)myShared % 64) == 0 && myShared[3] == 0);
    long* leftShared = nullptr;
    MPI_Aint leftSize = 0;
    int leftUnit = 0;
    CHECK_RANK(MPI_Win_shared_query(shared, left, &leftSize, &leftUnit, &leftShared) == MPI_SUCCESS);
    CHECK_RANK(leftSize == (MPI_Aint)(8 * sizeof(long)) && leftUnit == (int)sizeof(long) && leftShared != nullptr);
    CHECK_RANK(size == 1 || leftShared != myShared);
    CHECK_RANK(MPI_Win_shared_query(shared, size, &leftSize, &leftUnit, &leftShared) == MPI_ERR_RANK);
    MPI_Win_shared_query(shared, left, &leftSize, &leftUnit, &leftShared);
    MPI_Win_fence(0, shared);
    leftShared[5] = 5000 + rank;
    MPI_Win_fence(0, shared);
    CHECK_RANK(myShared[5] == 5000 + right);
    // two windows can be alive at once and keep separate ids
    std::vector<int> small(4, rank);
    MPI_Win second = nullptr;
    MPI_Win_create(small.data(), 4 * sizeof(int), sizeof(int), MPI_INFO_NULL, MPI_COMM_WORLD, &second);
    CHECK_RANK(second->id != shared->id);
    int got = -1;
    MPI_Win_fence(0, second);
    MPI_Get(&got, 1, MPI_INT, right, 2, 1, MPI_INT, second);
    MPI_Win_fence(0, second);
    CHECK_RANK(got == right);
    MPI_Win_free(&second);
    MPI_Win_free(&shared);

    MPI_Comm dup = nullptr;
    CHECK_RANK(MPI_Comm_dup(MPI_COMM_WORLD, &dup) == MPI_SUCCESS && dup == MPI_COMM_WORLD);
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}
}

TEST_CASE("mpi: one-sided windows, put/get/fence and shared segments", "[mpi][rma]")
{
    runMpi("rma-local", 5, 1, bodyRma);
    runMpi("rma-gpuhosts", 6, 3, bodyRma);
    runMpi("rma-one", 1, 1, bodyRma);
}

namespace {
// ---- sub-communicators ----
int bodySubComms(int rank, int size)
{
    // Even / odd halves, ranked in REVERSE world order through the key
    MPI_Comm half = nullptr;
    CHECK_RANK(MPI_Comm_split(MPI_COMM_WORLD, rank % 2, -rank, &half) == MPI_SUCCESS);
    int hRank = -1, hSize = -1;
    MPI_Comm_rank(half, &hRank);
    MPI_Comm_size(half, &hSize);
    const int nEven = (size + 1) / 2, nOdd = size / 2;
    CHECK_RANK(hSize == (rank % 2 == 0 ? nEven : nOdd));
    // members sorted by descending world rank
    std::vector<int> members;
    for (int r = size - 1; r >= 0; r--) {
        if (r % 2 == rank % 2) {
            members.push_back(r);
        }
    }
    CHECK_RANK(members[hRank] == rank);

    // point to point in communicator ranks
    int right = (hRank + 1) % hSize, left = (hRank + hSize - 1) % hSize;
    int token = rank, got = -1;
    MPI_Status status{};
    MPI_Sendrecv(&token, 1, MPI_INT, right, 0, &got, 1, MPI_INT, left, 0, half, &status);
    CHECK_RANK(got == members[left]);
    CHECK_RANK(status.MPI_SOURCE == left);

    // collectives stay inside the half
    int root = hSize - 1;
    std::vector<int> b(100, hRank == root ? 1000 + rank % 2 : -1);
    MPI_Bcast(b.data(), 100, MPI_INT, root, half);
    CHECK_RANK(b[0] == 1000 + rank % 2 && b[99] == 1000 + rank % 2);
    long mine = rank + 1, sum = 0;
    MPI_Allreduce(&mine, &sum, 1, MPI_LONG, MPI_SUM, half);
    long expected = 0;
    for (int m : members) {
        expected += m + 1;
    }
    CHECK_RANK(sum == expected);
    long inPlace = rank + 1;
    MPI_Allreduce(MPI_IN_PLACE, &inPlace, 1, MPI_LONG, MPI_SUM, half);
    CHECK_RANK(inPlace == expected);
    long maxAtRoot = -1;
    MPI_Reduce(&mine, &maxAtRoot, 1, MPI_LONG, MPI_MAX, 0, half);
    if (hRank == 0) {
        CHECK_RANK(maxAtRoot == members[0] + 1);
    }
    long prefix = 0;
    MPI_Scan(&mine, &prefix, 1, MPI_LONG, MPI_SUM, half);
    long expectedPrefix = 0;
    for (int i = 0; i <= hRank; i++) {
        expectedPrefix += members[i] + 1;
    }
    CHECK_RANK(prefix == expectedPrefix);

    std::vector<int> gathered(2 * hSize, -1);
    int pair[2] = { rank, rank * 10 };
    MPI_Gather(pair, 2, MPI_INT, gathered.data(), 2, MPI_INT, root, half);
    if (hRank == root) {
        for (int i = 0; i < hSize; i++) {
            CHECK_RANK(gathered[2 * i] == members[i] && gathered[2 * i + 1] == members[i] * 10);
        }
    }
    std::vector<int> toScatter(hSize);
    for (int i = 0; i < hSize; i++) {
        toScatter[i] = 500 + i;
    }
    int piece = -1;
    MPI_Scatter(toScatter.data(), 1, MPI_INT, &piece, 1, MPI_INT, 0, half);
    CHECK_RANK(piece == 500 + hRank);
    std::vector<int> everyone(hSize, -1);
    MPI_Allgather(&rank, 1, MPI_INT, everyone.data(), 1, MPI_INT, half);
    CHECK_RANK(everyone == members);
    std::vector<int> everyoneInPlace(hSize, -1);
    everyoneInPlace[hRank] = rank;
    MPI_Allgather(MPI_IN_PLACE, 0, MPI_DATATYPE_NULL, everyoneInPlace.data(), 1, MPI_INT, half);
    CHECK_RANK(everyoneInPlace == members);
    std::vector<int> out(hSize), in(hSize, -1);
    for (int i = 0; i < hSize; i++) {
        out[i] = rank * 100 + members[i];
    }
    MPI_Alltoall(out.data(), 1, MPI_INT, in.data(), 1, MPI_INT, half);
    for (int i = 0; i < hSize; i++) {
        CHECK_RANK(in[i] == members[i] * 100 + rank);
    }
    MPI_Barrier(half);

    // splitting a sub-communicator again
    MPI_Comm quarter = nullptr;
    MPI_Comm_split(half, hRank < hSize / 2 ? 0 : 1, hRank, &quarter);
    int qRank = -1, qSize = -1;
    MPI_Comm_rank(quarter, &qRank);
    MPI_Comm_size(quarter, &qSize);
    CHECK_RANK(qSize == (hRank < hSize / 2 ? hSize / 2 : hSize - hSize / 2));
    int qSum = 0, one = 1;
    MPI_Allreduce(&one, &qSum, 1, MPI_INT, MPI_SUM, quarter);
    CHECK_RANK(qSum == qSize);
    MPI_Comm_free(&quarter);
    CHECK_RANK(quarter == MPI_COMM_NULL);

    // MPI_UNDEFINED opts out
    MPI_Comm notZero = nullptr;
    MPI_Comm_split(MPI_COMM_WORLD, rank == 0 ? MPI_UNDEFINED : 7, rank, &notZero);
    if (rank == 0) {
        CHECK_RANK(notZero == MPI_COMM_NULL);
    } else {
        int nzSize = -1, nzRank = -1;
        MPI_Comm_size(notZero, &nzSize);
        MPI_Comm_rank(notZero, &nzRank);
        CHECK_RANK(nzSize == size - 1 && nzRank == rank - 1);
        MPI_Comm_free(&notZero);
    }

    // groups: the first three world ranks, reversed
    MPI_Group worldGroup = nullptr, firstThree = nullptr;
    MPI_Comm_group(MPI_COMM_WORLD, &worldGroup);
    int pick[3] = { 2, 1, 0 };
    CHECK_RANK(MPI_Group_incl(worldGroup, 3, pick, &firstThree) == MPI_SUCCESS);
    int bad[1] = { size };
    MPI_Group broken = nullptr;
    CHECK_RANK(MPI_Group_incl(worldGroup, 1, bad, &broken) == MPI_ERR_RANK);
    MPI_Comm three = nullptr;
    MPI_Comm_create(MPI_COMM_WORLD, firstThree, &three);
    if (rank < 3) {
        int tRank = -1, tSize = -1;
        MPI_Comm_rank(three, &tRank);
        MPI_Comm_size(three, &tSize);
        CHECK_RANK(tSize == 3 && tRank == 2 - rank);
        int v = rank, total = 0;
        MPI_Allreduce(&v, &total, 1, MPI_INT, MPI_SUM, three);
        CHECK_RANK(total == 3);
        // only the members take part in this one
        MPI_Comm again = nullptr;
        MPI_Comm_create_group(MPI_COMM_WORLD, firstThree, 42, &again);
        int aSize = -1;
        MPI_Comm_size(again, &aSize);
        CHECK_RANK(aSize == 3 && again->id != three->id);
        MPI_Comm_free(&again);
        MPI_Comm_free(&three);
    } else {
        CHECK_RANK(three == MPI_COMM_NULL);
    }
    MPI_Group_free(&firstThree);
    MPI_Group_free(&worldGroup);
    CHECK_RANK(worldGroup == nullptr);

    // ranks that share this process's address space
    MPI_Comm node = nullptr;
    MPI_Comm_split_type(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, rank, MPI_INFO_NULL, &node);
    int nodeSize = -1;
    MPI_Comm_size(node, &nodeSize);
    CHECK_RANK(nodeSize == size);
    MPI_Comm_free(&node);
    MPI_Comm_free(&half);
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}
}

TEST_CASE("mpi: sub-communicators, groups and their collectives", "[mpi][subcomm]")
{
    runMpi("subcomm-local", 7, 1, bodySubComms);
    runMpi("subcomm-gpuhosts", 6, 3, bodySubComms);
    runMpi("subcomm-four", 4, 1, bodySubComms);
}

namespace {
// ---- derived datatypes, handle conversion, irregular reduce-scatter ----
int bodyTypesAndHandles(int rank, int size)
{
    // three ints travel as one element
    MPI_Datatype triple = nullptr;
    CHECK_RANK(MPI_Type_contiguous(3, MPI_INT, &triple) == MPI_SUCCESS);
    MPI_Type_commit(&triple);
    int typeSize = 0;
    MPI_Type_size(triple, &typeSize);
    CHECK_RANK(typeSize == 3 * (int)sizeof(int));
    int right = (rank + 1) % size, left = (rank + size - 1) % size;
    int out[6] = { rank, rank + 1, rank + 2, rank + 3, rank + 4, rank + 5 }, in[6] = { 0 };
    MPI_Status status{};
    MPI_Sendrecv(out, 2, triple, right, 0, in, 2, triple, left, 0, MPI_COMM_WORLD, &status);
    CHECK_RANK(in[0] == left && in[5] == left + 5);
    int count = 0;
    MPI_Get_count(&status, triple, &count);
    CHECK_RANK(count == 2);
    // reductions see through the derived type
    int sums[6] = { 0 };
    MPI_Allreduce(out, sums, 2, triple, MPI_SUM, MPI_COMM_WORLD);
    int base = size * (size - 1) / 2;
    CHECK_RANK(sums[0] == base && sums[5] == base + 5 * size);
    // nested: two triples
    MPI_Datatype six = nullptr;
    MPI_Type_contiguous(2, triple, &six);
    MPI_Type_size(six, &typeSize);
    CHECK_RANK(typeSize == 6 * (int)sizeof(int));
    int maxes[6] = { 0 };
    MPI_Allreduce(out, maxes, 1, six, MPI_MAX, MPI_COMM_WORLD);
    CHECK_RANK(maxes[0] == size - 1 && maxes[5] == size + 4);
    CHECK_RANK(MPI_Type_free(&six) == MPI_SUCCESS && six == MPI_DATATYPE_NULL);
    CHECK_RANK(MPI_Type_free(&triple) == MPI_SUCCESS);
    MPI_Datatype predefined = MPI_INT;
    CHECK_RANK(MPI_Type_free(&predefined) == MPI_ERR_ARG);

    // Fortran handles
    CHECK_RANK(MPI_Comm_f2c(MPI_Comm_c2f(MPI_COMM_WORLD)) == MPI_COMM_WORLD);
    CHECK_RANK(MPI_Comm_f2c(MPI_Comm_c2f(MPI_COMM_NULL)) == MPI_COMM_NULL);
    MPI_Comm half = nullptr;
    MPI_Comm_split(MPI_COMM_WORLD, rank % 2, rank, &half);
    MPI_Comm same = MPI_Comm_f2c(MPI_Comm_c2f(half));
    int a = -1, b = -2;
    MPI_Comm_rank(half, &a);
    MPI_Comm_rank(same, &b);
    CHECK_RANK(a == b);
    MPI_Comm_free(&same);
    MPI_Comm_free(&half);
    CHECK_RANK(MPI_Comm_f2c(123456) == MPI_COMM_NULL);

    // non-blocking all-reduce on host buffers completes at the wait
    std::vector<int> nbIn(100, rank), nbOut(100, -1);
    MPI_Request nbReq = nullptr;
    CHECK_RANK(MPI_Iallreduce(nbIn.data(), nbOut.data(), 100, MPI_INT, MPI_SUM, MPI_COMM_WORLD, &nbReq) == MPI_SUCCESS);
    MPI_Wait(&nbReq, MPI_STATUS_IGNORE);
    CHECK_RANK(nbOut[0] == base && nbOut[99] == base);

    // reduce-scatter with a different block per rank: rank r gets r + 1 sums
    std::vector<int> counts(size);
    int total = 0;
    for (int r = 0; r < size; r++) {
        counts[r] = r + 1;
        total += r + 1;
    }
    std::vector<long> contrib(total), block(rank + 1, -1);
    for (int i = 0; i < total; i++) {
        contrib[i] = i + rank;
    }
    MPI_Reduce_scatter(contrib.data(), block.data(), counts.data(), MPI_LONG, MPI_SUM, MPI_COMM_WORLD);
    int myOffset = rank * (rank + 1) / 2;
    for (int i = 0; i <= rank; i++) {
        CHECK_RANK(block[i] == (long)(myOffset + i) * size + base);
    }
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}
}

TEST_CASE("mpi: derived datatypes, handle conversion, irregular reduce-scatter", "[mpi]")
{
    runMpi("types-local", 5, 1, bodyTypesAndHandles);
    runMpi("types-gpuhosts", 4, 2, bodyTypesAndHandles);
}

TEST_CASE("mpi: point-to-point on one host", "[mpi]")
{
    runMpi("p2p-local", 4, 1, bodyPointToPoint);
}

TEST_CASE("mpi: collectives on one host", "[mpi]")
{
    runMpi("coll-local", 5, 1, bodyCollectives);
}

TEST_CASE("mpi: misc API (cartesian, types, alloc, gatherv)", "[mpi]")
{
    runMpi("misc-local", 4, 1, bodyMisc);
}

TEST_CASE("mpi: ranks on several virtual GPU hosts of one worker", "[mpi]")
{
    runMpi("coll-gpuhosts", 8, 4, bodyCollectives);
    runMpi("p2p-gpuhosts", 6, 3, bodyPointToPoint);
}

TEST_CASE("mpi: ping-pong and all-reduce bursts", "[mpi][bench]")
{
    runMpi("pingpong", 2, 1, bodyPingPongAndAllreduce);
    runMpi("allreduce8", 8, 1, bodyPingPongAndAllreduce);
}

// ---------------------------------------------------------------------------
// Migration (strategy: reference tests/dist/mpi/test_mpi_functions.cpp
// "Test triggering an MPI migration", tests/test/planner migration tests)
// ---------------------------------------------------------------------------
#include <faabric/mpi/migration.h>

namespace {
std::atomic<int> migratedExecutions{ 0 };
// Lets a test change the cluster while the ranks wait just before their
// migration point
std::atomic<bool> holdBeforeMigrationPoint{ false };
std::atomic<int> ranksWaitingAtGate{ 0 };

// Ranks iterate; halfway through they hit a migration point.  Ranks that move
// re-enter the function with the loop index as input and carry on.
int migrationBody(faabric::Message& msg, faabric::executor::Executor* exec)
{
    const int nLoops = 6, checkAt = 3;
    int start = msg.inputdata().empty() ? 0 : std::stoi(msg.inputdata());
    if (start > 0) {
        migratedExecutions++;
    }
    MPI_Init(nullptr, nullptr);
    int rank = -1, size = -1;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    // Something in memory that must survive the move
    auto mem = exec->getMemoryView();
    if (start == 0) {
        *(int*)(mem.data() + 128) = 1000 + rank;
    } else if (*(int*)(mem.data() + 128) != 1000 + rank) {
        printf("         rank %d: memory not restored after migration (%d)\n", rank, *(int*)(mem.data() + 128));
        return 1;
    }
    for (int i = start; i < nLoops; i++) {
        if (i == checkAt && start == 0) {
            ranksWaitingAtGate++;
            while (holdBeforeMigrationPoint.load()) {
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
            MPI_Barrier(MPI_COMM_WORLD);
            faabric::mpi::mpiMigrationPoint(i);
        }
        int v = rank + i, sum = 0;
        MPI_Allreduce(&v, &sum, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
        if (sum != size * (size - 1) / 2 + i * size) {
            printf("         rank %d: bad all-reduce at loop %d: %d\n", rank, i, sum);
            return 1;
        }
    }
    MPI_Barrier(MPI_COMM_WORLD);
    MPI_Finalize();
    return 0;
}
}

TEST_CASE("mpi: an app is migrated onto fewer hosts at a migration point", "[mpi][migration]")
{
    // Two (virtual) hosts with 4 slots each; the world starts 2 + 2 and the
    // bin-pack policy consolidates it onto one host at the migration point
    ClusterFixture f(0, 2, 4);
    migratedExecutions = 0;
    registerTestFunction("mpi", "migrate", [&](auto* exec, int, int idx, auto req) {
        return migrationBody(*req->mutable_messages(idx), exec);
    });
    auto req = faabric::util::batchExecFactory("mpi", "migrate", 1);
    auto& msg = *req->mutable_messages(0);
    msg.set_ismpi(true);
    msg.set_mpiworldsize(4);
    msg.set_recordexecgraph(true);

    // Pin the initial layout: ranks 0,1 on gpu0 and 2,3 on gpu1
    auto preload = std::make_shared<faabric::batch_scheduler::SchedulingDecision>(req->appid(), 0);
    std::vector<std::string> initial = { "gpu0", "gpu0", "gpu1", "gpu1" };
    for (int r = 0; r < 4; r++) {
        preload->addMessage(initial[r], 0, r, r);
    }
    f.plannerCli.preloadSchedulingDecision(preload);

    f.plannerCli.callFunctions(req);
    auto status = f.awaitBatch(req, 60000);
    REQUIRE_EQ(f.planner.getNumMigrations(), 1);
    // 4 ranks + the re-executions of those that moved
    std::map<int, std::string> finalHost;
    for (auto& m : status->messageresults()) {
        if (m.returnvalue() != 0) {
            fbtest::fail(__FILE__, __LINE__, "rank " + std::to_string(m.mpirank()) + " failed: " + m.outputdata());
        }
        finalHost[m.mpirank()] = m.executedhost();
    }
    REQUIRE_EQ(finalHost.size(), 4u);
    std::set<std::string> hosts;
    for (auto& [r, h] : finalHost) {
        hosts.insert(h);
    }
    REQUIRE_EQ(hosts.size(), 1u);
    REQUIRE_EQ(migratedExecutions.load(), 2);
    // Slots are all free again
    for (auto& h : f.plannerCli.getAvailableHosts()) {
        REQUIRE_EQ(h.usedslots(), 0);
    }
    faabric::mpi::getMpiWorldRegistry().clear();
}

TEST_CASE("mpi: spot eviction freezes an app, it thaws when capacity returns", "[mpi][migration]")
{
    ClusterFixture f(0, 2, 2);
    f.planner.setPolicy("spot");
    migratedExecutions = 0;
    registerTestFunction("mpi", "freeze", [&](auto* exec, int, int idx, auto req) {
        return migrationBody(*req->mutable_messages(idx), exec);
    });
    auto req = faabric::util::batchExecFactory("mpi", "freeze", 1);
    auto& msg = *req->mutable_messages(0);
    msg.set_ismpi(true);
    msg.set_mpiworldsize(4);
    holdBeforeMigrationPoint = true;
    ranksWaitingAtGate = 0;
    f.plannerCli.callFunctions(req);
    for (int i = 0; i < 5000 && ranksWaitingAtGate.load() < 4; i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
    REQUIRE_EQ(ranksWaitingAtGate.load(), 4);
    // gpu1 is going away and nothing else has room: the app must freeze
    f.planner.setNextEvictedVm({ "gpu1" });
    holdBeforeMigrationPoint = false;

    // Wait until all four ranks have reported FROZEN
    bool frozen = false;
    for (int i = 0; i < 2000 && !frozen; i++) {
        auto evicted = f.planner.getEvictedReqs();
        auto it = evicted.find(req->appid());
        if (it != evicted.end() && it->second->messages_size() == 4) {
            frozen = true;
            for (auto& m : it->second->messages()) {
                frozen = frozen && m.returnvalue() == FROZEN_FUNCTION_RETURN_VALUE;
            }
        }
        if (!frozen) {
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
    }
    REQUIRE(frozen);
    REQUIRE_EQ(f.planner.getInFlightReqs().size(), 0u);
    for (auto& h : f.plannerCli.getAvailableHosts()) {
        REQUIRE_EQ(h.usedslots(), 0);
    }
    // Polling while there is no capacity keeps it frozen
    auto status = f.plannerCli.getBatchResults(req);
    REQUIRE(status != nullptr);
    REQUIRE(!status->finished());
    REQUIRE_EQ(f.planner.getEvictedReqs().size(), 1u);

    // The eviction passes: the next poll thaws the app and it runs to the end
    f.planner.setNextEvictedVm({});
    status = f.awaitBatch(req, 60000);
    REQUIRE_EQ(f.planner.getEvictedReqs().size(), 0u);
    std::set<int> ranks;
    for (auto& m : status->messageresults()) {
        if (m.returnvalue() != 0) {
            fbtest::fail(__FILE__, __LINE__, "rank " + std::to_string(m.mpirank()) + " failed: " + m.outputdata());
        }
        ranks.insert(m.mpirank());
    }
    REQUIRE_EQ(ranks.size(), 4u);
    REQUIRE_EQ(migratedExecutions.load(), 4);
    f.planner.setPolicy("bin-pack");
    faabric::mpi::getMpiWorldRegistry().clear();
}

// ---------------------------------------------------------------------------
// Worlds spanning two hosts, driven directly through MpiWorld in mock mode:
// the two-level algorithms send exactly the reference's remote messages
// (strategy: reference tests/test/mpi/test_remote_mpi_worlds.cpp:34-430)
// ---------------------------------------------------------------------------
namespace {
struct TwoHostWorlds
{
    // ranks 0,1 on this host, 2,3 on `otherHost`
    static constexpr int worldId = 4242;
    static constexpr int groupId = 8484;
    static constexpr int worldSize = 4;
    std::string thisHost = faabric::util::getSystemConfig().endpointHost;
    std::string otherHost = "192.0.2.201";
    faabric::Message msg = faabric::util::messageFactory("mpi", "two-hosts");
    faabric::mpi::MpiWorld thisWorld;
    faabric::mpi::MpiWorld otherWorld;

    TwoHostWorlds()
    {
        faabric::util::setMockMode(true);
        faabric::mpi::clearMpiMockedMessages();
        msg.set_ismpi(true);
        msg.set_mpiworldid(worldId);
        msg.set_mpiworldsize(worldSize);
        msg.set_groupid(groupId);
        faabric::batch_scheduler::SchedulingDecision decision(msg.appid(), groupId);
        for (int r = 0; r < worldSize; r++) {
            decision.addMessage(r < 2 ? thisHost : otherHost, msg.id() + r, r, r);
        }
        faabric::transport::getPointToPointBroker().setUpLocalMappingsFromSchedulingDecision(decision);
        thisWorld.initialiseFromMsg(msg);
        otherWorld.overrideHost(otherHost);
        otherWorld.initialiseFromMsg(msg);
    }

    ~TwoHostWorlds()
    {
        faabric::mpi::clearMpiMockedMessages();
        faabric::transport::getPointToPointBroker().clear();
        faabric::util::setMockMode(false);
    }
};
}

TEST_CASE("mpi: two-host worlds place ranks and leaders", "[mpi][mock]")
{
    TwoHostWorlds w;
    REQUIRE_EQ(w.thisWorld.getSize(), 4);
    REQUIRE_EQ(w.thisWorld.getId(), (int)TwoHostWorlds::worldId);
    REQUIRE_EQ(w.thisWorld.getHostForRank(0), w.thisHost);
    REQUIRE_EQ(w.thisWorld.getHostForRank(1), w.thisHost);
    REQUIRE_EQ(w.thisWorld.getHostForRank(2), w.otherHost);
    REQUIRE_EQ(w.otherWorld.getHostForRank(3), w.otherHost);
    REQUIRE_EQ(w.thisWorld.getUser(), std::string("mpi"));
    REQUIRE_EQ(w.thisWorld.getFunction(), std::string("two-hosts"));
    REQUIRE(w.thisWorld.getWTime() >= 0.0);
}

TEST_CASE("mpi: sends are captured and receives return at once (mock mode)", "[mpi][mock]")
{
    TwoHostWorlds w;
    std::vector<int> data = { 1, 2, 3 };
    // Local and remote destinations alike are recorded, nothing is queued
    w.thisWorld.send(0, 1, BYTES(data.data()), MPI_INT, 3);
    w.thisWorld.send(0, 3, BYTES(data.data()), MPI_INT, 3, faabric::mpi::MpiMessageType::SENDRECV);
    REQUIRE_EQ(w.thisWorld.getLocalQueueSize(0, 1), 0L);
    auto sent = faabric::mpi::getMpiMockedMessages(0);
    REQUIRE_EQ(sent.size(), (size_t)2);
    REQUIRE_EQ(sent[0].worldId, (int)TwoHostWorlds::worldId);
    REQUIRE_EQ(sent[0].sendRank, 0);
    REQUIRE_EQ(sent[0].recvRank, 1);
    REQUIRE_EQ(sent[0].count, 3);
    REQUIRE_EQ(sent[0].typeSize, (int)sizeof(int));
    REQUIRE(sent[0].messageType == faabric::mpi::MpiMessageType::NORMAL);
    REQUIRE(memcmp(sent[0].buffer, data.data(), 3 * sizeof(int)) == 0);
    REQUIRE_EQ(sent[1].recvRank, 3);
    REQUIRE(sent[1].messageType == faabric::mpi::MpiMessageType::SENDRECV);
    REQUIRE(faabric::mpi::getMpiMockedMessages(1).empty());
    // A receive that could never be satisfied does not block
    std::vector<int> got(3, -1);
    w.thisWorld.recv(2, 0, BYTES(got.data()), MPI_INT, 3, nullptr);
    REQUIRE_EQ(got[0], -1);
    // isend is an eager send: recorded too
    int reqId = w.thisWorld.isend(1, 2, BYTES(data.data()), MPI_INT, 3);
    w.thisWorld.awaitAsyncRequest(reqId);
    REQUIRE_EQ(faabric::mpi::getMpiMockedMessages(1).size(), (size_t)1);
    // out-of-range ranks are rejected
    REQUIRE_THROWS(w.thisWorld.send(0, 4, BYTES(data.data()), MPI_INT, 3));
    REQUIRE_THROWS(w.thisWorld.recv(-1, 0, BYTES(got.data()), MPI_INT, 3, nullptr));
}

TEST_CASE("mpi: broadcast fans out through one leader per host (mock mode)", "[mpi][mock]")
{
    TwoHostWorlds w;
    std::vector<int> data = { 7, 8, 9, 10 };
    auto bcast = faabric::mpi::MpiMessageType::BROADCAST;
    // Root 0: one message to its co-located rank, one to the remote leader
    w.thisWorld.broadcast(0, 0, BYTES(data.data()), MPI_INT, 4, bcast);
    auto fromRoot = faabric::mpi::getMpiMockedMessages(0);
    REQUIRE_EQ(fromRoot.size(), (size_t)2);
    std::set<int> dests = { fromRoot[0].recvRank, fromRoot[1].recvRank };
    REQUIRE(dests == (std::set<int>{ 1, 2 }));
    REQUIRE(fromRoot[0].messageType == bcast);
    // The remote leader forwards to its host, the leaf forwards nothing
    w.otherWorld.broadcast(0, 2, BYTES(data.data()), MPI_INT, 4, bcast);
    auto fromLeader = faabric::mpi::getMpiMockedMessages(2);
    REQUIRE_EQ(fromLeader.size(), (size_t)1);
    REQUIRE_EQ(fromLeader[0].recvRank, 3);
    w.otherWorld.broadcast(0, 3, BYTES(data.data()), MPI_INT, 4, bcast);
    REQUIRE(faabric::mpi::getMpiMockedMessages(3).empty());
    w.thisWorld.broadcast(0, 1, BYTES(data.data()), MPI_INT, 4, bcast);
    REQUIRE(faabric::mpi::getMpiMockedMessages(1).empty());

    // A root that is not its host's lowest rank still feeds each host once
    faabric::mpi::clearMpiMockedMessages();
    w.otherWorld.broadcast(3, 3, BYTES(data.data()), MPI_INT, 4, bcast);
    auto fromThree = faabric::mpi::getMpiMockedMessages(3);
    REQUIRE_EQ(fromThree.size(), (size_t)2);
    dests = { fromThree[0].recvRank, fromThree[1].recvRank };
    REQUIRE(dests == (std::set<int>{ 0, 2 }));
}

TEST_CASE("mpi: reduce, gather and barrier go through the local leader (mock mode)", "[mpi][mock]")
{
    TwoHostWorlds w;
    // Remote host: rank 3 hands its data to leader 2, which sends ONE message
    // to the root
    std::vector<int> three = { 30, 31 }, two = { 20, 21 };
    w.otherWorld.reduce(3, 0, BYTES(three.data()), nullptr, MPI_INT, 2, MPI_SUM);
    auto fromLeaf = faabric::mpi::getMpiMockedMessages(3);
    REQUIRE_EQ(fromLeaf.size(), (size_t)1);
    REQUIRE_EQ(fromLeaf[0].recvRank, 2);
    REQUIRE(fromLeaf[0].messageType == faabric::mpi::MpiMessageType::REDUCE);
    w.otherWorld.reduce(2, 0, BYTES(two.data()), nullptr, MPI_INT, 2, MPI_SUM);
    auto fromLeader = faabric::mpi::getMpiMockedMessages(2);
    REQUIRE_EQ(fromLeader.size(), (size_t)1);
    REQUIRE_EQ(fromLeader[0].recvRank, 0);
    REQUIRE_EQ(fromLeader[0].count, 2);
    // the leader's own send buffer is untouched
    REQUIRE(two[0] == 20 && two[1] == 21);
    // Root's host: the co-located rank sends straight to the root
    w.thisWorld.reduce(1, 0, BYTES(two.data()), nullptr, MPI_INT, 2, MPI_SUM);
    auto fromOne = faabric::mpi::getMpiMockedMessages(1);
    REQUIRE_EQ(fromOne.size(), (size_t)1);
    REQUIRE_EQ(fromOne[0].recvRank, 0);

    // gather: the leader packs its host's chunks into one message
    faabric::mpi::clearMpiMockedMessages();
    w.otherWorld.gather(3, 0, BYTES(three.data()), MPI_INT, 2, nullptr, MPI_INT, 2);
    fromLeaf = faabric::mpi::getMpiMockedMessages(3);
    REQUIRE_EQ(fromLeaf.size(), (size_t)1);
    REQUIRE_EQ(fromLeaf[0].recvRank, 2);
    REQUIRE_EQ(fromLeaf[0].count, 2);
    w.otherWorld.gather(2, 0, BYTES(two.data()), MPI_INT, 2, nullptr, MPI_INT, 2);
    auto packed = faabric::mpi::getMpiMockedMessages(2);
    REQUIRE_EQ(packed.size(), (size_t)1);
    REQUIRE_EQ(packed[0].recvRank, 0);
    REQUIRE_EQ(packed[0].count, 4);
    REQUIRE(((int*)packed[0].buffer)[0] == 20 && ((int*)packed[0].buffer)[1] == 21);

    // barrier: everyone joins at rank 0 with an empty message
    faabric::mpi::clearMpiMockedMessages();
    w.otherWorld.barrier(3);
    auto join = faabric::mpi::getMpiMockedMessages(3);
    REQUIRE(!join.empty());
    REQUIRE_EQ(join[0].recvRank, 0);
    REQUIRE_EQ(join[0].count, 0);
    REQUIRE(join[0].messageType == faabric::mpi::MpiMessageType::BARRIER_JOIN);
}

TEST_CASE("mpi: cartesian topology through the world API", "[mpi]")
{
    TwoHostWorlds w;
    // 2 x 2 periodic grid
    int dims[2] = { 2, 2 };
    int periods[2] = { 0, 0 };
    int coords[2] = { -1, -1 };
    w.thisWorld.getCartesianRank(3, 2, dims, periods, coords);
    REQUIRE(coords[0] == 1 && coords[1] == 1);
    REQUIRE(periods[0] == 1 && periods[1] == 1);
    int rank = -1;
    w.thisWorld.getRankFromCoords(&rank, coords);
    REQUIRE_EQ(rank, 3);
    int src = -1, dst = -1;
    w.thisWorld.shiftCartesianCoords(0, 0, 1, &src, &dst);
    REQUIRE(src == 2 && dst == 2);
    w.thisWorld.shiftCartesianCoords(0, 1, 1, &src, &dst);
    REQUIRE(src == 1 && dst == 1);
    // zero displacement: both ends are the rank itself
    w.thisWorld.shiftCartesianCoords(1, 1, 0, &src, &dst);
    REQUIRE(src == 1 && dst == 1);
    // 4 x 1
    int dims41[2] = { 4, 1 };
    w.thisWorld.getCartesianRank(2, 2, dims41, periods, coords);
    REQUIRE(coords[0] == 2 && coords[1] == 0);
    // wrong grid size / more than two real dimensions
    int bad[2] = { 3, 2 };
    REQUIRE_THROWS(w.thisWorld.getCartesianRank(0, 2, bad, periods, coords));
    int three[3] = { 2, 1, 2 };
    int periods3[3] = { 0, 0, 0 };
    int coords3[3] = { 0, 0, 0 };
    REQUIRE_THROWS(w.thisWorld.getCartesianRank(0, 3, three, periods3, coords3));
}
