// Environment-driven global configuration singleton.
// Same knobs and defaults as the reference's SystemConfig
// (include/faabric/util/config.h:12-73, src/util/config.cpp:19-84) plus the
// GPU-specific settings of this implementation.
#pragma once

#include <string>

#define MPI_HOST_STATE_LEN 20
#define DEFAULT_TIMEOUT 60000
#define RESULT_KEY_EXPIRY 30000
#define STATUS_KEY_EXPIRY 300000

namespace faabric::util {

class SystemConfig
{
  public:
    // System
    std::string serialisation;
    std::string logLevel;
    std::string logFile;
    std::string stateMode;
    std::string deltaSnapshotEncoding;

    // Redis-compatible store (in-process here; host/port kept for parity)
    std::string redisStateHost;
    std::string redisQueueHost;
    std::string redisPort;

    // Scheduling
    int overrideCpuCount;
    int overrideFreeCpuStart;
    std::string batchSchedulerMode;

    // Worker-related timeouts (all in milliseconds unless stated)
    int globalMessageTimeout;
    int boundTimeout;
    int reaperIntervalSeconds;

    // MPI
    int defaultMpiWorldSize;

    // Endpoint
    std::string endpointInterface;
    std::string endpointHost;
    int endpointPort;
    int endpointNumThreads;

    // Transport
    int functionServerThreads;
    int stateServerThreads;
    int snapshotServerThreads;
    int pointToPointServerThreads;

    // Dirty tracking
    std::string dirtyTrackingMode;
    std::string diffingMode;

    // Planner
    std::string plannerHost;
    int plannerPort;

    // ---- B200 additions ----
    // Comma separated GPU ordinals this worker may use ("" = all visible)
    std::string gpus;
    // cuda | loopback  (loopback = host memory, no GPU required)
    std::string deviceBackend;
    // auto | oneshot | twoshot | nvls | ll | nccl
    std::string allreduceAlgo;
    int useNvls;
    int commStreams;
    long symmHeapBytes;
    // Execution slots exposed per GPU "host"
    int slotsPerGpu;
    // Offset added to every well-known port (several workers on one box)
    int portOffset;

    SystemConfig();

    void print();

    void reset();

  private:
    int getSystemConfIntParam(const char* name, const char* defaultValue);
    long getSystemConfLongParam(const char* name, const char* defaultValue);

    void initialise();
};

SystemConfig& getSystemConfig();

} // namespace faabric::util
