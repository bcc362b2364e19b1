#include <faabric/util/config.h>
#include <faabric/util/environment.h>
#include <faabric/util/logging.h>
#include <faabric/util/network.h>

#include <cstdlib>

namespace faabric::util {

SystemConfig& getSystemConfig()
{
    static SystemConfig conf;
    return conf;
}

SystemConfig::SystemConfig()
{
    this->initialise();
}

void SystemConfig::initialise()
{
    // System
    serialisation = getEnvVar("SERIALISATION", "json");
    logLevel = getEnvVar("LOG_LEVEL", "info");
    logFile = getEnvVar("LOG_FILE", "off");
    stateMode = getEnvVar("STATE_MODE", "inmemory");
    deltaSnapshotEncoding =
      getEnvVar("DELTA_SNAPSHOT_ENCODING", "pages=4096;xor;zstd=1");

    // Redis-compatible store
    redisStateHost = getEnvVar("REDIS_STATE_HOST", "localhost");
    redisQueueHost = getEnvVar("REDIS_QUEUE_HOST", "localhost");
    redisPort = getEnvVar("REDIS_PORT", "6379");

    // Scheduling
    overrideCpuCount = getSystemConfIntParam("OVERRIDE_CPU_COUNT", "0");
    overrideFreeCpuStart = getSystemConfIntParam("OVERRIDE_FREE_CPU_START", "0");
    batchSchedulerMode = getEnvVar("BATCH_SCHEDULER_MODE", "bin-pack");

    // Worker-related timeouts (all in ms)
    globalMessageTimeout =
      getSystemConfIntParam("GLOBAL_MESSAGE_TIMEOUT", "60000");
    boundTimeout = getSystemConfIntParam("BOUND_TIMEOUT", "30000");
    reaperIntervalSeconds = getSystemConfIntParam("REAPER_INTERVAL_SECS", "30");

    // MPI
    defaultMpiWorldSize = getSystemConfIntParam("DEFAULT_MPI_WORLD_SIZE", "5");

    // Endpoint
    endpointInterface = getEnvVar("ENDPOINT_INTERFACE", "");
    endpointHost = getEnvVar("ENDPOINT_HOST", "");
    endpointPort = getSystemConfIntParam("ENDPOINT_PORT", "8080");
    endpointNumThreads = getSystemConfIntParam("ENDPOINT_NUM_THREADS", "4");
    if (endpointHost.empty()) {
        // Default to the primary IP of this machine
        endpointHost = getPrimaryIPForThisHost(endpointInterface);
    }

    // Transport
    functionServerThreads = getSystemConfIntParam("FUNCTION_SERVER_THREADS", "2");
    stateServerThreads = getSystemConfIntParam("STATE_SERVER_THREADS", "2");
    snapshotServerThreads = getSystemConfIntParam("SNAPSHOT_SERVER_THREADS", "2");
    pointToPointServerThreads =
      getSystemConfIntParam("POINT_TO_POINT_SERVER_THREADS", "8");

    // Dirty tracking
    dirtyTrackingMode = getEnvVar("DIRTY_TRACKING_MODE", "segfault");
    diffingMode = getEnvVar("DIFFING_MODE", "xor");

    // Planner
    plannerHost = getEnvVar("PLANNER_HOST", "planner");
    plannerPort = getSystemConfIntParam("PLANNER_PORT", "8080");

    // B200
    gpus = getEnvVar("FAABRIC_GPUS", "");
    deviceBackend = getEnvVar("FAABRIC_DEVICE_BACKEND", "cuda");
    allreduceAlgo = getEnvVar("FAABRIC_ALLREDUCE_ALGO", "auto");
    useNvls = getSystemConfIntParam("FAABRIC_USE_NVLS", "1");
    commStreams = getSystemConfIntParam("FAABRIC_COMM_STREAMS", "2");
    symmHeapBytes = getSystemConfLongParam("FAABRIC_SYMM_HEAP_BYTES", "1073741824");
    slotsPerGpu = getSystemConfIntParam("FAABRIC_SLOTS_PER_GPU", "8");
    portOffset = getSystemConfIntParam("FAABRIC_PORT_OFFSET", "0");
}

int SystemConfig::getSystemConfIntParam(const char* name,
                                        const char* defaultValue)
{
    return (int)strtol(getEnvVar(name, defaultValue).c_str(), nullptr, 10);
}

long SystemConfig::getSystemConfLongParam(const char* name,
                                          const char* defaultValue)
{
    return strtol(getEnvVar(name, defaultValue).c_str(), nullptr, 10);
}

void SystemConfig::reset()
{
    this->initialise();
}

void SystemConfig::print()
{
    SPDLOG_INFO("--- System ---");
    SPDLOG_INFO("SERIALISATION              {}", serialisation);
    SPDLOG_INFO("LOG_LEVEL                  {}", logLevel);
    SPDLOG_INFO("LOG_FILE                   {}", logFile);
    SPDLOG_INFO("STATE_MODE                 {}", stateMode);
    SPDLOG_INFO("DELTA_SNAPSHOT_ENCODING    {}", deltaSnapshotEncoding);
    SPDLOG_INFO("--- Store ---");
    SPDLOG_INFO("REDIS_STATE_HOST           {}", redisStateHost);
    SPDLOG_INFO("REDIS_QUEUE_HOST           {}", redisQueueHost);
    SPDLOG_INFO("REDIS_PORT                 {}", redisPort);
    SPDLOG_INFO("--- Scheduling ---");
    SPDLOG_INFO("OVERRIDE_CPU_COUNT         {}", overrideCpuCount);
    SPDLOG_INFO("OVERRIDE_FREE_CPU_START    {}", overrideFreeCpuStart);
    SPDLOG_INFO("BATCH_SCHEDULER_MODE       {}", batchSchedulerMode);
    SPDLOG_INFO("--- Timeouts ---");
    SPDLOG_INFO("GLOBAL_MESSAGE_TIMEOUT     {}", globalMessageTimeout);
    SPDLOG_INFO("BOUND_TIMEOUT              {}", boundTimeout);
    SPDLOG_INFO("REAPER_INTERVAL_SECS       {}", reaperIntervalSeconds);
    SPDLOG_INFO("--- MPI ---");
    SPDLOG_INFO("DEFAULT_MPI_WORLD_SIZE     {}", defaultMpiWorldSize);
    SPDLOG_INFO("--- Endpoint ---");
    SPDLOG_INFO("ENDPOINT_INTERFACE         {}", endpointInterface);
    SPDLOG_INFO("ENDPOINT_HOST              {}", endpointHost);
    SPDLOG_INFO("ENDPOINT_PORT              {}", endpointPort);
    SPDLOG_INFO("ENDPOINT_NUM_THREADS       {}", endpointNumThreads);
    SPDLOG_INFO("--- Transport ---");
    SPDLOG_INFO("FUNCTION_SERVER_THREADS    {}", functionServerThreads);
    SPDLOG_INFO("STATE_SERVER_THREADS       {}", stateServerThreads);
    SPDLOG_INFO("SNAPSHOT_SERVER_THREADS    {}", snapshotServerThreads);
    SPDLOG_INFO("POINT_TO_POINT_SERVER_THREADS {}", pointToPointServerThreads);
    SPDLOG_INFO("--- Dirty tracking ---");
    SPDLOG_INFO("DIRTY_TRACKING_MODE        {}", dirtyTrackingMode);
    SPDLOG_INFO("DIFFING_MODE               {}", diffingMode);
    SPDLOG_INFO("--- Planner ---");
    SPDLOG_INFO("PLANNER_HOST               {}", plannerHost);
    SPDLOG_INFO("PLANNER_PORT               {}", plannerPort);
    SPDLOG_INFO("--- B200 ---");
    SPDLOG_INFO("FAABRIC_GPUS               {}", gpus);
    SPDLOG_INFO("FAABRIC_DEVICE_BACKEND     {}", deviceBackend);
    SPDLOG_INFO("FAABRIC_ALLREDUCE_ALGO     {}", allreduceAlgo);
    SPDLOG_INFO("FAABRIC_USE_NVLS           {}", useNvls);
    SPDLOG_INFO("FAABRIC_COMM_STREAMS       {}", commStreams);
    SPDLOG_INFO("FAABRIC_SYMM_HEAP_BYTES    {}", symmHeapBytes);
    SPDLOG_INFO("FAABRIC_SLOTS_PER_GPU      {}", slotsPerGpu);
    SPDLOG_INFO("FAABRIC_PORT_OFFSET        {}", portOffset);
}

} // namespace faabric::util
