#pragma once

#include <faabric/util/snapshot.h>

#include <memory>
#include <shared_mutex>
#include <string>
#include <unordered_map>

namespace faabric::snapshot {

class DeviceSnapshot;

// key -> snapshot (host images and device images live side by side)
class SnapshotRegistry
{
  public:
    SnapshotRegistry() = default;

    std::shared_ptr<faabric::util::SnapshotData> getSnapshot(
      const std::string& key);

    bool snapshotExists(const std::string& key);

    void registerSnapshot(const std::string& key,
                          std::shared_ptr<faabric::util::SnapshotData> data);

    void deleteSnapshot(const std::string& key);

    size_t getSnapshotCount();

    // ---- device-resident images ----
    std::shared_ptr<DeviceSnapshot> getDeviceSnapshot(const std::string& key);

    bool deviceSnapshotExists(const std::string& key);

    void registerDeviceSnapshot(const std::string& key,
                                std::shared_ptr<DeviceSnapshot> data);

    void deleteDeviceSnapshot(const std::string& key);

    void clear();

  private:
    std::shared_mutex snapshotsMx;
    std::unordered_map<std::string, std::shared_ptr<faabric::util::SnapshotData>>
      snapshotMap;
    std::unordered_map<std::string, std::shared_ptr<DeviceSnapshot>> deviceMap;
};

SnapshotRegistry& getSnapshotRegistry();

}
