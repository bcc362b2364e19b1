// shared_mutex-protected hash map with the closure-based access API of the
// reference's ConcurrentMap (include/faabric/util/concurrent_map.h:39-304).
#pragma once

#include <faabric/util/locks.h>

#include <algorithm>
#include <functional>
#include <optional>
#include <unordered_map>
#include <utility>
#include <vector>

namespace faabric::util {

template<typename Key, typename Value>
class ConcurrentMap
{
  public:
    ConcurrentMap() = default;

    explicit ConcurrentMap(size_t initialCapacity)
    {
        map.reserve(initialCapacity);
    }

    bool isEmpty() const
    {
        SharedLock lock(mx);
        return map.empty();
    }

    size_t size() const
    {
        SharedLock lock(mx);
        return map.size();
    }

    size_t capacity() const
    {
        SharedLock lock(mx);
        return map.bucket_count();
    }

    void reserve(size_t count)
    {
        FullLock lock(mx);
        map.reserve(count);
    }

    void rehash(size_t count)
    {
        FullLock lock(mx);
        map.rehash(count);
    }

    void clear()
    {
        FullLock lock(mx);
        map.clear();
    }

    bool contains(const Key& key) const
    {
        SharedLock lock(mx);
        return map.find(key) != map.end();
    }

    // Inserts a default-constructible / argument-constructed value if absent.
    // Returns true if this call inserted it.
    template<typename... Args>
    bool tryEmplace(const Key& key, Args&&... args)
    {
        FullLock lock(mx);
        return map.try_emplace(key, std::forward<Args>(args)...).second;
    }

    // Fast path takes only the shared lock when the key already exists.
    // Returns (inserted, copy of value)
    template<typename... Args>
    std::pair<bool, Value> tryEmplaceShared(const Key& key, Args&&... args)
    {
        {
            SharedLock lock(mx);
            auto it = map.find(key);
            if (it != map.end()) {
                return { false, it->second };
            }
        }
        FullLock lock(mx);
        auto [it, inserted] = map.try_emplace(key, std::forward<Args>(args)...);
        return { inserted, it->second };
    }

    // Emplace then run `mutator(inserted, value&)` under the exclusive lock
    template<typename F, typename... Args>
    bool tryEmplaceThenMutate(const Key& key, F&& mutator, Args&&... args)
    {
        FullLock lock(mx);
        auto [it, inserted] = map.try_emplace(key, std::forward<Args>(args)...);
        mutator(inserted, it->second);
        return inserted;
    }

    template<typename V>
    void insertOrAssign(const Key& key, V&& value)
    {
        FullLock lock(mx);
        map.insert_or_assign(key, std::forward<V>(value));
    }

    bool erase(const Key& key)
    {
        FullLock lock(mx);
        return map.erase(key) > 0;
    }

    // Remove every element for which pred(key, value) is true; returns count
    template<typename F>
    size_t eraseIf(F&& pred)
    {
        FullLock lock(mx);
        size_t n = 0;
        for (auto it = map.begin(); it != map.end();) {
            if (pred(it->first, it->second)) {
                it = map.erase(it);
                n++;
            } else {
                ++it;
            }
        }
        return n;
    }

    std::optional<Value> get(const Key& key) const
    {
        SharedLock lock(mx);
        auto it = map.find(key);
        if (it == map.end()) {
            return std::nullopt;
        }
        return it->second;
    }

    // Read access to one element; returns false if missing
    template<typename F>
    bool inspect(const Key& key, F&& inspector) const
    {
        SharedLock lock(mx);
        auto it = map.find(key);
        if (it == map.end()) {
            return false;
        }
        inspector(it->second);
        return true;
    }

    template<typename F>
    bool mutate(const Key& key, F&& mutator)
    {
        FullLock lock(mx);
        auto it = map.find(key);
        if (it == map.end()) {
            return false;
        }
        mutator(it->second);
        return true;
    }

    template<typename F>
    void inspectAll(F&& inspector) const
    {
        SharedLock lock(mx);
        for (const auto& [k, v] : map) {
            inspector(k, v);
        }
    }

    template<typename F>
    void mutateAll(F&& mutator)
    {
        FullLock lock(mx);
        for (auto& [k, v] : map) {
            mutator(k, v);
        }
    }

    std::vector<std::pair<Key, Value>> sortedKvPairs() const
    {
        std::vector<std::pair<Key, Value>> out;
        {
            SharedLock lock(mx);
            out.reserve(map.size());
            for (const auto& kv : map) {
                out.emplace_back(kv.first, kv.second);
            }
        }
        std::sort(out.begin(), out.end(), [](const auto& a, const auto& b) {
            return a.first < b.first;
        });
        return out;
    }

  private:
    mutable std::shared_mutex mx;
    std::unordered_map<Key, Value> map;
};

} // namespace faabric::util
