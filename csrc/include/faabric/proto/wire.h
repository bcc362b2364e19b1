// Runtime support for the generated message classes (faabric.pb.h):
//  * Writer / Reader: protobuf wire format (varint + length-delimited)
//  * JsonWriter / JsonValue: minimal JSON encoder / parser
//  * RepeatedField<T>: std::vector with the few protobuf-isms callers use
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace faabric::proto {

template<typename T>
class RepeatedField : public std::vector<T>
{
  public:
    using std::vector<T>::vector;

    T* Add()
    {
        this->emplace_back();
        return &this->back();
    }

    void Add(const T& v) { this->push_back(v); }

    const T& Get(int i) const { return this->at(i); }

    T* Mutable(int i) { return &this->at(i); }

    void Clear() { this->clear(); }

    void Reserve(int n) { this->reserve(n); }

    void DeleteSubrange(int start, int num)
    {
        this->erase(this->begin() + start, this->begin() + start + num);
    }

    void RemoveLast() { this->pop_back(); }
};

// ---------------------------------------------------------------- binary ---
class Writer
{
  public:
    void rawVarint(uint64_t v)
    {
        while (v >= 0x80) {
            buf.push_back((char)((v & 0x7f) | 0x80));
            v >>= 7;
        }
        buf.push_back((char)v);
    }

    void tag(uint32_t field, int wireType)
    {
        rawVarint(((uint64_t)field << 3) | (uint64_t)wireType);
    }

    // proto3: default values are not written unless `always`
    void varint(uint32_t field, uint64_t v, bool always)
    {
        if (v == 0 && !always) {
            return;
        }
        tag(field, 0);
        rawVarint(v);
    }

    void str(uint32_t field, std::string_view s, bool always)
    {
        if (s.empty() && !always) {
            return;
        }
        tag(field, 2);
        rawVarint(s.size());
        buf.append(s.data(), s.size());
    }

    const std::string& data() const { return buf; }

    std::string take() { return std::move(buf); }

  private:
    std::string buf;
};

class Reader
{
  public:
    explicit Reader(std::string_view s)
      : p(s.data())
      , end(s.data() + s.size())
    {}

    bool ok() const { return good; }

    bool rawVarint(uint64_t& out)
    {
        uint64_t v = 0;
        int shift = 0;
        while (p < end && shift < 64) {
            uint8_t b = (uint8_t)*p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if ((b & 0x80) == 0) {
                out = v;
                return true;
            }
            shift += 7;
        }
        if (shift > 0 || p > end) {
            good = false;
        }
        return false;
    }

    bool next(uint32_t& field, int& wireType)
    {
        if (p >= end) {
            return false;
        }
        uint64_t t;
        if (!rawVarint(t)) {
            good = false;
            return false;
        }
        field = (uint32_t)(t >> 3);
        wireType = (int)(t & 7);
        return true;
    }

    bool varint(int wireType, uint64_t& out)
    {
        if (wireType != 0 || !rawVarint(out)) {
            good = false;
            return false;
        }
        return true;
    }

    bool bytes(int wireType, std::string_view& out)
    {
        uint64_t len;
        if (wireType != 2 || !rawVarint(len) || (uint64_t)(end - p) < len) {
            good = false;
            return false;
        }
        out = std::string_view(p, (size_t)len);
        p += len;
        return true;
    }

    bool skip(int wireType)
    {
        uint64_t tmp;
        std::string_view sv;
        switch (wireType) {
            case 0:
                return varint(0, tmp);
            case 1:
                if (end - p < 8) {
                    good = false;
                    return false;
                }
                p += 8;
                return true;
            case 2:
                return bytes(2, sv);
            case 5:
                if (end - p < 4) {
                    good = false;
                    return false;
                }
                p += 4;
                return true;
            default:
                good = false;
                return false;
        }
    }

  private:
    const char* p;
    const char* end;
    bool good = true;
};

// ------------------------------------------------------------------ JSON ---
std::string base64Encode(std::string_view in);

std::string base64Decode(std::string_view in);

class JsonWriter
{
  public:
    void beginObject()
    {
        sep();
        out += '{';
        first.push_back(true);
    }

    void endObject()
    {
        out += '}';
        first.pop_back();
    }

    void beginArray()
    {
        sep();
        out += '[';
        first.push_back(true);
    }

    void endArray()
    {
        out += ']';
        first.pop_back();
    }

    void key(std::string_view k)
    {
        sep();
        quote(k);
        out += ':';
        afterKey = true;
    }

    void value(std::string_view s)
    {
        sep();
        quote(s);
    }

    void value(const std::string& s) { value(std::string_view(s)); }

    void value(const char* s) { value(std::string_view(s)); }

    void value(bool b)
    {
        sep();
        out += b ? "true" : "false";
    }

    void value(int64_t v)
    {
        sep();
        out += std::to_string(v);
    }

    void value(uint64_t v)
    {
        sep();
        out += std::to_string(v);
    }

    void value(int32_t v) { value((int64_t)v); }

    void value(uint32_t v) { value((uint64_t)v); }

    void value(double v)
    {
        sep();
        out += std::to_string(v);
    }

    void bytesValue(std::string_view b) { value(base64Encode(b)); }

    void raw(std::string_view json)
    {
        sep();
        out += json;
    }

    const std::string& str() const { return out; }

  private:
    std::string out;
    std::vector<bool> first;
    bool afterKey = false;

    void sep()
    {
        if (afterKey) {
            afterKey = false;
            return;
        }
        if (!first.empty()) {
            if (!first.back()) {
                out += ',';
            }
            first.back() = false;
        }
    }

    void quote(std::string_view s);
};

class JsonValue
{
  public:
    enum Kind
    {
        Null,
        Bool,
        Number,
        String,
        Array,
        Object
    };

    Kind kind = Null;

    bool isObject() const { return kind == Object; }
    bool isArray() const { return kind == Array; }
    bool isString() const { return kind == String; }
    bool isNumber() const { return kind == Number; }
    bool isBool() const { return kind == Bool; }
    bool isNull() const { return kind == Null; }

    // Numbers given as strings are accepted (protobuf JSON does that for
    // 64-bit ints)
    int64_t asInt() const;
    double asDouble() const;
    bool asBool() const;
    const std::string& asString() const { return str; }
    std::string asBytes() const { return base64Decode(str); }

    const std::vector<std::pair<std::string, JsonValue>>& members() const
    {
        return obj;
    }
    const std::vector<JsonValue>& elements() const { return arr; }

    const JsonValue* find(const std::string& key) const;

    // Throws std::runtime_error on malformed input
    static JsonValue parse(std::string_view text);

    bool boolean = false;
    double number = 0;
    bool numberIsInt = false;
    int64_t intValue = 0;
    std::string str;
    std::vector<JsonValue> arr;
    std::vector<std::pair<std::string, JsonValue>> obj;
};

} // namespace faabric::proto
