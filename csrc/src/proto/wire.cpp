#include <faabric/proto/wire.h>

#include <cctype>
#include <cmath>
#include <cstdlib>

namespace faabric::proto {

static const char* B64 =
  "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

std::string base64Encode(std::string_view in)
{
    std::string out;
    out.reserve((in.size() + 2) / 3 * 4);
    size_t i = 0;
    while (i + 2 < in.size()) {
        uint32_t v = ((uint8_t)in[i] << 16) | ((uint8_t)in[i + 1] << 8) |
                     (uint8_t)in[i + 2];
        out += B64[(v >> 18) & 63];
        out += B64[(v >> 12) & 63];
        out += B64[(v >> 6) & 63];
        out += B64[v & 63];
        i += 3;
    }
    if (i + 1 == in.size()) {
        uint32_t v = (uint8_t)in[i] << 16;
        out += B64[(v >> 18) & 63];
        out += B64[(v >> 12) & 63];
        out += "==";
    } else if (i + 2 == in.size()) {
        uint32_t v = ((uint8_t)in[i] << 16) | ((uint8_t)in[i + 1] << 8);
        out += B64[(v >> 18) & 63];
        out += B64[(v >> 12) & 63];
        out += B64[(v >> 6) & 63];
        out += '=';
    }
    return out;
}

std::string base64Decode(std::string_view in)
{
    static int8_t table[256];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; i++) {
            table[i] = -1;
        }
        for (int i = 0; i < 64; i++) {
            table[(uint8_t)B64[i]] = (int8_t)i;
        }
        // url-safe alphabet too
        table[(uint8_t)'-'] = 62;
        table[(uint8_t)'_'] = 63;
        init = true;
    }
    std::string out;
    uint32_t acc = 0;
    int bits = 0;
    for (char c : in) {
        int8_t v = table[(uint8_t)c];
        if (v < 0) {
            continue; // padding / whitespace
        }
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out += (char)((acc >> bits) & 0xff);
        }
    }
    return out;
}

void JsonWriter::quote(std::string_view s)
{
    out += '"';
    for (char ch : s) {
        unsigned char c = (unsigned char)ch;
        switch (c) {
            case '"':
                out += "\\\"";
                break;
            case '\\':
                out += "\\\\";
                break;
            case '\n':
                out += "\\n";
                break;
            case '\r':
                out += "\\r";
                break;
            case '\t':
                out += "\\t";
                break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    snprintf(buf, sizeof(buf), "\\u%04x", c);
                    out += buf;
                } else {
                    out += ch;
                }
        }
    }
    out += '"';
}

int64_t JsonValue::asInt() const
{
    switch (kind) {
        case Number:
            return numberIsInt ? intValue : (int64_t)number;
        case String:
            return strtoll(str.c_str(), nullptr, 10);
        case Bool:
            return boolean ? 1 : 0;
        default:
            return 0;
    }
}

double JsonValue::asDouble() const
{
    switch (kind) {
        case Number:
            return number;
        case String:
            return strtod(str.c_str(), nullptr);
        default:
            return 0;
    }
}

bool JsonValue::asBool() const
{
    switch (kind) {
        case Bool:
            return boolean;
        case Number:
            return number != 0;
        case String:
            return str == "true" || str == "1";
        default:
            return false;
    }
}

const JsonValue* JsonValue::find(const std::string& key) const
{
    for (const auto& kv : obj) {
        if (kv.first == key) {
            return &kv.second;
        }
    }
    return nullptr;
}

namespace {
struct Parser
{
    const char* p;
    const char* end;

    [[noreturn]] void fail(const char* what)
    {
        throw std::runtime_error(std::string("JSON parse error: ") + what);
    }

    void ws()
    {
        while (p < end && isspace((unsigned char)*p)) {
            p++;
        }
    }

    std::string parseString()
    {
        if (p >= end || *p != '"') {
            fail("expected string");
        }
        p++;
        std::string out;
        while (p < end && *p != '"') {
            char c = *p++;
            if (c == '\\') {
                if (p >= end) {
                    fail("bad escape");
                }
                char e = *p++;
                switch (e) {
                    case 'n':
                        out += '\n';
                        break;
                    case 't':
                        out += '\t';
                        break;
                    case 'r':
                        out += '\r';
                        break;
                    case 'b':
                        out += '\b';
                        break;
                    case 'f':
                        out += '\f';
                        break;
                    case 'u': {
                        if (end - p < 4) {
                            fail("bad unicode escape");
                        }
                        unsigned cp = (unsigned)strtoul(std::string(p, 4).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) {
                            out += (char)cp;
                        } else if (cp < 0x800) {
                            out += (char)(0xc0 | (cp >> 6));
                            out += (char)(0x80 | (cp & 0x3f));
                        } else {
                            out += (char)(0xe0 | (cp >> 12));
                            out += (char)(0x80 | ((cp >> 6) & 0x3f));
                            out += (char)(0x80 | (cp & 0x3f));
                        }
                        break;
                    }
                    default:
                        out += e;
                }
            } else {
                out += c;
            }
        }
        if (p >= end) {
            fail("unterminated string");
        }
        p++;
        return out;
    }

    JsonValue parseValue(int depth)
    {
        if (depth > 64) {
            fail("too deep");
        }
        ws();
        if (p >= end) {
            fail("unexpected end");
        }
        JsonValue v;
        char c = *p;
        if (c == '{') {
            p++;
            v.kind = JsonValue::Object;
            ws();
            if (p < end && *p == '}') {
                p++;
                return v;
            }
            while (true) {
                ws();
                std::string k = parseString();
                ws();
                if (p >= end || *p != ':') {
                    fail("expected ':'");
                }
                p++;
                v.obj.emplace_back(std::move(k), parseValue(depth + 1));
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == '}') {
                    p++;
                    return v;
                }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            p++;
            v.kind = JsonValue::Array;
            ws();
            if (p < end && *p == ']') {
                p++;
                return v;
            }
            while (true) {
                v.arr.push_back(parseValue(depth + 1));
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == ']') {
                    p++;
                    return v;
                }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v.kind = JsonValue::String;
            v.str = parseString();
            return v;
        }
        if (end - p >= 4 && strncmp(p, "true", 4) == 0) {
            p += 4;
            v.kind = JsonValue::Bool;
            v.boolean = true;
            return v;
        }
        if (end - p >= 5 && strncmp(p, "false", 5) == 0) {
            p += 5;
            v.kind = JsonValue::Bool;
            return v;
        }
        if (end - p >= 4 && strncmp(p, "null", 4) == 0) {
            p += 4;
            return v;
        }
        // number
        const char* start = p;
        bool isInt = true;
        if (p < end && (*p == '-' || *p == '+')) {
            p++;
        }
        while (p < end && (isdigit((unsigned char)*p) || *p == '.' || *p == 'e' ||
                           *p == 'E' || *p == '-' || *p == '+')) {
            if (*p == '.' || *p == 'e' || *p == 'E') {
                isInt = false;
            }
            p++;
        }
        if (p == start) {
            fail("unexpected character");
        }
        std::string num(start, p);
        v.kind = JsonValue::Number;
        v.number = strtod(num.c_str(), nullptr);
        v.numberIsInt = isInt;
        if (isInt) {
            v.intValue = strtoll(num.c_str(), nullptr, 10);
        }
        return v;
    }
};
}

JsonValue JsonValue::parse(std::string_view text)
{
    Parser ps{ text.data(), text.data() + text.size() };
    JsonValue v = ps.parseValue(0);
    ps.ws();
    if (ps.p != ps.end) {
        ps.fail("trailing characters");
    }
    return v;
}

} // namespace faabric::proto
