/*
 * Small JSON document model, parser and writer for the host side: the scene file format of the reference
 * (src/lib/vengine/core/io/Import.cpp, Export.cpp — rapidjson there) and glTF 2.0 (io_gltf.cpp).
 * Objects keep their member order; numbers are doubles (ints up to 2^53 are exact, which covers glTF
 * byte offsets); strings are UTF-8 with \uXXXX escapes (surrogate pairs included) decoded on parse.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace vengine {
namespace json {

class Value {
public:
    enum class Type { Null, Bool, Number, String, Array, Object };
    typedef std::pair<std::string, Value> Member;

    Value() {}
    Value(bool b) : m_type(Type::Bool), m_bool(b) {}
    Value(double d) : m_type(Type::Number), m_num(d) {}
    Value(float d) : m_type(Type::Number), m_single(true), m_num((double)d) {}
    Value(int d) : m_type(Type::Number), m_num(d) {}
    Value(unsigned d) : m_type(Type::Number), m_num(d) {}
    Value(const char *s) : m_type(Type::String), m_str(s) {}
    Value(const std::string &s) : m_type(Type::String), m_str(s) {}
    static Value array() {
        Value v;
        v.m_type = Type::Array;
        return v;
    }
    static Value object() {
        Value v;
        v.m_type = Type::Object;
        return v;
    }

    Type type() const { return m_type; }
    bool isNull() const { return m_type == Type::Null; }
    bool isBool() const { return m_type == Type::Bool; }
    bool isNumber() const { return m_type == Type::Number; }
    bool isString() const { return m_type == Type::String; }
    bool isArray() const { return m_type == Type::Array; }
    bool isObject() const { return m_type == Type::Object; }

    bool getBool() const {
        need(Type::Bool, "bool");
        return m_bool;
    }
    double getDouble() const {
        need(Type::Number, "number");
        return m_num;
    }
    float getFloat() const { return (float)getDouble(); }
    int64_t getInt() const { return (int64_t)getDouble(); }
    const std::string &getString() const {
        need(Type::String, "string");
        return m_str;
    }

    /* arrays */
    size_t size() const { return m_type == Type::Array ? m_arr.size() : (m_type == Type::Object ? m_obj.size() : 0); }
    const Value &operator[](size_t i) const {
        need(Type::Array, "array");
        if (i >= m_arr.size()) throw std::runtime_error("json: array index out of range");
        return m_arr[i];
    }
    const Value &operator[](int i) const { return (*this)[(size_t)i]; }
    Value &push(Value v) {
        need(Type::Array, "array");
        m_arr.push_back(std::move(v));
        return m_arr.back();
    }
    const std::vector<Value> &items() const { return m_arr; }

    /* objects */
    bool has(const std::string &key) const { return find(key) != nullptr; }
    const Value *find(const std::string &key) const {
        if (m_type != Type::Object) return nullptr;
        for (const Member &m : m_obj)
            if (m.first == key) return &m.second;
        return nullptr;
    }
    const Value &operator[](const std::string &key) const {
        const Value *v = find(key);
        if (!v) throw std::runtime_error("json: missing member \"" + key + "\"");
        return *v;
    }
    const Value &operator[](const char *key) const { return (*this)[std::string(key)]; }
    Value &set(const std::string &key, Value v) {
        need(Type::Object, "object");
        for (Member &m : m_obj)
            if (m.first == key) {
                m.second = std::move(v);
                return m.second;
            }
        m_obj.emplace_back(key, std::move(v));
        return m_obj.back().second;
    }
    const std::vector<Member> &members() const { return m_obj; }

    /* convenience getters with defaults */
    double number(const std::string &key, double def) const {
        const Value *v = find(key);
        return v && v->isNumber() ? v->m_num : def;
    }
    std::string string(const std::string &key, const std::string &def) const {
        const Value *v = find(key);
        return v && v->isString() ? v->m_str : def;
    }
    bool boolean(const std::string &key, bool def) const {
        const Value *v = find(key);
        return v && v->isBool() ? v->m_bool : def;
    }

    bool isSingle() const { return m_single; }

private:
    void need(Type t, const char *what) const {
        if (m_type != t) throw std::runtime_error(std::string("json: value is not a ") + what);
    }
    Type m_type = Type::Null;
    bool m_bool = false;
    bool m_single = false; /* written with float precision (9 significant digits) */
    double m_num = 0.0;
    std::string m_str;
    std::vector<Value> m_arr;
    std::vector<Member> m_obj;
};

/* ------------------------------------------------------------------ parser */
class Parser {
public:
    Parser(const char *begin, const char *end) : p(begin), e(end), b(begin) {}
    Value parseDocument() {
        /* UTF-8 byte order mark */
        if (e - p >= 3 && (uint8_t)p[0] == 0xEF && (uint8_t)p[1] == 0xBB && (uint8_t)p[2] == 0xBF) p += 3;
        Value v = parseValue(0);
        skipWs();
        if (p != e) fail("trailing characters");
        return v;
    }

private:
    const char *p, *e, *b;
    [[noreturn]] void fail(const char *msg) const {
        throw std::runtime_error("json: " + std::string(msg) + " at byte " + std::to_string((long long)(p - b)));
    }
    void skipWs() {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    bool literal(const char *s) {
        size_t n = std::strlen(s);
        if ((size_t)(e - p) >= n && std::memcmp(p, s, n) == 0) {
            p += n;
            return true;
        }
        return false;
    }
    static void appendUtf8(std::string &out, uint32_t cp) {
        if (cp < 0x80) {
            out.push_back((char)cp);
        } else if (cp < 0x800) {
            out.push_back((char)(0xC0 | (cp >> 6)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char)(0xF0 | (cp >> 18)));
            out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    uint32_t hex4() {
        if (e - p < 4) fail("truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string parseString() {
        if (p >= e || *p != '"') fail("expected string");
        ++p;
        std::string out;
        /* long strings (base64 payloads of glTF) have no escapes: copy runs */
        while (true) {
            const char *run = p;
            while (p < e && *p != '"' && *p != '\\' && (uint8_t)*p >= 0x20) ++p;
            out.append(run, p);
            if (p >= e) fail("unterminated string");
            if (*p == '"') {
                ++p;
                return out;
            }
            if ((uint8_t)*p < 0x20) fail("control character in string");
            ++p; /* backslash */
            if (p >= e) fail("unterminated escape");
            char c = *p++;
            switch (c) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF) {
                        if (e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            uint32_t lo = hex4();
                            if (lo < 0xDC00 || lo > 0xDFFF) fail("bad surrogate pair");
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        } else {
                            fail("lone surrogate");
                        }
                    }
                    appendUtf8(out, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
    }
    Value parseNumber() {
        const char *s = p;
        if (p < e && *p == '-') ++p;
        if (p >= e) fail("bad number");
        if (*p == '0') {
            ++p;
        } else if (*p >= '1' && *p <= '9') {
            while (p < e && *p >= '0' && *p <= '9') ++p;
        } else {
            fail("bad number");
        }
        if (p < e && *p == '.') {
            ++p;
            if (p >= e || *p < '0' || *p > '9') fail("bad fraction");
            while (p < e && *p >= '0' && *p <= '9') ++p;
        }
        if (p < e && (*p == 'e' || *p == 'E')) {
            ++p;
            if (p < e && (*p == '+' || *p == '-')) ++p;
            if (p >= e || *p < '0' || *p > '9') fail("bad exponent");
            while (p < e && *p >= '0' && *p <= '9') ++p;
        }
        std::string tmp(s, p);
        return Value(std::strtod(tmp.c_str(), nullptr));
    }
    Value parseValue(int depth) {
        if (depth > 256) fail("nesting too deep");
        skipWs();
        if (p >= e) fail("unexpected end");
        switch (*p) {
            case '{': {
                ++p;
                Value v = Value::object();
                skipWs();
                if (p < e && *p == '}') {
                    ++p;
                    return v;
                }
                while (true) {
                    skipWs();
                    std::string key = parseString();
                    skipWs();
                    if (p >= e || *p != ':') fail("expected ':'");
                    ++p;
                    v.set(key, parseValue(depth + 1));
                    skipWs();
                    if (p < e && *p == ',') {
                        ++p;
                        continue;
                    }
                    if (p < e && *p == '}') {
                        ++p;
                        return v;
                    }
                    fail("expected ',' or '}'");
                }
            }
            case '[': {
                ++p;
                Value v = Value::array();
                skipWs();
                if (p < e && *p == ']') {
                    ++p;
                    return v;
                }
                while (true) {
                    v.push(parseValue(depth + 1));
                    skipWs();
                    if (p < e && *p == ',') {
                        ++p;
                        continue;
                    }
                    if (p < e && *p == ']') {
                        ++p;
                        return v;
                    }
                    fail("expected ',' or ']'");
                }
            }
            case '"': return Value(parseString());
            case 't':
                if (literal("true")) return Value(true);
                fail("bad literal");
            case 'f':
                if (literal("false")) return Value(false);
                fail("bad literal");
            case 'n':
                if (literal("null")) return Value();
                fail("bad literal");
            default: return parseNumber();
        }
    }
};

inline Value parse(const std::string &text) {
    Parser p(text.data(), text.data() + text.size());
    return p.parseDocument();
}

/* ------------------------------------------------------------------ writer */
inline void writeString(std::string &out, const std::string &s) {
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    std::snprintf(buf, sizeof(buf), "\\u%04X", c);
                    out += buf;
                } else {
                    out.push_back((char)c);
                }
        }
    }
    out.push_back('"');
}

inline void writeValue(std::string &out, const Value &v, int indent, int level) {
    auto newline = [&](int lv) {
        if (indent <= 0) return;
        out.push_back('\n');
        out.append((size_t)(indent * lv), ' ');
    };
    switch (v.type()) {
        case Value::Type::Null: out += "null"; break;
        case Value::Type::Bool: out += v.getBool() ? "true" : "false"; break;
        case Value::Type::Number: {
            double d = v.getDouble();
            char buf[40];
            if (!std::isfinite(d)) {
                out += "null"; /* JSON has no inf / nan */
            } else if (d == std::floor(d) && std::fabs(d) < 1e15) {
                std::snprintf(buf, sizeof(buf), v.isSingle() ? "%.1f" : "%.0f", d);
                out += buf;
            } else {
                /* shortest representation that round-trips the stored precision */
                std::snprintf(buf, sizeof(buf), v.isSingle() ? "%.9g" : "%.17g", d);
                out += buf;
            }
            break;
        }
        case Value::Type::String: writeString(out, v.getString()); break;
        case Value::Type::Array: {
            if (v.size() == 0) {
                out += "[]";
                break;
            }
            out.push_back('[');
            bool first = true;
            for (const Value &it : v.items()) {
                if (!first) out.push_back(',');
                first = false;
                newline(level + 1);
                writeValue(out, it, indent, level + 1);
            }
            newline(level);
            out.push_back(']');
            break;
        }
        case Value::Type::Object: {
            if (v.size() == 0) {
                out += "{}";
                break;
            }
            out.push_back('{');
            bool first = true;
            for (const Value::Member &m : v.members()) {
                if (!first) out.push_back(',');
                first = false;
                newline(level + 1);
                writeString(out, m.first);
                out += indent > 0 ? ": " : ":";
                writeValue(out, m.second, indent, level + 1);
            }
            newline(level);
            out.push_back('}');
            break;
        }
    }
}

inline std::string dump(const Value &v, int indent = 4) {
    std::string out;
    writeValue(out, v, indent, 0);
    return out;
}

}  // namespace json
}  // namespace vengine
