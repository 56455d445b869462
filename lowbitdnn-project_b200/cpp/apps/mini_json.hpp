// mini_json.hpp — a small JSON reader/writer for the benchmark app (the reference uses Boost.PropertyTree,
// cpp/apps/benchmark.cpp:10-11; Boost is not a dependency here).  Supports objects, arrays, strings, numbers,
// true/false/null — enough for cpp/apps/config.json.
#pragma once
#include <cctype>
#include <cstdio>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;   // insertion order kept

    const Value& at(const std::string& k) const
    {
        for (auto& kv : obj) if (kv.first == k) return kv.second;
        throw std::runtime_error("missing key '" + k + "'");
    }
    bool has(const std::string& k) const { for (auto& kv : obj) if (kv.first == k) return true; return false; }
    // scalars may be written as a single value or as a list (config.json mixes both)
    std::vector<double> numbers() const
    {
        std::vector<double> out;
        if (kind == Number) out.push_back(num);
        else for (auto& v : arr) out.push_back(v.num);
        return out;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    Value parse() { Value v = value(); ws(); if (i_ != s_.size()) fail("trailing characters"); return v; }
private:
    const std::string& s_;
    size_t i_ = 0;
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m + " at offset " + std::to_string(i_)); }
    void ws() { while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_; }
    char peek() { ws(); if (i_ >= s_.size()) fail("unexpected end"); return s_[i_]; }
    Value value()
    {
        char c = peek();
        Value v;
        if (c == '{') {
            v.kind = Value::Object; ++i_;
            if (peek() == '}') { ++i_; return v; }
            while (true) {
                if (peek() != '"') fail("expected key");
                std::string k = string();
                if (peek() != ':') fail("expected ':'");
                ++i_;
                v.obj.emplace_back(k, value());
                char d = peek(); ++i_;
                if (d == '}') break;
                if (d != ',') fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = Value::Array; ++i_;
            if (peek() == ']') { ++i_; return v; }
            while (true) {
                v.arr.push_back(value());
                char d = peek(); ++i_;
                if (d == ']') break;
                if (d != ',') fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.kind = Value::String; v.str = string();
        } else if (s_.compare(i_, 4, "true") == 0) { v.kind = Value::Bool; v.b = true; i_ += 4; }
        else if (s_.compare(i_, 5, "false") == 0) { v.kind = Value::Bool; i_ += 5; }
        else if (s_.compare(i_, 4, "null") == 0) { i_ += 4; }
        else {
            size_t n = 0;
            v.kind = Value::Number;
            try { v.num = std::stod(s_.substr(i_), &n); } catch (...) { fail("bad number"); }
            i_ += n;
        }
        return v;
    }
    std::string string()
    {
        std::string out; ++i_;
        while (i_ < s_.size() && s_[i_] != '"') {
            if (s_[i_] == '\\' && i_ + 1 < s_.size()) { ++i_; out += (s_[i_] == 'n' ? '\n' : s_[i_]); }
            else out += s_[i_];
            ++i_;
        }
        if (i_ >= s_.size()) fail("unterminated string");
        ++i_;
        return out;
    }
};

inline std::string quote(const std::string& s)
{
    std::string o = "\"";
    for (char c : s) { if (c == '"' || c == '\\') o += '\\'; o += c; }
    return o + "\"";
}

// compact serialisation (numbers that are whole print without a fraction)
inline std::string dump(const Value& v)
{
    switch (v.kind) {
        case Value::Null: return "null";
        case Value::Bool: return v.b ? "true" : "false";
        case Value::Number: {
            char buf[64];
            if (v.num == (double)(long long)v.num) std::snprintf(buf, sizeof buf, "%lld", (long long)v.num);
            else std::snprintf(buf, sizeof buf, "%.17g", v.num);
            return buf;
        }
        case Value::String: return quote(v.str);
        case Value::Array: {
            std::string o = "[";
            for (size_t i = 0; i < v.arr.size(); ++i) o += (i ? ", " : "") + dump(v.arr[i]);
            return o + "]";
        }
        default: {
            std::string o = "{";
            for (size_t i = 0; i < v.obj.size(); ++i) o += (i ? ", " : "") + quote(v.obj[i].first) + ": " + dump(v.obj[i].second);
            return o + "}";
        }
    }
}

}  // namespace mini_json
