// Minimal JSON reader for the tiny-cuda-nn style model config the reference builds in
// src/NeuralRadianceCache.cu:16-37 (objects, arrays, strings, numbers, booleans, null; no escapes beyond \" \\ \/ \n \t).
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json {

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;

    bool has(const std::string& k) const { return type == Object && obj.count(k) != 0; }
    const Value& at(const std::string& k) const {
        if (type != Object) throw std::runtime_error("json: not an object while looking up '" + k + "'");
        auto it = obj.find(k);
        if (it == obj.end()) throw std::runtime_error("json: missing key '" + k + "'");
        return it->second;
    }
    double number(const std::string& k, double dflt) const { return has(k) && at(k).type == Number ? at(k).num : dflt; }
    std::string string(const std::string& k, const std::string& dflt) const { return has(k) && at(k).type == String ? at(k).str : dflt; }
    bool boolean(const std::string& k, bool dflt) const {
        if (!has(k)) return dflt;
        const Value& v = at(k);
        return v.type == Bool ? v.b : v.type == Number ? v.num != 0 : dflt;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    Value parse() {
        Value v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_ = 0;
    [[noreturn]] void fail(const std::string& m) const { throw std::runtime_error("json: " + m + " at offset " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) p_++; }
    char peek() { ws(); if (p_ >= s_.size()) fail("unexpected end"); return s_[p_]; }
    void expect(char c) { if (peek() != c) fail(std::string("expected '") + c + "'"); p_++; }
    Value value() {
        char c = peek();
        Value v;
        if (c == '{') {
            v.type = Value::Object; p_++;
            if (peek() == '}') { p_++; return v; }
            while (true) {
                std::string k = string_lit();
                expect(':');
                v.obj[k] = value();
                char d = peek(); p_++;
                if (d == '}') break;
                if (d != ',') fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.type = Value::Array; p_++;
            if (peek() == ']') { p_++; return v; }
            while (true) {
                v.arr.push_back(value());
                char d = peek(); p_++;
                if (d == ']') break;
                if (d != ',') fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = Value::String; v.str = string_lit();
        } else if (s_.compare(p_, 4, "true") == 0) { v.type = Value::Bool; v.b = true; p_ += 4;
        } else if (s_.compare(p_, 5, "false") == 0) { v.type = Value::Bool; v.b = false; p_ += 5;
        } else if (s_.compare(p_, 4, "null") == 0) { p_ += 4;
        } else {
            const char* start = s_.c_str() + p_;
            char* end = nullptr;
            v.num = std::strtod(start, &end);
            if (end == start) fail("bad value");
            v.type = Value::Number;
            p_ += (size_t)(end - start);
        }
        return v;
    }
    std::string string_lit() {
        expect('"');
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c == '\\') {
                if (p_ >= s_.size()) fail("bad escape");
                char e = s_[p_++];
                out.push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e);
            } else out.push_back(c);
        }
        if (p_ >= s_.size()) fail("unterminated string");
        p_++;
        return out;
    }
};

inline Value parse(const std::string& s) { return Parser(s).parse(); }

}  // namespace mini_json
