// Minimal JSON reader (objects, arrays, strings, numbers, true/false/null) for the verification-key
// files -- the role serde_json plays at AL/operator/mina/lib/src/verifier_index.rs:115-116.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pasta {
namespace json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    const Value *get(const std::string &key) const {
        if (kind != Object) return nullptr;
        for (auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Value &at(const std::string &key) const {
        const Value *v = get(key);
        if (!v) throw std::runtime_error("json: missing key '" + key + "'");
        return *v;
    }
};

class Parser {
   public:
    explicit Parser(const std::string &s) : s_(s) {}
    Value parse() {
        Value v = value(0);
        ws();
        if (i_ != s_.size()) fail("trailing characters");
        return v;
    }

   private:
    [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("json: ") + what + " at offset " + std::to_string(i_)); }
    void ws() {
        while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) i_++;
    }
    bool lit(const char *w) {
        size_t n = std::char_traits<char>::length(w);
        if (s_.compare(i_, n, w) == 0) {
            i_ += n;
            return true;
        }
        return false;
    }
    std::string string() {
        if (s_[i_] != '"') fail("expected string");
        i_++;
        std::string out;
        while (i_ < s_.size() && s_[i_] != '"') {
            char c = s_[i_++];
            if (c == '\\') {
                if (i_ >= s_.size()) fail("bad escape");
                char e = s_[i_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u':
                        if (i_ + 4 > s_.size()) fail("bad \\u escape");
                        out += '?';  // non-ASCII never occurs in the key files; keep the parser total
                        i_ += 4;
                        break;
                    default: out += e; break;
                }
            } else {
                out += c;
            }
        }
        if (i_ >= s_.size()) fail("unterminated string");
        i_++;
        return out;
    }
    Value value(int depth) {
        if (depth > 64) fail("nesting too deep");
        ws();
        if (i_ >= s_.size()) fail("unexpected end");
        Value v;
        char c = s_[i_];
        if (c == '{') {
            v.kind = Value::Object;
            i_++;
            ws();
            if (i_ < s_.size() && s_[i_] == '}') {
                i_++;
                return v;
            }
            for (;;) {
                ws();
                std::string k = string();
                ws();
                if (i_ >= s_.size() || s_[i_] != ':') fail("expected ':'");
                i_++;
                v.obj.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (i_ < s_.size() && s_[i_] == ',') {
                    i_++;
                    continue;
                }
                if (i_ < s_.size() && s_[i_] == '}') {
                    i_++;
                    return v;
                }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v.kind = Value::Array;
            i_++;
            ws();
            if (i_ < s_.size() && s_[i_] == ']') {
                i_++;
                return v;
            }
            for (;;) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (i_ < s_.size() && s_[i_] == ',') {
                    i_++;
                    continue;
                }
                if (i_ < s_.size() && s_[i_] == ']') {
                    i_++;
                    return v;
                }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v.kind = Value::String;
            v.str = string();
            return v;
        }
        if (lit("true")) {
            v.kind = Value::Bool;
            v.b = true;
            return v;
        }
        if (lit("false")) {
            v.kind = Value::Bool;
            return v;
        }
        if (lit("null")) return v;
        char *end = nullptr;
        v.num = std::strtod(s_.c_str() + i_, &end);
        if (end == s_.c_str() + i_) fail("unexpected character");
        i_ = (size_t)(end - s_.c_str());
        v.kind = Value::Number;
        return v;
    }
    const std::string &s_;
    size_t i_ = 0;
};

inline Value parse(const std::string &text) { return Parser(text).parse(); }

}  // namespace json
}  // namespace pasta
