// Minimal JSON reader for the cells' parameter strings (search_json_params, db, object id lists).
// Handles objects, arrays, strings, numbers, true/false/null — enough for conf/detection.ork's `search:` subtree
// as ORK core hands it to DescriptorMatcher::configure (reference: src/detection/DescriptorMatcher.cpp:159-181,
// which uses or_json / json_spirit from the un-vendored object_recognition_core).
#ifndef TOD_JSON_MIN_H_
#define TOD_JSON_MIN_H_

#include <cctype>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace tod {
namespace json {

struct Value {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Value> arr;
  std::map<std::string, Value> obj;

  bool has(const std::string &k) const { return type == Object && obj.count(k) != 0; }
  const Value &at(const std::string &k) const {
    static const Value null_value;
    auto it = obj.find(k);
    return it == obj.end() ? null_value : it->second;
  }
};

// true / non-zero number
inline bool truthy(const Value &v) { return (v.type == Value::Bool && v.b) || (v.type == Value::Number && v.num != 0); }

class Parser {
 public:
  explicit Parser(const std::string &s) : s_(s) {}
  bool parse(Value &out) {
    skip();
    if (!value(out)) return false;
    skip();
    return i_ == s_.size();
  }

 private:
  void skip() {
    while (i_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[i_]))) ++i_;
  }
  bool lit(const char *w) {
    size_t n = std::char_traits<char>::length(w);
    if (s_.compare(i_, n, w) != 0) return false;
    i_ += n;
    return true;
  }
  bool string(std::string &out) {
    if (i_ >= s_.size() || (s_[i_] != '"' && s_[i_] != '\'')) return false;
    const char quote = s_[i_++];
    out.clear();
    while (i_ < s_.size() && s_[i_] != quote) {
      char c = s_[i_++];
      if (c == '\\' && i_ < s_.size()) {
        char e = s_[i_++];
        switch (e) {
          case 'n': c = '\n'; break;
          case 't': c = '\t'; break;
          case 'r': c = '\r'; break;
          case 'b': c = '\b'; break;
          case 'f': c = '\f'; break;
          default: c = e; break;
        }
      }
      out.push_back(c);
    }
    if (i_ >= s_.size()) return false;
    ++i_;
    return true;
  }
  bool value(Value &v) {
    skip();
    if (i_ >= s_.size()) return false;
    const char c = s_[i_];
    if (c == '{') {
      ++i_;
      v.type = Value::Object;
      skip();
      if (i_ < s_.size() && s_[i_] == '}') { ++i_; return true; }
      while (true) {
        skip();
        std::string key;
        if (!string(key)) return false;
        skip();
        if (i_ >= s_.size() || s_[i_] != ':') return false;
        ++i_;
        Value child;
        if (!value(child)) return false;
        v.obj[key] = child;
        skip();
        if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
        if (i_ < s_.size() && s_[i_] == '}') { ++i_; return true; }
        return false;
      }
    }
    if (c == '[') {
      ++i_;
      v.type = Value::Array;
      skip();
      if (i_ < s_.size() && s_[i_] == ']') { ++i_; return true; }
      while (true) {
        Value child;
        if (!value(child)) return false;
        v.arr.push_back(child);
        skip();
        if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
        if (i_ < s_.size() && s_[i_] == ']') { ++i_; return true; }
        return false;
      }
    }
    if (c == '"' || c == '\'') {
      v.type = Value::String;
      return string(v.str);
    }
    if (lit("true")) { v.type = Value::Bool; v.b = true; return true; }
    if (lit("false")) { v.type = Value::Bool; v.b = false; return true; }
    if (lit("null")) { v.type = Value::Null; return true; }
    char *end = nullptr;
    const double d = std::strtod(s_.c_str() + i_, &end);
    if (end == s_.c_str() + i_) return false;
    i_ = size_t(end - s_.c_str());
    v.type = Value::Number;
    v.num = d;
    return true;
  }

  const std::string &s_;
  size_t i_ = 0;
};

inline bool parse(const std::string &s, Value &out) { return Parser(s).parse(out); }

}  // namespace json
}  // namespace tod
#endif
