// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for folly::dynamic + folly::parseJson / folly::json::serialize as source/util/Camera.h/.cpp use them
// (rig JSON in and out).  A small recursive-descent JSON reader over a tagged value; numbers are kept as double
// (strtod, like folly's double parsing) plus an "was integral" flag.
#pragma once

#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace folly {

class dynamic {
 public:
  enum Kind { NUL, BOOL, NUMBER, STRING, ARRAY, OBJECT };
  dynamic() : kind_(NUL) {}
  dynamic(double d) : kind_(NUMBER), num_(d) {}
  dynamic(int i) : kind_(NUMBER), num_(i), integral_(true) {}
  dynamic(bool b) : kind_(BOOL), num_(b) {}
  dynamic(const char* s) : kind_(STRING), str_(s) {}
  dynamic(const std::string& s) : kind_(STRING), str_(s) {}

  // dynamic::array(a, b, ...) / dynamic::array(begin, end is not a folly API: Camera.h passes two pointers,
  // which folly treats as ... two elements? no: Camera::serializeVector relies on the iterator-pair constructor
  // of dynamic::array being selected — folly has `dynamic(Iterator first, Iterator last)`; mirrored here.
  static dynamic array() {
    dynamic d;
    d.kind_ = ARRAY;
    return d;
  }
  template <class It>
  static dynamic array(It first, It last) {
    dynamic d = array();
    for (; first != last; ++first) d.arr_.push_back(dynamic(*first));
    return d;
  }
  struct ObjectMaker;
  static ObjectMaker object(const std::string& k, const dynamic& v);
  static dynamic emptyObject() {
    dynamic d;
    d.kind_ = OBJECT;
    return d;
  }

  Kind kind() const { return kind_; }
  bool isObject() const { return kind_ == OBJECT; }
  bool isArray() const { return kind_ == ARRAY; }
  double asDouble() const {
    if (kind_ == NUMBER || kind_ == BOOL) return num_;
    if (kind_ == STRING) return std::strtod(str_.c_str(), nullptr);
    throw std::runtime_error("refshim: folly::dynamic is not a number");
  }
  const std::string& getString() const {
    if (kind_ != STRING) throw std::runtime_error("refshim: folly::dynamic is not a string");
    return str_;
  }
  size_t size() const { return kind_ == ARRAY ? arr_.size() : kind_ == OBJECT ? obj_.size() : 0; }
  size_t count(const std::string& k) const { return kind_ == OBJECT ? obj_.count(k) : 0; }
  const dynamic& operator[](const std::string& k) const {
    auto it = obj_.find(k);
    if (kind_ != OBJECT || it == obj_.end()) throw std::runtime_error("refshim: missing JSON key " + k);
    return it->second;
  }
  dynamic& operator[](const std::string& k) {
    if (kind_ != OBJECT) throw std::runtime_error("refshim: not an object");
    return obj_[k];
  }
  const dynamic& operator[](const char* k) const { return (*this)[std::string(k)]; }
  dynamic& operator[](const char* k) { return (*this)[std::string(k)]; }
  const dynamic& operator[](int i) const { return arr_.at((size_t)i); }
  const dynamic& operator[](size_t i) const { return arr_.at(i); }
  void push_back(const dynamic& d) {
    if (kind_ != ARRAY) throw std::runtime_error("refshim: not an array");
    arr_.push_back(d);
  }
  std::vector<dynamic>::const_iterator begin() const { return arr_.begin(); }
  std::vector<dynamic>::const_iterator end() const { return arr_.end(); }

  void write(std::ostream& os) const {
    switch (kind_) {
      case NUL: os << "null"; break;
      case BOOL: os << (num_ != 0 ? "true" : "false"); break;
      case NUMBER: {
        std::ostringstream t;
        t.precision(17);
        t << num_;
        os << t.str();
        break;
      }
      case STRING: os << '"' << str_ << '"'; break;
      case ARRAY: {
        os << "[";
        for (size_t i = 0; i < arr_.size(); ++i) {
          if (i) os << ",";
          arr_[i].write(os);
        }
        os << "]";
        break;
      }
      case OBJECT: {
        os << "{";
        bool first = true;
        for (const auto& kv : obj_) {
          if (!first) os << ",";
          first = false;
          os << '"' << kv.first << "\":";
          kv.second.write(os);
        }
        os << "}";
        break;
      }
    }
  }

 private:
  Kind kind_;
  double num_ = 0;
  bool integral_ = false;
  std::string str_;
  std::vector<dynamic> arr_;
  std::map<std::string, dynamic> obj_;
  friend class JsonReader;
};

struct dynamic::ObjectMaker {
  dynamic d;
  ObjectMaker& operator()(const std::string& k, const dynamic& v) {
    d[k] = v;
    return *this;
  }
  operator dynamic() const { return d; }
};
inline dynamic::ObjectMaker dynamic::object(const std::string& k, const dynamic& v) {
  ObjectMaker m;
  m.d = emptyObject();
  m.d[k] = v;
  return m;
}
inline std::ostream& operator<<(std::ostream& os, const dynamic& d) {
  d.write(os);
  return os;
}

class JsonReader {
 public:
  explicit JsonReader(const std::string& s) : s_(s) {}
  dynamic parse() {
    dynamic d = value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return d;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const char* what) const {
    throw std::runtime_error(std::string("refshim: JSON parse error: ") + what + " at offset " + std::to_string(i_));
  }
  void ws() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) ++i_;
  }
  dynamic value() {
    ws();
    if (i_ >= s_.size()) fail("unexpected end");
    const char c = s_[i_];
    if (c == '{') {
      dynamic d = dynamic::emptyObject();
      ++i_;
      ws();
      if (s_[i_] == '}') {
        ++i_;
        return d;
      }
      while (true) {
        ws();
        const std::string k = string();
        ws();
        if (s_[i_] != ':') fail("expected ':'");
        ++i_;
        d[k] = value();
        ws();
        if (s_[i_] == ',') {
          ++i_;
          continue;
        }
        if (s_[i_] == '}') {
          ++i_;
          return d;
        }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      dynamic d = dynamic::array();
      ++i_;
      ws();
      if (s_[i_] == ']') {
        ++i_;
        return d;
      }
      while (true) {
        d.push_back(value());
        ws();
        if (s_[i_] == ',') {
          ++i_;
          continue;
        }
        if (s_[i_] == ']') {
          ++i_;
          return d;
        }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') return dynamic(string());
    if (s_.compare(i_, 4, "true") == 0) {
      i_ += 4;
      return dynamic(true);
    }
    if (s_.compare(i_, 5, "false") == 0) {
      i_ += 5;
      return dynamic(false);
    }
    if (s_.compare(i_, 4, "null") == 0) {
      i_ += 4;
      return dynamic();
    }
    char* end = nullptr;
    const double v = std::strtod(s_.c_str() + i_, &end);
    if (end == s_.c_str() + i_) fail("unexpected character");
    i_ = (size_t)(end - s_.c_str());
    return dynamic(v);
  }
  std::string string() {
    if (s_[i_] != '"') fail("expected string");
    ++i_;
    std::string out;
    while (i_ < s_.size() && s_[i_] != '"') {
      if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
        ++i_;
        const char e = s_[i_];
        out += e == 'n' ? '\n' : e == 't' ? '\t' : e;
      } else {
        out += s_[i_];
      }
      ++i_;
    }
    if (i_ >= s_.size()) fail("unterminated string");
    ++i_;
    return out;
  }
};

inline dynamic parseJson(const std::string& s) { return JsonReader(s).parse(); }

}  // namespace folly
