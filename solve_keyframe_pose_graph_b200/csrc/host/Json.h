// Minimal JSON value / parser / writer for the reference's on-disk formats.  The reference
// uses nlohmann::json (vendored, out of scope); the writer reproduces what `dump(4)` / `std::setw(4) << j` emit —
// keys in alphabetical order (nlohmann's object is a std::map), 4-space indent, integers as integers, doubles through
// the library's own Grisu2 + layout rule (Grisu2.h) — so files written here are byte-identical to what the reference's
// build writes for the same values (tests/test_reference_json.py checks that against the reference's vendored library).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "Grisu2.h"

namespace pgs {

class Json {
 public:
  enum Kind { Null, Bool, Int, Double, String, Array, Object };
  Json() : kind_(Null) {}
  Json(bool b) : kind_(Bool), b_(b) {}
  Json(int v) : kind_(Int), i_(v) {}
  Json(int64_t v) : kind_(Int), i_(v) {}
  Json(double v) : kind_(Double), d_(v) {}
  Json(const char* s) : kind_(String), s_(s) {}
  Json(const std::string& s) : kind_(String), s_(s) {}
  static Json array() { Json j; j.kind_ = Array; return j; }
  static Json object() { Json j; j.kind_ = Object; return j; }

  Kind kind() const { return kind_; }
  bool is_null() const { return kind_ == Null; }
  bool is_number() const { return kind_ == Int || kind_ == Double; }
  // obj["key"] creates (null becomes an object, as in nlohmann)
  Json& operator[](const std::string& k) { if (kind_ == Null) kind_ = Object; return o_[k]; }
  const Json& at(const std::string& k) const { static const Json none; auto it = o_.find(k); return it == o_.end() ? none : it->second; }
  bool contains(const std::string& k) const { return kind_ == Object && o_.count(k); }
  void push_back(const Json& v) { if (kind_ == Null) kind_ = Array; a_.push_back(v); }
  size_t size() const { return kind_ == Array ? a_.size() : (kind_ == Object ? o_.size() : 0); }
  const Json& operator[](size_t i) const { return a_[i]; }
  double as_double() const { return kind_ == Int ? (double)i_ : (kind_ == Double ? d_ : 0.0); }
  int64_t as_int() const { return kind_ == Int ? i_ : (kind_ == Double ? (int64_t)std::llround(d_) : 0); }
  bool as_bool() const { return kind_ == Bool ? b_ : false; }
  const std::string& as_string() const { return s_; }

  std::string dump(int indent = 4) const { std::string out; write(out, indent, 0); return out; }

  // returns false on malformed input; *err (if given) says where
  static bool parse(const std::string& text, Json* out, std::string* err = nullptr) {
    size_t p = 0;
    if (!parse_value(text, p, out, err)) return false;
    skip(text, p);
    if (p != text.size()) { if (err) *err = "trailing characters at offset " + std::to_string(p); return false; }
    return true;
  }

 private:
  Kind kind_;
  bool b_ = false; int64_t i_ = 0; double d_ = 0.0; std::string s_;
  std::vector<Json> a_; std::map<std::string, Json> o_;

  static void write_string(std::string& out, const std::string& s) {
    out.push_back('"');
    for (unsigned char c : s) {
      switch (c) {
        case '"': out += "\\\""; break; case '\\': out += "\\\\"; break; case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break; case '\t': out += "\\t"; break; case '\b': out += "\\b"; break; case '\f': out += "\\f"; break;
        default: if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); out += b; } else out.push_back((char)c);
      }
    }
    out.push_back('"');
  }
  void write(std::string& out, int indent, int level) const {
    const std::string pad((size_t)indent * (level + 1), ' '), pad0((size_t)indent * level, ' ');
    switch (kind_) {
      case Null: out += "null"; break;
      case Bool: out += b_ ? "true" : "false"; break;
      case Int: out += std::to_string(i_); break;
      case Double: {
        if (!std::isfinite(d_)) { out += "null"; break; }          // nlohmann writes null for NaN / inf
        out += grisu2::to_string(d_); break;
      }
      case String: write_string(out, s_); break;
      case Array:
        if (a_.empty()) { out += "[]"; break; }
        out += "[\n";
        for (size_t i = 0; i < a_.size(); ++i) { out += pad; a_[i].write(out, indent, level + 1); out += i + 1 < a_.size() ? ",\n" : "\n"; }
        out += pad0 + "]"; break;
      case Object: {
        if (o_.empty()) { out += "{}"; break; }
        out += "{\n";
        size_t i = 0;
        for (const auto& kv : o_) { out += pad; write_string(out, kv.first); out += ": "; kv.second.write(out, indent, level + 1); out += ++i < o_.size() ? ",\n" : "\n"; }
        out += pad0 + "}"; break;
      }
    }
  }
  static void skip(const std::string& t, size_t& p) { while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\r' || t[p] == '\t')) ++p; }
  static bool fail(std::string* err, const char* what, size_t p) { if (err) *err = std::string(what) + " at offset " + std::to_string(p); return false; }
  static bool parse_string(const std::string& t, size_t& p, std::string* out, std::string* err) {
    if (t[p] != '"') return fail(err, "expected string", p);
    ++p; out->clear();
    while (p < t.size() && t[p] != '"') {
      if (t[p] == '\\') {
        if (++p >= t.size()) return fail(err, "bad escape", p);
        switch (t[p]) {
          case 'n': out->push_back('\n'); break; case 't': out->push_back('\t'); break; case 'r': out->push_back('\r'); break;
          case 'b': out->push_back('\b'); break; case 'f': out->push_back('\f'); break;
          case 'u': { if (p + 4 >= t.size()) return fail(err, "bad \\u", p); unsigned v = (unsigned)std::stoul(t.substr(p + 1, 4), nullptr, 16); p += 4;
                      if (v < 0x80) out->push_back((char)v); else if (v < 0x800) { out->push_back((char)(0xC0 | (v >> 6))); out->push_back((char)(0x80 | (v & 0x3F))); }
                      else { out->push_back((char)(0xE0 | (v >> 12))); out->push_back((char)(0x80 | ((v >> 6) & 0x3F))); out->push_back((char)(0x80 | (v & 0x3F))); } break; }
          default: out->push_back(t[p]);
        }
        ++p;
      } else out->push_back(t[p++]);
    }
    if (p >= t.size()) return fail(err, "unterminated string", p);
    ++p; return true;
  }
  static bool parse_value(const std::string& t, size_t& p, Json* out, std::string* err) {
    skip(t, p);
    if (p >= t.size()) return fail(err, "unexpected end", p);
    const char c = t[p];
    if (c == '{') {
      *out = object(); ++p; skip(t, p);
      if (p < t.size() && t[p] == '}') { ++p; return true; }
      for (;;) {
        skip(t, p); std::string k;
        if (p >= t.size() || !parse_string(t, p, &k, err)) return false;
        skip(t, p); if (p >= t.size() || t[p] != ':') return fail(err, "expected ':'", p); ++p;
        Json v; if (!parse_value(t, p, &v, err)) return false;
        out->o_[k] = v; skip(t, p);
        if (p < t.size() && t[p] == ',') { ++p; continue; }
        if (p < t.size() && t[p] == '}') { ++p; return true; }
        return fail(err, "expected ',' or '}'", p);
      }
    }
    if (c == '[') {
      *out = array(); ++p; skip(t, p);
      if (p < t.size() && t[p] == ']') { ++p; return true; }
      for (;;) {
        Json v; if (!parse_value(t, p, &v, err)) return false;
        out->a_.push_back(v); skip(t, p);
        if (p < t.size() && t[p] == ',') { ++p; continue; }
        if (p < t.size() && t[p] == ']') { ++p; return true; }
        return fail(err, "expected ',' or ']'", p);
      }
    }
    if (c == '"') { std::string s; if (!parse_string(t, p, &s, err)) return false; *out = Json(s); return true; }
    if (t.compare(p, 4, "true") == 0) { *out = Json(true); p += 4; return true; }
    if (t.compare(p, 5, "false") == 0) { *out = Json(false); p += 5; return true; }
    if (t.compare(p, 4, "null") == 0) { *out = Json(); p += 4; return true; }
    const size_t s0 = p; bool is_int = true;
    if (p < t.size() && (t[p] == '-' || t[p] == '+')) ++p;
    while (p < t.size() && (isdigit((unsigned char)t[p]) || t[p] == '.' || t[p] == 'e' || t[p] == 'E' || t[p] == '-' || t[p] == '+')) { if (!isdigit((unsigned char)t[p])) is_int = false; ++p; }
    if (p == s0) return fail(err, "unexpected character", p);
    const std::string num = t.substr(s0, p - s0);
    try { if (is_int) *out = Json((int64_t)std::stoll(num)); else *out = Json(std::stod(num)); } catch (...) { try { *out = Json(std::stod(num)); } catch (...) { return fail(err, "bad number", s0); } }
    return true;
  }
};

}  // namespace pgs
