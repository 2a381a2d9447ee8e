// ingest.hpp — control-plane ingest batching, C++ side.
//
// Restates Endpoints.PostMetric (reference control/endpoints.go:109-129): the request body is decoded like
// json.Unmarshal into map[string]interface{}, every key becomes events.Event{Metric, fmt.Sprintf("%v|%v", key, value)},
// and — the point of this row — the whole request is handed to the bus as ONE batch (EventBus::PublishMany -> one
// cpbus_publish call per batch_cap events) instead of one Publish per key.  HTTP and the unix socket stay in the control
// plane.  The Python twin is containerpilot_b200/ingest.py; both are pinned to control/endpoints_test.go:104-145.
#pragma once
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "events.hpp"

namespace events {
namespace ingest {

constexpr int StatusOK = 200, StatusUnprocessableEntity = 422;   // net/http

// What encoding/json puts into an interface{}: nil, bool, float64, string, []interface{}, map[string]interface{}
struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;   // document order; a repeated key keeps its place and takes the last value
};

namespace detail {

inline void AppendUtf8(std::string* s, uint32_t cp) {
  if (cp < 0x80) s->push_back((char)cp);
  else if (cp < 0x800) { s->push_back((char)(0xC0 | (cp >> 6))); s->push_back((char)(0x80 | (cp & 0x3F))); }
  else if (cp < 0x10000) { s->push_back((char)(0xE0 | (cp >> 12))); s->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s->push_back((char)(0x80 | (cp & 0x3F))); }
  else { s->push_back((char)(0xF0 | (cp >> 18))); s->push_back((char)(0x80 | ((cp >> 12) & 0x3F))); s->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s->push_back((char)(0x80 | (cp & 0x3F))); }
}

// length of the well-formed UTF-8 sequence at p (0 = ill-formed): encoding/json replaces ill-formed bytes in strings by U+FFFD
inline int Utf8Len(const unsigned char* p, const unsigned char* end) {
  const unsigned c = p[0];
  if (c < 0x80) return 1;
  auto cont = [&](int i) { return p + i < end && (p[i] & 0xC0) == 0x80; };
  if (c >= 0xC2 && c <= 0xDF) return cont(1) ? 2 : 0;
  if (c >= 0xE0 && c <= 0xEF) {
    if (!cont(1) || !cont(2)) return 0;
    if (c == 0xE0 && p[1] < 0xA0) return 0;          // overlong
    if (c == 0xED && p[1] >= 0xA0) return 0;         // surrogates
    return 3;
  }
  if (c >= 0xF0 && c <= 0xF4) {
    if (!cont(1) || !cont(2) || !cont(3)) return 0;
    if (c == 0xF0 && p[1] < 0x90) return 0;
    if (c == 0xF4 && p[1] >= 0x90) return 0;
    return 4;
  }
  return 0;
}

class Parser {
 public:
  Parser(const std::string& text) : p_((const unsigned char*)text.data()), end_(p_ + text.size()) {}
  bool ParseDocument(Json* out) {
    SkipWs();
    if (!ParseValue(out, 0)) return false;
    SkipWs();
    return p_ == end_;
  }

 private:
  static constexpr int kMaxDepth = 10000;   // encoding/json: "exceeded max depth"
  void SkipWs() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++; }
  bool Literal(const char* lit) {
    const unsigned char* q = p_;
    for (; *lit; lit++, q++) if (q >= end_ || *q != (unsigned char)*lit) return false;
    p_ = q;
    return true;
  }
  bool ParseValue(Json* out, int depth) {
    if (p_ >= end_ || depth > kMaxDepth) return false;
    switch (*p_) {
      case 'n': out->kind = Json::Null; return Literal("null");
      case 't': out->kind = Json::Bool; out->b = true; return Literal("true");
      case 'f': out->kind = Json::Bool; out->b = false; return Literal("false");
      case '"': out->kind = Json::String; return ParseString(&out->str);
      case '[': return ParseArray(out, depth);
      case '{': return ParseObject(out, depth);
      default: return ParseNumber(out);
    }
  }
  bool ParseNumber(Json* out) {   // RFC 8259 grammar; the value must fit a float64 (Unmarshal fails otherwise)
    const unsigned char* s = p_;
    if (p_ < end_ && *p_ == '-') p_++;
    if (p_ >= end_) return false;
    if (*p_ == '0') p_++;
    else if (*p_ >= '1' && *p_ <= '9') { while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++; }
    else return false;
    if (p_ < end_ && *p_ == '.') {
      p_++;
      if (p_ >= end_ || *p_ < '0' || *p_ > '9') return false;
      while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
    }
    if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
      p_++;
      if (p_ < end_ && (*p_ == '+' || *p_ == '-')) p_++;
      if (p_ >= end_ || *p_ < '0' || *p_ > '9') return false;
      while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
    }
    double v = 0;
    auto r = std::from_chars((const char*)s, (const char*)p_, v);
    if (r.ec == std::errc::result_out_of_range) {
      // from_chars reports underflow the same way; strconv.ParseFloat rounds those to 0 without error
      const std::string t((const char*)s, (const char*)p_);
      const size_t e = t.find_first_of("eE");
      const bool tiny = e != std::string::npos && t[e + 1] == '-';
      if (!tiny) return false;
      v = 0;
    } else if (r.ec != std::errc() || r.ptr != (const char*)p_) return false;
    if (std::isinf(v)) return false;
    out->kind = Json::Number; out->num = v;
    return true;
  }
  bool Hex4(uint32_t* out) {
    if (end_ - p_ < 4) return false;
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) {
      const unsigned c = p_[i];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else return false;
    }
    p_ += 4; *out = v;
    return true;
  }
  bool ParseString(std::string* out) {
    out->clear();
    p_++;   // opening quote
    while (p_ < end_) {
      const unsigned c = *p_;
      if (c == '"') { p_++; return true; }
      if (c < 0x20) return false;   // control characters must be escaped
      if (c == '\\') {
        if (++p_ >= end_) return false;
        const unsigned e = *p_++;
        switch (e) {
          case '"': out->push_back('"'); break;
          case '\\': out->push_back('\\'); break;
          case '/': out->push_back('/'); break;
          case 'b': out->push_back('\b'); break;
          case 'f': out->push_back('\f'); break;
          case 'n': out->push_back('\n'); break;
          case 'r': out->push_back('\r'); break;
          case 't': out->push_back('\t'); break;
          case 'u': {
            uint32_t cp = 0;
            if (!Hex4(&cp)) return false;
            if (cp >= 0xD800 && cp <= 0xDBFF) {   // high surrogate: needs a low one right behind it, else U+FFFD
              uint32_t lo = 0;
              const unsigned char* save = p_;
              if (end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u' && (p_ += 2, Hex4(&lo)) && lo >= 0xDC00 && lo <= 0xDFFF)
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              else { p_ = save; cp = 0xFFFD; }
            } else if (cp >= 0xDC00 && cp <= 0xDFFF) cp = 0xFFFD;
            AppendUtf8(out, cp);
            break;
          }
          default: return false;
        }
        continue;
      }
      const int n = Utf8Len(p_, end_);
      if (n == 0) { AppendUtf8(out, 0xFFFD); p_++; continue; }
      out->append((const char*)p_, (size_t)n);
      p_ += n;
    }
    return false;   // unterminated
  }
  bool ParseArray(Json* out, int depth) {
    out->kind = Json::Array;
    p_++;
    SkipWs();
    if (p_ < end_ && *p_ == ']') { p_++; return true; }
    for (;;) {
      out->arr.emplace_back();
      SkipWs();
      if (!ParseValue(&out->arr.back(), depth + 1)) return false;
      SkipWs();
      if (p_ >= end_) return false;
      if (*p_ == ',') { p_++; continue; }
      if (*p_ == ']') { p_++; return true; }
      return false;
    }
  }
  bool ParseObject(Json* out, int depth) {
    out->kind = Json::Object;
    p_++;
    SkipWs();
    if (p_ < end_ && *p_ == '}') { p_++; return true; }
    std::unordered_map<std::string, size_t> index;   // duplicate keys: the last value wins, the first position stays (O(1) per key)
    for (;;) {
      SkipWs();
      if (p_ >= end_ || *p_ != '"') return false;
      std::string key;
      if (!ParseString(&key)) return false;
      SkipWs();
      if (p_ >= end_ || *p_ != ':') return false;
      p_++;
      SkipWs();
      Json val;
      if (!ParseValue(&val, depth + 1)) return false;
      auto found = index.find(key);
      if (found != index.end()) out->obj[found->second].second = std::move(val);
      else { index.emplace(key, out->obj.size()); out->obj.emplace_back(std::move(key), std::move(val)); }
      SkipWs();
      if (p_ >= end_) return false;
      if (*p_ == ',') { p_++; continue; }
      if (*p_ == '}') { p_++; return true; }
      return false;
    }
  }
  const unsigned char* p_;
  const unsigned char* end_;
};

}  // namespace detail

inline bool ParseJson(const std::string& text, Json* out) { return detail::Parser(text).ParseDocument(out); }

// fmt's %v of a float64 = strconv.FormatFloat(x, 'g', -1, 64): shortest round-trip digits; the %e form when the decimal
// exponent is < -4 or >= 6 (the precision used for that decision when the shortest form was asked for)
inline std::string GoFloat(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "+Inf" : "-Inf";
  if (x == 0) return std::signbit(x) ? "-0" : "0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // shortest digits: d[.ddd]e[+-]XX
  std::string s(buf, r.ptr);
  const bool neg = s[0] == '-';
  if (neg) s.erase(0, 1);
  const size_t e = s.find('e');
  std::string digits = s.substr(0, e);
  const int e10 = std::stoi(s.substr(e + 1));
  digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int nd = (int)digits.size(), dp = e10 + 1;
  std::string out;
  if (e10 < -4 || e10 >= 6) {
    out = digits.substr(0, 1);
    if (nd > 1) out += "." + digits.substr(1);
    const int a = e10 < 0 ? -e10 : e10;
    out += e10 < 0 ? "e-" : "e+";
    if (a < 10) out += "0";
    out += std::to_string(a);
  } else if (dp <= 0) out = "0." + std::string((size_t)-dp, '0') + digits;
  else if (dp >= nd) out = digits + std::string((size_t)(dp - nd), '0');
  else out = digits.substr(0, (size_t)dp) + "." + digits.substr((size_t)dp);
  return neg ? "-" + out : out;
}

// fmt.Sprintf("%v", v) for the dynamic types above (maps print with sorted keys, as Go >= 1.12 does)
inline std::string GoSprintV(const Json& v) {
  switch (v.kind) {
    case Json::Null: return "<nil>";
    case Json::Bool: return v.b ? "true" : "false";
    case Json::Number: return GoFloat(v.num);
    case Json::String: return v.str;
    case Json::Array: {
      std::string s = "[";
      for (size_t i = 0; i < v.arr.size(); i++) { if (i) s += " "; s += GoSprintV(v.arr[i]); }
      return s + "]";
    }
    case Json::Object: {
      std::vector<const std::pair<std::string, Json>*> items;
      for (auto& kv : v.obj) items.push_back(&kv);
      std::sort(items.begin(), items.end(), [](auto* a, auto* b) { return a->first < b->first; });
      std::string s = "map[";
      for (size_t i = 0; i < items.size(); i++) { if (i) s += " "; s += items[i]->first + ":" + GoSprintV(items[i]->second); }
      return s + "]";
    }
  }
  return "";
}

// The events PostMetric would publish for this body, in document order.  false: the body does not decode into a
// map[string]interface{} (the handler answers 422).  `null` decodes into a nil map: no events, status 200.
inline bool MetricEvents(const std::string& body, std::vector<Event>* out) {
  out->clear();
  Json doc;
  if (!ParseJson(body, &doc)) return false;
  if (doc.kind == Json::Null) return true;
  if (doc.kind != Json::Object) return false;
  for (auto& kv : doc.obj) out->push_back(Event{Metric, kv.first + "|" + GoSprintV(kv.second)});
  return true;
}

// Endpoints.PostMetric (control/endpoints.go:112-129): the request's events as one batch
inline int PostMetric(EventBus* bus, const std::string& body) {
  std::vector<Event> evs;
  if (!MetricEvents(body, &evs)) return StatusUnprocessableEntity;
  if (!evs.empty()) bus->PublishMany(evs);
  return StatusOK;
}

}  // namespace ingest
}  // namespace events
