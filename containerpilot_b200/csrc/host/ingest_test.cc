// CPU-only test of the /v3/metric event construction (ingest.hpp): no bus, no GPU.  Vectors: the reference's own
// TestPostMetric (control/endpoints_test.go:104-145) plus the documented behaviour of fmt's %v and encoding/json.
#include <cstdio>
#include <cstdlib>
#include <map>

#include "ingest.hpp"

using namespace events;
using namespace events::ingest;

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("  FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

static std::map<std::string, int> Multiset(const std::string& body, bool* ok) {
  std::vector<Event> evs;
  *ok = MetricEvents(body, &evs);
  std::map<std::string, int> m;
  for (auto& e : evs) { EXPECT(e.Code == Metric); m[e.Source]++; }
  return m;
}

// `ingest_test --events`: one request body per stdin line -> "status<TAB>source<TAB>source..." per line
// (used by tests/test_ingest.py to compare this mirror with containerpilot_b200/ingest.py on random documents)
static int EventsMode() {
  char* line = nullptr; size_t cap = 0; ssize_t n;
  while ((n = getline(&line, &cap, stdin)) >= 0) {
    std::string body(line, (size_t)n);
    if (!body.empty() && body.back() == '\n') body.pop_back();
    std::vector<Event> evs;
    const bool ok = MetricEvents(body, &evs);
    std::printf("%d", ok ? StatusOK : StatusUnprocessableEntity);
    for (auto& e : evs) { std::fputc('\t', stdout); std::fwrite(e.Source.data(), 1, e.Source.size(), stdout); }
    std::fputc('\n', stdout);
  }
  free(line);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--events") return EventsMode();
  bool ok = false;
  // endpoints_test.go:125-144
  EXPECT(Multiset("{{\n", &ok).empty() && !ok);
  EXPECT((Multiset("{\"mymetric\": 1.0}", &ok) == std::map<std::string, int>{{"mymetric|1", 1}}) && ok);
  EXPECT((Multiset("{\"mymetric\": 1.5, \"myothermetric\": 2}", &ok) ==
          std::map<std::string, int>{{"mymetric|1.5", 1}, {"myothermetric|2", 1}}) && ok);

  // fmt %v of float64
  const std::pair<double, const char*> floats[] = {
      {1.0, "1"}, {2, "2"}, {1.5, "1.5"}, {-3.25, "-3.25"}, {0, "0"}, {100000.0, "100000"}, {123456.0, "123456"},
      {1000000.0, "1e+06"}, {1234567.0, "1.234567e+06"}, {123456789.0, "1.23456789e+08"}, {1e21, "1e+21"},
      {0.0001, "0.0001"}, {0.00001, "1e-05"}, {0.000012345, "1.2345e-05"}, {0.1, "0.1"}, {2.5e-7, "2.5e-07"},
      {1e100, "1e+100"}, {4611686018427387904.0, "4.611686018427388e+18"}, {49.75, "49.75"}};
  for (auto& f : floats) {
    if (GoFloat(f.first) != f.second) { std::printf("  FAIL GoFloat(%g) = %s, want %s\n", f.first, GoFloat(f.first).c_str(), f.second); failures++; }
  }

  // other dynamic types
  auto sprint = [](const char* json) { Json j; return ParseJson(json, &j) ? GoSprintV(j) : std::string("<parse error>"); };
  EXPECT(sprint("\"up\"") == "up");
  EXPECT(sprint("true") == "true" && sprint("false") == "false" && sprint("null") == "<nil>");
  EXPECT(sprint("[1, 2.5, \"x\"]") == "[1 2.5 x]" && sprint("[]") == "[]");
  EXPECT(sprint("{\"b\": 1, \"a\": [true]}") == "map[a:[true] b:1]");
  EXPECT(sprint(" \t\n{ \"k\" : { } } \r\n") == "map[k:map[]]");

  // bodies the Go decoder rejects (422) or accepts without events
  for (const char* bad : {"", "[1, 2]", "3", "\"x\"", "{\"a\": NaN}", "{\"a\": Infinity}", "{\"a\": 1e999}", "\xff\xfe",
                          "{\"a\": 1,}", "{\"a\": 01}", "{\"a\": 1.}", "{\"a\": .5}", "{\"a\": +1}", "{a: 1}", "{\"a\" 1}",
                          "{\"a\": \"\n\"}", "{\"a\": \"\\x\"}", "{\"a\": \"\\u12\"}", "{\"a\": 1} x", "{\"a\": tru}"}) {
    Multiset(bad, &ok);
    if (ok) { std::printf("  FAIL accepted: %s\n", bad); failures++; }
  }
  EXPECT(Multiset("null", &ok).empty() && ok);
  EXPECT(Multiset("{}", &ok).empty() && ok);
  EXPECT(Multiset("{\"a\": 1e-999}", &ok) == (std::map<std::string, int>{{"a|0", 1}}) && ok);   // underflow rounds to 0

  // duplicate keys: the last value wins, the key keeps its place; '|' inside values is not special
  std::vector<Event> evs;
  EXPECT(MetricEvents("{\"a\": 1, \"a\": 2, \"b\": \"x|y\"}", &evs));
  EXPECT(evs.size() == 2 && evs[0].Source == "a|2" && evs[1].Source == "b|x|y");

  // strings: escapes, surrogate pairs, lone surrogates and ill-formed UTF-8 become U+FFFD like encoding/json
  EXPECT(sprint("\"a\\n\\t\\\"\\\\\\/\\u0041\"") == "a\n\t\"\\/A");
  EXPECT(sprint("\"\\ud83d\\ude00\"") == "\xF0\x9F\x98\x80");
  EXPECT(sprint("\"\\ud800z\"") == "\xEF\xBF\xBDz");
  EXPECT(sprint("\"\\ud800\\u0041\"") == "\xEF\xBF\xBD" "A");
  EXPECT(sprint("\"\\udc00\"") == "\xEF\xBF\xBD");
  EXPECT(sprint("\"x\xfe\"") == "x\xEF\xBF\xBD");
  EXPECT(sprint("\"\xe2\x82\xac\"") == "\xe2\x82\xac");

  std::printf(failures ? "FAILED (%d)\n" : "PASS\n", failures);
  return failures ? 1 : 0;
}
