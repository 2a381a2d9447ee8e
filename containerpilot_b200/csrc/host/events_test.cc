// events_test.cc — the reference's bus tests restated against the C++ mirror (events.hpp) on the CUDA bus.
// Each test cites the Go test it follows.  Exit code 0 = all passed.  Needs a GPU (libcpbus has no CPU fallback).
#include <cstdio>
#include <map>
#include <thread>

#include "events.hpp"

using namespace events;

static int failures = 0;
#define EXPECT(cond)                                                           \
  do {                                                                         \
    if (!(cond)) { std::printf("  FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)

static std::string Show(const std::vector<Event>& v) {
  std::string s = "[";
  for (auto& e : v) s += "{" + String(e.Code) + " " + e.Source + "} ";
  return s + "]";
}

struct TestPublisher : Publisher {   // events/events_test.go:13-21
  explicit TestPublisher(EventBus* bus) { Register(bus); }
};

struct TestSubscriber : Subscriber {   // events/events_test.go:23-61
  std::vector<Event> results;
  std::mutex lock;
  std::thread th;
  TestSubscriber() { Rx = MakeChan(100); }
  void Run(Context& ctx, EventBus* bus) {
    Subscribe(bus);
    th = std::thread([this, &ctx] {
      for (;;) {
        Event e;
        if (Rx->Recv(&e, std::chrono::milliseconds(2))) { std::lock_guard<std::mutex> l(lock); results.push_back(e); continue; }
        if (Rx->Closed() || ctx.Done()) {
          while (Rx->Recv(&e)) { std::lock_guard<std::mutex> l(lock); results.push_back(e); }
          break;
        }
      }
      Unsubscribe();   // deferred in Go: ts.Unsubscribe(); ts.Wait(); close(ts.Rx)
      Rx->Close();
    });
  }
  void Join() { if (th.joinable()) th.join(); }
};

// events/events_test.go:64-89
static void TestPubSubInterfaces() {
  std::printf("TestPubSubInterfaces\n");
  EventBus bus(EventBus::Clock::Monotonic);
  TestPublisher tp(&bus);
  TestSubscriber ts;
  Context ctx;
  ts.Run(ctx, &bus);
  std::vector<Event> expected{Event{Startup, "serviceA"}};
  for (auto& e : expected) tp.Publish(e);
  for (int spin = 0; spin < 2000; spin++) {
    { std::lock_guard<std::mutex> l(ts.lock); if (ts.results.size() == expected.size()) break; }
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  ctx.Cancel();
  auto results = bus.DebugEvents();
  EXPECT(results == expected);
  ts.Join();
  EXPECT(ts.results == expected);   // collected but never asserted in Go; asserted here
  tp.Unregister();
  EXPECT(bus.Wait() == false);
  if (results != expected) std::printf("  expected %s got %s\n", Show(expected).c_str(), Show(results).c_str());
}

// events/events_test.go:91-113
static void TestPublishSignal() {
  std::printf("TestPublishSignal\n");
  EventBus bus(EventBus::Clock::Monotonic);
  TestSubscriber ts;
  Context ctx;
  ts.Run(ctx, &bus);
  std::vector<std::string> signals{"SIGHUP", "SIGUSR2"};
  std::vector<Event> expected;
  for (auto& s : signals) { expected.push_back(Event{Signal, s}); bus.PublishSignal(s); }
  for (int spin = 0; spin < 2000; spin++) {
    { std::lock_guard<std::mutex> l(ts.lock); if (ts.results.size() == expected.size()) break; }
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  ctx.Cancel();
  auto results = bus.DebugEvents();
  EXPECT(results == expected);
  ts.Join();
  EXPECT(ts.results == expected);
}

// jobs/jobs_test.go:15-48: publishing after the only subscriber unsubscribed must not panic
static void TestJobRunSafeClose() {
  std::printf("TestJobRunSafeClose\n");
  EventBus bus(EventBus::Clock::Virtual);
  Subscriber job; job.Rx = MakeChan(1000);
  Publisher pub;
  job.Subscribe(&bus); pub.Register(&bus);
  bus.Publish(GlobalStartup);
  pub.Publish(Event{Stopping, "myjob"});     // jobs/jobs.go:390
  job.Unsubscribe(); pub.Unregister();       // jobs/jobs.go:411-412
  bus.Publish(Event{Stopped, "myjob"});      // jobs/jobs.go:415
  EXPECT(bus.Wait() == false);
  std::vector<Event> expected{GlobalStartup, Event{Stopping, "myjob"}, Event{Stopped, "myjob"}};
  EXPECT(bus.DebugEvents() == expected);
  bool panicked = false;
  try { pub.Bus->Publish(GlobalStartup); } catch (const Panic&) { panicked = true; }
  EXPECT(!panicked);
  std::vector<Event> got; Event e;
  while (job.Rx->Recv(&e)) got.push_back(e);
  EXPECT((got == std::vector<Event>{GlobalStartup, Event{Stopping, "myjob"}}));
}

// control/endpoints_test.go:103-145 (multiset through DebugEvents) + Metric excluded from the counter (bus.go:130-132)
static void TestPostMetricMultiset() {
  std::printf("TestPostMetricMultiset\n");
  EventBus bus(EventBus::Clock::Virtual);
  bus.Publish(Event{Metric, "mymetric|1.5"});
  bus.Publish(Event{Metric, "myothermetric|2"});
  bus.Publish(GlobalEnterMaintenance);
  std::map<Event, int> got;
  for (auto& e : bus.DebugEvents()) got[e]++;
  std::map<Event, int> want{{Event{Metric, "mymetric|1.5"}, 1}, {Event{Metric, "myothermetric|2"}, 1}, {GlobalEnterMaintenance, 1}};
  EXPECT(got == want);
  EXPECT(bus.CounterValue("Metric", "mymetric|1.5") == 0);
  EXPECT(bus.CounterValue("EnterMaintenance", "global") == 1);
}

// events/bus.go:135-137 (closed Rx panics), :121 (double Unsubscribe panics), :108 (type assertion)
static void TestPanics() {
  std::printf("TestPanics\n");
  EventBus bus(EventBus::Clock::Virtual);
  Subscriber sub; sub.Rx = MakeChan(10);
  sub.Subscribe(&bus);
  sub.Rx->Close();
  bool p = false;
  try { bus.Publish(GlobalStartup); } catch (const Panic&) { p = true; }
  EXPECT(p);
  sub.Unsubscribe();
  p = false;
  try { sub.Unsubscribe(); } catch (const Panic&) { p = true; }
  EXPECT(p);
  struct Other : EventSubscriber { void Subscribe(EventBus*) override {} void Unsubscribe() override {} void Receive(const Event&) override {} } other;
  p = false;
  try { bus.Subscribe(&other); } catch (const Panic&) { p = true; }
  EXPECT(p);
}

// events/timer.go:12-71 under the virtual clock; names as in jobs/jobs.go:147-158; direct Receive as watches_test.go:48-50
static void TestTimers() {
  std::printf("TestTimers\n");
  EventBus bus(EventBus::Clock::Virtual);
  Subscriber job, other;
  job.Rx = MakeChan(1000); other.Rx = MakeChan(1000);
  job.Subscribe(&bus); other.Subscribe(&bus);
  Context ctx;
  using namespace std::chrono_literals;
  NewEventTimer(ctx, job.Rx, 1s, "myjob.heartbeat");
  NewEventTimeout(ctx, job.Rx, 2500ms, "myjob.wait-timeout");
  bus.Publish(GlobalStartup);
  bus.Advance(3000000000ull);
  bus.Publish(Event{StatusHealthy, "myjob"});
  job.Receive(QuitByTest);
  Event hb{TimerExpired, "myjob.heartbeat"}, to{TimerExpired, "myjob.wait-timeout"};
  std::vector<Event> got, got2; Event e;
  while (job.Rx->Recv(&e)) got.push_back(e);
  while (other.Rx->Recv(&e)) got2.push_back(e);
  EXPECT((got == std::vector<Event>{GlobalStartup, hb, hb, to, hb, Event{StatusHealthy, "myjob"}, QuitByTest}));
  EXPECT((got2 == std::vector<Event>{GlobalStartup, Event{StatusHealthy, "myjob"}}));
  ctx.Cancel();
  bus.Advance(10000000000ull);
  bus.Publish(GlobalShutdown);
  got.clear();
  while (job.Rx->Recv(&e)) got.push_back(e);
  EXPECT((got == std::vector<Event>{GlobalShutdown}));
  EXPECT((bus.DebugEvents() == std::vector<Event>{GlobalStartup, Event{StatusHealthy, "myjob"}, GlobalShutdown}));
}

// watches/watches.go:65-101 as an actor on the mirror: the Watch keeps a PRIVATE channel that is never subscribed
// (watches.go:37), hands it to NewEventTimer (watches.go:71) and publishes Status* events when the backend reports a change.
struct NoopDiscoveryBackend {   // tests/mocks/discovery.go:6-22
  bool Val = false, lastVal = false;
  std::pair<bool, bool> CheckForUpstreamChanges() { bool changed = lastVal != Val; lastVal = Val; return {changed, Val}; }
};
struct Watch : Publisher {
  std::string Name; ChanPtr rx = MakeChan(1000); NoopDiscoveryBackend disc; Context ctx; std::thread th; std::string timerSource;
  void Run(EventBus* bus, std::chrono::nanoseconds poll) {   // watches.go:65-96
    Register(bus);
    timerSource = Name + ".poll";
    NewEventTimer(ctx, rx, poll, timerSource);   // any `chan Event`: this one is not a bus subscriber
    th = std::thread([this] {
      for (;;) {
        Event e;
        if (!rx->Recv(&e, std::chrono::milliseconds(2))) { if (ctx.Done()) break; continue; }
        if (e == QuitByTest) break;
        if (e == Event{TimerExpired, timerSource}) {
          auto [didChange, isHealthy] = disc.CheckForUpstreamChanges();
          if (didChange) {
            Publish(Event{StatusChanged, Name});
            Publish(Event{isHealthy ? StatusHealthy : StatusUnhealthy, Name});
          }
        }
      }
      ctx.Cancel(); Unregister();
    });
  }
  void Receive(const Event& e) { rx->Send(e); }   // watches.go:99-101: `watch.rx <- event`
};

// watches/watches_test.go:13-58
static std::map<Event, int> RunWatchTest(const std::string& name, bool val) {
  EventBus bus(EventBus::Clock::Monotonic);
  Watch watch; watch.Name = "watch." + name; watch.disc.Val = val;
  watch.Run(&bus, std::chrono::seconds(1));
  Event poll{TimerExpired, watch.Name + ".poll"};
  watch.Receive(poll);
  watch.Receive(poll);   // "Ensure we can run it more than once"
  watch.Receive(QuitByTest);
  watch.th.join();
  bus.Wait();
  std::map<Event, int> got;
  for (auto& e : bus.DebugEvents()) got[e]++;
  return got;
}
static void TestWatchPoll() {
  std::printf("TestWatchPoll\n");
  auto ok = RunWatchTest("mywatchOk", true);
  EXPECT((ok[Event{StatusChanged, "watch.mywatchOk"}] == 1 && ok[Event{StatusHealthy, "watch.mywatchOk"}] == 1));
  auto fail = RunWatchTest("mywatchFail", false);
  EXPECT((fail[Event{StatusChanged, "watch.mywatchFail"}] == 0 && fail[Event{StatusUnhealthy, "watch.mywatchFail"}] == 0));
}

// events/timer.go:40-71 on a channel nobody subscribed: real ticks under the virtual clock, interleaved with direct sends;
// broadcasts never land there; a late Subscribe keeps the mailbox; closing the channel ends the timer (timer.go:50-54)
static void TestTimerOnPrivateChannel() {
  std::printf("TestTimerOnPrivateChannel\n");
  EventBus bus(EventBus::Clock::Virtual);
  Subscriber other; other.Rx = MakeChan(1000); other.Subscribe(&bus);
  ChanPtr rx = MakeChan(1000);
  Context ctx;
  using namespace std::chrono_literals;
  NewEventTimer(ctx, rx, 1000ns, "w.poll");
  NewEventTimeout(ctx, rx, 2500ns, "w.once");
  bus.Advance(1000);
  bus.Publish(Event{Startup, "everyone"});
  bus.Advance(3000);
  Event tick{TimerExpired, "w.poll"}, once{TimerExpired, "w.once"}, e;
  std::vector<Event> got, got2;
  while (rx->Recv(&e)) got.push_back(e);
  while (other.Rx->Recv(&e)) got2.push_back(e);
  EXPECT((got == std::vector<Event>{tick, tick, once, tick}));
  EXPECT((got2 == std::vector<Event>{Event{Startup, "everyone"}}));
  ctx.Cancel();
  bus.Advance(6000);
  EXPECT(!rx->Recv(&e));
  Context ctx2;
  NewEventTimer(ctx2, rx, 1000ns, "w.again");
  Subscriber late; late.Rx = rx; late.Subscribe(&bus);   // the channel becomes a real subscriber: mailbox and timer are kept
  bus.Advance(7000);
  bus.Publish(Event{Signal, "SIGHUP"});
  got.clear();
  while (rx->Recv(&e)) got.push_back(e);
  EXPECT((got == std::vector<Event>{Event{TimerExpired, "w.again"}, Event{Signal, "SIGHUP"}}));
  ChanPtr rx2 = MakeChan(10);
  NewEventTimer(ctx2, rx2, 1000ns, "w.closed");
  rx2->Close();
  bus.Advance(9000);   // the tick meets a closed channel: the implicit mailbox is released, nothing panics
  bus.Advance(10000);
  ctx2.Cancel();
  late.Unsubscribe(); other.Unsubscribe();
}

// events/events.go:52-86 and eventcode_string.go:9-15
static void TestNames() {
  std::printf("TestNames\n");
  EXPECT(FromString("exitSuccess").first == ExitSuccess && FromString("SIGUSR2").first == Signal && FromString("changed").first == StatusChanged);
  auto bad = FromString("bogus");
  EXPECT(bad.first == None && bad.second == "bogus is not a valid event code");
  EXPECT(String(StatusUnhealthy) == "StatusUnhealthy" && String((EventCode)42) == "EventCode(42)");
}

// 10,000 events, 8 subscribers with consumer threads on real channels (BASELINE config 1, Go-path analogue)
static void TestConfig1Plumbing() {
  std::printf("TestConfig1Plumbing\n");
  EventBus bus(EventBus::Clock::Monotonic);
  const int N = 8, E = 10000;
  std::vector<std::unique_ptr<TestSubscriber>> subs;
  Context ctx;
  for (int i = 0; i < N; i++) { subs.emplace_back(new TestSubscriber()); subs.back()->Run(ctx, &bus); }
  Publisher pub; pub.Register(&bus);
  std::vector<Event> sent;
  for (int i = 0; i < E; i++) { Event e{(EventCode)(1 + i % 16), "src" + std::to_string(i % 64)}; sent.push_back(e); pub.Publish(e); }
  for (int spin = 0; spin < 10000; spin++) {   // the consumers run concurrently; give them up to 10 s
    bool all = true;
    for (auto& s : subs) { std::lock_guard<std::mutex> l(s->lock); all = all && s->results.size() == sent.size(); }
    if (all) break;
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  ctx.Cancel();
  for (auto& s : subs) { s->Join(); EXPECT(s->results == sent); }
  pub.Unregister();
  EXPECT(bus.Wait() == false);
}

// many subscribers through the bulk drain bridge (cpbus_drain_many): 300 mailboxes, 600 events, virtual clock
static void TestManySubscribers() {
  std::printf("TestManySubscribers\n");
  EventBus bus(EventBus::Clock::Virtual, 512);
  const int N = 300, E = 600;
  std::vector<std::unique_ptr<Subscriber>> subs;
  for (int i = 0; i < N; i++) { subs.emplace_back(new Subscriber()); subs.back()->Rx = MakeChan(1000); subs.back()->Subscribe(&bus); }
  std::vector<Event> sent;
  for (int i = 0; i < E; i++) { Event e{(EventCode)(1 + i % 16), "s" + std::to_string(i % 37)}; sent.push_back(e); bus.Publish(e); }
  bool ok = true;
  for (auto& s : subs) {
    std::vector<Event> got; Event e;
    while (s->Rx->Recv(&e)) got.push_back(e);
    ok = ok && got == sent;
  }
  EXPECT(ok);
  for (auto& s : subs) s->Unsubscribe();
  EXPECT(bus.Wait() == false);
}

// a Job-shaped consumer subscribes with the exact cases of its switch (jobs/jobs.go:197-231)
static void TestSubscribeWithCases() {
  std::printf("TestSubscribeWithCases\n");
  EventBus bus(EventBus::Clock::Virtual, 16);
  Subscriber job, all;
  job.Rx = MakeChan(1000); all.Rx = MakeChan(1000);
  const std::vector<Event> cases{{ExitSuccess, "check.web"}, {ExitFailed, "check.web"}, {Quit, "web"}, GlobalShutdown,
                                 {Signal, "SIGHUP"}, {StatusHealthy, "watch.db"}};
  job.Subscribe(&bus, 0u, cases);
  all.Subscribe(&bus);
  const std::vector<Event> stream{{ExitSuccess, "check.web"}, {ExitSuccess, "check.api"}, {StatusHealthy, "watch.db"},
                                  {StatusHealthy, "watch.cache"}, {Metric, "m|1"}, {Signal, "SIGHUP"}, {Signal, "SIGTERM"},
                                  {Quit, "web"}, {Quit, "api"}, GlobalShutdown};
  std::vector<Event> want;
  for (auto& e : stream) {
    bus.Publish(e);
    if (std::find(cases.begin(), cases.end(), e) != cases.end()) want.push_back(e);
  }
  std::vector<Event> got, got_all; Event e;
  while (job.Rx->Recv(&e)) got.push_back(e);
  while (all.Rx->Recv(&e)) got_all.push_back(e);
  EXPECT(got == want);
  EXPECT(got_all == stream);
  job.Unsubscribe(); all.Unsubscribe();
  EXPECT(bus.Wait() == false);
}

int main() {
  TestNames();
  TestPubSubInterfaces();
  TestPublishSignal();
  TestJobRunSafeClose();
  TestPostMetricMultiset();
  TestPanics();
  TestTimers();
  TestWatchPoll();
  TestTimerOnPrivateChannel();
  TestConfig1Plumbing();
  TestManySubscribers();
  TestSubscribeWithCases();
  std::printf(failures ? "FAILED (%d)\n" : "PASS\n", failures);
  return failures ? 1 : 0;
}
