/* abi_smoke.c — the C-ABI of libcpbus used from PLAIN C, the way cgo-generated code calls it:
 * include/cpbus.h only, no C++, no Python.  Publishes the reference's own test sequence
 * (jobs/jobs_test.go:15-48: GlobalStartup, {Stopping,"myjob"}, {Stopped,"myjob"}) to two subscribers — one with the
 * reference's behaviour (all-ones mask), one with a code mask — with a periodic timer on the first, drains the mailboxes
 * and checks sequences, DebugEvents and the {code, source} publish counts.  Exit code 0 = all checks passed.
 * Build: gcc -std=c99 -Wall -Wextra -Werror -pedantic -I include tests/c/abi_smoke.c -L containerpilot_b200 -lcpbus */
#include <stdio.h>
#include <string.h>

#include "cpbus.h"

static int failures = 0;
#define EXPECT(c) do { if (!(c)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)
#define OK(call) do { int rc_ = (call); if (rc_ != CPBUS_OK) { printf("FAIL %s:%d: %s -> %s %s\n", __FILE__, __LINE__, #call, cpbus_strerror(rc_), cpbus_last_cuda_error()); return 1; } } while (0)

static uint32_t intern(cpbus_t* bus, const char* s) {
  uint32_t id = 0;
  if (cpbus_intern(bus, s, strlen(s), &id) != CPBUS_OK) failures++;
  return id;
}

int main(void) {
  cpbus_config cfg;
  cpbus_t* bus = NULL;
  uint32_t all = 0, masked = 0, timer = 0, g, job, hb;
  cpbus_event ev, got[64];
  size_t n = 0, i;
  uint64_t lost = 0;
  cpbus_pair_count counts[16];
  cpbus_stats_t st;

  memset(&cfg, 0, sizeof(cfg));
  cfg.n_max_subs = 8; cfg.ring_cap = 1024; cfg.batch_cap = 256; cfg.timers_per_sub = 2;
  cfg.flags = CPBUS_CFG_LOSSLESS | CPBUS_CFG_DIGEST; cfg.device = -1;
  if (cpbus_abi_version() != 2) { printf("FAIL: ABI version %u\n", cpbus_abi_version()); return 1; }
  OK(cpbus_create(&cfg, &bus));
  g = intern(bus, "global"); job = intern(bus, "myjob"); hb = intern(bus, "myjob.heartbeat");
  OK(cpbus_subscribe(bus, CPBUS_MASK_ALL, &all));
  OK(cpbus_subscribe(bus, (1u << CPBUS_STOPPED) | (1u << CPBUS_SHUTDOWN), &masked));
  OK(cpbus_timer_add(bus, all, 1000, hb, 0, &timer));                 /* NewEventTimer(ctx, job.Rx, 1us, "myjob.heartbeat") */

  memset(&ev, 0, sizeof(ev));
  ev.code = CPBUS_STARTUP; ev.source_id = g;   OK(cpbus_publish(bus, &ev, 1));
  OK(cpbus_advance(bus, 2500));                                       /* two ticks are due: 1000, 2000 */
  ev.code = CPBUS_STOPPING; ev.source_id = job; OK(cpbus_publish(bus, &ev, 1));
  ev.code = CPBUS_STOPPED; ev.source_id = job;  OK(cpbus_publish(bus, &ev, 1));
  OK(cpbus_timer_cancel(bus, timer));
  EXPECT(cpbus_timer_cancel(bus, timer) == CPBUS_ENOENT);            /* generation-checked id */
  OK(cpbus_flush(bus)); OK(cpbus_sync(bus));

  OK(cpbus_drain(bus, all, got, 64, &n, &lost));
  EXPECT(n == 5 && lost == 0);
  if (n == 5) {
    EXPECT(got[0].code == CPBUS_STARTUP && got[0].source_id == g);
    EXPECT(got[1].code == CPBUS_TIMER_EXPIRED && got[1].source_id == hb && (got[1].flags & CPBUS_F_TICK) && got[1].ts_ns == 1000);
    EXPECT(got[2].code == CPBUS_TIMER_EXPIRED && got[2].ts_ns == 2000 && got[2].seq == 1);
    EXPECT(got[3].code == CPBUS_STOPPING && got[3].source_id == job);
    EXPECT(got[4].code == CPBUS_STOPPED && got[4].source_id == job && got[4].seq == 2);
  }
  OK(cpbus_drain(bus, masked, got, 64, &n, &lost));
  EXPECT(n == 1 && got[0].code == CPBUS_STOPPED);                    /* the filter kept one of three; ticks are unicast to `all` */

  OK(cpbus_debug_events(bus, got, 10, &n));                          /* jobs/jobs_test.go:40-47: the ordered expectation */
  EXPECT(n == 3 && got[0].code == CPBUS_STARTUP && got[1].code == CPBUS_STOPPING && got[2].code == CPBUS_STOPPED);

  OK(cpbus_publish_counts(bus, counts, 16, &n));                     /* containerpilot_events{code, source} (bus.go:131) */
  EXPECT(n == 3);
  for (i = 0; i < n && i < 16; i++) EXPECT(counts[i].count == 1);
  OK(cpbus_stats(bus, &st));
  EXPECT(st.publishes == 3 && st.deliveries == 6 && st.ticks == 2 && st.published_by_code[CPBUS_STOPPED] == 1);
  EXPECT(strcmp(cpbus_code_name(CPBUS_STATUS_CHANGED), "StatusChanged") == 0 && cpbus_code_from_string("SIGUSR2") == CPBUS_SIGNAL);
  OK(cpbus_destroy(bus));
  printf(failures ? "FAILED (%d)\n" : "PASS\n", failures);
  return failures ? 1 : 0;
}
