/*
 * cpbus_oracle.h — CPU oracle for the ContainerPilot `events` bus hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call it, and only as the checker or the
 * timed CPU baseline.  libcpbus never links or calls it.
 *
 * What it is: a plain-C restatement, event at a time, of the reference's Go
 * code in /root/reference/events/ (bus.go, subscriber.go, timer.go, events.go,
 * eventcode_string.go).  Every function cites the lines it follows.
 *
 * Pinning status: the reference is Go 1.9 and there is no Go toolchain in this
 * image, so the reference itself cannot be run here.  The oracle IS pinned
 * against every known-answer vector the reference's own tests hold for this
 * path (tests/golden/reference_vectors.json, transcribed with file:line from
 * events/events_test.go, jobs/jobs_test.go, core/signals_test.go,
 * control/endpoints_test.go, commands/commands_test.go, watches/watches_test.go,
 * events/events.go, events/eventcode_string.go).  Those vectors pin the publish
 * order as seen through DebugEvents and the enum tables.  No reference test
 * asserts the per-subscriber received sequence, so for that
 * property parity is UNPINNED beyond the code reading (bus.go:126-138 holds an
 * exclusive lock and channels are FIFO => per-subscriber order == publish order).
 *
 * Determinism rules where Go is nondeterministic:
 *  - cross-subscriber order inside one Publish (map iteration, bus.go:134) is
 *    not observable: only per-subscriber sequences are compared;
 *  - time is virtual: orc_advance(now) plays the role of the runtime clock.  A
 *    timer firing due at d is appended to its owner's mailbox at the first
 *    advance with now >= d, i.e. after every event published before that call
 *    and before every event published after it.  Simultaneous firings of one
 *    subscriber's timers are ordered by (due, slot).
 */
#ifndef CPBUS_ORACLE_H
#define CPBUS_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_event {   /* same 32-byte layout as cpbus_event */
  uint64_t seq, ts_ns;
  uint32_t code, source_id, target, flags;
} orc_event;

typedef struct orc_bus orc_bus;

#define ORC_OK 0
#define ORC_EINVAL (-1)
#define ORC_EAGAIN (-4)
#define ORC_ENOSPC (-5)
#define ORC_ENOENT (-6)
#define ORC_ECLOSED (-7)

/* keep_window: 0 = keep every delivered record per subscriber (small cases);
 *              W > 0 = keep only the last W records (+ count + digest).
 * mailbox_cap: 0 = unbounded; C > 0 = a send into a mailbox holding C
 *              unconsumed records "blocks" (the call returns ORC_EAGAIN and
 *              delivers to nobody: the Go publisher would sit in chansend,
 *              events/subscriber.go:31, until the consumer runs). */
orc_bus* orc_new(uint32_t n_max_subs, uint32_t timers_per_sub, uint32_t keep_window,
                 uint32_t mailbox_cap, uint32_t sub_id_base);
void orc_free(orc_bus*);

int orc_subscribe(orc_bus*, uint32_t mask, uint32_t* sub_id);
/* mask (any source) OR one of up to ORC_MAX_PAIRS exact {code, source_id} pairs (jobs/jobs.go:188-231) */
#define ORC_MAX_PAIRS 16
int orc_subscribe_pairs(orc_bus*, uint32_t mask, const uint32_t* codes, const uint32_t* sources, uint32_t n_pairs,
                        uint32_t* sub_id);
int orc_unsubscribe(orc_bus*, uint32_t sub_id);
int orc_register(orc_bus*);
int orc_unregister(orc_bus*);
void orc_set_reload(orc_bus*);
int orc_wait(orc_bus*);          /* 0/1 = reload flag; -1 = would block (done counter > 0) */

int orc_publish(orc_bus*, uint32_t code, uint32_t source_id);
int orc_publish_many(orc_bus*, const uint32_t* codes, const uint32_t* sources, size_t n,
                     uint64_t dt_ns /* advance by dt before each publish; 0 = none */);
/* complete records (cpbus_publish_device / CPBUS_PUT_RAW semantics): clock -> each ts, deliver verbatim, clock -> watermark */
int orc_publish_records(orc_bus*, const orc_event* recs, size_t n, uint64_t watermark_ns);
int orc_receive(orc_bus*, uint32_t sub_id, uint32_t code, uint32_t source_id);
int orc_advance(orc_bus*, uint64_t now_ns);
int orc_timer_add(orc_bus*, uint32_t sub_id, uint64_t period_ns, uint32_t source_id, int oneshot, uint32_t* timer_id);
int orc_timer_cancel(orc_bus*, uint32_t timer_id);

size_t orc_debug_events(orc_bus*, orc_event* out, size_t cap);

uint64_t orc_count(orc_bus*, uint32_t sub_id);
uint64_t orc_digest(orc_bus*, uint32_t sub_id);
/* retained records, oldest first (all of them when keep_window == 0) */
size_t orc_mailbox(orc_bus*, uint32_t sub_id, orc_event* out, size_t cap);
/* consume up to cap records FIFO (`<-sub.Rx`); returns the number consumed */
size_t orc_consume(orc_bus*, uint32_t sub_id, orc_event* out, size_t cap);
uint64_t orc_now(orc_bus*);
uint64_t orc_total_deliveries(orc_bus*);
uint64_t orc_total_ticks(orc_bus*);
uint64_t orc_published_by_code(orc_bus*, uint32_t code);

const char* orc_code_name(int code);
int orc_code_from_string(const char* name);
uint64_t orc_record_hash(const orc_event* ev);
uint64_t orc_digest_multiplier(void);

/* ---- timed CPU baseline (gobus_baseline.c): a restatement of the Go bus's
 *      cost model, NOT a checker.  Returns deliveries per second. ---- */
double gobus_bench(uint32_t n_subs, uint32_t n_events, uint32_t mailbox_cap,
                   uint32_t n_threads, uint64_t* checksum_out);
/* step-structured variant (threads/channels created once; `warmup` untimed steps, then `steps` timed ones) */
double gobus_bench_steps(uint32_t n_subs, uint32_t events_per_step, uint32_t steps, uint32_t warmup, uint32_t mailbox_cap,
                         uint32_t n_threads, double* seconds_out);

/* round 2: pinned threads, NUMA-local mailboxes, per-step times (warmup + steps doubles), optional send-only
 * (overwrite-oldest, no consumer: what the GPU arm's throughput mode does) */
double gobus_bench_steps2(uint32_t n_subs, uint32_t events_per_step, uint32_t steps, uint32_t warmup, uint32_t mailbox_cap,
                          uint32_t n_threads, int send_only, double* step_seconds_out);

#ifdef __cplusplus
}
#endif
#endif
