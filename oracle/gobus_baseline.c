/*
 * gobus_baseline.c — TIMED CPU BASELINE (test/bench infrastructure only).
 *
 * The reference bus is Go and cannot be built in this image (no Go toolchain;
 * SURVEY.md F1).  This file restates the *cost model* of its publish path so a
 * CPU number can be printed beside the GPU number.  It is a
 * reported baseline, not the optimisation target and not a checker.
 *
 * What is restated, per delivery:
 *   EventBus.Publish   events/bus.go:125-140   exclusive bus lock, then for every
 *                                              registered subscriber:
 *   Subscriber.Receive events/subscriber.go:30 `sub.Rx <- event` = runtime.chansend
 *       on a buffered channel: take the channel lock, copy the 24-byte Event
 *       {int Code; string Source(ptr,len)} into buf[sendx], bump sendx/qcount,
 *       release the lock;
 *   consumer           jobs/jobs.go:173 `<-job.Rx` = runtime.chanrecv: take the
 *       lock, copy the 24 bytes out, bump recvx/qcount, release; then the
 *       consumer-side switch (jobs/jobs.go:195-233) looks at the code.
 * Mailbox capacity 1000 (jobs/jobs.go:23).  When a mailbox is full the publisher
 * blocks and the consumers run; with GOMAXPROCS(1) (main.go:19) both sides share
 * one core, which is what n_threads == 1 models.  n_threads > 1 shards the
 * subscribers over threads, each doing its shard's sends and receives — more
 * parallelism than the Go bus has (its lock serialises publishers), i.e. a
 * generous baseline.
 */
#define _GNU_SOURCE                /* pthread_setaffinity_np, sched_getaffinity; pthread_barrier_t, clock_gettime */
#include "cpbus_oracle.h"

#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct go_event { int64_t code; const char* src; int64_t len; } go_event; /* 24 B */

typedef struct go_chan {           /* runtime.hchan, the fields chansend touches */
  atomic_flag lock;
  uint32_t qcount, dataqsiz, sendx, recvx;
  go_event* buf;
} go_chan;

typedef struct shard {
  go_chan** subs; uint32_t n; uint32_t n_events, cap; uint64_t checksum, deliveries;
  const go_event* trace;
} shard;

static inline void ch_lock(go_chan* c) { while (atomic_flag_test_and_set_explicit(&c->lock, memory_order_acquire)) {} }
static inline void ch_unlock(go_chan* c) { atomic_flag_clear_explicit(&c->lock, memory_order_release); }

static inline int chansend(go_chan* c, const go_event* e) {
  ch_lock(c);
  if (c->qcount == c->dataqsiz) { ch_unlock(c); return 0; }   /* would block */
  c->buf[c->sendx] = *e;
  if (++c->sendx == c->dataqsiz) c->sendx = 0;
  c->qcount++;
  ch_unlock(c);
  return 1;
}
static inline int chanrecv(go_chan* c, go_event* out) {
  ch_lock(c);
  if (c->qcount == 0) { ch_unlock(c); return 0; }
  *out = c->buf[c->recvx];
  if (++c->recvx == c->dataqsiz) c->recvx = 0;
  c->qcount--;
  ch_unlock(c);
  return 1;
}

static void* run_shard(void* arg) {
  shard* sh = (shard*)arg;
  uint64_t sum = 0, deliv = 0;
  for (uint32_t i = 0; i < sh->n_events; i++) {
    const go_event* e = &sh->trace[i];
    for (uint32_t s = 0; s < sh->n; s++) {
      go_chan* c = sh->subs[s];
      while (!chansend(c, e)) {
        /* publisher blocked: the consumer goroutines get the core and drain */
        for (uint32_t d = 0; d < sh->n; d++) {
          go_event got;
          while (chanrecv(sh->subs[d], &got)) { sum += (uint64_t)got.code * 31u + (uint64_t)got.len; deliv++; }
        }
      }
    }
  }
  for (uint32_t d = 0; d < sh->n; d++) {
    go_event got;
    while (chanrecv(sh->subs[d], &got)) { sum += (uint64_t)got.code * 31u + (uint64_t)got.len; deliv++; }
  }
  sh->checksum = sum; sh->deliveries = deliv;
  return NULL;
}

double gobus_bench(uint32_t n_subs, uint32_t n_events, uint32_t mailbox_cap,
                   uint32_t n_threads, uint64_t* checksum_out) {
  if (!n_subs || !n_events || !mailbox_cap) return 0.0;
  if (n_threads == 0) n_threads = 1;
  if (n_threads > n_subs) n_threads = n_subs;
  static const char* const SRC[4] = { "global", "myjob", "SIGHUP", "watch.backend" };
  go_event* trace = (go_event*)malloc((size_t)n_events * sizeof(go_event));
  uint64_t x = 0xC0DEB200ull;
  for (uint32_t i = 0; i < n_events; i++) {
    x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z ^= z >> 27;
    trace[i].code = 1 + (int64_t)(z % 16); trace[i].src = SRC[(z >> 8) & 3]; trace[i].len = (int64_t)strlen(trace[i].src);
  }
  go_chan** all = (go_chan**)malloc((size_t)n_subs * sizeof(go_chan*));
  for (uint32_t s = 0; s < n_subs; s++) {       /* one heap object per subscriber, like Go */
    go_chan* c = (go_chan*)calloc(1, sizeof(go_chan));
    c->dataqsiz = mailbox_cap; c->buf = (go_event*)malloc((size_t)mailbox_cap * sizeof(go_event));
    memset(c->buf, 0, (size_t)mailbox_cap * sizeof(go_event));
    atomic_flag_clear(&c->lock);
    all[s] = c;
  }
  shard* sh = (shard*)calloc(n_threads, sizeof(shard));
  pthread_t* th = (pthread_t*)malloc(n_threads * sizeof(pthread_t));
  uint32_t per = n_subs / n_threads, extra = n_subs % n_threads, off = 0;
  for (uint32_t t = 0; t < n_threads; t++) {
    sh[t].n = per + (t < extra); sh[t].subs = all + off; off += sh[t].n;
    sh[t].n_events = n_events; sh[t].cap = mailbox_cap; sh[t].trace = trace;
  }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (n_threads == 1) run_shard(&sh[0]);
  else {
    for (uint32_t t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, run_shard, &sh[t]);
    for (uint32_t t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  uint64_t sum = 0, deliv = 0;
  for (uint32_t t = 0; t < n_threads; t++) { sum += sh[t].checksum; deliv += sh[t].deliveries; }
  if (checksum_out) *checksum_out = sum ^ deliv;
  for (uint32_t s = 0; s < n_subs; s++) { free(all[s]->buf); free(all[s]); }
  free(all); free(sh); free(th); free(trace);
  return sec > 0 ? (double)deliv / sec : 0.0;
}

/* ---- step-structured run for `bench.py --impl reference`: channels and threads are created ONCE; every step
 *      publishes `events_per_step` events to every subscriber (shards in parallel, a barrier between steps);
 *      `warmup` steps are untimed.  Returns deliveries per second over the timed steps. ---- */
typedef struct step_shard {
  go_chan** subs; uint32_t n, events_per_step, steps, warmup; const go_event* trace; uint32_t trace_len;
  pthread_barrier_t* bar; struct timespec* t0; struct timespec* t1; int lead; uint64_t checksum, deliveries;
} step_shard;

static void* run_step_shard(void* arg) {
  step_shard* sh = (step_shard*)arg;
  uint64_t sum = 0, deliv = 0;
  uint32_t pos = 0;
  for (uint32_t st = 0; st < sh->warmup + sh->steps; st++) {
    if (st == sh->warmup) {
      pthread_barrier_wait(sh->bar);
      if (sh->lead) clock_gettime(CLOCK_MONOTONIC, sh->t0);
      deliv = 0;
    }
    for (uint32_t i = 0; i < sh->events_per_step; i++) {
      const go_event* e = &sh->trace[pos]; if (++pos == sh->trace_len) pos = 0;
      for (uint32_t s = 0; s < sh->n; s++) {
        go_chan* c = sh->subs[s];
        while (!chansend(c, e)) {
          for (uint32_t d = 0; d < sh->n; d++) { go_event got; while (chanrecv(sh->subs[d], &got)) { sum += (uint64_t)got.code * 31u + (uint64_t)got.len; deliv++; } }
        }
      }
    }
    /* the consumers run at the end of every step: nothing is left queued across the timing boundary */
    for (uint32_t d = 0; d < sh->n; d++) { go_event got; while (chanrecv(sh->subs[d], &got)) { sum += (uint64_t)got.code * 31u + (uint64_t)got.len; deliv++; } }
    pthread_barrier_wait(sh->bar);
  }
  if (sh->lead) clock_gettime(CLOCK_MONOTONIC, sh->t1);
  sh->checksum = sum; sh->deliveries = deliv;
  return NULL;
}

double gobus_bench_steps(uint32_t n_subs, uint32_t events_per_step, uint32_t steps, uint32_t warmup, uint32_t mailbox_cap,
                         uint32_t n_threads, double* seconds_out) {
  if (!n_subs || !events_per_step || !steps || !mailbox_cap) return 0.0;
  if (n_threads == 0) n_threads = 1;
  if (n_threads > n_subs) n_threads = n_subs;
  static const char* const SRC[4] = { "global", "myjob", "SIGHUP", "watch.backend" };
  const uint32_t trace_len = 4096;
  go_event* trace = (go_event*)malloc((size_t)trace_len * sizeof(go_event));
  uint64_t x = 0xC0DEB200ull;
  for (uint32_t i = 0; i < trace_len; i++) {
    x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z ^= z >> 27;
    trace[i].code = 1 + (int64_t)(z % 16); trace[i].src = SRC[(z >> 8) & 3]; trace[i].len = (int64_t)strlen(trace[i].src);
  }
  go_chan** all = (go_chan**)malloc((size_t)n_subs * sizeof(go_chan*));
  for (uint32_t s = 0; s < n_subs; s++) {
    go_chan* c = (go_chan*)calloc(1, sizeof(go_chan));
    c->dataqsiz = mailbox_cap; c->buf = (go_event*)calloc(mailbox_cap, sizeof(go_event));
    atomic_flag_clear(&c->lock);
    all[s] = c;
  }
  pthread_barrier_t bar; pthread_barrier_init(&bar, NULL, n_threads);
  struct timespec t0, t1;
  step_shard* sh = (step_shard*)calloc(n_threads, sizeof(step_shard));
  pthread_t* th = (pthread_t*)malloc(n_threads * sizeof(pthread_t));
  uint32_t per = n_subs / n_threads, extra = n_subs % n_threads, off = 0;
  for (uint32_t t = 0; t < n_threads; t++) {
    sh[t].n = per + (t < extra); sh[t].subs = all + off; off += sh[t].n;
    sh[t].events_per_step = events_per_step; sh[t].steps = steps; sh[t].warmup = warmup; sh[t].trace = trace; sh[t].trace_len = trace_len;
    sh[t].bar = &bar; sh[t].t0 = &t0; sh[t].t1 = &t1; sh[t].lead = (t == 0);
  }
  for (uint32_t t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, run_step_shard, &sh[t]);
  run_step_shard(&sh[0]);
  for (uint32_t t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
  double sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  uint64_t deliv = 0;
  for (uint32_t t = 0; t < n_threads; t++) deliv += sh[t].deliveries;
  if (seconds_out) *seconds_out = sec;
  pthread_barrier_destroy(&bar);
  for (uint32_t s = 0; s < n_subs; s++) { free(all[s]->buf); free(all[s]); }
  free(all); free(sh); free(th); free(trace);
  return sec > 0 ? (double)deliv / sec : 0.0;
}

/* ---- round 2: `bench.py --impl reference`.  Same model as gobus_bench_steps, made repeatable across boxes:
 *  - every worker thread is pinned to one CPU of the process's affinity mask and allocates (first-touches) its own
 *    shard's channels, so the mailboxes are NUMA-local to the thread that uses them (round 1 swung 4x box to box);
 *  - per-step wall times are returned (lead thread, between barriers), so the caller can report a median;
 *  - send_only != 0: no consumer runs and a full mailbox overwrites its oldest entry — the GPU arm's throughput mode,
 *    i.e. the CPU is NOT charged for chanrecv.  send_only == 0 is the reference's semantics (lossless, consumers drain).
 * Returns deliveries per second over the timed steps (send_only: sends; otherwise receives). ---- */
typedef struct step2_shard {
  uint32_t n, events_per_step, steps, warmup, cap; const go_event* trace; uint32_t trace_len;
  pthread_barrier_t* bar; double* step_sec; int lead, cpu, send_only; uint64_t checksum, deliveries;
} step2_shard;

static double now_sec(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

static void* run_step2_shard(void* arg) {
  step2_shard* sh = (step2_shard*)arg;
  if (sh->cpu >= 0) { cpu_set_t set; CPU_ZERO(&set); CPU_SET(sh->cpu, &set); pthread_setaffinity_np(pthread_self(), sizeof(set), &set); }
  go_chan** subs = (go_chan**)malloc((size_t)sh->n * sizeof(go_chan*));
  for (uint32_t s = 0; s < sh->n; s++) {       /* one heap object per subscriber, like Go; touched by its own thread */
    go_chan* c = (go_chan*)calloc(1, sizeof(go_chan));
    c->dataqsiz = sh->cap; c->buf = (go_event*)malloc((size_t)sh->cap * sizeof(go_event));
    memset(c->buf, 0, (size_t)sh->cap * sizeof(go_event));
    atomic_flag_clear(&c->lock);
    subs[s] = c;
  }
  uint64_t sum = 0, deliv = 0;
  uint32_t pos = 0;
  pthread_barrier_wait(sh->bar);
  for (uint32_t st = 0; st < sh->warmup + sh->steps; st++) {
    const double t0 = sh->lead ? now_sec() : 0.0;
    if (st == sh->warmup) deliv = 0;
    for (uint32_t i = 0; i < sh->events_per_step; i++) {
      const go_event* e = &sh->trace[pos]; if (++pos == sh->trace_len) pos = 0;
      for (uint32_t s = 0; s < sh->n; s++) {
        go_chan* c = subs[s];
        if (sh->send_only) {
          ch_lock(c);
          c->buf[c->sendx] = *e;
          if (++c->sendx == c->dataqsiz) c->sendx = 0;
          if (c->qcount < c->dataqsiz) c->qcount++; else if (++c->recvx == c->dataqsiz) c->recvx = 0;   /* overwrite oldest */
          ch_unlock(c);
          deliv++;
        } else {
          while (!chansend(c, e))
            for (uint32_t d = 0; d < sh->n; d++) { go_event got; while (chanrecv(subs[d], &got)) { sum += (uint64_t)got.code * 31u + (uint64_t)got.len; deliv++; } }
        }
      }
    }
    if (!sh->send_only)   /* the consumers run at the end of every step: nothing is left queued across the timing boundary */
      for (uint32_t d = 0; d < sh->n; d++) { go_event got; while (chanrecv(subs[d], &got)) { sum += (uint64_t)got.code * 31u + (uint64_t)got.len; deliv++; } }
    pthread_barrier_wait(sh->bar);
    if (sh->lead) sh->step_sec[st] = now_sec() - t0;
  }
  for (uint32_t s = 0; s < sh->n; s++) { sum += subs[s]->qcount; free(subs[s]->buf); free(subs[s]); }
  free(subs);
  sh->checksum = sum; sh->deliveries = deliv;
  return NULL;
}

double gobus_bench_steps2(uint32_t n_subs, uint32_t events_per_step, uint32_t steps, uint32_t warmup, uint32_t mailbox_cap,
                          uint32_t n_threads, int send_only, double* step_seconds_out /* warmup + steps entries */) {
  if (!n_subs || !events_per_step || !steps || !mailbox_cap || !step_seconds_out) return 0.0;
  if (n_threads == 0) n_threads = 1;
  if (n_threads > n_subs) n_threads = n_subs;
  static const char* const SRC[4] = { "global", "myjob", "SIGHUP", "watch.backend" };
  const uint32_t trace_len = 4096;
  go_event* trace = (go_event*)malloc((size_t)trace_len * sizeof(go_event));
  uint64_t x = 0xC0DEB200ull;
  for (uint32_t i = 0; i < trace_len; i++) {
    x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z ^= z >> 27;
    trace[i].code = 1 + (int64_t)(z % 16); trace[i].src = SRC[(z >> 8) & 3]; trace[i].len = (int64_t)strlen(trace[i].src);
  }
  cpu_set_t avail; CPU_ZERO(&avail);
  int cpus[4096], n_cpus = 0;
  if (sched_getaffinity(0, sizeof(avail), &avail) == 0)
    for (int c = 0; c < CPU_SETSIZE && n_cpus < 4096; c++) if (CPU_ISSET(c, &avail)) cpus[n_cpus++] = c;
  pthread_barrier_t bar; pthread_barrier_init(&bar, NULL, n_threads);
  step2_shard* sh = (step2_shard*)calloc(n_threads, sizeof(step2_shard));
  pthread_t* th = (pthread_t*)malloc(n_threads * sizeof(pthread_t));
  const uint32_t per = n_subs / n_threads, extra = n_subs % n_threads;
  for (uint32_t t = 0; t < n_threads; t++) {
    sh[t].n = per + (t < extra); sh[t].events_per_step = events_per_step; sh[t].steps = steps; sh[t].warmup = warmup;
    sh[t].cap = mailbox_cap; sh[t].trace = trace; sh[t].trace_len = trace_len; sh[t].bar = &bar; sh[t].step_sec = step_seconds_out;
    sh[t].lead = (t == 0); sh[t].cpu = n_cpus ? cpus[t % (uint32_t)n_cpus] : -1; sh[t].send_only = send_only;
  }
  for (uint32_t t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, run_step2_shard, &sh[t]);
  run_step2_shard(&sh[0]);
  for (uint32_t t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
  cpu_set_t restore = avail;                                   /* thread 0 was the caller: give it its mask back */
  if (n_cpus) pthread_setaffinity_np(pthread_self(), sizeof(restore), &restore);
  double sec = 0.0;
  for (uint32_t st = warmup; st < warmup + steps; st++) sec += step_seconds_out[st];
  uint64_t deliv = 0;
  for (uint32_t t = 0; t < n_threads; t++) deliv += sh[t].deliveries;
  pthread_barrier_destroy(&bar);
  free(sh); free(th); free(trace);
  return sec > 0 ? (double)deliv / sec : 0.0;
}
