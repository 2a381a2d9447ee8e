// cpbus_kernels.cuh — sm_100a kernels of the event bus hot path.
//
// Replaces the inner loop of EventBus.Publish (reference events/bus.go:134-138:
// `for subscriber := range bus.registry { subscriber.Receive(event) }`, one
// runtime.chansend per subscriber per event, events/subscriber.go:30-32) and the
// per-timer goroutines of events/timer.go:12-71, for a whole batch of events and
// all subscribers of this GPU's shard in one launch.
//
// Shape of the work: pure integer / byte movement, HBM-write bound.  No tensor
// cores.  One warp owns one subscriber (mailbox) at a time; the batch of 32-byte
// records is staged once per CTA into shared memory with a 1-D TMA bulk copy
// (cp.async.bulk + mbarrier); matches are found with warp ballots; every record
// is written as one full, aligned 32-byte sector (st.global.v8.b32 or a v4 pair)
// or, for dense runs, by TMA bulk stores straight out of the staged batch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/cpbus.h"

namespace cpbus_dev {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr uint32_t kActiveBit = 0x80000000u;   // mask word: subscriber is subscribed
constexpr int kTimerHintShift = 24;            // mask word bits 24..27: #timer slots to look at
constexpr uint32_t kTimerActive = 1u, kTimerOneshot = 2u;
constexpr uint64_t kDigestP = 0x9E3779B97F4A7C15ull;

struct __align__(32) DevTimer {     // one timer slot (events/timer.go: one goroutine + ticker)
  uint64_t next_due;
  uint64_t period;
  uint32_t source_id;
  uint32_t fired;
  uint32_t flags;
  uint32_t pad;
};

struct DevStats {
  unsigned long long deliveries, ticks, overwritten, admit_overflow;
};

struct FanoutParams {
  const cpbus_event* batch;   // n_ev records, sorted by ts (HBM)
  cpbus_event* ring;          // [n_subs][R]
  unsigned long long* tail;   // records ever delivered, per subscriber
  unsigned long long* head;   // consumer cursor
  unsigned long long* digest;
  const uint32_t* mask;
  DevTimer* timers;           // [n_subs][K] or nullptr
  DevStats* stats;
  uint64_t w_now;             // watermark: timers due <= w_now fire in this launch
  uint32_t n_ev, n_subs, ring_cap, K, sub_base;
  uint32_t use_digest, lossless, timers_on;
  uint32_t smem_cap;          // n_ev rounded up to 32 (shared-memory carve-up)
};

// ---------------------------------------------------------------- helpers ---
__host__ __device__ inline uint64_t record_hash_words(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3) {
  const uint64_t K0 = 0x9E3779B97F4A7C15ull, K1 = 0xBF58476D1CE4E5B9ull,
                 K2 = 0x94D049BB133111EBull, K3 = 0xD6E8FEB86659FD93ull,
                 K4 = 0xA0761D6478BD642Full;
  uint64_t x = (w0 + K4) * K0; x ^= x >> 32;
  x = (x + w1) * K1; x ^= x >> 32;
  x = (x + w2) * K2; x ^= x >> 32;
  x = (x + w3) * K3; x ^= x >> 29;
  return x;
}

__host__ __device__ inline uint64_t pow_p(uint32_t e) {
  uint64_t r = 1, b = kDigestP;
  while (e) { if (e & 1u) r *= b; b *= b; e >>= 1; }
  return r;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, o);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), o);
    v += ((uint64_t)hi << 32) | lo;
  }
  return v;
}
// one full 32-byte sector per instruction (SASS: STG.E.ENL2.256)
__device__ __forceinline__ void st_v8(void* dst, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void st_v4(void* dst, const uint4& a) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
}
template <int STORE>
__device__ __forceinline__ void st_record(cpbus_event* dst, const uint4& a, const uint4& b) {
  if (STORE == CPBUS_STORE_V8) st_v8(dst, a, b);
  else { st_v4(dst, a); st_v4(reinterpret_cast<unsigned char*>(dst) + 16, b); }
}
// TMA 1-D bulk copies (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "CPBUS_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CPBUS_DONE;\n\t"
      "bra CPBUS_WAIT;\n\t"
      "CPBUS_DONE:\n\t}" ::"r"(smem_u32(mbar)),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------- the kernel ---
// Shared memory carve-up (cap = smem_cap records):
//   [0, 32cap)            staged batch (TMA destination)
//   [32cap, 40cap)        record hashes H(e_i)
//   [40cap, 48cap)        {codebit, target} per event
//   [48cap, 56cap + 520)  powers of the digest multiplier P^0 .. P^(cap+64)
//   then                  mbarrier, batch summary, per-warp tick scratch
struct BatchSummary {
  uint64_t mbar;
  uint64_t hfull;          // digest of the whole batch taken as one dense run
  uint32_t present;        // OR of codebits of the broadcast events
  uint32_t has_unicast;    // any record with a specific target
  uint64_t red[kWarpsPerCta];
};

__host__ __device__ inline size_t fanout_smem_bytes(uint32_t cap) {
  return (size_t)cap * 56 + (64 + 1) * 8 + sizeof(BatchSummary) + kWarpsPerCta * 32 * sizeof(uint32_t) + 128;
}

template <int STORE>
__global__ void __launch_bounds__(kThreads) fanout_kernel(const FanoutParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t cap = p.smem_cap;
  cpbus_event* s_batch = reinterpret_cast<cpbus_event*>(smem);
  uint64_t* s_rhash = reinterpret_cast<uint64_t*>(smem + (size_t)cap * 32);
  uint2* s_meta = reinterpret_cast<uint2*>(smem + (size_t)cap * 40);
  uint64_t* s_pow = reinterpret_cast<uint64_t*>(smem + (size_t)cap * 48);
  BatchSummary* s_sum = reinterpret_cast<BatchSummary*>(smem + (size_t)cap * 56 + 65 * 8);
  uint32_t* s_tick = reinterpret_cast<uint32_t*>(s_sum + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = p.n_ev;

  // ---- stage the batch: one elected thread drives the TMA engine ----
  if (tid == 0) {
    s_sum->present = 0; s_sum->has_unicast = 0; s_sum->hfull = 0;
    mbar_init(&s_sum->mbar, 1);
  }
  __syncthreads();
  if (n) {
    if (tid == 0) {
      mbar_expect_tx(&s_sum->mbar, n * 32u);
      bulk_g2s(s_batch, p.batch, n * 32u, &s_sum->mbar);
    }
    // powers of P do not depend on the batch: compute them while the copy is in flight
    for (uint32_t i = tid; i < cap + 65; i += kThreads) s_pow[i] = pow_p(i);
    mbar_wait(&s_sum->mbar, 0);
  } else {
    for (uint32_t i = tid; i < cap + 65; i += kThreads) s_pow[i] = pow_p(i);
  }

  // ---- per-batch precompute, once per CTA ----
  {
    uint32_t present = 0, uni = 0;
    for (uint32_t i = tid; i < n; i += kThreads) {
      const ulonglong4 w = *reinterpret_cast<const ulonglong4*>(&s_batch[i]);
      s_rhash[i] = record_hash_words(w.x, w.y, w.z, w.w);
      const uint32_t code = (uint32_t)w.z, target = (uint32_t)w.w;
      uint32_t codebit = 0;
      if (target == CPBUS_TARGET_ALL) { codebit = code < 32 ? (1u << code) : 0u; present |= codebit; }
      else uni = 1;
      s_meta[i] = make_uint2(codebit, target);
    }
    present = __reduce_or_sync(0xffffffffu, present);
    uni = __reduce_or_sync(0xffffffffu, uni);
    if (lane == 0) { if (present) atomicOr(&s_sum->present, present); if (uni) atomicOr(&s_sum->has_unicast, 1u); }
  }
  __syncthreads();
  if (p.use_digest) {   // H(batch as a dense run) = sum r_i * P^(n-1-i)
    uint64_t part = 0;
    for (uint32_t i = tid; i < n; i += kThreads) part += s_rhash[i] * s_pow[n - 1 - i];
    part = warp_sum64(part);
    if (lane == 0) s_sum->red[warp] = part;
    __syncthreads();
    if (tid == 0) { uint64_t t = 0; for (int w = 0; w < kWarpsPerCta; w++) t += s_sum->red[w]; s_sum->hfull = t; }
    __syncthreads();
  }
  const uint32_t present = s_sum->present;
  const bool has_unicast = s_sum->has_unicast != 0;
  const uint64_t hfull = s_sum->hfull;
  const uint32_t Rm = p.ring_cap - 1;
  const uint4* s4 = reinterpret_cast<const uint4*>(s_batch);
  uint32_t* my_tick = s_tick + warp * 32;

  unsigned long long acc_deliv = 0, acc_ticks = 0, acc_over = 0;
  bool bulk_pending = false;

  const uint32_t wstride = gridDim.x * kWarpsPerCta;
  for (uint32_t s = blockIdx.x * kWarpsPerCta + warp; s < p.n_subs; s += wstride) {
    const uint32_t m = p.mask[s];
    if (!(m & kActiveBit)) continue;
    const uint64_t tail = p.tail[s];
    cpbus_event* ring = p.ring + (size_t)s * p.ring_cap;
    const uint32_t gid = p.sub_base + s;
    const uint32_t nslots = p.timers_on ? min((m >> kTimerHintShift) & 0xFu, p.K) : 0u;
    // dense <=> this mailbox takes every record of the batch (the reference's only mode)
    const bool dense = !has_unicast && ((m & present) == present);

    // ---- timers: which ticks fire in (previous watermark, w_now] ----
    uint32_t n_ticks = 0;
    bool tk_valid = false; uint64_t tk_due = 0; uint32_t tk_slot = 0, tk_j = 0, tk_src = 0, tk_fired = 0, tk_rank = 0;
    uint32_t tk_mask = 0;
    if (nslots) {
      const uint32_t J = 32u / p.K;            // candidate firings per slot handled per launch (host bounds the window)
      tk_slot = lane / J; tk_j = lane % J;
      uint64_t due0 = 0, period = 0; uint32_t fl = 0;
      if (tk_slot < nslots) {
        const DevTimer t = p.timers[(size_t)s * p.K + tk_slot];
        due0 = t.next_due; period = t.period; fl = t.flags; tk_src = t.source_id; tk_fired = t.fired;
      }
      tk_due = due0 + (uint64_t)tk_j * period;
      tk_valid = (fl & kTimerActive) && tk_due <= p.w_now && (tk_j == 0 || !(fl & kTimerOneshot));
      tk_mask = __ballot_sync(0xffffffffu, tk_valid);
      n_ticks = __popc(tk_mask);
      if (n_ticks) {
        // order simultaneous firings by (due, slot): rank = #valid ticks with a smaller key
        if (J == 32 || (tk_mask >> J) == 0) tk_rank = tk_j;       // only slot 0 fired
        else {
          for (int t = 0; t < 32; t++) {
            if (!((tk_mask >> t) & 1u)) continue;      // warp-uniform
            const uint64_t od = shfl64(tk_due, t);
            const uint32_t os = __shfl_sync(0xffffffffu, tk_slot, t);
            tk_rank += (od < tk_due || (od == tk_due && os < tk_slot)) ? 1u : 0u;
          }
        }
      }
    }

    if (dense && n_ticks == 0) {
      // ================= dense run: copy the staged batch into the ring =================
      if (STORE == CPBUS_STORE_BULK) {
        if (lane == 0 && n) {
          const uint32_t slot0 = (uint32_t)tail & Rm;
          const uint32_t first = min(n, p.ring_cap - slot0);
          bulk_s2g(ring + slot0, s_batch, first * 32u);
          if (n > first) bulk_s2g(ring, s_batch + first, (n - first) * 32u);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        bulk_pending = true;
      } else if (STORE == CPBUS_STORE_V8) {
        for (uint32_t i = lane; i < n; i += 32) {
          const uint4 a = s4[2 * i], b = s4[2 * i + 1];
          st_v8(ring + (((uint32_t)tail + i) & Rm), a, b);
        }
      } else {
        for (uint32_t q = lane; q < 2 * n; q += 32) {   // lane pair per record: 512 contiguous bytes per instruction
          const uint4 v = s4[q];
          st_v4(reinterpret_cast<unsigned char*>(ring + (((uint32_t)tail + (q >> 1)) & Rm)) + (q & 1u) * 16u, v);
        }
      }
      if (lane == 0) {
        const uint64_t nt = tail + n;
        p.tail[s] = nt;
        if (p.use_digest) p.digest[s] = p.digest[s] * s_pow[n] + hfull;
        if (!p.lossless && nt > p.ring_cap) {
          const uint64_t h = p.head[s], floor_h = nt - p.ring_cap;
          if (h < floor_h) { p.head[s] = floor_h; acc_over += floor_h - h; }
        }
        acc_deliv += n;
      }
      continue;
    }

    // ================= general run: filter and/or interleaved ticks =================
    const uint32_t nchunks = (n + 31) >> 5;
    // pass A: match bitmap, 32 events per ballot; lane c keeps the word of chunk c
    uint32_t myword = 0;
    if (dense) {
      if ((uint32_t)lane < nchunks) myword = ((uint32_t)lane == nchunks - 1 && (n & 31u)) ? ((1u << (n & 31u)) - 1u) : 0xffffffffu;
    } else {
      for (uint32_t c = 0; c < nchunks; c++) {
        const uint32_t i = c * 32 + lane;
        bool match = false;
        if (i < n) {
          const uint2 mt = s_meta[i];
          match = (mt.y == CPBUS_TARGET_ALL) ? ((m & mt.x) != 0) : (mt.y == gid);
        }
        const uint32_t w = __ballot_sync(0xffffffffu, match);
        if ((uint32_t)lane == c) myword = w;
      }
    }
    // exclusive prefix of popcounts over chunks: wprefix(lane c) = matches before chunk c
    uint32_t wcount = __popc(myword), wprefix = wcount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wprefix, o);
      if (lane >= o) wprefix += t;
    }
    const uint32_t k_ev = __shfl_sync(0xffffffffu, wprefix, 31);
    wprefix -= wcount;

    // ticks: position among the matched events, then output slot
    uint32_t tk_mp = 0;
    if (n_ticks) {
      uint32_t pos = 0;
      if (tk_valid) {                         // lower_bound: events with ts < due stay in front of the tick
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s_batch[mid].ts_ns < tk_due) lo = mid + 1; else hi = mid;
        }
        pos = lo;
      }
      const uint32_t pc = pos >> 5;           // chunk of the first event behind the tick
      const uint32_t wsel = __shfl_sync(0xffffffffu, myword, pc & 31);
      const uint32_t psel = __shfl_sync(0xffffffffu, wprefix, pc & 31);
      tk_mp = (pos >= n) ? k_ev : psel + __popc(wsel & ((1u << (pos & 31u)) - 1u));
      if (tk_valid) my_tick[tk_rank] = tk_mp;
      __syncwarp();
    }
    const uint32_t k = k_ev + n_ticks;

    uint64_t dacc = 0;
    // pass B: every matched event goes to output index (matched rank + ticks in front of it)
    for (uint32_t c = 0; c < nchunks; c++) {
      const uint32_t w = __shfl_sync(0xffffffffu, myword, c);
      const uint32_t wp = __shfl_sync(0xffffffffu, wprefix, c);
      if ((w >> lane) & 1u) {
        const uint32_t i = c * 32 + lane;
        const uint32_t mrank = wp + __popc(w & ((1u << lane) - 1u));
        uint32_t out = mrank;
        for (uint32_t t = 0; t < n_ticks; t++) out += (my_tick[t] <= mrank) ? 1u : 0u;
        const uint4 a = s4[2 * i], b = s4[2 * i + 1];
        st_record<STORE>(ring + (((uint32_t)tail + out) & Rm), a, b);
        if (p.use_digest) dacc += s_rhash[i] * s_pow[k - 1 - out];
      }
    }
    if (tk_valid) {   // the tick records themselves: {TimerExpired, name} (events/timer.go:31,60)
      const uint32_t out = tk_mp + tk_rank;
      const uint64_t w0 = (uint64_t)tk_fired + tk_j, w1 = tk_due;
      const uint64_t w2 = (uint64_t)CPBUS_TIMER_EXPIRED | ((uint64_t)tk_src << 32);
      const uint64_t w3 = (uint64_t)gid | ((uint64_t)CPBUS_F_TICK << 32);
      const uint4 a = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
      const uint4 b = make_uint4((uint32_t)w2, (uint32_t)(w2 >> 32), (uint32_t)w3, (uint32_t)(w3 >> 32));
      st_record<STORE>(ring + (((uint32_t)tail + out) & Rm), a, b);
      if (p.use_digest) dacc += record_hash_words(w0, w1, w2, w3) * s_pow[k - 1 - out];
    }
    if (n_ticks) {   // re-arm: one lane per slot writes its timer back
      const uint32_t J = 32u / p.K;
      const uint32_t slotmask = (J == 32 ? 0xffffffffu : ((1u << J) - 1u)) << (tk_slot * J);
      const uint32_t fired_here = __popc(tk_mask & slotmask);
      if (tk_j == 0 && tk_slot < nslots && fired_here) {
        DevTimer* t = &p.timers[(size_t)s * p.K + tk_slot];
        const uint32_t fl = t->flags;
        if (fl & kTimerOneshot) t->flags = fl & ~kTimerActive;
        else t->next_due = t->next_due + (uint64_t)fired_here * t->period;
        t->fired = tk_fired + fired_here;
      }
      __syncwarp();
    }
    if (p.use_digest) dacc = warp_sum64(dacc);
    if (lane == 0 && k) {
      const uint64_t nt = tail + k;
      p.tail[s] = nt;
      if (p.use_digest) p.digest[s] = p.digest[s] * s_pow[k] + dacc;
      if (!p.lossless && nt > p.ring_cap) {
        const uint64_t h = p.head[s], floor_h = nt - p.ring_cap;
        if (h < floor_h) { p.head[s] = floor_h; acc_over += floor_h - h; }
      }
      acc_deliv += k; acc_ticks += n_ticks;
    }
  }

  if (STORE == CPBUS_STORE_BULK && bulk_pending && lane == 0)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staged batch must outlive the TMA reads
  if (lane == 0) {
    if (acc_deliv) atomicAdd(&p.stats->deliveries, acc_deliv);
    if (acc_ticks) atomicAdd(&p.stats->ticks, acc_ticks);
    if (acc_over) atomicAdd(&p.stats->overwritten, acc_over);
  }
}

// Lossless mode (reference semantics, events/subscriber.go:30-32: a full channel
// blocks the sender): before a batch is fanned out, count for every mailbox what
// the batch would append and flag any that lacks the room.  Thread per subscriber.
__global__ void admit_kernel(const cpbus_event* batch, uint32_t n_ev, uint64_t w_now, const uint32_t* mask,
                             const unsigned long long* tail, const unsigned long long* head, const DevTimer* timers,
                             uint32_t n_subs, uint32_t ring_cap, uint32_t K, uint32_t sub_base, uint32_t timers_on,
                             DevStats* stats) {
  __shared__ uint32_t hist[32];
  __shared__ uint32_t s_uni;
  if (threadIdx.x < 32) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_uni = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_ev; i += blockDim.x) {
    const uint32_t code = batch[i].code, target = batch[i].target;
    if (target == CPBUS_TARGET_ALL) { if (code < 32) atomicAdd(&hist[code], 1u); }
    else atomicAdd(&s_uni, 1u);
  }
  __syncthreads();
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_subs) return;
  const uint32_t m = mask[s];
  if (!(m & kActiveBit)) return;
  uint64_t k = 0;
  for (uint32_t c = 0; c < CPBUS_N_CODES; c++) if ((m >> c) & 1u) k += hist[c];
  if (s_uni) {
    const uint32_t gid = sub_base + s;
    for (uint32_t i = 0; i < n_ev; i++) if (batch[i].target == gid) k++;
  }
  const uint32_t nslots = timers_on ? min((m >> kTimerHintShift) & 0xFu, K) : 0u;
  for (uint32_t t = 0; t < nslots; t++) {
    const DevTimer tm = timers[(size_t)s * K + t];
    if ((tm.flags & kTimerActive) && tm.next_due <= w_now)
      k += (tm.flags & kTimerOneshot) ? 1u : (w_now - tm.next_due) / tm.period + 1u;
  }
  if (tail[s] - head[s] + k > ring_cap) atomicAdd(&stats->admit_overflow, 1ull);
}

// (count, digest) folds over a range of mailboxes: one 32-byte result instead of 16 B per subscriber
__global__ void digest_fold_kernel(const unsigned long long* tail, const unsigned long long* digest, uint32_t first,
                                   uint32_t n, uint32_t sub_base, unsigned long long* out4) {
  unsigned long long c = 0, d = 0, x = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long t = tail[first + i], g = digest[first + i];
    c += t; d += g;
    x ^= record_hash_words(g, t, sub_base + first + i, 0);
  }
  c = warp_sum64(c); d = warp_sum64(d);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)x, o), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(x >> 32), o);
    x ^= ((unsigned long long)hi << 32) | lo;
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out4[0], c); atomicAdd(&out4[1], d); atomicXor(&out4[2], x); }
  if (blockIdx.x == 0 && threadIdx.x == 0) out4[3] = n;
}
#endif  // __CUDACC__

}  // namespace cpbus_dev
